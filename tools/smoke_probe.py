"""How far is a tiny ResNet-50 training step (the shape of __graft_entry__.smoke()) from the fp32 oracle, next to PyTorch's own
bf16 autocast on the same weights and input?  Train-mode BatchNorm over a handful of values per channel (8 images x 2x2 pixels in
layer4 for 64x64 inputs) amplifies bf16 rounding; this table is what the smoke tolerance / shape was chosen from.
python tools/smoke_probe.py   (B200_RESNET_ALGEBRA=0 for the plain BatchNorm schedule)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import torchvision
from deeplearning_b200.classification.resnet.models.networks import resnet50
from oracle.resnet import train_step_grads

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
print(f"algebra={os.environ.get('B200_RESNET_ALGEBRA', '1')}")
for (B, S, seed) in ((8, 64, 1), (8, 64, 3), (8, 64, 5), (8, 96, 1), (8, 128, 1), (16, 64, 1), (16, 96, 1), (16, 128, 1), (16, 128, 3)):
    torch.manual_seed(0)
    m = resnet50()
    state = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(seed))
    y = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(seed + 1))
    ref_logits, ref_loss, _ = train_step_grads({k: v.clone() for k, v in state.items()}, x, y)
    mg = m.cuda().train()
    out = mg(x.cuda())
    loss = F.cross_entropy(out, y.cuda())
    loss.backward()
    ref = torchvision.models.ResNet(torchvision.models.resnet.Bottleneck, [3, 4, 6, 3]).cuda().train()
    ref.load_state_dict(state)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = ref(x.cuda()).float()
    amp_loss = F.cross_entropy(o, y.cuda())
    e_l = (out.detach().float().cpu() - ref_logits).abs().max()
    a_l = (o.detach().cpu() - ref_logits).abs().max()
    print(f"B={B:3d} {S:3d}x{S:<3d} seed {seed}: loss {float(loss):.4f} oracle {float(ref_loss):.4f} |d| {abs(float(loss) - float(ref_loss)):.4f}"
          f"  (torch autocast |d| {abs(float(amp_loss) - float(ref_loss)):.4f});  logits max|err| {float(e_l):.3f} (autocast {float(a_l):.3f})")
