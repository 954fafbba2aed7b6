"""Debug: gradients of a shallow Bottleneck ResNet on the algebra path / plain path vs the fp32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from deeplearning_b200.classification.resnet.models.networks import Bottleneck, ResNet
from oracle.resnet import train_step_grads

layers = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2,2,1,1").split(",")]
B, hw = 32, 128
x = torch.randn(B, 3, hw, hw, generator=torch.Generator().manual_seed(1))
y = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(2))
torch.manual_seed(0)
m0 = ResNet(Bottleneck, layers)
state = {k: v.clone() for k, v in m0.state_dict().items()}
ref_logits, ref_loss, ref_grads = train_step_grads({k: v.clone() for k, v in state.items()}, x, y)
res = {}
for mode in ("1", "0"):
    os.environ["B200_RESNET_ALGEBRA"] = mode
    m = ResNet(Bottleneck, layers)
    m.load_state_dict(state)
    m = m.cuda().train()
    out = m(x.cuda())
    loss = F.cross_entropy(out, y.cuda())
    loss.backward()
    res[mode] = {n: p.grad.detach().float().cpu() for n, p in m.named_parameters()}
    print("mode", mode, "loss", float(loss), "ref", float(ref_loss), "logit err", float((out.detach().float().cpu() - ref_logits).abs().max()))
names = [n for n in ref_grads]
for n in names:
    r = ref_grads[n]
    ea = float((res["1"][n] - r).norm() / (r.norm() + 1e-12))
    eb = float((res["0"][n] - r).norm() / (r.norm() + 1e-12))
    if max(ea, eb) > 0.03 or n.endswith("conv1.weight"):
        print(f"{n:34s} algebra {ea:.4f}  plain {eb:.4f}")
