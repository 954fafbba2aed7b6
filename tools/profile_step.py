"""One ResNet-50 training step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).
python tools/profile_step.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplearning_b200.engine.trainer import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else "resnet50"
torch.manual_seed(0)
if which == "resnet50":
    from deeplearning_b200.classification.resnet.models.networks import resnet50
    m = resnet50().cuda().train()
elif which == "convnext_tiny":
    from deeplearning_b200.classification.convNext.models.networks import ConvNeXt
    m = ConvNeXt(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], num_classes=1000, drop_path_rate=0.0).cuda().train()
elif which == "swin_tiny":
    from deeplearning_b200.classification.swin_transformer.models.swin_transformer import SwinTransformer
    m = SwinTransformer(drop_path_rate=0.0).cuda().train()
else:
    from deeplearning_b200.classification.vision_transformer.vit_model import vit_base_patch16_224_in21k
    m = vit_base_patch16_224_in21k(num_classes=1000, has_logits=False).cuda().train()
if which in ("convnext_tiny", "swin_tiny"):
    tr = TrainStep(m, lr=5e-4, weight_decay=5e-2, optimizer="adamw", clip_grad=5.0 if which == "swin_tiny" else None)
else:
    tr = TrainStep(m)
x = torch.randn(B, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss, _ = tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
