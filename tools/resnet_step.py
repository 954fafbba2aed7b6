"""Quick ResNet-50 step timing on the GPU box: python tools/resnet_step.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from deeplearning_b200.classification.resnet.models.networks import resnet50

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = resnet50().cuda().train()
opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=5e-5)
x = torch.randn(B, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")

def step():
    out = m(x)
    loss = F.cross_entropy(out, y)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
    return loss

for _ in range(3):
    l = step()
torch.cuda.synchronize()
print("loss", float(l), "mem GB", torch.cuda.max_memory_allocated() / 2**30)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
t0 = time.time()
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"B={B} {ms:.2f} ms/step  {B / ms * 1e3:.0f} img/s  (host wall {(time.time() - t0) / n * 1e3:.2f} ms/step)")
