"""Eager ResNet-50 step under ops.Profiler: per-op totals and the longest individual spans (debugging the bench roofline).
python tools/span_dump.py [run_ahead_ms]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deeplearning_b200 import ops
from deeplearning_b200.engine.trainer import TrainStep
from deeplearning_b200.classification.resnet.models.networks import resnet50

ahead = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
torch.manual_seed(0)
m = resnet50().cuda().train()
tr = TrainStep(m)
x = torch.randn(256, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (256,), device="cuda")
for _ in range(3):
    tr.step_eager(x, y)
torch.cuda.synchronize()
for ra in (0.0, ahead):
    with ops.Profiler(run_ahead_ms=ra) as prof:
        tr.step_eager(x, y)
    torch.cuda.synchronize()
    spans = [(e0.elapsed_time(e1), name, i) for i, (name, fl, nb, e0, e1) in enumerate(prof.records)]
    tot = {}
    for ms, name, _ in spans:
        tot[name] = tot.get(name, 0.0) + ms
    print(f"run_ahead={ra}: total {sum(tot.values()):.2f} ms", {k: round(v, 2) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
    print("  longest:", [(round(ms, 3), n, i) for ms, n, i in sorted(spans, reverse=True)[:8]])
