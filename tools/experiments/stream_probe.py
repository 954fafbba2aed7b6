"""Timing of the bottleneck-tail GEMMs (conv3 + BN + identity + ReLU; masked conv1 dgrad) on the streaming kernel vs the generic
implicit-GEMM kernel, ResNet-50 layer1 / layer2 shapes at bs 256.  B200_STREAM=0 in the environment selects the generic kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from deeplearning_b200 import ops

dev = torch.device("cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (B, H, K, N) in ((256, 56, 64, 256), (256, 28, 128, 512), (256, 14, 256, 1024)):
    g = torch.Generator(device=dev).manual_seed(0)
    y2 = torch.randn(B, H, H, K, device=dev, generator=g).to(torch.bfloat16)
    ident = torch.randn(B, H, H, N, device=dev, generator=g).to(torch.bfloat16)
    ymask = torch.randn(B, H, H, N, device=dev, generator=g).relu_().to(torch.bfloat16)
    w = torch.randn(N, K, 1, 1, device=dev, generator=g) * K ** -0.5
    wp = ops.pack_weight(w)
    w1 = torch.randn(K, N, 1, 1, device=dev, generator=g) * N ** -0.5
    wd = ops.pack_weight(w1, 1)
    co = ops.BnCoeffs(N, dev)
    co.scale.fill_(1.0), co.shift.fill_(0.1)
    dc = torch.randn(B, H, H, K, device=dev, generator=g).to(torch.bfloat16)
    t_f = timed(lambda: ops.conv1x1_bn_act(y2, wp, co, ident))
    t_b = timed(lambda: ops.conv1x1_dgrad_masked(dc, wd, residual=ident, mask_src=ymask))
    mb_f = (y2.numel() + 2 * ident.numel()) * 2 / 1e6
    mb_b = (dc.numel() + 3 * ident.numel()) * 2 / 1e6
    print(f"B={B} {H}x{H} K={K} N={N}: conv+bn+add+relu {t_f:7.1f} us ({mb_f / t_f * 1e-3 * 1e3:6.0f} GB/s)   "
          f"masked dgrad {t_b:7.1f} us ({mb_b / t_b * 1e-3 * 1e3:6.0f} GB/s)")
