"""ROUND-2 PROTOTYPE check: two-CTAs-per-SM attention backward (attention_bwd_v2.cuh) against the product kernel.

Build the wrapper first (repo root):
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC \
       -o tools/experiments/libattn_v2.so tools/experiments/attn_bwd_v2_lib.cu -lcuda
then on the GPU box (wrap in `timeout`: an untested pipeline can hang):
  timeout 120 python tools/experiments/attn_bwd_v2_check.py
The product kernel is itself checked against torch autograd by tests/test_gpu_transformer.py, so agreement here (bf16
rounding differences only: the partial dQ of the first key block is rounded to bf16 before the second is added) validates v2.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from deeplearning_b200 import _lib, ops

HERE = os.path.dirname(os.path.abspath(__file__))
v2 = ctypes.CDLL(os.path.join(HERE, "libattn_v2.so"))
P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
v2.b200x_attention_bwd_v2.argtypes = [P, P, P, P, P, I, I, I, F, P]
v2.b200x_attention_bwd_v2.restype = I


def run(B, T, H, iters=0):
    torch.manual_seed(B * 1000 + T)
    qkv = (torch.randn(B, T, 3 * H * 64, device="cuda") * 0.5).to(torch.bfloat16)
    dout = torch.randn(B, T, H * 64, device="cuda").to(torch.bfloat16)
    scale = 0.125
    out, lse = ops.attention_fwd(qkv, H, scale)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    ref = torch.empty_like(qkv)
    delta = torch.empty(B, H, T, dtype=torch.float32, device="cuda")
    rc = lib.b200_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                ref.data_ptr(), B, T, H, scale, st)
    assert rc == 0
    got = torch.full_like(qkv, float("nan"))
    rc = v2.b200x_attention_bwd_v2(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(), got.data_ptr(), B, T, H,
                                   scale, st)
    assert rc == 0, rc
    torch.cuda.synchronize()
    HD = H * 64
    for name, sl in (("dQ", slice(0, HD)), ("dK", slice(HD, 2 * HD)), ("dV", slice(2 * HD, 3 * HD))):
        a, r = got[..., sl].float(), ref[..., sl].float()
        err = (a - r).abs().max().item()
        rel = ((a - r).norm() / r.norm()).item()
        print(f"B={B} T={T} H={H} {name}: max abs err {err:.4g} (ref max {r.abs().max().item():.4g}), rel L2 {rel:.3g}")
        assert torch.isfinite(a).all() and rel < 1e-2, name
    if iters:
        def t(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        t_ref = t(lambda: lib.b200_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                                 ref.data_ptr(), B, T, H, scale, st))
        t_v2 = t(lambda: v2.b200x_attention_bwd_v2(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                                   got.data_ptr(), B, T, H, scale, st))
        print(f"B={B} T={T} H={H}: product {t_ref:.1f} us (incl. the delta kernel), v2 {t_v2:.1f} us")


if __name__ == "__main__":
    run(2, 197, 3)
    run(3, 64, 2)
    run(2, 128, 1)
    run(2, 256, 2)
    run(1, 130, 4)
    run(256, 197, 12, iters=10)   # ViT-B/16 bs 256
