"""GPU parity tests of the transformer-side kernels (LayerNorm, GEMM epilogues on strided / fp32 views, attention)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from deeplearning_b200 import ops

    return ops


def _rand(*shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


def _close(a, b, rtol, atol, what):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    bad = err > atol + rtol * b.abs()
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} bad, max abs err {float(err.max()):.4g} (ref max {float(b.abs().max()):.4g})"


@pytest.mark.parametrize("rows,C,dt", [(1000, 768, torch.float32), (513, 96, torch.bfloat16), (1001, 96, torch.float32), (777, 192, torch.float32), (64, 1536, torch.float32), (300, 384, torch.float32)])
def test_layernorm_fwd_bwd(rows, C, dt):
    ops = _ops()
    x = (_rand(rows, C, seed=1, dtype=torch.float32) * 2 + 0.5).to(dt)
    g = torch.rand(C, device="cuda") + 0.5
    b = torch.randn(C, device="cuda") * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
    xr = x.float().clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (C,), gr, br, 1e-6)
    _close(y, ref, 1e-2, 1e-2, "ln fwd")
    if C > 1024:
        return
    dy = _rand(rows, C, seed=2)
    add = _rand(rows, C, seed=3, dtype=torch.float32)
    gx, gg, gb = torch.autograd.grad(ref, (xr, gr, br), dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, mean, rstd, g, add=add, dx_dtype=torch.float32)
    _close(dx, gx + add, 1e-3, 2e-3, "ln dx")
    _close(dg, gg, 2e-3, 2e-3 * float(gg.abs().max()), "ln dgamma")
    _close(db, gb, 2e-3, 2e-3 * float(gb.abs().max()), "ln dbeta")
    dx16, _, _ = ops.layernorm_bwd(dy, x, mean, rstd, g, dx_dtype=torch.bfloat16)
    _close(dx16, gx, 1e-2, 1e-2, "ln dx bf16")


def test_gemm_views_fp32_residual_aux_and_gelu_grad():
    ops = _ops()
    M, K, N = 700, 256, 320
    a = _rand(M, K, seed=1)
    w = _rand(N, K, seed=2, scale=K ** -0.5)
    bias = torch.randn(N, device="cuda")
    res32 = _rand(M, N, seed=3, dtype=torch.float32)
    wp = ops.pack_weight(w.float())
    ref = a.float() @ w.float().t() + bias
    out, _ = ops.gemm(a, wp, bias=bias, residual=res32, out_f32=True)
    assert out.dtype == torch.float32
    _close(out, ref + res32, 1e-4, 2e-4, "fp32 out + fp32 residual")
    # GELU forward: the second output is GELU'(pre) (evaluated together with the value), which is all the backward needs
    post, dact = ops.gemm(a, wp, bias=bias, act=2, aux_out=True)
    prer = ref.clone().requires_grad_(True)
    (gelu_grad,) = torch.autograd.grad(F.gelu(prer).sum(), prer)
    _close(dact, gelu_grad, 1e-2, 1e-2, "aux GELU'(pre)")
    _close(post, F.gelu(ref), 1e-2, 1e-2, "gelu")
    post_only, _ = ops.gemm(a, wp, bias=bias, act=2)
    assert torch.equal(post_only, post), "GELU value must not depend on whether the derivative is saved"
    # backward through GELU fused into the dgrad GEMM of the following layer (multiplies by the saved derivative)
    dy = _rand(M, 96, seed=4)
    w2 = _rand(96, N, seed=5, scale=0.1)
    w2d = ops.pack_weight(w2.float(), mode=1)  # [N][96]
    d_pre, _ = ops.gemm(dy, w2d, act=3, aux_in=dact)
    gref = gelu_grad * (dy.float() @ w2.float())
    _close(d_pre, gref, 2e-2, 2e-2, "gelu grad epilogue")


def test_patch_embed_into_token_rows():
    """ViT PatchEmbed + pos_embed add written straight into rows 1.. of the fp32 [B,197,D] token tensor."""
    ops = _ops()
    B, D, ps = 3, 128, 16
    x = torch.randn(B, 3, 64, 64, device="cuda")
    P = (64 // ps) ** 2
    T = P + 1
    wconv = torch.randn(D, 3, ps, ps, device="cuda") * 0.05
    bias = torch.randn(D, device="cuda")
    pos = torch.randn(1, T, D, device="cuda")
    cls = torch.randn(1, 1, D, device="cuda")
    a = ops.patchify_nchw(x, ps)
    wp = ops.pack_weight(wconv.reshape(D, -1, 1, 1))
    tokens = torch.empty(B, T, D, dtype=torch.float32, device="cuda")
    ops.gemm(a, wp, bias=bias, out=tokens,
             a_view=((P, B, 1), (3 * ps * ps, P * 3 * ps * ps, 0)),
             out_view=((P, B, 1), (D, T * D, 0)), out_offset=D, residual=pos[0, 1:], residual_view=((P, B, 1), (D, 0, 0)))
    ops.cls_row_(tokens, cls.reshape(-1), pos.reshape(-1))
    ref = F.conv2d(x.to(torch.bfloat16).float(), wconv.to(torch.bfloat16).float(), bias, stride=ps).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.expand(B, -1, -1), ref], 1) + pos
    _close(tokens, ref, 1e-3, 5e-3, "patch embed tokens")


def _attn_ref(qkv, H, scale):
    B, T, _ = qkv.shape
    q, k, v = qkv.float().reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-2, -1)) * scale).softmax(-1)
    return (att @ v).transpose(1, 2).reshape(B, T, H * 64)


# (the last two: several (batch, head) items per SM - buffer reuse / barrier phases of the persistent forward kernel)
@pytest.mark.parametrize("B,T,H", [(2, 197, 12), (3, 49, 3), (1, 256, 2), (2, 130, 4), (5, 16, 1), (40, 197, 12), (64, 100, 6)])
def test_attention_fwd_bwd(B, T, H):
    ops = _ops()
    scale = 64 ** -0.5
    qkv = _rand(B, T, 3 * H * 64, seed=7)
    out, lse = ops.attention_fwd(qkv, H, scale)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, H, scale)
    _close(out, ref, 2e-2, 2e-2, "attention fwd")
    q, k = qr.detach().reshape(B, T, 3, H, 64)[:, :, 0], qr.detach().reshape(B, T, 3, H, 64)[:, :, 1]
    s = torch.einsum("bthd,bshd->bhts", q, k) * scale
    _close(lse, torch.logsumexp(s, -1), 1e-3, 1e-3, "lse")
    dout = _rand(B, T, H * 64, seed=8)
    (gref,) = torch.autograd.grad(ref, qr, dout.float())
    dqkv = ops.attention_bwd(qkv, out, dout, lse, H, scale)
    sc = float(gref.abs().max())
    _close(dqkv / sc, gref / sc, 2e-2, 2e-2, "attention bwd")
