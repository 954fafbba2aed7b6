"""GPU parity tests of the Swin kernels: window-process permutations (the reference's unit_test.py checks, bit-exact),
shifted-window attention forward/backward with relative-position bias and shift mask, patch-merging LayerNorm."""
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


def _window_partition(x, ws):
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _window_reverse(w, ws, H, W):
    B = int(w.shape[0] / (H * W / ws / ws))
    return w.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C,shift,ws", [(192, 56, 56, 96, 2, 7), (3, 14, 28, 32, 3, 7), (2, 8, 8, 8, 0, 4)])
def test_window_process_bit_exact(dtype, B, H, W, C, shift, ws):
    """Same construction as kernels/window_process/unit_test.py (B=192,H=W=56,C=96,shift=2,ws=7): torch.equal, all dtypes."""
    ops = _ops()
    x = torch.randn(B, H, W, C, device="cuda").to(dtype)
    ref = _window_partition(torch.roll(x, shifts=(-shift, -shift), dims=(1, 2)), ws)
    got = ops.window_partition(x, -shift, ws)
    assert torch.equal(got, ref)
    back_ref = torch.roll(_window_reverse(ref, ws, H, W), shifts=(shift, shift), dims=(1, 2))
    back = ops.window_merge(got, B, H, W, shift, ws)
    assert torch.equal(back, back_ref) and torch.equal(back, x)
    # the backward of each op is the other op with the negated shift
    g = torch.randn_like(ref)
    assert torch.equal(ops.window_merge(g, B, H, W, shift, ws), torch.roll(_window_reverse(g, ws, H, W), (shift, shift), (1, 2)))


def _swin_attn_ref(qkv, nH, table, index, mask, shift, scale):
    """fp32 restatement of roll -> window_partition -> WindowAttention (bias, mask) -> window_reverse -> roll."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    x = qkv.float()
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    xw = _window_partition(x, 7).view(-1, 49, 3, nH, 32).permute(2, 0, 3, 1, 4)
    q, k, v = xw[0] * scale, xw[1], xw[2]
    attn = q @ k.transpose(-2, -1)
    bias = table[index.view(-1)].view(49, 49, nH).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(-1, nW, nH, 49, 49) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, nH, 49, 49)
    attn = attn.softmax(-1)
    out = (attn @ v).transpose(1, 2).reshape(-1, 7, 7, C)
    out = _window_reverse(out, 7, H, W)
    if shift > 0:
        out = torch.roll(out, shifts=(shift, shift), dims=(1, 2))
    return out


def _rel_index():
    coords = torch.stack(torch.meshgrid([torch.arange(7), torch.arange(7)], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += 6
    rel[:, :, 1] += 6
    rel[:, :, 0] *= 13
    return rel.sum(-1)


def _shift_mask(H, W, shift):
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for h in (slice(0, -7), slice(-7, -shift), slice(-shift, None)):
        for w in (slice(0, -7), slice(-7, -shift), slice(-shift, None)):
            img[:, h, w, :] = cnt
            cnt += 1
    mw = _window_partition(img, 7).view(-1, 49)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


@pytest.mark.parametrize("B,H,W,nH,shift", [(2, 56, 56, 3, 0), (2, 56, 56, 3, 3), (3, 14, 14, 12, 3), (4, 7, 7, 24, 0), (1, 28, 14, 6, 3), (3, 7, 7, 2, 0), (1, 7, 7, 1, 0), (5, 14, 7, 3, 3)])
def test_window_attention_fwd_bwd(B, H, W, nH, shift):
    ops = _ops()
    C = nH * 32
    scale = 32 ** -0.5
    qkv = _rand(B, H, W, 3 * C, seed=1)
    table = (torch.randn(169, nH, device="cuda") * 0.5)
    index = _rel_index().cuda()
    mask = _shift_mask(H, W, shift).cuda() if shift > 0 else None
    bias = ops.window_bias_gather(table, index, nH, mask)   # [nH, nW or 1, 49 (i), 64 (j)]
    dense = table[index.view(-1)].view(49, 49, nH).permute(2, 0, 1)           # [nH, i, j]
    want = dense[:, None] + (mask[None] if mask is not None else 0)          # [nH, nW or 1, i, j]
    _close(bias[..., :49], want.expand_as(bias[..., :49]) * 1.4426950408889634, 1e-6, 1e-5, "bias table (log2 units)")
    assert float(bias[..., 49:].abs().max()) == 0.0
    out, lse = ops.window_attention_fwd(qkv, nH, bias, shift, scale)
    qr = qkv.float().requires_grad_(True)
    tr = table.clone().requires_grad_(True)
    ref = _swin_attn_ref(qr, nH, tr, index, mask, shift, scale)
    _close(out, ref, 2e-2, 2e-2, "window attention fwd")
    dout = _rand(B, H, W, C, seed=2)
    gq, gt = torch.autograd.grad(ref, (qr, tr), dout.float())
    dqkv, dbias = ops.window_attention_bwd(qkv, out, dout, bias, lse, nH, shift, scale)
    sc = float(gq.abs().max())
    _close(dqkv / sc, gq / sc, 2e-2, 2e-2, "window attention dqkv")
    dtable = ops.window_bias_scatter(dbias, index, torch.zeros_like(table))
    st = float(gt.abs().max())
    _close(dtable / st, gt / st, 2e-2, 2e-2, "relative position bias table grad")


@pytest.mark.parametrize("B,H,W,C", [(2, 56, 56, 96), (3, 14, 14, 384), (2, 28, 28, 192)])
def test_patch_merge_layernorm(B, H, W, C):
    ops = _ops()
    x = _rand(B, H, W, C, seed=1, dtype=torch.float32) * 2 + 0.3
    g = torch.rand(4 * C, device="cuda") + 0.5
    b = torch.randn(4 * C, device="cuda") * 0.1
    y, mean, rstd = ops.patch_merge_ln_fwd(x, g, b, 1e-5)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    cat = torch.cat([xr[:, 0::2, 0::2], xr[:, 1::2, 0::2], xr[:, 0::2, 1::2], xr[:, 1::2, 1::2]], -1).view(-1, 4 * C)
    ref = F.layer_norm(cat, (4 * C,), gr, br, 1e-5)
    _close(y, ref, 1e-2, 1e-2, "patch-merge LN fwd")
    dy = _rand(ref.shape[0], 4 * C, seed=2)
    gx, gg, gb = torch.autograd.grad(ref, (xr, gr, br), dy.float())
    dx, dg, db = ops.patch_merge_ln_bwd(dy, x, mean, rstd, g)
    _close(dx, gx, 1e-2, 1e-2, "patch-merge LN dx")
    _close(dg, gg, 2e-3, 2e-3 * float(gg.abs().max()), "dgamma")
    _close(db, gb, 2e-3, 2e-3 * float(gb.abs().max()), "dbeta")
