"""BatchNorm folded through a 1x1 convolution (csrc/bn_algebra.cuh, the ResNet bottleneck tail conv3 -> bn3 -> +identity ->
ReLU of classification/resnet/models/networks.py:116-124): every piece against plain fp32 PyTorch on the same bf16 operands."""
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


# (pixels % 128 == 0 with K in {64, 128, 256}, N % 256 == 0 run on the streaming kernel conv1x1_stream.cuh, the rest on the generic one)
SHAPES = [(4, 14, 14, 64, 256), (3, 10, 6, 128, 512), (2, 7, 7, 256, 1024), (8, 28, 28, 64, 256), (8, 16, 16, 128, 512),
          (37, 32, 32, 64, 256), (19, 16, 24, 128, 512), (2, 8, 8, 64, 512), (8, 16, 16, 256, 1024), (5, 16, 24, 256, 512)]


@pytest.mark.parametrize("B,H,W,K,N", SHAPES)
def test_forward_stats_and_fused_conv_bn_add_relu(B, H, W, K, N):
    ops = _ops()
    y2 = _rand(B, H, W, K, seed=1).relu_() + 0.25        # post-ReLU-like input with non-zero channel means
    y2 = y2.to(torch.bfloat16)
    w = torch.randn(N, K, 1, 1, device="cuda") * K ** -0.5
    ident = _rand(B, H, W, N, seed=2)
    gamma = torch.rand(N, device="cuda") + 0.5
    beta = torch.randn(N, device="cuda") * 0.2
    rm, rv, nb = torch.zeros(N, device="cuda"), torch.ones(N, device="cuda"), torch.zeros((), dtype=torch.long, device="cuda")
    wp = ops.pack_weight(w)
    G, s = ops.gram_colsum(y2)
    rows = B * H * W
    co = ops.bn_gram_stats(G, s, wp, rows, gamma, beta, 1e-5, 0.1, rm, rv, nb)
    y = ops.conv1x1_bn_act(y2, wp, co, ident)
    # reference: fp32 conv of the same bf16 operands, train-mode batch norm, + identity, ReLU
    c = y2.float().view(rows, K) @ wp.float().t()
    bn = torch.nn.BatchNorm1d(N).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(gamma), bn.bias.copy_(beta)
    ref = F.relu(bn(c) + ident.float().view(rows, N))
    _close(co.mean, c.mean(0), 1e-4, 1e-4 * float(c.abs().max()), "batch mean")
    _close(1.0 / co.invstd ** 2, c.var(0, unbiased=False) + 1e-5, 2e-4, 1e-6, "batch variance")
    _close(rm, bn.running_mean, 1e-4, 1e-5, "running_mean")
    _close(rv, bn.running_var, 2e-4, 1e-5, "running_var")
    assert int(nb) == 1
    _close(y.view(rows, N), ref, 1e-2, 2e-2, "relu(bn(conv) + identity)")


@pytest.mark.parametrize("B,H,W,K,N", SHAPES)
def test_masked_dgrad_with_column_sums(B, H, W, K, N):
    """dz = (y > 0) * (dc @ W + residual) and its per-CTA column sums (K = narrow conv1 output, N = block width)."""
    ops = _ops()
    dc = _rand(B, H, W, K, seed=1)
    w1 = torch.randn(K, N, 1, 1, device="cuda") * N ** -0.5          # conv1: N -> K
    res = _rand(B, H, W, N, seed=2)
    y = _rand(B, H, W, N, seed=3).relu_()                              # block output (the mask source): about half zeros
    wd = ops.pack_weight(w1, 1)                                        # [N][K]
    dz, stats = ops.conv1x1_dgrad_masked(dc, wd, residual=res, mask_src=y)
    rows = B * H * W
    ref = (dc.float().view(rows, K) @ wd.float().t() + res.float().view(rows, N)) * (y.float().view(rows, N) > 0)
    _close(dz.view(rows, N), ref, 1e-2, 2e-2, "masked dgrad")
    assert torch.equal(dz.view(rows, N) == 0, ref == 0) or float(((dz.view(rows, N) == 0) != (ref == 0)).float().mean()) < 1e-3
    sums = stats[:, 0, :].sum(0)
    _close(sums, dz.float().view(rows, N).sum(0), 1e-3, 1e-3 * float(dz.float().abs().sum(0).max()), "column sums of dz")


@pytest.mark.parametrize("B,H,W,K,N", SHAPES)
def test_backward_algebra_matches_autograd(B, H, W, K, N):
    """Given dz: dgamma, dbeta, dW and dL/dy2 of y = bn(conv1x1(y2, W)) (train mode) without ever forming the conv output."""
    ops = _ops()
    rows = B * H * W
    y2 = (_rand(B, H, W, K, seed=1).relu_() + 0.1).to(torch.bfloat16)
    w = torch.randn(N, K, 1, 1, device="cuda") * K ** -0.5
    gamma = torch.rand(N, device="cuda") + 0.5
    beta = torch.randn(N, device="cuda") * 0.2
    wp = ops.pack_weight(w)
    G, s = ops.gram_colsum(y2)
    co = ops.bn_gram_stats(G, s, wp, rows, gamma, beta, 1e-5, 0.1, None, None, None)
    dz = _rand(B, H, W, N, seed=5) * (_rand(B, H, W, N, seed=6) > 0)   # a masked upstream gradient
    dz = dz.to(torch.bfloat16)
    _, partial = ops.relu_mask_sum(dz, torch.ones_like(dz))             # column sums of dz through the reduce pass
    D = ops.conv2d_wgrad(dz, y2, 1, 1)
    dgamma, dbeta, dW, wcat, wbias = ops.bn_conv1x1_bwd(partial, D, G, s, wp, w, rows, gamma, co)
    g2 = ops.gemm_dual(dz, y2, wcat, wbias)
    # autograd reference on the bf16 operands the kernels saw (weights: bf16 forward copy for c, fp32 master for the data grad
    # differ by rounding only; use the bf16 copy throughout so that the reference is exact for what is being tested)
    y2r = y2.float().view(rows, K).requires_grad_(True)
    wr = wp.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    c = y2r @ wr.t()
    mu, var = c.mean(0), c.var(0, unbiased=False)
    out = (c - mu) / torch.sqrt(var + 1e-5) * gr + br
    gy, gw, gg, gb = torch.autograd.grad(out, (y2r, wr, gr, br), dz.float().view(rows, N))
    sc = float(gg.abs().max())
    _close(dgamma / sc, gg / sc, 2e-3, 2e-3, "dgamma")
    sc = float(gb.abs().max())
    _close(dbeta / sc, gb / sc, 1e-3, 1e-3, "dbeta")
    sc = float(gw.abs().max())
    _close(dW.view(N, K) / sc, gw / sc, 5e-3, 5e-3, "dW")
    sc = float(gy.abs().max())
    _close(g2.view(rows, K) / sc, gy / sc, 2e-2, 2e-2, "dL/dy2")


def test_algebra_path_is_as_close_to_the_oracle_as_the_plain_bn_schedule(monkeypatch):
    """The folded path and the conv -> BN pass schedule are two implementations of the same math.  At random init a bf16
    ResNet's gradients are noisy (train-mode BN amplifies rounding; both schedules sit ~0.5 rel-L2 from the fp32 oracle on
    this net, and the reference's own bf16 autocast does too, see test_gpu_resnet.py), so the paths are compared through
    their distance to the fp32 oracle, parameter by parameter: same loss, same logits, and the algebra path never
    noticeably further from the oracle than the plain one.  Layers [2,2,1,1] exercise downsample blocks, plain blocks
    (masked conv1 dgrad handing dz to the previous block) and the K = 64 / 128 streaming kernels."""
    from deeplearning_b200.classification.resnet.models.networks import Bottleneck, ResNet
    from oracle.resnet import train_step_grads

    layers = [2, 2, 1, 1]
    x = torch.randn(32, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    y = torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(2))
    torch.manual_seed(0)
    state = {k: v.clone() for k, v in ResNet(Bottleneck, layers).state_dict().items()}
    ref_logits, ref_loss, ref_grads = train_step_grads({k: v.clone() for k, v in state.items()}, x, y)
    errs, losses, logit_err = {}, {}, {}
    for mode in ("1", "0"):
        monkeypatch.setenv("B200_RESNET_ALGEBRA", mode)
        m = ResNet(Bottleneck, layers)
        m.load_state_dict(state)
        m = m.cuda().train()
        out = m(x.cuda())
        loss = F.cross_entropy(out, y.cuda())
        loss.backward()
        losses[mode] = float(loss.detach())
        logit_err[mode] = float((out.detach().float().cpu() - ref_logits).abs().max())
        errs[mode] = {n: float((p.grad.float().cpu() - ref_grads[n]).norm() / (ref_grads[n].norm() + 1e-12)) for n, p in m.named_parameters()}
    assert abs(losses["1"] - float(ref_loss)) < 1e-2 and abs(losses["0"] - float(ref_loss)) < 1e-2, (losses, float(ref_loss))
    assert logit_err["1"] <= 1.25 * logit_err["0"] + 1e-2, logit_err
    bad = [(n, round(errs["1"][n], 3), round(errs["0"][n], 3)) for n in errs["1"] if errs["1"][n] > 1.3 * errs["0"][n] + 0.02]
    assert not bad, bad[:8]
    mean_a = sum(errs["1"].values()) / len(errs["1"])
    mean_b = sum(errs["0"].values()) / len(errs["0"])
    assert mean_a <= 1.05 * mean_b + 0.005, (mean_a, mean_b)
