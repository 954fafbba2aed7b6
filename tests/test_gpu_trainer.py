"""TrainStep: fused optimizer kernels vs torch.optim on the same gradients, and CUDA-graph replay vs eager stepping."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _small_resnet(seed=0):
    from deeplearning_b200.classification.resnet.models.networks import Bottleneck, ResNet

    torch.manual_seed(seed)
    return ResNet(Bottleneck, [1, 1, 1, 1], num_classes=16).cuda().train()


@pytest.mark.parametrize("opt", ["sgd", "adamw"])
def test_fused_optimizer_matches_torch(opt):
    from deeplearning_b200.engine.trainer import TrainStep, no_decay_rule

    m = _small_resnet()
    ref = copy.deepcopy(m)
    tr = TrainStep(m, lr=0.05, momentum=0.9, weight_decay=5e-2, optimizer=opt)
    named = dict(ref.named_parameters())
    if opt == "sgd":
        ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9, weight_decay=5e-2)
    else:
        decay = [p for n, p in named.items() if not no_decay_rule(n, p)]
        nodecay = [p for n, p in named.items() if no_decay_rule(n, p)]
        ropt = torch.optim.AdamW([{"params": decay, "weight_decay": 5e-2}, {"params": nodecay, "weight_decay": 0.0}], lr=0.05)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    for step in range(3):
        before = [p.detach().clone() for p in m.parameters()]
        tr.step_eager(x, y)
        # replay the same gradients through torch's optimizer on the reference copy
        for (n, p), b in zip(m.named_parameters(), before):
            named[n].data.copy_(b)
            named[n].grad = p.grad.detach().clone()
        ropt.step()
        for (n, p) in m.named_parameters():
            assert torch.allclose(p.detach(), named[n].detach(), rtol=2e-4, atol=2e-6), (opt, step, n)


@pytest.mark.parametrize("opt,max_norm", [("adamw", 0.05), ("adamw", 1e6), ("sgd", 0.05)])
def test_grad_clipping_matches_clip_grad_norm(opt, max_norm):
    """TrainStep(clip_grad=...) == torch.nn.utils.clip_grad_norm_ + torch optimizer (the Swin recipe,
    classification/swin_transformer/utils/torch_utils.py:303-317): clipped (small max_norm) and not clipped (huge)."""
    from deeplearning_b200.engine.trainer import TrainStep, no_decay_rule

    m = _small_resnet(3)
    ref = copy.deepcopy(m)
    tr = TrainStep(m, lr=0.02, momentum=0.9, weight_decay=5e-2, optimizer=opt, clip_grad=max_norm)
    named = dict(ref.named_parameters())
    if opt == "sgd":
        ropt = torch.optim.SGD(ref.parameters(), lr=0.02, momentum=0.9, weight_decay=5e-2)
    else:
        decay = [p for n, p in named.items() if not no_decay_rule(n, p)]
        nodecay = [p for n, p in named.items() if no_decay_rule(n, p)]
        ropt = torch.optim.AdamW([{"params": decay, "weight_decay": 5e-2}, {"params": nodecay, "weight_decay": 0.0}], lr=0.02)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    for step in range(2):
        before = [p.detach().clone() for p in m.parameters()]
        tr.step_eager(x, y)
        for (n, p), b in zip(m.named_parameters(), before):
            named[n].data.copy_(b)
            named[n].grad = p.grad.detach().clone()
        total = torch.nn.utils.clip_grad_norm_(list(named.values()), max_norm)
        assert abs(float(tr.grad_norm) - float(total)) <= 1e-4 * float(total)
        if max_norm < 1:
            assert float(total) > max_norm   # the clipped branch is really exercised
        ropt.step()
        for (n, p) in m.named_parameters():
            assert torch.allclose(p.detach(), named[n].detach(), rtol=3e-4, atol=3e-6), (opt, step, n)


def test_graph_replay_equals_eager():
    from deeplearning_b200.engine.trainer import TrainStep

    a, b = _small_resnet(1), _small_resnet(1)
    ta, tb = TrainStep(a, lr=0.02), TrainStep(b, lr=0.02)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    tb.capture(x, y)           # capture's warm-up steps are rolled back: both models are still at the same point
    for _ in range(3):
        la, _ = ta.step_eager(x, y)
        lb, _ = tb.step(x, y)
    assert abs(float(la) - float(lb)) < 1e-3
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, rtol=1e-3, atol=1e-5)
    for ba, bb in zip(a.buffers(), b.buffers()):
        assert torch.allclose(ba.float(), bb.float(), rtol=1e-3, atol=1e-5)


def test_capture_is_side_effect_free():
    """capture() warms up with real steps; parameters, momentum, BN running statistics and the step count must come back."""
    from deeplearning_b200.engine.trainer import TrainStep

    for opt in ("sgd", "adamw"):
        m = _small_resnet(2)
        tr = TrainStep(m, lr=0.02, optimizer=opt)
        x = torch.randn(8, 3, 64, 64, device="cuda")
        y = torch.randint(0, 16, (8,), device="cuda")
        tr.step_eager(x, y)     # non-trivial optimizer state
        before_p = [p.detach().clone() for p in m.parameters()]
        before_b = [b.detach().clone() for b in m.buffers()]
        before_m = tr.arena.flat_m.clone()
        steps = tr.steps
        tr.capture(torch.randn_like(x), y)   # a dummy batch
        assert tr.steps == steps
        assert torch.equal(tr.arena.flat_m, before_m)
        for p, b in zip(m.parameters(), before_p):
            assert torch.equal(p.detach(), b)
        for p, b in zip(m.buffers(), before_b):
            assert torch.equal(p, b)


def test_forward_after_graph_replay_sees_updated_weights():
    """the captured graph repacks the bf16 operands at its START (pre-update values); a forward outside the graph
    after N replays must repack, i.e. equal the forward of an eagerly trained twin."""
    from deeplearning_b200.engine.trainer import TrainStep

    a, b = _small_resnet(5), _small_resnet(5)
    ta, tb = TrainStep(a, lr=0.05), TrainStep(b, lr=0.05)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    tb.capture(x, y)
    for _ in range(3):
        ta.step_eager(x, y)
        tb.step(x, y)
    a.eval(), b.eval()
    with torch.no_grad():
        oa, ob = a(x), b(x)
    assert torch.allclose(oa, ob, rtol=1e-3, atol=1e-3), float((oa - ob).abs().max())
    # and the stale-operand failure mode is detectable: one more replay changes the eval output
    b.train()
    tb.step(x, y)
    b.eval()
    with torch.no_grad():
        ob2 = b(x)
    assert float((ob2 - ob).abs().max()) > 0


@pytest.mark.parametrize("B,N", [(16, 10), (64, 1000), (7, 37)])
def test_soft_target_and_label_smoothing_cross_entropy(B, N):
    """The fused loss kernel with a target distribution == timm SoftTargetCrossEntropy (sum(-t log_softmax(x)).mean(),
    swin_transformer/main.py:111-113) and with smoothed hard labels == LabelSmoothingCrossEntropy (main.py:114-115, the same
    value as F.cross_entropy(label_smoothing=eps)); gradients against autograd."""
    from deeplearning_b200 import ops
    from deeplearning_b200.engine.mixup import mixup_target

    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N)
    x = torch.randn(B, N, device="cuda", generator=g) * 3
    y = torch.randint(0, N, (B,), device="cuda", generator=g)
    n_pad = (N + 7) // 8 * 8
    # soft targets (mixup of two smoothed one-hots)
    t = mixup_target(y, N, lam=0.3, smoothing=0.1)
    xr = x.clone().requires_grad_(True)
    ref = torch.sum(-t * torch.log_softmax(xr, dim=-1), dim=-1).mean()
    ref.backward()
    loss, d, _ = ops.softmax_xent(x, t, ld_d=n_pad)
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))
    assert torch.allclose(d[:, :N].float(), xr.grad, rtol=1e-2, atol=1e-4), float((d[:, :N].float() - xr.grad).abs().max())
    assert float(d[:, N:].float().abs().max() if n_pad > N else 0.0) == 0.0
    # label smoothing on hard labels, gradient scaled for 4 accumulation steps
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(xr, y, label_smoothing=0.1)
    (ref / 4).backward()
    loss, d, correct = ops.softmax_xent(x, y, ld_d=n_pad, label_smoothing=0.1, loss_scale=0.25)
    assert abs(float(loss) - float(ref)) < 1e-4 * max(1.0, abs(float(ref)))
    assert torch.allclose(d[:, :N].float(), xr.grad, rtol=1e-2, atol=1e-4)
    assert torch.equal(correct.bool(), x.argmax(1) == y)


def _small_vit(seed=0):
    from deeplearning_b200.classification.vision_transformer.vit_model import VisionTransformer

    torch.manual_seed(seed)
    return VisionTransformer(img_size=224, patch_size=16, embed_dim=768, depth=2, num_heads=12, num_classes=16).cuda().train()


@pytest.mark.parametrize("graph", [False, True])
def test_gradient_accumulation_equals_one_large_batch(graph):
    """accum_steps=2 over two half batches == one step on the whole batch (BatchNorm-free model, mean loss): the Swin loop's
    ``loss / ACCUMULATION_STEPS`` with an update every second micro-batch (swin_transformer/main.py:190-199)."""
    from deeplearning_b200.engine.trainer import TrainStep

    m1, m2 = _small_vit(5), _small_vit(5)
    t1 = TrainStep(m1, lr=0.05, momentum=0.9, weight_decay=1e-2, label_smoothing=0.1)
    t2 = TrainStep(m2, lr=0.05, momentum=0.9, weight_decay=1e-2, label_smoothing=0.1, accum_steps=2)
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(16, 3, 224, 224, device="cuda", generator=g)
    y = torch.randint(0, 16, (16,), device="cuda", generator=g)
    if graph:
        t2.capture(x[:8], y[:8])
    for _ in range(2):
        t1.step_eager(x, y)
        before = [p.detach().clone() for p in m2.parameters()]
        t2.step(x[:8], y[:8])
        assert all(torch.equal(a, p.detach()) for a, p in zip(before, m2.parameters())), "no update inside a group"
        t2.step(x[8:], y[8:])
    assert t1.steps == t2.steps == 2
    for (n, p), q in zip(m1.named_parameters(), m2.parameters()):
        assert torch.allclose(p.detach(), q.detach(), rtol=2e-2, atol=2e-4), (n, float((p - q).abs().max()))


@pytest.mark.parametrize("opt", ["sgd", "adamw"])
def test_optimizer_state_dict_round_trips_through_torch_optim(opt):
    """TrainStep.optimizer_state_dict() loads into the reference's torch optimizer (and back): after two fused steps a torch
    optimizer resumed from the exported state takes the same third step; a fresh TrainStep resumed from it does too."""
    from deeplearning_b200.engine.trainer import TrainStep, no_decay_rule

    m = _small_resnet(7)
    tr = TrainStep(m, lr=0.03, momentum=0.9, weight_decay=5e-2, optimizer=opt)
    x = torch.randn(8, 3, 64, 64, device="cuda")
    y = torch.randint(0, 16, (8,), device="cuda")
    for _ in range(2):
        tr.step_eager(x, y)
    sd = tr.optimizer_state_dict()
    ref = copy.deepcopy(m)
    named = dict(ref.named_parameters())
    if opt == "sgd":
        ropt = torch.optim.SGD(ref.parameters(), lr=0.03, momentum=0.9, weight_decay=5e-2)
    else:
        plist = list(ref.parameters())
        ropt = torch.optim.AdamW([{"params": [plist[i] for i in g["params"]], "weight_decay": g["weight_decay"]}
                                  for g in sd["param_groups"]], lr=0.03)
        # (torch numbers the parameters group by group: re-key the exported state accordingly)
        order = [i for g in sd["param_groups"] for i in g["params"]]
        sd_t = {"state": {k: sd["state"][i] for k, i in enumerate(order)},
                "param_groups": [dict(g, params=list(range(s0, s0 + len(g["params"]))))
                                 for g, s0 in zip(sd["param_groups"], [0, len(sd["param_groups"][0]["params"])])]}
    ropt.load_state_dict(sd if opt == "sgd" else {**ropt.state_dict(), "state": sd_t["state"]})
    resumed = copy.deepcopy(m)
    tr2 = TrainStep(resumed, lr=0.5, momentum=0.9, weight_decay=5e-2, optimizer=opt)
    tr2.load_optimizer_state_dict(sd)
    # third step: fused, torch (on the fused step's gradients), resumed-fused
    tr.step_eager(x, y)
    for (n, p) in m.named_parameters():
        named[n].grad = p.grad.detach().clone()
    ropt.step()
    tr2.step_eager(x, y)
    for (n, p), q in zip(m.named_parameters(), resumed.parameters()):
        assert torch.allclose(p.detach(), named[n].detach(), rtol=2e-4, atol=2e-6), (opt, "torch", n)
        # (AdamW: the resumed bias corrections 1 - beta^t come from a double-precision power, the running ones from t fp32
        #  multiplications - a few 1e-6 relative on the step)
        assert torch.allclose(p.detach(), q.detach(), rtol=1e-4, atol=1e-6), (opt, "resumed", n)
