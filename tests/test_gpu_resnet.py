"""End-to-end parity of the B200 ResNet path against the CPU oracle (fp32) on the same weights and inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _models(seed=0, **kw):
    from deeplearning_b200.classification.resnet.models.networks import resnet50

    torch.manual_seed(seed)
    m = resnet50(**kw)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    return m.cuda(), state


def _autocast_yardstick(state, x, train):
    """Max-abs logit error of PyTorch's own bf16 autocast on the same weights/input (what bf16 storage costs the reference)."""
    import torchvision

    ref = torchvision.models.resnet50().cuda()
    ref.load_state_dict(state)
    ref.train(train)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        full = ref(x.cuda()).float()
        ref.load_state_dict(state)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            half = ref(x.cuda()).float()
    return float((half - full).abs().max())


def test_resnet50_eval_logits_parity():
    """Eval-mode logits vs the fp32 oracle, after calibrating the running statistics on one batch (at the raw init
    running_var = 1 makes eval-mode activations explode to |logit| ~ 100, where 1e-2 absolute is below bf16 resolution)."""
    from oracle.resnet import resnet_forward

    m, state = _models()
    xc = torch.randn(32, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        resnet_forward(state, xc, train=True, momentum=1.0)  # running stats := batch stats
    m.load_state_dict(state)
    m.eval()
    x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = resnet_forward(state, x, train=False)
        got = m(x.cuda()).float().cpu()
    err = float((got - ref).abs().max())
    yard = _autocast_yardstick(state, x, False)
    print(f"eval logits max-abs err {err:.4g} (|ref| max {float(ref.abs().max()):.3g}); torch bf16 autocast on the same input: {yard:.4g}")
    assert err <= max(1e-2, 1.5 * yard), (err, yard)  # north_star: 1e-2 for bf16, or no worse than the reference's own bf16


def test_resnet50_train_step_parity():
    from oracle.resnet import train_step_grads

    m, state = _models()
    m.train()
    B = 64
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    labels = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(2))
    state_before = {k: v.clone() for k, v in state.items()}
    ref_logits, ref_loss, ref_grads = train_step_grads(state, x, labels)
    out = m(x.cuda())
    loss = F.cross_entropy(out, labels.cuda())
    loss.backward()
    err = float((out.detach().float().cpu() - ref_logits).abs().max())
    print(f"train logits max-abs err {err:.4g} (|ref| max {float(ref_logits.abs().max()):.3g}); loss {float(loss):.5f} vs {float(ref_loss):.5f}")
    yard = _autocast_yardstick(state_before, x, True)
    print(f"torch bf16 autocast train-mode logits error on the same input: {yard:.4g}")
    assert err <= max(1e-2, 1.5 * yard), (err, yard)
    assert abs(float(loss) - float(ref_loss)) <= 1e-2
    worst = 0.0
    for name, p in m.named_parameters():
        g, r = p.grad.float().cpu(), ref_grads[name]
        rel = float((g - r).norm() / (r.norm() + 1e-12))
        worst = max(worst, rel)
        assert rel < 0.15, f"{name}: grad rel-L2 error {rel:.3g}"
    print(f"worst grad rel-L2 error {worst:.3g}")
    sd = m.state_dict()
    for k in state:
        if "running_" in k:
            assert torch.allclose(sd[k].cpu(), state[k], rtol=2e-2, atol=2e-3), k
        if "num_batches" in k:
            assert int(sd[k]) == int(state[k])


def test_resnet50_head_surgery_and_small_classes():
    """model.fc = nn.Linear(2048, 5) as the reference fine-tune script does (classification/resnet/train.py:79-80)."""
    m, _ = _models()
    m.fc = torch.nn.Linear(2048, 5).cuda()
    m.train()
    x = torch.randn(4, 3, 64, 64, device="cuda")
    out = m(x)
    assert out.shape == (4, 5)
    out.sum().backward()
    assert m.fc.weight.grad.shape == (5, 2048) and torch.isfinite(m.fc.weight.grad).all()
    assert m.conv1.weight.grad.shape == (64, 3, 7, 7)


def test_cpu_tensor_raises():
    m, _ = _models()
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 3, 32, 32))
