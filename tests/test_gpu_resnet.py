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


def test_resnet50_eval_logits_parity():
    from oracle.resnet import resnet_forward

    m, state = _models()
    m.eval()
    x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = resnet_forward(state, x, train=False)
        got = m(x.cuda()).float().cpu()
    err = float((got - ref).abs().max())
    print(f"eval logits max-abs err {err:.4g} (|ref| max {float(ref.abs().max()):.3g})")
    assert err <= 1e-2, err  # north_star tolerance for bf16


def test_resnet50_train_step_parity():
    from oracle.resnet import train_step_grads

    m, state = _models()
    m.train()
    B = 64
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    labels = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(2))
    ref_logits, ref_loss, ref_grads = train_step_grads(state, x, labels)
    out = m(x.cuda())
    loss = F.cross_entropy(out, labels.cuda())
    loss.backward()
    err = float((out.detach().float().cpu() - ref_logits).abs().max())
    print(f"train logits max-abs err {err:.4g} (|ref| max {float(ref_logits.abs().max()):.3g}); loss {float(loss):.5f} vs {float(ref_loss):.5f}")
    assert err <= 5e-2
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
