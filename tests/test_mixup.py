"""engine/mixup.py on the CPU: the batch-mode Mixup / CutMix step of the Swin recipe (swin_transformer/main.py:187-188)."""
import numpy as np
import torch

from deeplearning_b200.engine.mixup import Mixup, mixup_target, rand_bbox, smooth_one_hot


def test_smooth_one_hot_and_mixup_target():
    y = torch.tensor([0, 3, 1, 2])
    t = smooth_one_hot(y, 4, 0.1)
    assert torch.allclose(t.sum(1), torch.ones(4))
    assert torch.allclose(t[1], torch.tensor([0.025, 0.025, 0.025, 0.925]))
    m = mixup_target(y, 4, lam=0.3, smoothing=0.0)
    # sample 0 pairs with the last sample (label 2): 0.3 on class 0, 0.7 on class 2
    assert torch.allclose(m[0], torch.tensor([0.3, 0.0, 0.7, 0.0]))
    assert torch.allclose(m.sum(1), torch.ones(4))


def test_mixup_batch_mode_is_lam_blend_with_the_flipped_batch():
    np.random.seed(0)
    fn = Mixup(mixup_alpha=0.8, cutmix_alpha=0.0, prob=1.0, label_smoothing=0.1, num_classes=10)
    np.random.seed(0)
    np.random.rand()                       # the "apply at all" draw
    lam = float(np.random.beta(0.8, 0.8))  # the draw the call below makes
    np.random.seed(0)
    x = torch.randn(4, 3, 8, 8)
    y = torch.tensor([1, 2, 3, 4])
    x0 = x.clone()
    xm, t = fn(x, y)
    assert torch.allclose(xm, lam * x0 + (1 - lam) * x0.flip(0), atol=1e-6)
    assert torch.allclose(t, mixup_target(y, 10, lam, 0.1))


def test_cutmix_pastes_a_box_and_corrects_lam():
    np.random.seed(3)
    fn = Mixup(mixup_alpha=0.0, cutmix_alpha=1.0, prob=1.0, label_smoothing=0.0, num_classes=5)
    x = torch.arange(2 * 1 * 16 * 16, dtype=torch.float32).view(2, 1, 16, 16)
    x0 = x.clone()
    xm, t = fn(x, torch.tensor([0, 1]))
    pasted = (xm != x0)[0, 0]
    area = int(pasted.sum())
    assert torch.equal(xm[0][:, pasted], x0[1][:, pasted])
    lam = 1.0 - area / 256.0
    assert torch.allclose(t[0], torch.tensor([lam, 1 - lam, 0, 0, 0]), atol=1e-6)
    yl, yh, xl, xh = rand_bbox((16, 16), 0.5, np.random.RandomState(0))
    assert 0 <= yl <= yh <= 16 and 0 <= xl <= xh <= 16


def test_prob_zero_leaves_the_batch_alone():
    fn = Mixup(mixup_alpha=0.8, cutmix_alpha=1.0, prob=0.0, label_smoothing=0.0, num_classes=3)
    x = torch.randn(2, 3, 4, 4)
    x0 = x.clone()
    xm, t = fn(x, torch.tensor([2, 0]))
    assert torch.equal(xm, x0) and torch.equal(t, smooth_one_hot(torch.tensor([2, 0]), 3))
