"""Batch-mode Mixup / CutMix on the device: the ``mixup_fn(samples, targets)`` step the Swin recipe runs between the loader
and the model (classification/swin_transformer/main.py:187-188; built by dataLoader/build.py:86-95 from ``timm.data.Mixup``).

timm is a third-party dependency that is absent from the reference tree (requirements: timm==0.4.12); this restates its
published batch-mode algorithm ("mode='batch'", the recipe's default):

    lam ~ Beta(alpha, alpha) drawn with numpy's global RNG (so a seeded numpy reproduces timm's draws), CutMix instead of
    Mixup with probability ``switch_prob`` when both are enabled, nothing at all with probability ``1 - prob``;
    Mixup:  x <- lam x + (1 - lam) x.flip(0)
    CutMix: a (H sqrt(1-lam)) x (W sqrt(1-lam)) box around a uniform centre is pasted from x.flip(0), lam corrected to the
            clipped box area;
    targets: lam * smooth_one_hot(y) + (1 - lam) * smooth_one_hot(y.flip(0)), smooth_one_hot = eps/N off, 1 - eps + eps/N on.

The mixed batch and the [B, num_classes] target distribution go straight into ``TrainStep.step`` (soft targets select the
SoftTargetCrossEntropy form of the fused loss kernel).  Data plumbing only - a handful of torch ops on the input batch.
"""
import numpy as np
import torch


def smooth_one_hot(labels, num_classes, smoothing=0.0):
    off = smoothing / num_classes
    on = 1.0 - smoothing + off
    y = torch.full((labels.shape[0], num_classes), off, dtype=torch.float32, device=labels.device)
    return y.scatter_(1, labels.view(-1, 1).long(), on)


def mixup_target(labels, num_classes, lam=1.0, smoothing=0.0):
    y1 = smooth_one_hot(labels, num_classes, smoothing)
    y2 = smooth_one_hot(labels.flip(0), num_classes, smoothing)
    return y1 * lam + y2 * (1.0 - lam)


def rand_bbox(img_hw, lam, rng=np.random):
    """CutMix box (yl, yh, xl, xh) for mixing ratio lam (box centre uniform over the image, clipped at the border)."""
    ratio = np.sqrt(1.0 - lam)
    img_h, img_w = img_hw
    cut_h, cut_w = int(img_h * ratio), int(img_w * ratio)
    cy = rng.randint(0, img_h)
    cx = rng.randint(0, img_w)
    yl = int(np.clip(cy - cut_h // 2, 0, img_h))
    yh = int(np.clip(cy + cut_h // 2, 0, img_h))
    xl = int(np.clip(cx - cut_w // 2, 0, img_w))
    xh = int(np.clip(cx + cut_w // 2, 0, img_w))
    return yl, yh, xl, xh


class Mixup:
    """Same constructor arguments and call convention as ``timm.data.Mixup`` in batch mode."""

    def __init__(self, mixup_alpha=1.0, cutmix_alpha=0.0, cutmix_minmax=None, prob=1.0, switch_prob=0.5, mode="batch",
                 correct_lam=True, label_smoothing=0.1, num_classes=1000):
        if mode != "batch" or cutmix_minmax is not None:
            raise NotImplementedError("only timm's batch mode without cutmix_minmax (the reference recipe) is mirrored")
        self.mixup_alpha, self.cutmix_alpha = mixup_alpha, cutmix_alpha
        self.mix_prob, self.switch_prob = prob, switch_prob
        self.correct_lam = correct_lam
        self.label_smoothing, self.num_classes = label_smoothing, num_classes
        self.mixup_enabled = True

    def _params_per_batch(self):
        lam, use_cutmix = 1.0, False
        if self.mixup_enabled and np.random.rand() < self.mix_prob:
            if self.mixup_alpha > 0.0 and self.cutmix_alpha > 0.0:
                use_cutmix = np.random.rand() < self.switch_prob
                lam_mix = (np.random.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix
                           else np.random.beta(self.mixup_alpha, self.mixup_alpha))
            elif self.mixup_alpha > 0.0:
                lam_mix = np.random.beta(self.mixup_alpha, self.mixup_alpha)
            elif self.cutmix_alpha > 0.0:
                use_cutmix = True
                lam_mix = np.random.beta(self.cutmix_alpha, self.cutmix_alpha)
            else:
                raise ValueError("one of mixup_alpha > 0, cutmix_alpha > 0 is required")
            lam = float(lam_mix)
        return lam, use_cutmix

    def __call__(self, x, target):
        if x.shape[0] % 2 != 0:
            raise ValueError("batch size should be even when using this")
        lam, use_cutmix = self._params_per_batch()
        if lam != 1.0:
            if use_cutmix:
                yl, yh, xl, xh = rand_bbox(x.shape[-2:], lam)
                if self.correct_lam:
                    lam = 1.0 - (yh - yl) * (xh - xl) / float(x.shape[-2] * x.shape[-1])
                x[:, :, yl:yh, xl:xh] = x.flip(0)[:, :, yl:yh, xl:xh]
            else:
                x_flipped = x.flip(0).mul_(1.0 - lam)
                x.mul_(lam).add_(x_flipped)
        return x, mixup_target(target, self.num_classes, lam, self.label_smoothing)
