"""Oracle: the reference ConvNeXt forward restated functionally in fp32 PyTorch (classification/convNext/models/networks.py):
stem conv4x4/4 -> LN(channels_first) (:127-128, LN impl :61-67); Block = x + gamma * pwconv2(GELU(pwconv1(LN(dwconv7x7(x)))))
(:92-105; stochastic depth ``drop_path`` :11-26 is identity in eval / at rate 0, otherwise the per-sample masks are passed in);
downsample LN -> conv2x2/2 (:133-134); mean over H,W -> LayerNorm -> head (:165,169).

``drop``: optional list with one entry per Block, in forward order: None (rate 0) or ``(random_tensor [B] of 0/1, keep_prob)``
- the reference's ``floor(keep_prob + rand)`` - applied exactly as the reference does: ``x.div(keep_prob) * random_tensor``."""
import torch
import torch.nn.functional as F


def _ln_cf(x, w, b, eps=1e-6):
    mean = x.mean(1, keepdim=True)
    var = (x - mean).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - mean) / torch.sqrt(var + eps)) + b[:, None, None]


def _drop_path(y, entry):
    if entry is None:
        return y
    r, keep = entry
    return y.div(keep) * r.to(y.dtype).view((-1,) + (1,) * (y.dim() - 1))


def convnext_forward(s, x, train=False, drop=None):
    drop = list(drop) if (drop is not None and train) else None
    x = _ln_cf(F.conv2d(x, s["downsample_layers.0.0.weight"], s["downsample_layers.0.0.bias"], stride=4),
               s["downsample_layers.0.1.weight"], s["downsample_layers.0.1.bias"])
    for i in range(4):
        if i > 0:
            p = f"downsample_layers.{i}."
            x = F.conv2d(_ln_cf(x, s[p + "0.weight"], s[p + "0.bias"]), s[p + "1.weight"], s[p + "1.bias"], stride=2)
        j = 0
        while f"stages.{i}.{j}.dwconv.weight" in s:
            p = f"stages.{i}.{j}."
            C = x.shape[1]
            y = F.conv2d(x, s[p + "dwconv.weight"], s[p + "dwconv.bias"], padding=3, groups=C).permute(0, 2, 3, 1)
            y = F.layer_norm(y, (C,), s[p + "norm.weight"], s[p + "norm.bias"], 1e-6)
            y = F.linear(F.gelu(F.linear(y, s[p + "pwconv1.weight"], s[p + "pwconv1.bias"])), s[p + "pwconv2.weight"], s[p + "pwconv2.bias"])
            if (p + "gamma") in s:
                y = s[p + "gamma"] * y
            y = y.permute(0, 3, 1, 2)
            if drop is not None:
                y = _drop_path(y, drop.pop(0))
            x = x + y
            j += 1
    x = F.layer_norm(x.mean([-2, -1]), (x.shape[1],), s["norm.weight"], s["norm.bias"], 1e-6)
    return F.linear(x, s["head.weight"], s["head.bias"])


def train_step_grads(state, x, labels, drop=None):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state.items() if v.is_floating_point()}
    logits = convnext_forward(params, x, train=True, drop=drop)
    loss = F.cross_entropy(logits, labels)
    grads = torch.autograd.grad(loss, list(params.values()))
    return logits.detach(), loss.detach(), dict(zip(params.keys(), grads))
