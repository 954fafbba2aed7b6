"""Oracle: the reference ViT forward restated functionally in fp32 PyTorch
(classification/vision_transformer/vit_model.py): PatchEmbed conv -> flatten -> transpose (:59-68); cls concat + pos add
(:244-250); Block = x + attn(LN(x)); x + mlp(LN(x)) (:158-161) with Attention (:88-111: qkv Linear, (q@k^T)*scale, softmax,
@v, proj) and Mlp (:127-133: fc1, exact-erf GELU, fc2); final LN eps 1e-6 (:194,252), cls row, optional pre_logits
Linear+Tanh (:218-221), head (:268).  Dropouts are p=0 in the BASELINE configuration; stochastic depth (``drop_path``
:12-40, applied to both branches of a Block :159-160) takes its per-sample masks from ``drop``: a list with two entries per
block in forward order (attention branch, MLP branch), each None or ``(random_tensor [B] of 0/1, keep_prob)``, applied as
the reference does: ``x.div(keep_prob) * random_tensor``."""
import torch
import torch.nn.functional as F


def _drop_path(y, entry):
    if entry is None:
        return y
    r, keep = entry
    return y.div(keep) * r.to(y.dtype).view((-1,) + (1,) * (y.dim() - 1))


def vit_forward(s, x, num_heads=12, patch=16, eps=1e-6, train=False, drop=None):
    B = x.shape[0]
    drop = list(drop) if (drop is not None and train) else None
    h = F.conv2d(x, s["patch_embed.proj.weight"], s["patch_embed.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    h = torch.cat([s["cls_token"].expand(B, -1, -1), h], 1) + s["pos_embed"]
    D = h.shape[-1]
    hd = D // num_heads
    i = 0
    while f"blocks.{i}.norm1.weight" in s:
        p = f"blocks.{i}."
        y = F.layer_norm(h, (D,), s[p + "norm1.weight"], s[p + "norm1.bias"], eps)
        qkv = F.linear(y, s[p + "attn.qkv.weight"], s.get(p + "attn.qkv.bias"))
        T = qkv.shape[1]
        q, k, v = qkv.reshape(B, T, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
        att = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1)
        y = (att @ v).transpose(1, 2).reshape(B, T, D)
        y = F.linear(y, s[p + "attn.proj.weight"], s[p + "attn.proj.bias"])
        h = h + (_drop_path(y, drop.pop(0)) if drop is not None else y)
        y = F.layer_norm(h, (D,), s[p + "norm2.weight"], s[p + "norm2.bias"], eps)
        y = F.linear(F.gelu(F.linear(y, s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"])), s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"])
        h = h + (_drop_path(y, drop.pop(0)) if drop is not None else y)
        i += 1
    h = F.layer_norm(h, (D,), s["norm.weight"], s["norm.bias"], eps)[:, 0]
    if "pre_logits.fc.weight" in s:
        h = torch.tanh(F.linear(h, s["pre_logits.fc.weight"], s["pre_logits.fc.bias"]))
    return F.linear(h, s["head.weight"], s["head.bias"])


def train_step_grads(state, x, labels, **kw):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state.items() if v.is_floating_point()}
    logits = vit_forward(params, x, train=True, **kw)
    loss = F.cross_entropy(logits, labels)
    grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    return logits.detach(), loss.detach(), {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in zip(params, grads)}
