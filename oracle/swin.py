"""Oracle: the reference Swin Transformer forward restated functionally in fp32 PyTorch
(classification/swin_transformer/models/swin_transformer.py): PatchEmbed conv4x4/4 -> flatten -> LN (:460-468);
SwinTransformerBlock (:241-287): LN -> roll(-s) -> window_partition (:38-50) -> WindowAttention (:118-149: qkv, q*scale,
q k^T + relative_position_bias_table[relative_position_index] + attn_mask (0/-100), softmax, @v, proj) -> window_reverse
(:53-67) -> roll(+s) -> residual; LN -> Mlp (fc1, exact GELU, fc2) -> residual; PatchMerging (:324-345): 2x2 gather-concat
[x0,x1,x2,x3] -> LN(4C) -> Linear(4C,2C, no bias); head: LN -> mean over tokens -> Linear (:590-597).
Stochastic depth (timm 0.4.12 ``DropPath`` = the rand/floor ``drop_path`` function, applied to both branches :282,285) takes
its per-sample masks from ``drop``: two entries per block in forward order, each None or ``(random_tensor [B], keep_prob)``."""
import torch
import torch.nn.functional as F


def _partition(x, ws):
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _reverse(w, ws, H, W):
    B = int(w.shape[0] / (H * W / ws / ws))
    return w.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def _drop_path(y, entry):
    if entry is None:
        return y
    r, keep = entry
    return y.div(keep) * r.to(y.dtype).view((-1,) + (1,) * (y.dim() - 1))


def swin_forward(s, x, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), patch=4, eps=1e-5, train=False, drop=None):
    B = x.shape[0]
    drop = list(drop) if (drop is not None and train) else None
    h = F.conv2d(x, s["patch_embed.proj.weight"], s["patch_embed.proj.bias"], stride=patch)
    H, W = h.shape[2], h.shape[3]
    h = h.flatten(2).transpose(1, 2)
    if "patch_embed.norm.weight" in s:
        h = F.layer_norm(h, (h.shape[-1],), s["patch_embed.norm.weight"], s["patch_embed.norm.bias"], eps)
    if "absolute_pos_embed" in s:
        h = h + s["absolute_pos_embed"]
    for li, depth in enumerate(depths):
        C = h.shape[-1]
        nH = num_heads[li]
        for bi in range(depth):
            p = f"layers.{li}.blocks.{bi}."
            ws = min(7, H, W)
            mask = s.get(p + "attn_mask")
            shift = 0 if mask is None else ws // 2
            y = F.layer_norm(h, (C,), s[p + "norm1.weight"], s[p + "norm1.bias"], eps).view(B, H, W, C)
            if shift > 0:
                y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
            yw = _partition(y, ws).view(-1, ws * ws, C)
            N = ws * ws
            qkv = F.linear(yw, s[p + "attn.qkv.weight"], s.get(p + "attn.qkv.bias")).reshape(-1, N, 3, nH, C // nH).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0] * (C // nH) ** -0.5, qkv[1], qkv[2]
            attn = q @ k.transpose(-2, -1)
            bias = s[p + "attn.relative_position_bias_table"][s[p + "attn.relative_position_index"].view(-1)].view(N, N, -1)
            attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
            if mask is not None:
                nW = mask.shape[0]
                attn = (attn.view(-1, nW, nH, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nH, N, N)
            attn = attn.softmax(-1)
            yw = F.linear((attn @ v).transpose(1, 2).reshape(-1, N, C), s[p + "attn.proj.weight"], s[p + "attn.proj.bias"])
            y = _reverse(yw.view(-1, ws, ws, C), ws, H, W)
            if shift > 0:
                y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
            y = y.view(B, H * W, C)
            h = h + (_drop_path(y, drop.pop(0)) if drop is not None else y)
            y = F.layer_norm(h, (C,), s[p + "norm2.weight"], s[p + "norm2.bias"], eps)
            y = F.linear(F.gelu(F.linear(y, s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"])), s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"])
            h = h + (_drop_path(y, drop.pop(0)) if drop is not None else y)
        p = f"layers.{li}.downsample."
        if (p + "reduction.weight") in s:
            y = h.view(B, H, W, C)
            y = torch.cat([y[:, 0::2, 0::2], y[:, 1::2, 0::2], y[:, 0::2, 1::2], y[:, 1::2, 1::2]], -1).view(B, -1, 4 * C)
            y = F.layer_norm(y, (4 * C,), s[p + "norm.weight"], s[p + "norm.bias"], eps)
            h = F.linear(y, s[p + "reduction.weight"])
            H, W = H // 2, W // 2
    h = F.layer_norm(h, (h.shape[-1],), s["norm.weight"], s["norm.bias"], eps)
    return F.linear(h.mean(1), s["head.weight"], s["head.bias"])


def train_step_grads(state, x, labels, **kw):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state.items()
              if v.is_floating_point() and "attn_mask" not in k}
    work = dict(state)
    work.update(params)
    logits = swin_forward(work, x, train=True, **kw)
    loss = F.cross_entropy(logits, labels)
    grads = torch.autograd.grad(loss, list(params.values()))
    return logits.detach(), loss.detach(), dict(zip(params.keys(), grads))
