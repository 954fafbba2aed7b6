"""Forward / backward schedule of the ViT family on the sm_100a kernels (one autograd.Function for the whole network).

Mirrors ``VisionTransformer.forward_features`` / ``Block.forward`` / ``Attention.forward`` / ``Mlp.forward`` of the
reference (classification/vision_transformer/vit_model.py:240-268, :158-161, :88-111, :127-133).

Data flow per block (residual stream ``h`` fp32 [B,T,D], everything feeding a tensor core bf16):
    LN1(h) -> qkv GEMM(+bias) -> tcgen05 attention -> proj GEMM(+bias, +h, fp32 out) = h2
    LN2(h2) -> fc1 GEMM(+bias, GELU; also writes GELU'(pre) for the backward) -> fc2 GEMM(+bias, +h2, fp32 out) = h3
Residual adds, biases, GELU, GELU' (backward) and the pos-embed add of the patch embedding all live in GEMM epilogues; the
attention scores never touch HBM.  The gradient of the residual stream is carried in bf16 and accumulated inside the
LayerNorm-backward kernel.
"""
import torch
import torch.nn as nn

from .. import ops
from . import droppath
from .packing import weight_cache
from .resnet import _Grads  # same gradient-sink protocol as the ResNet engine

BF16 = torch.bfloat16
F32 = torch.float32


def _linears(model):
    out = [model.head]
    for blk in model.blocks:
        out += [blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2]
    if model.has_logits:
        out.append(model.pre_logits.fc)   # Linear + Tanh on the class-token row (vit_model.py:218-221)
    return out


class _PackSpec:
    @staticmethod
    def key(model):
        return (len(model.blocks), id(model.head), model.head.out_features, bool(model.has_logits))

    def __call__(self, model):
        specs = []
        pe = model.patch_embed.proj.weight  # [D, C, p, p]: conv weight.view(D, -1) is already the GEMM operand order
        k0 = pe.numel() // pe.shape[0]
        specs.append((pe, 0, k0, pe.shape[0], (pe.shape[0], k0, 1)))
        for lin in _linears(model)[1:]:
            w = lin.weight
            specs.append((w, 0, w.shape[1], w.shape[0]))
            specs.append((w, 1, w.shape[0], w.shape[1]))
        head = model.head
        n_pad = (head.out_features + 7) // 8 * 8
        specs.append((head.weight, 0, head.in_features, n_pad))
        specs.append((head.weight, 1, n_pad, head.in_features))
        return specs


_pack_spec = _PackSpec()


def _check(model):
    if model.dist_token is not None:
        raise NotImplementedError("distilled ViT (dist_token) is not implemented on the B200 engine")
    if model.has_logits and not (isinstance(getattr(model.pre_logits, "fc", None), nn.Linear)
                                 and isinstance(getattr(model.pre_logits, "act", None), nn.Tanh)):
        raise NotImplementedError("pre_logits must be the reference's Sequential(fc=Linear, act=Tanh) (vit_model.py:218-221)")
    if not isinstance(model.head, nn.Linear):
        raise NotImplementedError("model.head must be an nn.Linear (num_classes > 0)")
    for m in model.modules():
        if isinstance(m, nn.Dropout) and m.p != 0 and model.training:
            raise NotImplementedError("dropout > 0 is not implemented on the B200 engine")
    for blk in model.blocks:
        if blk.attn.qkv.in_features // blk.attn.num_heads != 64:
            raise NotImplementedError("the attention kernel is built for head_dim 64")
        if not isinstance(blk.mlp.act, nn.GELU):
            raise NotImplementedError("Mlp activation must be nn.GELU (exact erf)")


def forward(model, x, train, want_tape):
    _check(model)
    if x.dtype == torch.uint8:      # GPU input pipeline: decoded uint8 NHWC batch -> ToTensor + Normalize on the device
        x = ops.normalize_u8_nhwc(x, *getattr(model, "input_norm", (ops.IMAGENET_MEAN, ops.IMAGENET_STD)))
    x = x.contiguous().float()
    B, Cin, Hh, Ww = x.shape
    pe = model.patch_embed
    ps = pe.patch_size[0]
    if (Hh, Ww) != tuple(pe.img_size):
        raise AssertionError(f"Input image size ({Hh}*{Ww}) doesn't match model ({pe.img_size[0]}*{pe.img_size[1]}).")
    D = model.embed_dim
    P = pe.num_patches
    T = P + 1
    pack = weight_cache.model_pack(model, _pack_spec)
    tape = {"blocks": [], "pack": pack} if want_tape else None
    # ---- patch embedding: patch matrix GEMM writing rows 1.. of the token tensor, + bias + pos_embed in the epilogue
    a = ops.patchify_nchw(x, ps)                       # bf16 [B, P, Cin*ps*ps]
    K0 = a.shape[-1]
    tokens = torch.empty(B, T, D, dtype=F32, device=x.device)
    pos = model.pos_embed.detach()
    ops.gemm(a, pack.get(pe.proj.weight, 0), bias=pe.proj.bias.detach() if pe.proj.bias is not None else None, out=tokens,
             a_view=((P, B, 1), (K0, P * K0, 0)), out_view=((P, B, 1), (D, T * D, 0)), out_offset=D,
             residual=pos.reshape(T, D)[1:], residual_view=((P, B, 1), (D, 0, 0)))
    ops.cls_row_(tokens, model.cls_token.detach().reshape(-1), pos.reshape(-1))
    h = tokens
    if want_tape:
        tape["patches"] = a
    for blk in model.blocks:
        att_m, mlp = blk.attn, blk.mlp
        H = att_m.num_heads
        y1, m1, r1 = ops.layernorm_fwd(h, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv, _ = ops.gemm(y1, pack.get(att_m.qkv.weight, 0), bias=att_m.qkv.bias)
        att, lse = ops.attention_fwd(qkv, H, float(att_m.scale))
        dp = droppath.drop_prob_of(blk, train)
        dp1 = droppath.sample_scale(dp, B, 3, x.device)     # x = x + drop_path(attn(norm1(x)))   (vit_model.py:159)
        h2, _ = ops.gemm(att, pack.get(att_m.proj.weight, 0), bias=att_m.proj.bias, residual=h, out_f32=True,
                         rowscale=None if dp1 is None else (dp1, T))
        y2, m2, r2 = ops.layernorm_fwd(h2, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        post, dact = ops.gemm(y2, pack.get(mlp.fc1.weight, 0), bias=mlp.fc1.bias, act=2, aux_out=want_tape)
        dp2 = droppath.sample_scale(dp, B, 3, x.device)     # x = x + drop_path(mlp(norm2(x)))    (vit_model.py:160)
        h3, _ = ops.gemm(post, pack.get(mlp.fc2.weight, 0), bias=mlp.fc2.bias, residual=h2, out_f32=True,
                         rowscale=None if dp2 is None else (dp2, T))
        if want_tape:
            tape["blocks"].append((blk, h, y1, m1, r1, qkv, att, lse, h2, y2, m2, r2, dact, post, dp1, dp2))
        h = h3
    # ---- head: final LayerNorm on the class-token rows only, then the classifier (fp32 logits)
    cls_rows = torch.empty(B, D, dtype=F32, device=x.device)
    ops.copy_rows(h, 0, T * D, cls_rows, 0, D, B, D)
    yc, mc, rc = ops.layernorm_fwd(cls_rows, model.norm.weight, model.norm.bias, model.norm.eps)
    head = model.head
    feat, t32 = yc, None          # classifier input (bf16 [B, R])
    if model.has_logits:          # pre_logits: tanh(fc(cls row))  (vit_model.py:218-221,254)
        fc = model.pre_logits.fc
        u, _ = ops.gemm(yc, pack.get(fc.weight, 0), bias=fc.bias, out_f32=True)
        t32, feat = ops.tanh_fwd(u)
    R = feat.shape[1]
    n_cls = head.out_features
    n_pad = (n_cls + 7) // 8 * 8
    bias = None
    if head.bias is not None:
        bias = head.bias.detach()
        if n_pad != n_cls:
            bias = torch.cat([bias, bias.new_zeros(n_pad - n_cls)])
    logits, _ = ops.conv2d_fwd(feat.view(B, 1, 1, R), pack.get(head.weight, 0), bias=bias, out_f32=True)
    logits = logits.view(B, n_pad)
    if want_tape:
        tape["head"] = (cls_rows, yc, mc, rc, n_cls, n_pad, (B, T, D, P), feat, t32)
    return (logits[:, :n_cls] if n_pad != n_cls else logits), tape


def _lin_grads(grads, lin, dy2d, x2d, dy_stats=None):
    """Weight / bias gradient of a Linear layer from dy [M, N] and its input x [M, K] (both bf16).
    dy_stats: epilogue column-sum partials of dy when the GEMM that produced dy already summed its columns."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    dst = grads.dest(lin.weight)
    gb = None
    if lin.bias is not None and dy_stats is None:
        # bias gradient = column sums of dy: summed inside the wgrad kernel from the dy tiles it already holds
        gb = grads.dest(lin.bias)
        if gb is None:
            gb = torch.empty(N, dtype=F32, device=dy2d.device)
    gw = ops.conv2d_wgrad(dy2d.view(M, 1, 1, N), x2d.view(M, 1, 1, K), out=dst.view(N, K, 1, 1) if dst is not None else None,
                          bias_out=gb)
    grads.put(lin.weight, gw)
    if lin.bias is not None:
        if dy_stats is not None:
            grads.put(lin.bias, ops.stats_colsum(dy_stats, out=grads.dest(lin.bias)))
        else:
            grads.put(lin.bias, gb)


def backward(model, tape, dlogits, sink=None):
    grads = _Grads(sink)
    pack = tape["pack"]
    cls_rows, yc, mc, rc, n_cls, n_pad, (B, T, D, P), feat, t32 = tape["head"]
    head = model.head
    R = feat.shape[1]
    if dlogits.dtype == BF16 and dlogits.shape[1] == n_pad and dlogits.is_contiguous():
        dl16 = dlogits
    else:
        dl = dlogits.contiguous().float()
        if n_pad != n_cls:
            dl = torch.cat([dl, dl.new_zeros(B, n_pad - n_cls)], 1).contiguous()
        dl16 = ops.cast_bf16(dl)
    dst = grads.dest(head.weight)
    if dst is not None and n_pad == n_cls:
        grads.put(head.weight, ops.conv2d_wgrad(dl16.view(B, 1, 1, n_pad), feat.view(B, 1, 1, R), out=dst.view(n_cls, R, 1, 1)))
    else:
        gw = ops.conv2d_wgrad(dl16.view(B, 1, 1, n_pad), feat.view(B, 1, 1, R)).view(n_pad, R)[:n_cls]
        if dst is not None:
            dst.copy_(gw)
            gw = dst
        grads.put(head.weight, gw)
    if head.bias is not None:
        grads.put(head.bias, ops.colsum(dl16, cols=n_cls, out=grads.dest(head.bias)))
    d_yc = ops.conv2d_dgrad(dl16.view(B, 1, 1, n_pad), pack.get(head.weight, 1), (1, 1)).view(B, R)
    if model.has_logits:
        fc = model.pre_logits.fc
        du = ops.tanh_bwd(d_yc, t32)                      # bf16 [B, R]: d tanh
        _lin_grads(grads, fc, du, yc)
        d_yc, _ = ops.gemm(du, pack.get(fc.weight, 1))    # bf16 [B, D]
    d_cls, dgn, dbn = ops.layernorm_bwd(d_yc, cls_rows, mc, rc, model.norm.weight, dx_dtype=BF16,
                                        dgamma=grads.dest(model.norm.weight), dbeta=grads.dest(model.norm.bias))
    grads.put(model.norm.weight, dgn)
    grads.put(model.norm.bias, dbn)
    g = torch.zeros(B, T, D, dtype=BF16, device=d_cls.device)   # gradient of the residual stream
    ops.copy_rows(d_cls, 0, D, g, 0, T * D, B, D)
    M = B * T
    for (blk, h, y1, m1, r1, qkv, att, lse, h2, y2, m2, r2, dact, post, dp1, dp2) in reversed(tape["blocks"]):
        att_m, mlp = blk.attn, blk.mlp
        H = att_m.num_heads
        # (stochastic depth: the branch sees the per-sample scaled gradient, the identity path - `add=g` below - the full one)
        g2 = (g if dp2 is None else ops.rowscale(g, dp2)).view(M, D)
        # h3 = h2 + fc2(gelu(fc1(LN2(h2))))
        _lin_grads(grads, mlp.fc2, g2, post.view(M, -1))
        # dgrad + GELU' in the epilogue, which also sums the columns of d_pre (= fc1 bias gradient) on the way out
        d_pre, _, st_pre = ops.gemm(g2, pack.get(mlp.fc2.weight, 1), act=3, aux_in=dact.view(M, -1), want_stats=True)
        _lin_grads(grads, mlp.fc1, d_pre, y2.view(M, D), dy_stats=st_pre)
        d_y2, _ = ops.gemm(d_pre, pack.get(mlp.fc1.weight, 1))
        g, dg2, db2 = ops.layernorm_bwd(d_y2, h2, m2, r2, blk.norm2.weight, add=g, dx_dtype=BF16,
                                        dgamma=grads.dest(blk.norm2.weight), dbeta=grads.dest(blk.norm2.bias))
        grads.put(blk.norm2.weight, dg2)
        grads.put(blk.norm2.bias, db2)
        # h2 = h + proj(attention(qkv(LN1(h))))
        g2 = (g if dp1 is None else ops.rowscale(g, dp1)).view(M, D)
        _lin_grads(grads, att_m.proj, g2, att.view(M, D))
        d_att, _ = ops.gemm(g2, pack.get(att_m.proj.weight, 1))
        dqkv = ops.attention_bwd(qkv, att, d_att.view(B, T, D), lse, H, float(att_m.scale))
        _lin_grads(grads, att_m.qkv, dqkv.view(M, 3 * D), y1.view(M, D))
        d_y1, _ = ops.gemm(dqkv.view(M, 3 * D), pack.get(att_m.qkv.weight, 1))
        g, dg1, db1 = ops.layernorm_bwd(d_y1, h, m1, r1, blk.norm1.weight, add=g, dx_dtype=BF16,
                                        dgamma=grads.dest(blk.norm1.weight), dbeta=grads.dest(blk.norm1.bias))
        grads.put(blk.norm1.weight, dg1)
        grads.put(blk.norm1.bias, db1)
        g = g.view(B, T, D)
    # ---- embedding: tokens = [cls ; patches W^T + b] + pos
    grads.put(model.pos_embed, ops.batch_rowsum(g, T * D, B, T * D, out=_flat(grads.dest(model.pos_embed))))
    grads.put(model.cls_token, ops.batch_rowsum(g, T * D, B, D, out=_flat(grads.dest(model.cls_token))))
    gp = torch.empty(B, P, D, dtype=BF16, device=g.device)
    ops.copy_rows(g, D, T * D, gp, 0, P * D, B, P * D)          # drop the class-token rows
    a = tape["patches"]
    pe = model.patch_embed.proj
    K0 = a.shape[-1]
    dst = grads.dest(pe.weight)
    gb = None
    if pe.bias is not None:
        gb = grads.dest(pe.bias)
        if gb is None:
            gb = torch.empty(D, dtype=F32, device=gp.device)
    gw = ops.conv2d_wgrad(gp.view(B * P, 1, 1, D), a.view(B * P, 1, 1, K0), out=dst.view(D, K0, 1, 1) if dst is not None else None,
                          bias_out=gb)
    grads.put(pe.weight, gw)
    if pe.bias is not None:
        grads.put(pe.bias, gb)
    return grads


def _flat(t):
    return None if t is None else t.view(-1)


class _VitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, *params):
        want_tape = any(ctx.needs_input_grad[2:])
        logits, tape = forward(model, x, model.training, want_tape)
        ctx.model, ctx.tape, ctx.params = model, tape, params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        if ctx.tape is None:
            raise RuntimeError("backward called on a forward that recorded no tape")
        grads = backward(ctx.model, ctx.tape, dlogits)
        ctx.tape = None
        out = []
        for p, need in zip(ctx.params, ctx.needs_input_grad[2:]):
            gp = grads.get(p.data_ptr()) if need else None
            out.append(gp.reshape(p.shape) if gp is not None else None)
        return (None, None, *out)


def apply(model, x):
    if not x.is_cuda:
        raise RuntimeError("deeplearning_b200 ViT runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
    params = tuple(model.parameters())
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return _VitFunction.apply(x, model, *params)
    logits, _ = forward(model, x, model.training, False)
    return logits
