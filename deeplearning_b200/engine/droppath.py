"""Stochastic depth (drop_path) for the residual branches of ConvNeXt / ViT / Swin.

Reference semantics (classification/convNext/models/networks.py:11-26, classification/vision_transformer/vit_model.py:12-40,
timm 0.4.12 ``DropPath`` used by classification/swin_transformer/models/swin_transformer.py:210,282,285):

    keep = 1 - drop_prob;  r = floor(keep + U[0,1))  per SAMPLE;  branch output = branch / keep * r

The engine draws ``U`` with ``torch.rand`` on the activation's device, one call per ``drop_path`` application and in the
reference's call order (same shape, dtype and generator, so a run seeded like the reference consumes the identical Philox
stream; under CUDA-graph capture torch's graph-safe generator advances the offset per replay).  The resulting per-sample
multiplier ``r / keep`` (fp32 ``[B]``) is applied in the epilogue of the GEMM that closes the branch (``rowscale``), is kept
on the tape, and scales the gradient entering the branch in the backward pass (``ops.rowscale``).

``replay(scales)`` is the test hook that shares masks with the CPU oracle: inside the context the engine
consumes the given multipliers (in call order) instead of drawing new ones.
"""
import contextlib

import torch
import torch.nn as nn

_replay = None  # list of fp32 [B] tensors being consumed, or None
_record = None  # list collecting the multipliers drawn, or None


def drop_prob_of(blk, train):
    """Drop probability of a block's ``drop_path`` member (0 when it is nn.Identity, in eval mode or at rate 0)."""
    dp = getattr(blk, "drop_path", None)
    if not train or dp is None or isinstance(dp, nn.Identity):
        return 0.0
    return float(getattr(dp, "drop_prob", 0.0) or 0.0)


def sample_scale(drop_prob, batch, ndim, device):
    """fp32 [batch] multiplier of one drop_path application on an ``ndim``-dimensional activation, or None (rate 0)."""
    if drop_prob <= 0.0:
        return None
    if _replay is not None:
        if not _replay:
            raise RuntimeError("droppath.replay: more drop_path applications than recorded multipliers")
        s = _replay.pop(0).to(device=device, dtype=torch.float32).contiguous()
        if s.numel() != batch:
            raise RuntimeError("droppath.replay: multiplier of the wrong batch size")
    else:
        keep = 1.0 - drop_prob
        r = keep + torch.rand((batch,) + (1,) * (ndim - 1), dtype=torch.float32, device=device)
        r.floor_()
        s = (r / keep).view(batch)
    if _record is not None:
        _record.append(s.detach().clone())
    return s


@contextlib.contextmanager
def replay(scales):
    """Consume the given per-sample multipliers (an iterable of fp32 [B] tensors, reference call order)."""
    global _replay
    prev, _replay = _replay, [s for s in scales]
    try:
        yield
    finally:
        _replay = prev


@contextlib.contextmanager
def record():
    """Collect the multipliers drawn inside the context (list of fp32 [B] tensors, call order)."""
    global _record
    prev, _record = _record, []
    try:
        yield _record
    finally:
        _record = prev
