"""Derived bf16 GEMM operands of the fp32 master parameters.

The reference keeps OIHW fp32 ``nn.Parameter``s (checkpoints are plain ``state_dict()``s, classification/resnet/train.py:130),
so the packed [O][taps*I] / [I][taps*O] bf16 copies the tensor cores read are caches, rebuilt whenever the parameter's
version counter or storage changes (i.e. after every optimizer step or ``load_state_dict``).
"""
import weakref

import torch

from .. import ops


class _WeightCache:
    def __init__(self):
        self._store = {}
        self.generation = 0  # bumped by code that updates parameters behind autograd's back (fused optimizer kernels)

    def bump(self):
        self.generation += 1

    def get(self, param, mode, ld=None, pad_rows=None, pad_cols=None):
        key = (id(param), mode, ld, pad_rows, pad_cols)
        hit = self._store.get(key)
        stamp = (param._version, param.data_ptr(), self.generation)
        if hit is not None and hit[0] == stamp and hit[2]() is param:
            return hit[1]
        w = param.detach()
        if mode == 0:
            packed = ops.pack_weight(w, 0, ld=ld)
            if pad_rows is not None and pad_rows != packed.shape[0]:
                full = torch.zeros(pad_rows, packed.shape[1], dtype=packed.dtype, device=packed.device)
                full[: packed.shape[0]] = packed
                packed = full
        else:
            packed = ops.pack_weight(w, 1, ld=pad_cols if pad_cols is not None else ld)
        self._store[key] = (stamp, packed, weakref.ref(param))
        if len(self._store) > 4096:
            self._store = {k: v for k, v in self._store.items() if v[2]() is not None}
        return packed

    def clear(self):
        self._store.clear()


weight_cache = _WeightCache()
