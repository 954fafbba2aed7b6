"""Data-parallel training step: flat parameter / gradient arenas, ONE gradient all-reduce, one fused optimizer launch.

B200-native counterpart of the reference's DDP recipe (others/train_with_DDP/train.py:106-111,188-201,245-253 and
classification/swin_transformer/main.py:101-103): one process per GPU, identical replicas, per-GPU BatchNorm statistics
(DDP *without* SyncBN, as the north-star asks), gradients averaged across ranks after backward, identical update on every
rank.  Instead of DDP's bucketed reducer the backward kernels write straight into one contiguous fp32 arena which is
all-reduced with a single NCCL call over NVLink/NVSwitch; the 1/world scale is folded into the fused SGD kernel.

``model.parameters()`` keep their identity: each ``p.data`` becomes a view into the parameter arena and ``p.grad`` a
view into the gradient arena, so ``state_dict()`` / checkpoints / user code reading ``.grad`` behave as in the reference.
"""
import torch
import torch.distributed as dist

from .. import ops
from .packing import weight_cache


def _engine_for(model):
    from ..classification.resnet.models.networks import ResNet

    if isinstance(model, ResNet):
        from . import resnet

        return resnet
    raise NotImplementedError(f"no B200 engine schedule for {type(model).__name__}")


class TrainStep:
    def __init__(self, model, lr=0.01, momentum=0.9, weight_decay=5e-5, process_group=None, world_size=None,
                 broadcast=True):
        self.model = model
        self.engine = _engine_for(model)
        self.lr, self.momentum, self.weight_decay = lr, momentum, weight_decay
        self.group = process_group
        if world_size is None:
            world_size = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.world = world_size
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("TrainStep needs the model on a CUDA (sm_100a) device")
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every slot 16-byte aligned
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self._gviews = {}
        with torch.no_grad():
            for p, o in zip(params, offs):
                pv = self.flat_p[o:o + p.numel()].view(p.shape)
                pv.copy_(p.data)
                p.data = pv
                gv = self.flat_g[o:o + p.numel()].view(p.shape)
                p.grad = gv
                self._gviews[p.data_ptr()] = gv
        self.params = params
        self.steps = 0
        if self.world > 1 and broadcast:
            dist.broadcast(self.flat_p, src=0, group=self.group)  # identical replicas (train_with_DDP/train.py:171-176)
            for b in model.buffers():
                dist.broadcast(b, src=0, group=self.group)
        weight_cache.bump()

    def _sink(self, param):
        return self._gviews.get(param.data_ptr())

    def step(self, images, labels, lr=None):
        """One training step on this rank's shard. Returns (loss [1] fp32 device tensor, correct int32 [B])."""
        model = self.model
        if not model.training:
            model.train()
        logits, tape = self.engine.forward(model, images, True, True)
        n_pad = (logits.shape[1] + 7) // 8 * 8
        loss, dlogits, correct = ops.softmax_xent(logits, labels, want_grad=True, ld_d=n_pad)
        self.engine.backward(model, tape, dlogits, sink=self._sink)
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
        ops.sgd_momentum_(self.flat_p, self.flat_g, self.flat_m, self.lr if lr is None else lr, self.momentum,
                          self.weight_decay, gscale=1.0 / self.world, first_step=(self.steps == 0))
        self.steps += 1
        weight_cache.bump()  # parameters changed behind autograd's back -> repack bf16 operands on next use
        return loss, correct
