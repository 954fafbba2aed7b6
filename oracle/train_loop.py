"""Oracle: the reference's epoch loop restated (classification/resnet/utils.py:28-57 ``train_one_epoch``):
for each (images, labels): pred = model(images); acc += (argmax == labels); loss = CE(pred, labels); loss.backward();
optimizer.step(); optimizer.zero_grad().  The optimizer is SGD(momentum 0.9, weight_decay 5e-5) as built at
classification/resnet/train.py:96.  Runs on CPU in fp32 on the oracle's functional models; used as the parity checker of the
trainer tests and as bench.py's CPU baseline / ``--impl reference`` arm (kind "port": the reference's Python modules cannot
travel to the GPU box, and ``tests/golden/make_golden.py`` shows this restatement is bit-identical to them).
"""
import torch
import torch.nn.functional as F


class CpuSgdTrainer:
    def __init__(self, forward_fn, state, lr=0.01, momentum=0.9, weight_decay=5e-5):
        self.forward_fn = forward_fn
        self.state = state
        self.param_names = [k for k, v in state.items() if v.is_floating_point() and "running_" not in k]
        for k in self.param_names:
            state[k] = state[k].detach().clone().requires_grad_(True)
        self.opt = torch.optim.SGD([state[k] for k in self.param_names], lr=lr, momentum=momentum,
                                   weight_decay=weight_decay)

    def step(self, images, labels):
        pred = self.forward_fn(self.state, images, train=True)
        correct = int((pred.argmax(1) == labels).sum())
        loss = F.cross_entropy(pred, labels)
        loss.backward()
        self.opt.step()
        self.opt.zero_grad()
        return float(loss.detach()), correct
