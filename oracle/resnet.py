"""Oracle: ResNet forward (torchvision-style v1.5) restated functionally in fp32 PyTorch.

Follows classification/resnet/models/networks.py of the reference:
  stem conv7x7/2 -> BN -> ReLU -> maxpool3x3/2      :206-209
  Bottleneck: 1x1 -> BN -> ReLU -> 3x3(stride) -> BN -> ReLU -> 1x1 -> BN -> (+downsample 1x1/BN) -> add -> ReLU   :104-124
  BasicBlock: 3x3(stride) -> BN -> ReLU -> 3x3 -> BN -> (+downsample) -> add -> ReLU                              :59-75
  avgpool -> flatten -> fc                                                                                        :216-218
BatchNorm uses nn.BatchNorm2d defaults (eps 1e-5, momentum 0.1); in train mode the running statistics in ``state`` are
updated in place exactly like the module would.
"""
import torch
import torch.nn.functional as F


def _bn(state, prefix, x, train, eps=1e-5, momentum=0.1):
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    if train and (prefix + ".num_batches_tracked") in state:
        state[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, state[prefix + ".weight"], state[prefix + ".bias"], train, momentum, eps)


def _block(state, p, x, stride, train, momentum=0.1):
    bottleneck = (p + ".conv3.weight") in state
    if bottleneck:
        out = F.relu(_bn(state, p + ".bn1", F.conv2d(x, state[p + ".conv1.weight"]), train, momentum=momentum))
        out = F.relu(_bn(state, p + ".bn2", F.conv2d(out, state[p + ".conv2.weight"], stride=stride, padding=1), train, momentum=momentum))
        out = _bn(state, p + ".bn3", F.conv2d(out, state[p + ".conv3.weight"]), train, momentum=momentum)
    else:
        out = F.relu(_bn(state, p + ".bn1", F.conv2d(x, state[p + ".conv1.weight"], stride=stride, padding=1), train, momentum=momentum))
        out = _bn(state, p + ".bn2", F.conv2d(out, state[p + ".conv2.weight"], padding=1), train, momentum=momentum)
    if (p + ".downsample.0.weight") in state:
        x = _bn(state, p + ".downsample.1", F.conv2d(x, state[p + ".downsample.0.weight"], stride=stride), train, momentum=momentum)
    return F.relu(out + x)


def resnet_forward(state, x, train=False, momentum=0.1):
    """state: dict with the reference's state_dict keys (tensors may require grad); x: [B,3,H,W] fp32."""
    h = F.conv2d(x, state["conv1.weight"], stride=2, padding=3)
    h = F.max_pool2d(F.relu(_bn(state, "bn1", h, train, momentum=momentum)), 3, 2, 1)
    for li in range(1, 5):
        bi = 0
        while f"layer{li}.{bi}.conv1.weight" in state:
            stride = 2 if (li > 1 and bi == 0) else 1
            h = _block(state, f"layer{li}.{bi}", h, stride, train, momentum)
            bi += 1
    h = torch.flatten(F.adaptive_avg_pool2d(h, 1), 1)
    return F.linear(h, state["fc.weight"], state["fc.bias"])


def train_step_grads(state, x, labels):
    """One reference training-step's worth of math (classification/resnet/utils.py:38-43): logits, CE loss, gradients.
    Returns (logits, loss, {name: grad}) with ``state`` updated like the module's buffers would be."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state.items() if v.is_floating_point()
              and "running_" not in k}
    work = dict(state)
    work.update(params)
    logits = resnet_forward(work, x, train=True)
    loss = F.cross_entropy(logits, labels)
    grads = torch.autograd.grad(loss, list(params.values()))
    for k in state:
        if "running_" in k or "num_batches" in k:
            state[k] = work[k]
    return logits.detach(), loss.detach(), dict(zip(params.keys(), grads))
