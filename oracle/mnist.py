"""Oracle: the reference's two MNIST toy nets (classification/mnist/models/network.py:7-31 mnist_cnn, :34-70 mnist_fcn)
restated functionally.  Conv2d(3->32,3x3,p1)+ReLU+MaxPool2 ; (32->64) ; (64->64) ; then either Linear(576,128)+ReLU+
Linear(128,C) (cnn) or conv3x3 valid (64->128)+ReLU + conv1x1 (128->C) (fcn)."""
import torch
import torch.nn.functional as F


def mnist_fcn_forward(s, x):
    for i in (1, 2, 3):
        x = F.max_pool2d(F.relu(F.conv2d(x, s[f"conv{i}.0.weight"], s[f"conv{i}.0.bias"], padding=1)), 2, 2)
    x = F.relu(F.conv2d(x, s["conv4.0.weight"], s["conv4.0.bias"]))
    x = F.conv2d(x, s["conv5.0.weight"], s["conv5.0.bias"])
    return torch.flatten(x, 1)


def mnist_cnn_forward(s, x):
    for i in (0, 3, 6):
        x = F.max_pool2d(F.relu(F.conv2d(x, s[f"backbone.{i}.weight"], s[f"backbone.{i}.bias"], padding=1)), 2, 2)
    x = torch.flatten(x, 1)
    x = F.relu(F.linear(x, s["fc.0.weight"], s["fc.0.bias"]))
    return F.linear(x, s["fc.2.weight"], s["fc.2.bias"])
