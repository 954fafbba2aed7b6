"""Host-side mirrors of the reference constructors: names, shapes, surgery, and the no-CPU-fallback rule."""
import pytest
import torch
import torch.nn as nn


def test_resnet_family_state_dict_layout():
    from deeplearning_b200.classification.resnet.models import networks as N

    m = N.resnet50(num_classes=7)
    sd = m.state_dict()
    assert sd["conv1.weight"].shape == (64, 3, 7, 7)
    assert sd["layer1.0.downsample.0.weight"].shape == (256, 64, 1, 1)
    assert sd["layer4.2.conv3.weight"].shape == (2048, 512, 1, 1)
    assert sd["fc.weight"].shape == (7, 2048)
    assert len(list(m.parameters())) == 161 and len(list(m.buffers())) == 159
    assert sum(p.numel() for p in N.resnet50().parameters()) == 25557032
    assert N.resnet18().state_dict()["layer2.0.conv1.weight"].shape == (128, 64, 3, 3)
    assert N.wide_resnet50_2().state_dict()["layer1.0.conv2.weight"].shape == (128, 128, 3, 3)
    # strict round trip like classification/resnet/test.py:69
    N.resnet50(num_classes=7).load_state_dict(sd, strict=True)


def test_resnet_matches_torchvision_init_bit_for_bit():
    torchvision = pytest.importorskip("torchvision")
    from deeplearning_b200.classification.resnet.models import networks as N

    for ours, theirs in ((N.resnet50, torchvision.models.resnet50), (N.resnet34, torchvision.models.resnet34)):
        torch.manual_seed(5)
        a = ours(zero_init_residual=True).state_dict()
        torch.manual_seed(5)
        b = theirs(zero_init_residual=True).state_dict()
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_cpu_input_raises_instead_of_falling_back():
    from deeplearning_b200.classification.resnet.models.networks import resnet50

    m = resnet50()
    m.fc = nn.Linear(2048, 5)  # head surgery as in classification/resnet/train.py:79-80
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 32, 32))


def test_pretrained_needs_network():
    from deeplearning_b200.classification.resnet.models.networks import resnet50

    with pytest.raises(RuntimeError):
        resnet50(pretrained=True)
