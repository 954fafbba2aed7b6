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


def test_swin_state_dict_layout_and_decay_groups():
    from deeplearning_b200.classification.swin_transformer.models.build import build_model
    from deeplearning_b200.classification.swin_transformer.models.swin_transformer import SwinTransformer
    from deeplearning_b200.engine.trainer import model_no_decay_rule
    from types import SimpleNamespace as NS

    m = SwinTransformer(drop_path_rate=0.0)
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 28288354
    assert sd["patch_embed.proj.weight"].shape == (96, 3, 4, 4) and sd["patch_embed.norm.weight"].shape == (96,)
    assert sd["layers.0.blocks.0.attn.relative_position_bias_table"].shape == (169, 3)
    assert sd["layers.0.blocks.0.attn.relative_position_index"].shape == (49, 49)
    assert sd["layers.0.blocks.1.attn_mask"].shape == (64, 49, 49) and "layers.0.blocks.0.attn_mask" not in sd
    assert "layers.3.blocks.1.attn_mask" not in sd          # 7x7 stage: window == resolution, shift forced to 0 (:187-190)
    assert sd["layers.2.downsample.reduction.weight"].shape == (768, 1536) and sd["layers.2.downsample.norm.weight"].shape == (1536,)
    assert sd["head.weight"].shape == (1000, 768)
    mask = sd["layers.1.blocks.1.attn_mask"]
    assert set(mask.unique().tolist()) == {0.0, -100.0}
    rule = model_no_decay_rule(m)
    named = dict(m.named_parameters())
    nd = {n for n, p in named.items() if rule(n, p)}
    assert "layers.0.blocks.0.attn.relative_position_bias_table" in nd and "norm.weight" in nd and "head.bias" in nd
    assert "head.weight" not in nd and "layers.0.blocks.0.attn.qkv.weight" not in nd
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 224, 224))
    cfg = NS(MODEL=NS(TYPE="swin", NUM_CLASSES=10, DROP_RATE=0.0, DROP_PATH_RATE=0.0,
                      SWIN=NS(PATCH_SIZE=4, IN_CHANS=3, EMBED_DIM=96, DEPTHS=[2, 2], NUM_HEADS=[3, 6], WINDOW_SIZE=7, MLP_RATIO=4.,
                              QKV_BIAS=True, QK_SCALE=None, APE=False, PATCH_NORM=True)),
             DATA=NS(IMG_SIZE=224), TRAIN=NS(USE_CHECKPOINT=False), FUSED_WINDOW_PROCESS=True)
    small = build_model(cfg)
    assert small.head.weight.shape == (10, 192) and len(small.layers) == 2
    cfg.MODEL.TYPE = "swinv2"
    with pytest.raises(NotImplementedError):
        build_model(cfg)
