from oracle import backbones
from oracle.backbones import EfficientNet  # noqa: F401  (film.py:35 isinstance target)


def tf_efficientnet_b0(pretrained=False, pretrained_cfg=None, num_classes=0, **kw):
    assert num_classes == 0
    return backbones.seeded_init(backbones.build('efficientnet_b0'), *_SEED_ARGS)


def tf_efficientnetv2_s_in21k(pretrained=False, pretrained_cfg=None, num_classes=0, **kw):
    assert num_classes == 0
    return backbones.seeded_init(backbones.build('efficientnet_v2_s'), *_SEED_ARGS)


_SEED_ARGS = (1991, None)  # (seed, calib_input): set by make_golden before construction
