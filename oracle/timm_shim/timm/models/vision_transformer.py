from oracle import backbones
from . import efficientnet as _e


def vit_small_patch32_224_in21k(pretrained=False, pretrained_cfg=None, num_classes=0, **kw):
    return backbones.seeded_init(backbones.build('vit_s_32'), *_e._SEED_ARGS)


def vit_base_patch32_224_in21k(pretrained=False, pretrained_cfg=None, num_classes=0, **kw):
    return backbones.seeded_init(backbones.build('vit_b_32'), *_e._SEED_ARGS)


def vit_base_patch32_224_clip_laion2b(pretrained=False, pretrained_cfg=None, num_classes=0, **kw):
    return backbones.seeded_init(backbones.build('vit_b_32_clip'), *_e._SEED_ARGS)
