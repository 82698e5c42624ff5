"""The reference's six registries and ``build_*`` one-liners (models/{backbone,decoder,encoder,head,loss,refiner}/builder.py:1-6)."""
from .registry import Registry, build_from_cfg

REFINERS = Registry('refiner')
DECODERS = Registry('decoder')
ENCODERS = Registry('encoder')
HEAD = Registry('head')
LOSSES = Registry('loss')
BACKBONES = Registry('backbone')


def build_refiner(cfg):
    return build_from_cfg(cfg, REFINERS)


def build_decoder(cfg):
    return build_from_cfg(cfg, DECODERS)


def build_encoder(cfg):
    return build_from_cfg(cfg, ENCODERS)


def build_head(cfg):
    return build_from_cfg(cfg, HEAD)


def build_loss(cfg):
    return build_from_cfg(cfg, LOSSES)


def build_backbone(cfg):
    return build_from_cfg(cfg, BACKBONES)
