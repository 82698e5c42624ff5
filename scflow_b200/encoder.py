"""RAFTEncoder ('Basic' ResNet-ish feature / context encoder) with the reference's state-dict keys.

NOT part of the replaced hot path: it feeds the decoder and runs as stock
cuDNN convolutions through PyTorch, exactly like the reference (models/encoder/raft_encoder.py:286-314,
models/backbone/resnet.py:14-94,678-773).  It exists so that ``SCFlowRefiner.get_pose`` is runnable end to end.
"""
from typing import Optional, Sequence, Union

import torch
import torch.nn as nn

from .builder import ENCODERS
from .cnn import BaseModule


def _norm(cfg: dict, channels: int, postfix=''):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    if typ in ('BN', 'BN2d', 'SyncBN'):
        name, layer = 'bn', nn.BatchNorm2d(channels, eps=cfg.get('eps', 1e-5))
    elif typ == 'IN':
        name, layer = 'in', nn.InstanceNorm2d(channels, eps=cfg.get('eps', 1e-5))
    elif typ == 'GN':
        name, layer = 'gn', nn.GroupNorm(cfg['num_groups'], channels, eps=cfg.get('eps', 1e-5))
    else:
        raise KeyError(f'unsupported norm {typ}')
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name + str(postfix), layer


class BasicBlock(BaseModule):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, norm_cfg=dict(type='BN')):
        super().__init__()
        self.norm1_name, norm1 = _norm(norm_cfg, planes, 1)
        self.norm2_name, norm2 = _norm(norm_cfg, planes, 2)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=True)
        self.add_module(self.norm1_name, norm1)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=True)
        self.add_module(self.norm2_name, norm2)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(getattr(self, self.norm1_name)(self.conv1(x)))
        out = getattr(self, self.norm2_name)(self.conv2(out))
        identity = x if self.downsample is None else self.downsample(x)
        return self.relu(out + identity)


class ResLayer(nn.Sequential):
    def __init__(self, inplanes, planes, num_blocks, stride, norm_cfg):
        downsample = None
        if stride != 1 or inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride=stride, bias=True), _norm(norm_cfg, planes)[1])
        layers = [BasicBlock(inplanes, planes, stride, downsample, norm_cfg)]
        for _ in range(1, num_blocks):
            layers.append(BasicBlock(planes, planes, 1, None, norm_cfg))
        super().__init__(*layers)


@ENCODERS.register_module()
class RAFTEncoder(BaseModule):
    _stem_channels = {'Basic': 64}
    _base_channels = {'Basic': (64, 96, 128)}
    _strides = {'Basic': (1, 2, 2)}

    def __init__(self, in_channels: int, out_channels: int, scale: float = 1 / 8, net_type: str = 'Basic',
                 norm_cfg: dict = dict(type='BN', requires_grad=True), norm_eval: bool = False,
                 init_cfg: Optional[Union[dict, list]] = None, **unsupported) -> None:
        super().__init__(init_cfg=init_cfg)
        if net_type != 'Basic':
            raise NotImplementedError('only the shipped net_type="Basic" encoder is provided')
        extra = {k: v for k, v in unsupported.items() if v not in (None, False, -1)}
        if extra:
            raise NotImplementedError(f'unsupported RAFTEncoder options: {sorted(extra)}')
        self.in_channels, self.out_channels, self.scale = in_channels, out_channels, scale
        self.norm_eval = norm_eval
        stem = self._stem_channels[net_type]
        self.conv1 = nn.Conv2d(in_channels, stem, 7, stride=1 if scale == 1 / 4 else 2, padding=3, bias=True)
        self.norm1_name, norm1 = _norm(norm_cfg, stem, 1)
        self.add_module(self.norm1_name, norm1)
        self.relu = nn.ReLU(inplace=True)
        self.res_layers = []
        inplanes = stem
        for i, (planes, stride) in enumerate(zip(self._base_channels[net_type], self._strides[net_type])):
            name = f'res_layer{i + 1}'
            self.add_module(name, ResLayer(inplanes, planes, 2, stride, norm_cfg))
            self.res_layers.append(name)
            inplanes = planes
        self.conv2 = nn.Conv2d(inplanes, out_channels, 1)
        self.init_weights()

    def init_weights(self):
        """Kaiming(fan_out, relu) for convs, constant 1/0 for norms (configs/refine_models/scflow.py:29-36)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)) and m.weight is not None:
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.relu(getattr(self, self.norm1_name)(self.conv1(x)))
        for name in self.res_layers:
            x = getattr(self, name)(x)
        return self.conv2(x)

    def train(self, mode: bool = True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self
