"""ConvModule / BaseModule stand-ins with mmcv 1.3.16 parameter naming, executing on the scflow_b200 kernels.

The reference builds every layer of the hot path from ``mmcv.cnn.ConvModule`` (models/decoder/raft_decoder.py:140-148,
201-221, 272-277; models/head/pose_head.py:150-159).  Only its wiring matters: order
conv -> norm -> act, ``bias='auto'`` => ``bias = not with_norm``, default activation ReLU, ``act_cfg=None`` => none;
sub-module names ``conv`` / ``gn`` so that state-dict keys match released checkpoints.
"""
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from . import ops


class BaseModule(nn.Module):
    """mmcv.runner.BaseModule: nn.Module that remembers ``init_cfg`` and has ``init_weights``."""

    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()


_ACT_NAMES = {'ReLU': 'relu', 'Sigmoid': 'sigmoid', 'Tanh': 'tanh'}


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


class PackedCache:
    """Caches a derived device tensor (packed weights) until any source parameter changes."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, params, make):
        key = tuple((p.data_ptr(), p._version, p.device) for p in params)
        if key != self._key:
            self._val = make()
            self._key = key
        return self._val


class ConvModule(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, Tuple[int, int]], stride=1,
                 padding=0, dilation=1, groups=1, bias='auto', conv_cfg: Optional[dict] = None,
                 norm_cfg: Optional[dict] = None, act_cfg: Optional[dict] = dict(type='ReLU'), inplace: bool = True,
                 order: tuple = ('conv', 'norm', 'act')):
        super().__init__()
        if conv_cfg is not None and conv_cfg.get('type', 'Conv2d') not in ('Conv2d', 'Conv'):
            raise NotImplementedError(f'conv_cfg {conv_cfg} is not supported')
        if dilation != 1 or groups != 1 or tuple(order) != ('conv', 'norm', 'act'):
            raise NotImplementedError('only dense, undilated conv->norm->act ConvModules are on the SCFlow hot path')
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        self.in_channels, self.out_channels = in_channels, out_channels
        # parameter holder only: its own forward is never used
        self.conv = nn.Conv2d(in_channels, out_channels, self.kernel_size, self.stride, self.padding, bias=bias)
        self.norm_name = None
        if self.with_norm:
            cfg = dict(norm_cfg)
            typ = cfg.pop('type')
            if typ != 'GN':
                raise NotImplementedError(f'norm {typ} is not on the SCFlow decoder path (only GN in the pose head)')
            requires_grad = cfg.pop('requires_grad', True)
            self.num_groups = cfg.pop('num_groups')
            self.norm_eps = cfg.pop('eps', 1e-5)
            self.norm_name = 'gn'
            self.gn = nn.GroupNorm(self.num_groups, out_channels, eps=self.norm_eps)
            for p in self.gn.parameters():
                p.requires_grad = requires_grad
        self.act = 'none'
        if self.with_activation:
            typ = act_cfg['type']
            if typ not in _ACT_NAMES:
                raise NotImplementedError(f'activation {typ} is not supported')
            self.act = _ACT_NAMES[typ]
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)
        self._packed = PackedCache()

    def packed_weight(self) -> torch.Tensor:
        return self._packed.get([self.conv.weight], lambda: ops.pack_conv_weight([self.conv.weight.detach()]))

    def forward_nhwc(self, segs, out=None, out_coff=0) -> torch.Tensor:
        bias = None if self.conv.bias is None else self.conv.bias.detach()
        act = self.act if not self.with_norm else 'none'
        y = ops.conv2d_nhwc(segs, self.packed_weight(), bias, self.out_channels, self.kernel_size, self.stride,
                            self.padding, act=act, out=out, out_coff=out_coff)
        if self.with_norm:
            if self.act != 'relu' or out_coff != 0 or y.shape[-1] != self.out_channels:
                raise NotImplementedError('GroupNorm ConvModules are fused with ReLU only')
            ops.group_norm_relu_(y, self.gn.weight.detach(), self.gn.bias.detach(), self.num_groups, self.norm_eps)
        return y

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """NCHW in / NCHW out, like the reference module (layout converted around the NHWC kernel)."""
        xh = ops.nchw_to_nhwc(x.contiguous())
        y = self.forward_nhwc([(xh, 0, self.in_channels)])
        return ops.nhwc_to_nchw(y)
