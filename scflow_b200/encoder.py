"""RAFTEncoder ('Basic' ResNet-ish feature / context encoder) with the reference's state-dict keys
(models/encoder/raft_encoder.py:286-314, models/backbone/resnet.py:14-94,678-773).

SURVEY.md §8(f) rank 1 - the component feeding the refinement loop.  In inference (``eval()`` on a CUDA device,
``norm_cfg`` IN or BN) ``forward`` is ONE call into the C ABI (``scf_encoder_forward``): tcgen05 split-bf16
convolutions, fused InstanceNorm / folded BatchNorm.  In training mode (batch statistics, autograd) it falls through
to the plain nn.Module graph below, which is also what defines the parameters / state-dict keys.
"""
from typing import Optional, Sequence, Union

import ctypes as C

import os

import torch
import torch.nn as nn

from . import _lib
from .builder import ENCODERS
from .cnn import BaseModule, PackedCache


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
        self.norm_type = norm_cfg['type']
        self.norm_eps = float(norm_cfg.get('eps', 1e-5))
        self.use_native = True            # set False to force the nn.Module graph (stock cuDNN) in eval mode too
        self._arena = PackedCache()
        self._ws = {}
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

    # ------------------------------------------------------------------ native (C ABI) inference path
    def _native_refusal(self, x: torch.Tensor) -> Optional[str]:
        """Why the native (C ABI) path cannot take this call, or None when it can."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return ('autograd is enabled with tensors that require grad; the native encoder is inference-only - call it under '
                    'torch.no_grad(), or put the module in train() mode to use the differentiable nn.Module graph')
        if not x.is_cuda:
            return 'the input is not a CUDA tensor (scflow_b200 has no CPU path)'
        if self.norm_type not in ('IN', 'BN') or self.in_channels != 3 or self.out_channels != 256 or self.scale != 1 / 8:
            return 'only the shipped configuration (3 -> 256 channels, scale 1/8, IN or BN) is implemented natively'
        if self.norm_eps != 1e-5:
            return f'norm eps {self.norm_eps} != 1e-5 (the native kernels fold the default eps)'
        if x.shape[-2] % 8 or x.shape[-1] % 8:
            return f'input size {tuple(x.shape[-2:])} is not a multiple of 8'
        return None

    def _native_ok(self, x: torch.Tensor) -> bool:
        return self.use_native and not self.training and self._native_refusal(x) is None

    def _forward_native(self, x: torch.Tensor, ex: Optional['_lib.EncoderOut'] = None) -> Optional[torch.Tensor]:
        """One scf_encoder_forward call; with ``ex`` the final convolution writes the consumer's buffers described by the
        ``EncoderOut`` structure (scf_encoder_forward_ex) and nothing is returned."""
        lib = _lib.load()
        norm = _lib.ENC_NORM_IN if self.norm_type == 'IN' else _lib.ENC_NORM_BN
        mods = dict(self.named_modules())
        tensors = []
        for conv_name, norm_name in _lib.ENCODER_UNITS:
            conv = mods[conv_name]
            unit = [conv.weight, conv.bias, None, None, None, None]
            if norm == _lib.ENC_NORM_BN and norm_name is not None:
                bn = mods[norm_name.format(n='bn')]
                unit[2:] = [bn.weight, bn.bias, bn.running_mean, bn.running_var]
            tensors.extend(unit)
        live = [t for t in tensors if t is not None]

        def make():
            arena = torch.empty(lib.scf_encoder_packed_bytes(), device=x.device, dtype=torch.uint8)
            srcs = [None if t is None else t.detach().contiguous().float() for t in tensors]
            arr = (C.c_void_p * len(srcs))(*[None if t is None else t.data_ptr() for t in srcs])
            _lib.check(lib.scf_encoder_pack(norm, arr, _lib.ptr(arena), _lib.stream_ptr()), 'scf_encoder_pack')
            return arena

        arena = self._arena.get(live, make)
        n, _, h, w = x.shape
        key = (n, h, w, str(x.device))
        ws = self._ws.get(key)
        if ws is None:
            self._ws = {key: torch.empty(lib.scf_encoder_workspace_bytes(n, h, w), device=x.device, dtype=torch.uint8)}
            ws = self._ws[key]
        xin = x.detach().contiguous().float()
        if ex is not None:
            _lib.check(lib.scf_encoder_forward_ex(norm, _lib.ptr(arena), _lib.ptr(xin), n, h, w, C.byref(ex), _lib.ptr(ws), ws.numel(),
                                                  _lib.stream_ptr()), 'scf_encoder_forward_ex')
            return None
        out = torch.empty(n, self.out_channels, h // 8, w // 8, device=x.device, dtype=torch.float32)
        _lib.check(lib.scf_encoder_forward(norm, _lib.ptr(arena), _lib.ptr(x.detach().contiguous().float()), n, h, w,
                                           _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), 'scf_encoder_forward')
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """eval(): ONE C-ABI call, or an error saying why not - never a silent switch to stock PyTorch kernels.  train(): the
        differentiable nn.Module graph below (batch statistics, autograd).  ``use_native = False`` is the explicit opt-out used
        by the tests that compare the native path with cuDNN."""
        if not self.training and self.use_native:
            why = self._native_refusal(x)
            if why is not None:
                raise RuntimeError(f'RAFTEncoder (eval): {why}')
            with torch.cuda.device(x.device):
                return self._forward_native(x)
        if os.environ.get('SCFLOW_TRAIN_CHANNELS_LAST', '0') != '0':
            x = x.contiguous(memory_format=torch.channels_last)       # cuDNN's tensor-core kernels are NHWC: no per-layer layout launches
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
