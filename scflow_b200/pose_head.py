"""Pose regressors registered under the reference's names (models/head/pose_head.py:11-104, 110-211)."""
import torch
import torch.nn as nn

from . import ops
from .builder import HEAD
from .cnn import BaseModule, ConvModule, PackedCache


class _PoseHeadBase(BaseModule):
    _conv_feat_channels = {'Basic': [128, 128, 128], 'Large': [128, 128, 128]}
    _conv_strides = {'Basic': [2, 2, 2], 'Large': [2, 2, 2]}
    _conv_paddings = {'Basic': [1, 1, 1], 'Large': [1, 1, 1]}
    _conv_kernel_sizes = {'Basic': [3, 3, 3], 'Large': [3, 3, 3]}
    _fc_feat_channels = {'Basic': [1024, 256], 'Large': [1024, 256]}
    _feat_size = {'Basic': (32, 32), 'Large': (64, 64)}

    def _build(self, num_class, in_channels, net_type, norm_cfg, act_cfg, feat_size, rotation_mode):
        assert net_type in ['Basic', 'Small', 'Large']
        if net_type not in self._conv_feat_channels:
            raise KeyError(f'pose head has no {net_type} setting (same as the reference)')
        if feat_size is None:
            feat_size = self._feat_size[net_type]
        else:
            assert isinstance(feat_size, (list, tuple)) and len(feat_size) == 2
        self.num_class = num_class
        self.rotation_mode = rotation_mode
        layers = []
        conv_out_size = feat_size[0] * feat_size[1]
        ch = in_channels
        for ch, k, s, p in zip(self._conv_feat_channels[net_type], self._conv_kernel_sizes[net_type],
                               self._conv_strides[net_type], self._conv_paddings[net_type]):
            layers.append(ConvModule(in_channels, ch, k, stride=s, padding=p, norm_cfg=norm_cfg, act_cfg=act_cfg))
            in_channels = ch
            conv_out_size = int(conv_out_size / (s ** 2))
        self.conv_layers = nn.Sequential(*layers)
        self.conv_out_channels = ch
        fc_in = ch * conv_out_size
        self.fc_in_features = fc_in
        fcs = []
        for ch in self._fc_feat_channels[net_type]:
            fcs.append(nn.Sequential(nn.Linear(fc_in, ch), nn.ReLU()))
            fc_in = ch
        self.flatten_op = nn.Flatten(start_dim=1, end_dim=-1)
        self.fc_layers = nn.Sequential(*fcs)
        if rotation_mode == 'quaternion':
            self.rotation_out_channels = 4
        elif rotation_mode == 'ortho6d':
            self.rotation_out_channels = 6
        else:
            raise RuntimeError(f'Not supported rotation mode:{rotation_mode}')
        nc = num_class if num_class is not None else 1
        self.rotation_pred = nn.Linear(fc_in, self.rotation_out_channels * nc)
        self.translation_pred = nn.Linear(fc_in, 3 * nc)
        self._fc0_cache = PackedCache()
        self.init_weights()

    def init_weights(self):
        """Zero translation / identity rotation at init (pose_head.py:86-96, 187-198)."""
        nc = self.num_class if self.num_class is not None else 1
        nn.init.zeros_(self.translation_pred.weight)
        nn.init.zeros_(self.translation_pred.bias)
        nn.init.zeros_(self.rotation_pred.weight)
        with torch.no_grad():
            if self.rotation_mode == 'quaternion':
                self.rotation_pred.bias.copy_(torch.tensor([0., 0., 0., 1.] * nc))
            else:
                self.rotation_pred.bias.copy_(torch.tensor([1., 0., 0., 0., 1., 0.] * nc))

    def _fc0_nhwc(self, pix: int) -> torch.Tensor:
        """fc_layers.0 weight with columns permuted from the reference's NCHW flatten (c*pix + p) to NHWC (p*C + c)."""
        w = self.fc_layers[0][0].weight
        c = self.conv_out_channels
        return self._fc0_cache.get([w], lambda: w.detach().view(-1, c, pix).permute(0, 2, 1).reshape(w.shape[0], -1).contiguous())

    def _trunk(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise NotImplementedError('scflow_b200 pose head: backward is not implemented; run under torch.no_grad()')
        xh = ops.nchw_to_nhwc(x.contiguous())
        segs = [(xh, 0, x.shape[1])]
        for layer in self.conv_layers:
            y = layer.forward_nhwc(segs)
            segs = [(y, 0, y.shape[-1])]
        b, h, w, c = y.shape
        if h * w * c != self.fc_in_features:
            raise RuntimeError(f'pose head: flattened conv output has {h * w * c} features but fc expects '
                               f'{self.fc_in_features} (the reference fails the same way off its feat_size)')
        f = ops.linear(y.view(b, -1), self._fc0_nhwc(h * w), self.fc_layers[0][0].bias.detach(), 'relu')
        for fc in list(self.fc_layers)[1:]:
            f = ops.linear(f, fc[0].weight.detach(), fc[0].bias.detach(), 'relu')
        return f


@HEAD.register_module()
class SingleClassPoseHead(_PoseHeadBase):
    def __init__(self, in_channels: int, net_type: str, norm_cfg: dict, act_cfg: dict, feat_size: tuple = None,
                 rotation_mode: str = 'quaternion', init_cfg=None):
        super().__init__(init_cfg)
        self._build(None, in_channels, net_type, norm_cfg, act_cfg, feat_size, rotation_mode)

    def forward(self, x: torch.Tensor, label: torch.Tensor = None):
        f = self._trunk(x)
        return ops.pose_project(f, self.rotation_pred.weight.detach(), self.rotation_pred.bias.detach(),
                                self.translation_pred.weight.detach(), self.translation_pred.bias.detach(), None,
                                self.rotation_out_channels, 0)


@HEAD.register_module()
class MultiClassPoseHead(_PoseHeadBase):
    def __init__(self, num_class: int, in_channels: int, net_type: str, norm_cfg: dict, act_cfg: dict,
                 feat_size: tuple = None, rotation_mode: str = 'quaternion', init_cfg=None):
        super().__init__(init_cfg)
        self._build(num_class, in_channels, net_type, norm_cfg, act_cfg, feat_size, rotation_mode)

    def forward(self, x: torch.Tensor, label: torch.Tensor):
        """Returns (delta_rotation [B,rot], delta_translation [B,3]) of class ``label[0]`` for every row - the
        reference's ``index_select(dim=1, index=label)[:, 0]`` behaviour (pose_head.py:209-210), kept on purpose."""
        f = self._trunk(x)
        return ops.pose_project(f, self.rotation_pred.weight.detach(), self.rotation_pred.bias.detach(),
                                self.translation_pred.weight.detach(), self.translation_pred.bias.detach(),
                                label.contiguous(), self.rotation_out_channels, self.num_class)
