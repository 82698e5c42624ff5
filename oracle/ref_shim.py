"""Import shim that runs the UNMODIFIED reference hot-path files from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). Used in the build container by
``oracle/make_golden.py`` to (a) validate the restatement in ``oracle/scflow_oracle.py`` against the
reference's own code and (b) generate the committed fixtures under ``tests/golden/``.
Nothing in the product package, the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file:
``/root/reference`` does not exist on the GPU box.

The reference needs mmcv 1.3.16 / kornia / pytorch3d / trimesh, none of which is installed (no
network).  Only wiring is taken from them on this path, so we register minimal stand-in
modules in ``sys.modules`` and skip the reference's broken package ``__init__`` files
(models/__init__.py:1-6 imports names its sub-packages do not export).
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('SCFLOW_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'models', 'decoder'))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


class _Registry:
    """mmcv.utils.Registry: name -> class table with a decorator."""

    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def get(self, key):
        return self.module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop('type')
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f'{typ} is not in the {registry.name} registry')
    return cls(**args)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class _Sequential(_BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        _BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


def _build_conv_layer(cfg, *args, **kwargs):
    assert cfg is None or cfg.get('type') in (None, 'Conv2d', 'Conv')
    return nn.Conv2d(*args, **kwargs)


def _build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if typ in ('BN', 'BN2d', 'SyncBN'):
        name, layer = 'bn', nn.BatchNorm2d(num_features, **cfg)
    elif typ == 'IN':
        name, layer = 'in', nn.InstanceNorm2d(num_features, **cfg)
    elif typ == 'GN':
        name, layer = 'gn', nn.GroupNorm(num_channels=num_features, **cfg)
    else:
        raise KeyError(typ)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name + str(postfix), layer


_ACTS = {'ReLU': nn.ReLU, 'Sigmoid': nn.Sigmoid, 'Tanh': nn.Tanh, 'LeakyReLU': nn.LeakyReLU}


def _build_activation_layer(cfg):
    cfg = dict(cfg)
    typ = cfg.pop('type')
    if typ in ('ReLU', 'LeakyReLU'):
        cfg.setdefault('inplace', True)   # mmcv ConvModule default inplace=True
    return _ACTS[typ](**cfg)


class _ConvModule(nn.Module):
    """mmcv 1.3.16 ConvModule semantics: order conv->norm->act, bias='auto' => bias = not with_norm,
    default act_cfg=ReLU, act_cfg=None => no activation. Sub-module names conv / <norm_name> / activate."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias='auto', conv_cfg=None, norm_cfg=None, act_cfg=dict(type='ReLU'), inplace=True,
                 with_spectral_norm=False, padding_mode='zeros', order=('conv', 'norm', 'act')):
        super().__init__()
        assert order == ('conv', 'norm', 'act') and padding_mode == 'zeros'
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = _build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            act_cfg = dict(act_cfg)
            if act_cfg['type'] not in ('Tanh', 'PReLU', 'Sigmoid', 'HSigmoid', 'Swish'):
                act_cfg.setdefault('inplace', inplace)
            self.activate = _build_activation_layer(act_cfg)
        nn.init.kaiming_normal_(self.conv.weight, a=0, mode='fan_out', nonlinearity='relu')
        if self.conv.bias is not None:
            nn.init.constant_(self.conv.bias, 0)

    def forward(self, x):
        x = self.conv(x)
        if self.with_norm:
            x = getattr(self, self.norm_name)(x)
        if self.with_activation:
            x = self.activate(x)
        return x


def _dummy(name):
    return type(name, (), {'__init__': lambda self, *a, **k: None})


_INSTALLED = False


def _knn_points(p1, p2, K=1, **kwargs):
    """Stand-in for pytorch3d.ops.knn_points (pytorch3d is not installed; the reference depends on the fork
    YangHai-1218/pytorch3d, README.md:29, unpinned): brute-force K=1 nearest neighbour under the squared L2 distance,
    which is the published semantics of knn_points(norm=2). Returns an object with .dists [N,P1,K] and .idx [N,P1,K]."""
    assert K == 1
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    dists, idx = d.min(dim=2)
    return types.SimpleNamespace(dists=dists[..., None], idx=idx[..., None], knn=None)


def install():
    """Register stand-in third-party modules and namespace packages for the reference."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    hooks = _Registry('hook')
    _mod('mmcv', flow2rgb=None, mkdir_or_exist=lambda *a, **k: None, is_str=lambda s: isinstance(s, str))
    _mod('mmcv.utils', Registry=_Registry, build_from_cfg=_build_from_cfg)
    _mod('mmcv.runner', BaseModule=_BaseModule, Sequential=_Sequential)
    _mod('mmcv.runner.hooks', HOOKS=hooks, Hook=_dummy('Hook'))
    _mod('mmcv.runner.hooks.logger', TensorboardLoggerHook=_dummy('TensorboardLoggerHook'),
         TextLoggerHook=_dummy('TextLoggerHook'))
    _mod('mmcv.runner.dist_utils', master_only=lambda f: f)
    _mod('mmcv.cnn', ConvModule=_ConvModule, build_conv_layer=_build_conv_layer,
         build_norm_layer=_build_norm_layer, build_activation_layer=_build_activation_layer,
         build_plugin_layer=None)
    _mod('mmcv.ops', Correlation=_dummy('Correlation'), furthest_point_sample=None)
    _mod('kornia')
    _mod('kornia.geometry')
    _mod('kornia.geometry.conversions')
    _mod('kornia.augmentation', AugmentationSequential=_dummy('AugmentationSequential'))
    _mod('pytorch3d')
    _mod('pytorch3d.ops', knn_points=_knn_points)
    _mod('pytorch3d.structures', join_meshes_as_batch=None)
    names = ['PointLights', 'PerspectiveCameras', 'BlendParams', 'MeshRasterizer', 'RasterizationSettings',
             'HardPhongShader', 'SoftPhongShader', 'HardGouraudShader', 'SoftGouraudShader',
             'SoftSilhouetteShader', 'HardFlatShader']
    _mod('pytorch3d.renderer', **{n: _dummy(n) for n in names})
    _mod('pytorch3d.renderer.mesh')
    _mod('pytorch3d.renderer.mesh.renderer', MeshRendererWithFragments=_dummy('MeshRendererWithFragments'))
    _mod('pytorch3d.io')
    _mod('pytorch3d.io.ply_io', MeshPlyFormat=_dummy('MeshPlyFormat'))
    _mod('iopath')
    _mod('iopath.common')
    _mod('iopath.common.file_io', PathManager=_dummy('PathManager'))
    _mod('trimesh', load=None)
    _mod('turtle', forward=None)
    # namespace stand-ins so the broken/heavy package __init__ files never run
    for pkg in ('models', 'datasets'):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(REFERENCE_ROOT, pkg)]
        sys.modules[pkg] = m
    _INSTALLED = True


def load_reference():
    """Returns a namespace with the reference's own classes/functions for the hot path."""
    install()
    from models.decoder.scflow_decoder import SCFlowDecoder
    from models.decoder.raft_decoder import CorrelationPyramid, MotionEncoder, ConvGRU, XHead
    from models.utils.corr_lookup import CorrLookup
    from models.utils import pose as pose_mod
    from models.head.pose_head import MultiClassPoseHead, SingleClassPoseHead
    from models.encoder.raft_encoder import RAFTEncoder
    from models.utils import flow as flow_mod
    from models.loss import sequence_loss as seq_loss_mod
    from models.loss import point_matching_loss as pm_loss_mod
    return types.SimpleNamespace(
        flow=flow_mod, sequence_loss=seq_loss_mod, point_matching_loss=pm_loss_mod,
        SCFlowDecoder=SCFlowDecoder, CorrelationPyramid=CorrelationPyramid, MotionEncoder=MotionEncoder,
        ConvGRU=ConvGRU, XHead=XHead, CorrLookup=CorrLookup, pose=pose_mod,
        MultiClassPoseHead=MultiClassPoseHead, SingleClassPoseHead=SingleClassPoseHead,
        RAFTEncoder=RAFTEncoder)


# Constructor kwargs of the shipped config (configs/refine_models/scflow.py:51-74 and :23-50).
def decoder_cfg(iters=8, num_class=21):
    return dict(
        net_type='Basic', num_levels=4, radius=4, iters=iters, detach_flow=True, detach_mask=True,
        detach_pose=True, detach_depth_for_xy=True, mask_flow=False, mask_corr=False,
        pose_head_cfg=dict(type='MultiClassPoseHead', num_class=num_class, in_channels=224, net_type='Basic',
                           rotation_mode='ortho6d', norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
                           act_cfg=dict(type='ReLU')),
        corr_lookup_cfg=dict(align_corners=True), gru_type='SeqConv', act_cfg=dict(type='ReLU'))


def encoder_cfg(norm='IN'):
    return dict(in_channels=3, out_channels=256, net_type='Basic', norm_cfg=dict(type=norm))
