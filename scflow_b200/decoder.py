"""SCFlow decoder and its RAFT building blocks, registered under the reference's names with the reference's
constructor kwargs and state-dict keys (models/decoder/scflow_decoder.py, models/decoder/raft_decoder.py:19-294).

``SCFlowDecoder.forward`` is ONE call into the C ABI (``scf_decoder_forward``): the C++ side enqueues the pyramid
build and every kernel of every refinement iteration on the current CUDA stream; Python only allocates the
outputs / workspace and (optionally) replays the whole thing as a CUDA graph.
"""
import ctypes as C
import os
from typing import Optional, Sequence, Union

import torch
import torch.nn as nn

from . import _lib, ops
from .builder import DECODERS, build_head
from .cnn import BaseModule, ConvModule, PackedCache
from .corr_lookup import CorrLookup


class CorrelationPyramid(BaseModule):
    """All-pairs correlation volume + average-pooled pyramid (raft_decoder.py:19-58)."""

    def __init__(self, num_levels: int = 4, precision: int = ops.PRECISION_FP32) -> None:
        super().__init__()
        self.num_levels = num_levels
        self.precision = precision

    def forward(self, feat1: torch.Tensor, feat2: torch.Tensor) -> Sequence[torch.Tensor]:
        return ops.corr_build(feat1.contiguous(), feat2.contiguous(), self.num_levels, self.precision)


class MotionEncoder(BaseModule):
    """raft_decoder.py:61-166; only the channel tables of the reference are reproduced."""
    _corr_channels = {'Basic': (256, 192), 'Small': 96, 'Large': (256, 192)}
    _corr_kernel = {'Basic': (1, 3), 'Small': 1, 'Large': (1, 3)}
    _corr_padding = {'Basic': (0, 1), 'Small': 0, 'Large': (0, 1)}
    _flow_channels = {'Basic': (128, 64), 'Small': (64, 32), 'Large': (128, 64)}
    _flow_kernel = {'Basic': (7, 3), 'Small': (7, 3), 'Large': (7, 3)}
    _flow_padding = {'Basic': (3, 1), 'Small': (3, 1), 'Large': (3, 1)}
    _out_channels = {'Basic': 126, 'Small': 80, 'Large': 126}
    _out_kernel = {'Basic': 3, 'Small': 3, 'Large': 3}
    _out_padding = {'Basic': 1, 'Small': 1, 'Large': 1}

    def __init__(self, num_levels: int = 4, radius: int = 4, net_type: str = 'Basic', **kwargs) -> None:
        super().__init__()
        assert net_type in ['Basic', 'Small', 'Large']

        def as_list(v):
            return list(v) if isinstance(v, (tuple, list)) else [v]

        corr_channels = as_list(self._corr_channels[net_type])
        self.out_channels = as_list(self._out_channels[net_type])
        corr_inch = num_levels * (2 * radius + 1) ** 2
        self.corr_net = nn.Sequential(*self._make_encoder(corr_inch, corr_channels, as_list(self._corr_kernel[net_type]),
                                                          as_list(self._corr_padding[net_type]), **kwargs))
        flow_channels = as_list(self._flow_channels[net_type])
        self.flow_net = nn.Sequential(*self._make_encoder(2, flow_channels, as_list(self._flow_kernel[net_type]),
                                                          as_list(self._flow_padding[net_type]), **kwargs))
        self.out_net = nn.Sequential(*self._make_encoder(corr_channels[-1] + flow_channels[-1], self.out_channels,
                                                         as_list(self._out_kernel[net_type]),
                                                         as_list(self._out_padding[net_type]), **kwargs))

    @staticmethod
    def _make_encoder(in_channel, channels, kernels, paddings, conv_cfg=None, norm_cfg=None, act_cfg=None):
        layers = []
        for ch, k, p in zip(channels, kernels, paddings):
            layers.append(ConvModule(in_channel, ch, k, padding=p, conv_cfg=conv_cfg, norm_cfg=norm_cfg, act_cfg=act_cfg))
            in_channel = ch
        return layers

    def forward(self, corr: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
        """corr [B,324,H,W], flow [B,2,H,W] -> [B,128,H,W] (standalone NCHW form of raft_decoder.py:152-166)."""
        b, _, h, w = flow.shape
        x = ops.nchw_to_nhwc(corr.contiguous())
        for layer in self.corr_net:
            x = layer.forward_nhwc([(x, 0, layer.in_channels)])
        fl = ops.nchw_to_nhwc(flow.contiguous())
        f = fl
        for layer in self.flow_net:
            f = layer.forward_nhwc([(f, 0, layer.in_channels)])
        cout = self.out_channels[0]
        out = torch.empty(b, h, w, cout + 2, device=flow.device, dtype=torch.float32)
        self.out_net[0].forward_nhwc([(x, 0, x.shape[-1]), (f, 0, f.shape[-1])], out=out, out_coff=0)
        out[..., cout:] = fl
        return ops.nhwc_to_nchw(out)


class ConvGRU(BaseModule):
    """raft_decoder.py:168-253. z and r share one N=2*Ch convolution whose epilogue also forms r*h; the q
    convolution's epilogue applies the state update."""
    _kernel = {'Conv': 3, 'SeqConv': ((1, 5), (5, 1))}
    _padding = {'Conv': 1, 'SeqConv': ((0, 2), (2, 0))}

    def __init__(self, h_channels: int, x_channels: int, net_type: str = 'SeqConv') -> None:
        super().__init__()
        assert net_type in ['Conv', 'SeqConv']
        kernel_size = self._kernel[net_type] if isinstance(self._kernel[net_type], (tuple, list)) else [self._kernel[net_type]]
        padding = self._padding[net_type] if isinstance(self._padding[net_type], (tuple, list)) else [self._padding[net_type]]
        self.h_channels, self.x_channels = h_channels, x_channels
        conv_z, conv_r, conv_q = [], [], []
        for k, p in zip(kernel_size, padding):
            conv_z.append(ConvModule(h_channels + x_channels, h_channels, k, padding=p, act_cfg=dict(type='Sigmoid')))
            conv_r.append(ConvModule(h_channels + x_channels, h_channels, k, padding=p, act_cfg=dict(type='Sigmoid')))
            conv_q.append(ConvModule(h_channels + x_channels, h_channels, k, padding=p, act_cfg=dict(type='Tanh')))
        self.conv_z = nn.ModuleList(conv_z)
        self.conv_r = nn.ModuleList(conv_r)
        self.conv_q = nn.ModuleList(conv_q)
        self._zr = [PackedCache() for _ in kernel_size]

    def init_weights(self) -> None:
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.orthogonal_(m.weight)

    def forward(self, h: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        ch, cx = self.h_channels, self.x_channels
        hh = ops.nchw_to_nhwc(h.contiguous())
        xh = ops.nchw_to_nhwc(x.contiguous())
        for i, (cz, cr, cq) in enumerate(zip(self.conv_z, self.conv_r, self.conv_q)):
            wz, wr = cz.conv.weight, cr.conv.weight
            zr_w = self._zr[i].get([wz, wr, cz.conv.bias, cr.conv.bias], lambda: (
                ops.pack_conv_weight([wz.detach(), wr.detach()]), torch.cat([cz.conv.bias.detach(), cr.conv.bias.detach()])))
            z = torch.empty_like(hh)
            rh = torch.empty_like(hh)
            ops.conv2d_nhwc([(hh, 0, ch), (xh, 0, cx)], zr_w[0], zr_w[1], 2 * ch, cz.kernel_size, 1, cz.padding, act='sigmoid',
                            out=z, epi=_lib.EPI_GRU_ZR, aux0=hh, out2=rh)
            hn = torch.empty_like(hh)
            ops.conv2d_nhwc([(rh, 0, ch), (xh, 0, cx)], cq.packed_weight(), cq.conv.bias.detach(), ch, cq.kernel_size, 1,
                            cq.padding, act='tanh', out=hn, epi=_lib.EPI_GRU_Q, aux0=hh, aux1=z)
            hh = hn
        return ops.nhwc_to_nchw(hh)


class XHead(BaseModule):
    """Flow / mask prediction head (raft_decoder.py:256-294)."""

    def __init__(self, in_channels: int, feat_channels: Sequence[int], x_channels: int, x: str) -> None:
        super().__init__()
        layers = []
        for ch in feat_channels:
            layers.append(ConvModule(in_channels, ch, 3, padding=1))
            in_channels = ch
        self.layers = nn.Sequential(*layers)
        if x in ('flow', 'tradeoff'):
            self.predict_layer = nn.Conv2d(feat_channels[-1], x_channels, kernel_size=3, padding=1)
        elif x == 'mask':
            self.predict_layer = nn.Conv2d(feat_channels[-1], x_channels, kernel_size=1, padding=0)
        else:
            raise ValueError(f'x must be \'flow\' or \'mask\', but got {x}')
        self._pred = PackedCache()

    def forward(self, x: torch.Tensor, scale: float = 1.0, act: str = 'none') -> torch.Tensor:
        """``scale`` multiplies the prediction inside the convolution's epilogue (RAFTDecoder's ``.25 * mask_pred(h)``; for a
        power of two ``scale * acc + scale * bias`` equals ``scale * (acc + bias)`` exactly); ``act`` is applied there too
        (RAFTDecoderMask's ``sigmoid(occlusion_pred(h))``)."""
        y = ops.nchw_to_nhwc(x.contiguous())
        for layer in self.layers:
            y = layer.forward_nhwc([(y, 0, layer.in_channels)])
        p = self.predict_layer
        w = self._pred.get([p.weight], lambda: ops.pack_conv_weight([p.weight.detach()]))
        bias = p.bias.detach() if scale == 1.0 else p.bias.detach() * scale
        out = ops.conv2d_nhwc([(y, 0, y.shape[-1])], w, bias, p.out_channels, p.kernel_size, 1, p.padding, scale=scale, act=act)
        return ops.nhwc_to_nchw(out)


@DECODERS.register_module()
class RAFTDecoder(BaseModule):
    """The RAFT baseline decoder (models/decoder/raft_decoder.py:296-457): same constructor kwargs,
    parameter names and return value (the list of x8 up-sampled flow predictions).  Assembled from the same native
    operators as ``SCFlowDecoder`` - correlation pyramid, pyramid lookup, motion encoder, SepConvGRU, flow head - plus the
    mask head's 576-channel prediction and the convex up-sampling kernel; runs module by module (exact-fp32 convolutions),
    not as the fused loop, because it is a baseline next to the path rather than the path."""
    _h_channels = {'Basic': 128, 'Small': 96}
    _cxt_channels = {'Basic': 128, 'Small': 64}

    def __init__(self, net_type: str, num_levels: int, radius: int, iters: int,
                 corr_lookup_cfg: dict = dict(type='CorrLookup', align_corners=True), gru_type: str = 'SeqConv',
                 feat_channels: Union[int, Sequence[int]] = 256, mask_channels: int = 64, convex_unsample_flow: bool = True,
                 conv_cfg: Optional[dict] = None, norm_cfg: Optional[dict] = None, act_cfg: Optional[dict] = None) -> None:
        super().__init__()
        assert net_type in ['Basic', 'Small']
        assert type(feat_channels) in (int, tuple, list)
        if net_type != 'Basic' or gru_type != 'SeqConv' or num_levels != 4 or radius != 4:
            raise NotImplementedError('RAFTDecoder: net_type="Basic", gru_type="SeqConv", num_levels=4, radius=4 (the shipped '
                                      'configuration: x8 up-sampling over a 3x3 neighbourhood) is implemented')
        self.corr_block = CorrelationPyramid(num_levels=num_levels)
        feat_channels = [feat_channels]         # the reference's `isinstance(tuple, list)` quirk (raft_decoder.py:343-344)
        self.net_type, self.num_levels, self.radius = net_type, num_levels, radius
        self.h_channels = self._h_channels[net_type]
        self.cxt_channels = self._cxt_channels[net_type]
        self.iters = iters
        self.mask_channels = mask_channels * (2 * radius + 1)
        corr_lookup_cfg = {k: v for k, v in corr_lookup_cfg.items() if k != 'type'}
        corr_lookup_cfg['radius'] = radius
        self.corr_lookup = CorrLookup(**corr_lookup_cfg)
        self.encoder = MotionEncoder(num_levels=num_levels, radius=radius, net_type=net_type, conv_cfg=conv_cfg,
                                     norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.gru_type = gru_type
        self.gru = ConvGRU(self.h_channels, self.encoder.out_channels[0] + 2 + self.cxt_channels, net_type=gru_type)
        self.flow_pred = XHead(self.h_channels, feat_channels, 2, x='flow')
        self.mask_pred = XHead(self.h_channels, feat_channels, self.mask_channels, x='mask')
        self.convex_upsample_flow = convex_unsample_flow

    def _upsample(self, flow: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """raft_decoder.py:381-416."""
        scale = 2 ** (self.num_levels - 1)
        if mask is None:
            return ops.resize_bilinear_nchw(flow.contiguous(), scale * flow.shape[2], scale * flow.shape[3], scale=float(scale))
        return ops.convex_upsample(flow.contiguous(), mask.contiguous())

    def forward(self, feat1: torch.Tensor, feat2: torch.Tensor, flow: torch.Tensor, h_feat: torch.Tensor,
                cxt_feat: torch.Tensor) -> Sequence[torch.Tensor]:
        """raft_decoder.py:418-457 (inference: the reference's ``flow.detach()`` is a no-op without autograd)."""
        if torch.is_grad_enabled() and any(t.requires_grad for t in (feat1, feat2, flow, h_feat, cxt_feat)):
            raise NotImplementedError('RAFTDecoder: forward only (no backward yet); call under torch.no_grad()')
        corr_pyramid = self.corr_block(feat1, feat2)
        upflow_preds = []
        b, _, h, w = flow.shape
        flow = flow.contiguous()
        for _ in range(self.iters):
            corr = self.corr_lookup(corr_pyramid, flow)
            motion_feat = self.encoder(corr, flow)
            h_feat = self.gru(h_feat, torch.cat([cxt_feat, motion_feat], dim=1))
            delta_flow = self.flow_pred(h_feat)
            flow = ops.resize_bilinear_nchw(delta_flow, h, w, scale=1.0, add=flow)        # flow + delta_flow (identity resize)
            mask = self.mask_pred(h_feat, scale=0.25) if self.convex_upsample_flow else None
            upflow_preds.append(self._upsample(flow, mask))
        return upflow_preds


@DECODERS.register_module()
class RAFTDecoderMask(RAFTDecoder):
    """models/decoder/raft_decoder_mask.py:21-208: RAFTDecoder plus an occlusion head whose sigmoid output is up-sampled with
    the same convex combination (without the x8 factor); returns ``(upflow_preds, upocclusion_preds)``."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        # same parameter order as the reference (flow_pred, occlusion_pred, mask_pred)
        mask_pred = self.mask_pred
        del self.mask_pred
        self.occlusion_pred = XHead(self.h_channels, [256], 1, x='mask')
        self.mask_pred = mask_pred

    def upsample_flow(self, flow: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._upsample(flow, mask)

    def upsample_mask(self, occlusion: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """raft_decoder_mask.py:143-162."""
        scale = 2 ** (self.num_levels - 1)
        if mask is None:
            return ops.resize_bilinear_nchw(occlusion.contiguous(), scale * occlusion.shape[2], scale * occlusion.shape[3])
        return ops.convex_upsample(occlusion.contiguous(), mask.contiguous(), mul=1.0)

    def forward(self, feat1: torch.Tensor, feat2: torch.Tensor, flow: torch.Tensor, h_feat: torch.Tensor, cxt_feat: torch.Tensor):
        """raft_decoder_mask.py:165-208."""
        if torch.is_grad_enabled() and any(t.requires_grad for t in (feat1, feat2, flow, h_feat, cxt_feat)):
            raise NotImplementedError('RAFTDecoderMask: forward only (no backward yet); call under torch.no_grad()')
        corr_pyramid = self.corr_block(feat1, feat2)
        upflow_preds, upocclusion_preds = [], []
        b, _, h, w = flow.shape
        flow = flow.contiguous()
        for _ in range(self.iters):
            corr = self.corr_lookup(corr_pyramid, flow)
            motion_feat = self.encoder(corr, flow)
            h_feat = self.gru(h_feat, torch.cat([cxt_feat, motion_feat], dim=1))
            delta_flow = self.flow_pred(h_feat)
            flow = ops.resize_bilinear_nchw(delta_flow, h, w, scale=1.0, add=flow)        # flow + delta_flow
            occlusion = self.occlusion_pred(h_feat, act='sigmoid')
            mask = self.mask_pred(h_feat, scale=0.25) if self.convex_upsample_flow else None
            upflow_preds.append(self.upsample_flow(flow, mask))
            upocclusion_preds.append(self.upsample_mask(occlusion, mask))
        return upflow_preds, upocclusion_preds


def _on_tensor_device(argname_index):
    """Runs the method under ``torch.cuda.device`` of its ``argname_index``-th positional tensor argument, so that kernels, TMA
    descriptors and the library's side streams are issued on the device that owns the memory (not merely the current one)."""
    def deco(fn):
        import functools

        @functools.wraps(fn)
        def wrapper(self, *args, **kwargs):
            t = args[argname_index] if len(args) > argname_index else None
            if isinstance(t, torch.Tensor) and t.is_cuda:
                with torch.cuda.device(t.device):
                    return fn(self, *args, **kwargs)
            return fn(self, *args, **kwargs)
        return wrapper
    return deco


@DECODERS.register_module()
class SCFlowDecoder(BaseModule):
    """Drop-in for the reference's ``SCFlowDecoder`` (scflow_decoder.py:18-251): same constructor kwargs, same
    parameter names, same call signature, same 7-list return value.

    Extra (non-reference) attributes:
        precision: ops.PRECISION_BF16X3 (default: tcgen05 split-bf16, fp32-accurate - the benchmarked path) or
            ops.PRECISION_FP32 (exact fp32 CUDA-core convolutions; comparison / debugging).
        use_cuda_graph (default True): once a call repeats the previous call's shapes the step is captured and later calls
            replay it (inference only); the returned tensors are then views of static buffers that the next call overwrites.
    """
    _h_channels = {'Basic': 128, 'Small': 96}
    _cxt_channels = {'Basic': 128, 'Small': 64}

    def __init__(self, net_type: str, num_levels: int, radius: int, iters: int, detach_flow: bool, detach_mask: bool,
                 detach_pose: bool, mask_flow: bool, mask_corr: bool, pose_head_cfg: dict, depth_transform: str = 'exp',
                 detach_depth_for_xy: bool = False, corr_lookup_cfg: dict = dict(align_corners=True),
                 gru_type: str = 'SeqConv', feat_channels: Union[int, Sequence[int]] = 256,
                 conv_cfg: Optional[dict] = None, norm_cfg: Optional[dict] = None, act_cfg: Optional[dict] = None,
                 precision: int = ops.PRECISION_BF16X3, use_cuda_graph: bool = True) -> None:
        super().__init__()
        assert net_type in ['Basic', 'Small']
        assert type(feat_channels) in (int, tuple, list)
        if net_type != 'Basic' or gru_type != 'SeqConv':
            raise NotImplementedError('scflow_b200 implements the shipped configuration: net_type="Basic", gru_type="SeqConv"')
        if depth_transform != 'exp':
            raise NotImplementedError('only depth_transform="exp" is implemented')
        if act_cfg is None or act_cfg.get('type') != 'ReLU' or norm_cfg is not None:
            raise NotImplementedError('the fused loop implements the shipped act_cfg=dict(type="ReLU"), norm_cfg=None')
        self.corr_block = CorrelationPyramid(num_levels=num_levels, precision=precision)
        # the reference's `isinstance(tuple, list)` bug makes feat_channels always [feat_channels] (scflow_decoder.py:73-74)
        feat_channels = [feat_channels]
        if feat_channels != [256]:
            raise NotImplementedError('feat_channels must be 256 (the only value the reference can produce)')
        self.net_type, self.num_levels, self.radius = net_type, num_levels, radius
        self.detach_flow, self.detach_mask, self.detach_pose = detach_flow, detach_mask, detach_pose
        self.detach_depth_for_xy = detach_depth_for_xy
        self.mask_flow, self.mask_corr = mask_flow, mask_corr
        self.depth_transform = depth_transform
        self.h_channels = self._h_channels[net_type]
        self.cxt_channels = self._cxt_channels[net_type]
        self.iters = iters
        corr_lookup_cfg = dict(corr_lookup_cfg)
        corr_lookup_cfg['radius'] = radius
        self.corr_lookup = CorrLookup(**corr_lookup_cfg)
        self.encoder = MotionEncoder(num_levels=num_levels, radius=radius, net_type=net_type, conv_cfg=conv_cfg,
                                     norm_cfg=norm_cfg, act_cfg=act_cfg)
        self.gru_type = gru_type
        self.gru = ConvGRU(self.h_channels, self.encoder.out_channels[0] + 2 + self.cxt_channels, net_type=gru_type)
        self.pose_pred = build_head(pose_head_cfg)
        gn_eps = [getattr(l, 'norm_eps', 1e-5) for l in getattr(self.pose_pred, 'conv_layers', [])]
        if any(abs(e - 1e-5) > 1e-12 for e in gn_eps):
            raise NotImplementedError(f'the fused loop folds GroupNorm eps = 1e-5; pose_head_cfg asks for {gn_eps}')
        self.flow_pred = XHead(self.h_channels, feat_channels, 2, x='flow')
        self.mask_pred = XHead(self.h_channels, feat_channels, 1, x='mask')
        self.delta_flow_encoder = nn.Sequential(*MotionEncoder._make_encoder(2, [128, 64], [7, 3], [3, 1], conv_cfg, norm_cfg, act_cfg))
        self.mask_encoder = nn.Sequential(*MotionEncoder._make_encoder(1, [64, 32], [3, 3], [1, 1], conv_cfg, norm_cfg, act_cfg))
        self.precision = precision
        self.use_cuda_graph = use_cuda_graph
        self.identity_pose_head = False   # config-3 mode: skip the regressor (it cannot run off 256x256)
        # sub-batches run concurrently on separate streams; measured on B200 (B=32): 2 -> +2 % step time, 4 -> +11 %, so 1
        self.batch_splits = int(os.environ.get('SCFLOW_DEC_SPLITS', '1'))
        self._streams = []
        self._arena = PackedCache()
        self._workspaces = {}
        self._graphs = {}
        self._last_key = None

    # ------------------------------------------------------------------ C-ABI plumbing
    def _cfg(self) -> _lib.DecoderCfg:
        head = self.pose_pred
        num_class = getattr(head, 'num_class', None)
        return _lib.DecoderCfg(self.num_levels, self.radius, num_class if num_class else 0, head.rotation_out_channels,
                               int(self.mask_flow), int(self.mask_corr), 0 if self.identity_pose_head else 1,
                               int(self.precision))

    def _packed_arena(self, cfg) -> torch.Tensor:
        sd = {k: v for k, v in self.named_parameters()}
        params = [sd[k] for k in _lib.DECODER_WEIGHT_KEYS]

        def make():
            lib = _lib.load()
            dev = params[0].device
            arena = torch.empty(lib.scf_decoder_packed_bytes(C.byref(cfg)), device=dev, dtype=torch.uint8)
            srcs = [p.detach().contiguous().float() for p in params]
            arr = (C.c_void_p * _lib.SCF_W_COUNT)(*[t.data_ptr() for t in srcs])
            _lib.check(lib.scf_decoder_pack(C.byref(cfg), arr, _lib.ptr(arena), _lib.stream_ptr()), 'scf_decoder_pack')
            return (arena, int(cfg.pose_head), int(cfg.precision))

        arena, ph, prec = self._arena.get(params, make)
        if ph != int(cfg.pose_head) or prec != int(cfg.precision):
            self._arena = PackedCache()
            arena, _, _ = self._arena.get(params, make)
        return arena

    def _workspace(self, cfg, b, h, w, device):
        """One workspace per concurrently running sub-batch."""
        n = self._splits(b)
        key = (b, n, h, w, int(cfg.precision), str(device))
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = _lib.load().scf_decoder_workspace_bytes(C.byref(cfg), b // n, h, w)
            if nbytes == 0:
                raise _lib.ScfError('scf_decoder_workspace_bytes rejected the configuration')
            # keep only the latest shape
            self._workspaces = {key: [torch.empty(nbytes, device=device, dtype=torch.uint8) for _ in range(n)]}
            ws = self._workspaces[key]
        return ws

    def _run_one(self, cfg, arena, ws, ins, outs, b, h, w, iters, invalid, lo=0, total=0):
        """One scf_decoder_forward call on the current stream for samples lo..lo+b-1 of the batch."""
        io = _lib.DecoderIO()
        io.native_inputs = 1 if ins.get('feat_render') is None else 0
        for k in ('feat_render', 'feat_real', 'h_feat', 'cxt_feat', 'ref_rotation', 'ref_translation', 'depth', 'internel_k',
                  'init_flow'):
            t = ins.get(k)
            if t is None:
                continue
            io_ptr = t.data_ptr() + lo * t.stride(0) * t.element_size()
            setattr(io, k, io_ptr)
        io.label = ins['label'].data_ptr()        # full batch: only label[0] is read (pose_head.py:201-211 quirk)
        io.invalid_flow_num = float(invalid)
        for k in ('flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation'):
            setattr(io, k, outs[k].data_ptr())
        io.h_out = None
        io.out_batch_total, io.out_batch_offset = total, lo
        _lib.check(_lib.load().scf_decoder_forward(C.byref(cfg), _lib.ptr(arena), C.byref(io), b, h, w, iters, _lib.ptr(ws),
                                                   ws.numel(), _lib.stream_ptr()), 'scf_decoder_forward')

    def _splits(self, b: int) -> int:
        """Number of sub-batches run concurrently on separate streams (``self.batch_splits``)."""
        n = max(1, int(self.batch_splits))
        return n if b % n == 0 else 1

    def _run(self, cfg, arena, ws, ins, outs, b, h, w, iters, invalid):
        n = self._splits(b)
        if n == 1:
            return self._run_one(cfg, arena, ws[0], ins, outs, b, h, w, iters, invalid)
        # Independent sub-batches on separate streams: every layer is its own kernel, so one chain spends a sizeable part of
        # its time in per-kernel heads and tails (pipeline fill, last-tile epilogue) with idle tensor cores; a second,
        # independent chain fills those gaps.  Fork/join through events, so the whole thing still captures into one graph.
        main = torch.cuda.current_stream(ins['depth'].device)
        if len(self._streams) < n - 1:
            self._streams = [torch.cuda.Stream(device=ins['depth'].device) for _ in range(n - 1)]
        bs = b // n
        for i in range(1, n):
            self._streams[i - 1].wait_stream(main)
        self._run_one(cfg, arena, ws[0], ins, outs, bs, h, w, iters, invalid, 0, b)
        for i in range(1, n):
            with torch.cuda.stream(self._streams[i - 1]):
                self._run_one(cfg, arena, ws[i], ins, outs, bs, h, w, iters, invalid, i * bs, b)
        for i in range(1, n):
            main.wait_stream(self._streams[i - 1])

    @staticmethod
    def _alloc_outputs(iters, b, h, w, rot_dim, device):
        f32 = dict(device=device, dtype=torch.float32)
        return dict(flow_from_pose=torch.empty(iters, b, 2, h, w, **f32), flow_from_pred=torch.empty(iters, b, 2, h, w, **f32),
                    rotation=torch.empty(iters, b, 3, 3, **f32), translation=torch.empty(iters, b, 3, **f32),
                    mask=torch.empty(iters, b, 1, h, w, **f32), delta_rotation=torch.empty(iters, b, rot_dim, **f32),
                    delta_translation=torch.empty(iters, b, 3, **f32))

    @_on_tensor_device(0)
    def forward(self, feat_render: torch.Tensor, feat_real: torch.Tensor, h_feat: torch.Tensor, cxt_feat: torch.Tensor,
                ref_rotation: torch.Tensor, ref_translation: torch.Tensor, depth: torch.Tensor, internel_k: torch.Tensor,
                label: torch.Tensor, init_flow: torch.Tensor, invalid_flow_num: float):
        """Same contract as scflow_decoder.py:150-251. Returns (flow_from_pose, flow_from_pred, rotation_preds,
        translation_preds, mask_preds, delta_rotation_preds, delta_translation_preds), each a list of ``iters`` tensors."""
        if torch.is_grad_enabled() and (feat_render.requires_grad or any(p.requires_grad for p in self.parameters())):
            if feat_render.requires_grad or self.training:
                # training: the differentiable graph (native forward for the gradient-free kernels, torch-composed backward)
                from . import training
                return training.decoder_forward_train(self, feat_render, feat_real, h_feat, cxt_feat, ref_rotation, ref_translation,
                                                      depth, internel_k, label, init_flow, invalid_flow_num)
        ins = dict(feat_render=feat_render, feat_real=feat_real, h_feat=h_feat, cxt_feat=cxt_feat, ref_rotation=ref_rotation,
                   ref_translation=ref_translation, depth=depth, internel_k=internel_k, init_flow=init_flow)
        for k, t in ins.items():
            if not t.is_cuda:
                raise RuntimeError(f'SCFlowDecoder: {k} must be a CUDA tensor (scflow_b200 has no CPU path)')
            ins[k] = t.detach().contiguous().float()
        if label is None:
            label = torch.zeros(depth.shape[0], dtype=torch.int64, device=depth.device)
        ins['label'] = label.detach().to(torch.int64).contiguous()
        b, h, w = depth.shape
        h8, w8 = h // 2 ** (self.num_levels - 1), w // 2 ** (self.num_levels - 1)
        if tuple(feat_render.shape) != (b, 256, h8, w8) or feat_real.shape != feat_render.shape:
            raise ValueError(f'feature maps must be [{b},256,{h8},{w8}], got {tuple(feat_render.shape)} / {tuple(feat_real.shape)}')
        if tuple(h_feat.shape) != (b, self.h_channels, h8, w8) or tuple(cxt_feat.shape) != (b, self.cxt_channels, h8, w8):
            raise ValueError('h_feat / cxt_feat have the wrong shape')
        if tuple(init_flow.shape) != (b, 2, h, w):
            raise ValueError(f'init_flow must be [{b},2,{h},{w}]')
        iters = int(self.iters)
        cfg = self._cfg()
        dev = depth.device
        arena = self._packed_arena(cfg)
        ws = self._workspace(cfg, b, h, w, dev)
        rot_dim = cfg.rot_dim
        if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
            outs = self._forward_graph(cfg, arena, ws, ins, b, h, w, iters, invalid_flow_num, rot_dim, dev)
        else:
            outs = self._alloc_outputs(iters, b, h, w, rot_dim, dev)
            self._run(cfg, arena, ws, ins, outs, b, h, w, iters, invalid_flow_num)
        order = ('flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation')
        return tuple([outs[k][i] for i in range(iters)] for k in order)

    # ------------------------------------------------------------------ encoder -> loop without the NCHW round trip
    def native_slots(self, b: int, h: int, w: int, device):
        """(cfg, workspace tensor, byte offsets of the feature / h split / h fp32 / context slots) for encoders that write the
        loop's inputs directly (scf_encoder_forward_ex); None when this configuration cannot take them."""
        if int(self.precision) != 1 or self._splits(b) != 1:
            return None
        cfg = self._cfg()
        ws = self._workspace(cfg, b, h, w, device)[0]
        slots = (C.c_size_t * 4)()
        _lib.check(_lib.load().scf_decoder_workspace_slots(C.byref(cfg), b, h, w, slots), 'scf_decoder_workspace_slots')
        return cfg, ws, [int(v) for v in slots]

    @_on_tensor_device(0)
    def forward_prepared(self, ref_rotation, ref_translation, depth, internel_k, label, init_flow, invalid_flow_num,
                         allow_graph: bool = True):
        """Same as ``forward`` when the feature maps, hidden state and context already sit in the workspace slots
        (``native_slots``), written there by the encoders in the loop's own layout."""
        ins = dict(ref_rotation=ref_rotation, ref_translation=ref_translation, depth=depth, internel_k=internel_k, init_flow=init_flow)
        for k, t in ins.items():
            if not t.is_cuda:
                raise RuntimeError(f'SCFlowDecoder: {k} must be a CUDA tensor (scflow_b200 has no CPU path)')
            ins[k] = t.detach().contiguous().float()
        if label is None:
            label = torch.zeros(depth.shape[0], dtype=torch.int64, device=depth.device)
        ins['label'] = label.detach().to(torch.int64).contiguous()
        b, h, w = depth.shape
        iters = int(self.iters)
        cfg = self._cfg()
        dev = depth.device
        arena = self._packed_arena(cfg)
        ws = self._workspace(cfg, b, h, w, dev)
        if self.use_cuda_graph and allow_graph and not torch.cuda.is_current_stream_capturing():
            outs = self._forward_graph(cfg, arena, ws, ins, b, h, w, iters, invalid_flow_num, cfg.rot_dim, dev)
        else:
            outs = self._alloc_outputs(iters, b, h, w, cfg.rot_dim, dev)
            self._run(cfg, arena, ws, ins, outs, b, h, w, iters, invalid_flow_num)
        order = ('flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation')
        return tuple([outs[k][i] for i in range(iters)] for k in order)

    def _forward_graph(self, cfg, arena, ws, ins, b, h, w, iters, invalid, rot_dim, dev):
        key = (b, h, w, iters, float(invalid), int(cfg.precision), int(cfg.pose_head), arena.data_ptr(), ws[0].data_ptr(), len(ws),
               'feat_render' in ins)
        entry = self._graphs.get(key)
        if entry is None:
            # Cache miss: THIS call's result comes from one eager run (which is also the warm-up that loads modules and sets
            # function attributes); the graph is then captured - capture enqueues nothing - and only later calls replay it.
            # The loop consumes its hidden-state slots in place, so with encoder-written ("native") inputs a second run on the
            # live slots would start from the final hidden state instead of the encoder output; never run the loop twice.
            # A shape is captured only when it repeats (varying test-time batch sizes would pay a capture per call otherwise).
            static_in = {k: v.clone() for k, v in ins.items()}
            static_out = self._alloc_outputs(iters, b, h, w, rot_dim, dev)
            self._run(cfg, arena, ws, static_in, static_out, b, h, w, iters, invalid)
            seen, self._last_key = self._last_key == key, key
            if not seen:
                return static_out
            torch.cuda.current_stream(dev).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._run(cfg, arena, ws, static_in, static_out, b, h, w, iters, invalid)
            self._graphs = {key: (graph, static_in, static_out)}
            return static_out
        graph, static_in, static_out = entry
        for k, v in ins.items():
            static_in[k].copy_(v, non_blocking=True)
        graph.replay()
        return static_out   # NOTE: overwritten by the next call with the same shapes (documented in INTEGRATION.md)
