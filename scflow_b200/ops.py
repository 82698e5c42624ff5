"""Tensor-level wrappers over the C ABI (one function per exported kernel family).

All functions require contiguous fp32 CUDA tensors and launch on torch's current stream. They never fall back
to PyTorch arithmetic: bad devices/dtypes raise, library errors raise ``ScfError``.
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import ConvDesc, ConvSeg, TcConvDesc, TcSeg, check, ptr, stream_ptr

PRECISION_FP32 = 0       # exact-fp32 CUDA-core convolutions / build
PRECISION_BF16X3 = 1     # tcgen05 split-bf16 (3 MMAs, fp32 accumulate in TMEM)


def _req(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name} must be a torch.Tensor')
    if not t.is_cuda:
        raise RuntimeError(f'{name} must live on a CUDA device: scflow_b200 has no CPU path')
    if t.dtype != dtype:
        raise TypeError(f'{name} must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f'{name} lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: call '
                           f'scflow_b200.ops under torch.cuda.device({t.device.index}) (the modules do this themselves)')
    return t


def nchw_to_nhwc(x: torch.Tensor, out: Optional[torch.Tensor] = None, coff: int = 0) -> torch.Tensor:
    _req(x, 'x')
    b, c, h, w = x.shape
    if out is None:
        out = torch.empty(b, h, w, c, device=x.device, dtype=torch.float32)
    check(_lib.load().scf_nchw_to_nhwc(ptr(x), ptr(out), b, c, h, w, out.shape[-1], coff, stream_ptr()), 'scf_nchw_to_nhwc')
    return out


def nhwc_to_nchw(x: torch.Tensor, channels: Optional[int] = None, coff: int = 0) -> torch.Tensor:
    _req(x, 'x')
    b, h, w, stride = x.shape
    c = channels if channels is not None else stride - coff
    out = torch.empty(b, c, h, w, device=x.device, dtype=torch.float32)
    check(_lib.load().scf_nhwc_to_nchw(ptr(x), stride, coff, ptr(out), b, c, h, w, stream_ptr()), 'scf_nhwc_to_nchw')
    return out


def pack_conv_weight(weights: Sequence[torch.Tensor]) -> torch.Tensor:
    """OIHW conv weights (one or several convs with identical I,kh,kw, concatenated along O) -> packed [K][ldw]."""
    o_total = sum(int(w.shape[0]) for w in weights)
    _, i, kh, kw = weights[0].shape
    ldw = (o_total + 3) // 4 * 4
    packed = torch.zeros(kh * kw * i, ldw, device=weights[0].device, dtype=torch.float32)
    off = 0
    for w in weights:
        _req(w, 'weight')
        assert tuple(w.shape[1:]) == (i, kh, kw)
        check(_lib.load().scf_pack_conv_weight(ptr(w), ptr(packed), w.shape[0], i, kh, kw, ldw, off, stream_ptr()),
              'scf_pack_conv_weight')
        off += int(w.shape[0])
    return packed


def conv2d_nhwc(segs: Sequence, packed_w: torch.Tensor, bias: Optional[torch.Tensor], cout: int, kernel, stride=1,
                padding=None, act: str = 'none', out: Optional[torch.Tensor] = None, out_coff: int = 0,
                epi: int = _lib.EPI_ACT, aux0=None, aux1=None, out2=None, scale: float = 1.0,
                w_batch_stride: int = 0) -> torch.Tensor:
    """Generic convolution on NHWC buffers. ``segs`` = [(tensor[B,H,W,stride], coff, nch), ...] concatenated on C."""
    kh, kw = (kernel, kernel) if isinstance(kernel, int) else kernel
    ph, pw = (kh // 2, kw // 2) if padding is None else ((padding, padding) if isinstance(padding, int) else padding)
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    t0 = segs[0][0]
    b, hi, wi, _ = t0.shape
    ho = (hi + 2 * ph - kh) // sh + 1
    wo = (wi + 2 * pw - kw) // sw + 1
    if out is None:
        out = torch.empty(b, ho, wo, cout if epi != _lib.EPI_GRU_ZR else cout // 2, device=t0.device, dtype=torch.float32)
    d = ConvDesc()
    for n, (t, coff, nch) in enumerate(segs):
        _req(t, f'seg{n}')
        d.seg[n] = ConvSeg(t.data_ptr(), t.shape[-1], coff, nch)
    d.nseg = len(segs)
    d.B, d.Hi, d.Wi, d.Ho, d.Wo = b, hi, wi, ho, wo
    d.kh, d.kw, d.sh, d.sw, d.ph, d.pw = kh, kw, sh, sw, ph, pw
    d.w = packed_w.data_ptr()
    d.w_batch_stride = w_batch_stride
    d.ldw = packed_w.shape[-1]
    d.cout = cout
    d.bias = None if bias is None else bias.data_ptr()
    d.scale = scale
    d.epi, d.act = epi, _lib.ACT[act]
    d.out, d.out_stride, d.out_coff = out.data_ptr(), out.shape[-1], out_coff
    if aux0 is not None:
        d.aux0, d.aux0_stride = aux0.data_ptr(), aux0.shape[-1]
    if aux1 is not None:
        d.aux1, d.aux1_stride = aux1.data_ptr(), aux1.shape[-1]
    if out2 is not None:
        d.out2, d.out2_stride = out2.data_ptr(), out2.shape[-1]
    check(_lib.load().scf_conv2d(C.byref(d), stream_ptr()), 'scf_conv2d')
    return out


def level_shapes(h8: int, w8: int, num_levels: int):
    shapes = []
    for _ in range(num_levels):
        shapes.append((h8, w8))
        h8, w8 = h8 // 2, w8 // 2
    return shapes


def corr_build(feat_render: torch.Tensor, feat_real: torch.Tensor, num_levels: int = 4,
               precision: int = PRECISION_FP32) -> List[torch.Tensor]:
    """CorrelationPyramid.forward: returns the reference's list of [B*P, 1, Hl, Wl] tensors."""
    _req(feat_render, 'feat_render')
    _req(feat_real, 'feat_real')
    if feat_render.shape != feat_real.shape:
        raise ValueError('feature maps must have the same shape')
    b, c, h8, w8 = feat_render.shape
    lib = _lib.load()
    levels = [torch.empty(b * h8 * w8, 1, hl, wl, device=feat_render.device, dtype=torch.float32)
              for hl, wl in level_shapes(h8, w8, num_levels)]
    scratch = torch.empty(lib.scf_corr_build_scratch_bytes(b, c, h8, w8), device=feat_render.device, dtype=torch.uint8)
    arr = (C.c_void_p * num_levels)(*[t.data_ptr() for t in levels])
    check(lib.scf_corr_build(ptr(feat_render), ptr(feat_real), b, c, h8, w8, num_levels, arr, ptr(scratch), precision,
                             stream_ptr()), 'scf_corr_build')
    return levels


def corr_lookup_nhwc(levels: Sequence[torch.Tensor], flow8_nhwc: torch.Tensor, radius: int,
                     mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out_coff: int = 0) -> torch.Tensor:
    _req(flow8_nhwc, 'flow8')
    b, h8, w8, two = flow8_nhwc.shape
    assert two == 2
    n = len(levels)
    ch = n * (2 * radius + 1) ** 2
    for l, (t, (hl, wl)) in enumerate(zip(levels, level_shapes(h8, w8, n))):
        _req(t, f'level{l}')
        if t.numel() != b * h8 * w8 * hl * wl:
            raise ValueError(f'pyramid level {l} has {t.numel()} elements, expected {b * h8 * w8 * hl * wl}')
    if out is None:
        out = torch.empty(b, h8, w8, (ch + 3) // 4 * 4, device=flow8_nhwc.device, dtype=torch.float32)
    arr = (C.c_void_p * n)(*[t.data_ptr() for t in levels])
    check(_lib.load().scf_corr_lookup(arr, n, radius, ptr(flow8_nhwc), ptr(mask), ptr(out), out.shape[-1], out_coff, b, h8, w8,
                                      stream_ptr()), 'scf_corr_lookup')
    return out


def corr_lookup_taps(flow8_nhwc: torch.Tensor, level: int, radius: int):
    _req(flow8_nhwc, 'flow8')
    b, h8, w8, _ = flow8_nhwc.shape
    k = 2 * radius + 1
    x0 = torch.empty(b, h8, w8, k, device=flow8_nhwc.device, dtype=torch.int32)
    y0 = torch.empty_like(x0)
    check(_lib.load().scf_corr_lookup_taps(level, radius, ptr(flow8_nhwc), ptr(x0), ptr(y0), b, h8, w8, stream_ptr()),
          'scf_corr_lookup_taps')
    return x0, y0


def group_norm_relu_(x_nhwc: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, num_groups: int, eps: float = 1e-5):
    _req(x_nhwc, 'x')
    b, h, w, c = x_nhwc.shape
    check(_lib.load().scf_group_norm_relu(ptr(x_nhwc), ptr(_req(gamma, 'gamma')), ptr(_req(beta, 'beta')), b, h * w, c,
                                          num_groups, eps, stream_ptr()), 'scf_group_norm_relu')
    return x_nhwc


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: str = 'none') -> torch.Tensor:
    _req(x, 'x')
    _req(w, 'w')
    bsz, i = x.shape
    o = w.shape[0]
    y = torch.empty(bsz, o, device=x.device, dtype=torch.float32)
    check(_lib.load().scf_linear(ptr(x), ptr(w), ptr(bias), ptr(y), bsz, i, o, _lib.ACT[act], stream_ptr()), 'scf_linear')
    return y


def linear_tc(x: torch.Tensor, w: torch.Tensor, kr: int = 256, x_bias: Optional[torch.Tensor] = None, x_relu: bool = False,
              reduce: bool = False, bias: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """Split-K tcgen05 FC layer (scf_linear_tc): returns the raw partial sums [I / kr, B, O]; ``x`` is [B, I] or a stack of partial
    maps [S, B, I] of the previous layer (summed, + ``x_bias``, ReLU if ``x_relu``, while the operand is formed).
    ``reduce=True`` (I / kr must be 8): the K-range blocks reduce through distributed shared memory and the call returns the
    finished layer output act(W x + bias) [B, O]."""
    _req(x, 'x')
    nsplit = 1 if x.dim() == 2 else x.shape[0]
    bsz, i = x.shape[-2:]
    if w.dtype == torch.float32:                       # [O, I] fp32 -> split-bf16 [2, 1, O, I]
        _req(w, 'w')
        o = w.shape[0]
        packed = pack_conv_weight_tc([w.reshape(o, i, 1, 1).contiguous()], cin_pad=i)
    else:
        packed = _req(w, 'w', torch.bfloat16)
        o = w.shape[-2]
    part = None if reduce else torch.empty(i // kr, bsz, o, device=x.device, dtype=torch.float32)
    y = torch.empty(bsz, o, device=x.device, dtype=torch.float32) if reduce else None
    check(_lib.load().scf_linear_tc(ptr(x), nsplit, bsz * i, ptr(x_bias), int(x_relu), ptr(packed), ptr(part), ptr(y), ptr(bias), int(relu),
                                    bsz, i, o, kr, stream_ptr()), 'scf_linear_tc')
    return y if reduce else part


def pose_project(x, rot_w, rot_b, tr_w, tr_b, label, rot_dim: int, num_class: int):
    _req(x, 'x')
    bsz, i = x.shape
    d_rot = torch.empty(bsz, rot_dim, device=x.device, dtype=torch.float32)
    d_trs = torch.empty(bsz, 3, device=x.device, dtype=torch.float32)
    if num_class > 0:
        _req(label, 'label', torch.int64)
    check(_lib.load().scf_pose_project(ptr(x), ptr(rot_w), ptr(rot_b), ptr(tr_w), ptr(tr_b), ptr(label), ptr(d_rot), ptr(d_trs),
                                       bsz, i, rot_dim, num_class, stream_ptr()), 'scf_pose_project')
    return d_rot, d_trs


def pose_update(d_rot, d_trs, rot, trs):
    for n, t in (('d_rot', d_rot), ('d_trs', d_trs), ('rot', rot), ('trs', trs)):
        _req(t, n)
    b = rot.shape[0]
    rot_out, trs_out = torch.empty_like(rot), torch.empty_like(trs)
    check(_lib.load().scf_pose_update(ptr(d_rot), ptr(d_trs), ptr(rot), ptr(trs), ptr(rot_out), ptr(trs_out), b, stream_ptr()),
          'scf_pose_update')
    return rot_out, trs_out


def unproject(depth, k, rot, trs):
    for n, t in (('depth', depth), ('k', k), ('rot', rot), ('trs', trs)):
        _req(t, n)
    b, h, w = depth.shape
    pts4 = torch.empty(b, h, w, 4, device=depth.device, dtype=torch.float32)
    check(_lib.load().scf_unproject(ptr(depth), ptr(k), ptr(rot), ptr(trs), ptr(pts4), b, h, w, stream_ptr()), 'scf_unproject')
    return pts4


def reproject(pts4, k, rot, trs, invalid: float = 0.0):
    for n, t in (('pts4', pts4), ('k', k), ('rot', rot), ('trs', trs)):
        _req(t, n)
    b, h, w, _ = pts4.shape
    flow = torch.empty(b, 2, h, w, device=pts4.device, dtype=torch.float32)
    check(_lib.load().scf_reproject(ptr(pts4), ptr(k), ptr(rot), ptr(trs), float(invalid), ptr(flow), b, h, w, stream_ptr()),
          'scf_reproject')
    return flow


def reproject_down(pts4, k, rot, trs, factor: int = 8, invalid: float = 0.0):
    """``reproject`` plus, from the same launch, the next iteration's coarse flow (scflow_decoder.py:196-197):
    ``1/factor * F.interpolate(flow, 1/factor, bilinear, align_corners=True)`` as NHWC [B, H/factor, W/factor, 2]."""
    for n, t in (('pts4', pts4), ('k', k), ('rot', rot), ('trs', trs)):
        _req(t, n)
    b, h, w, _ = pts4.shape
    flow = torch.empty(b, 2, h, w, device=pts4.device, dtype=torch.float32)
    flow8 = torch.empty(b, h // factor, w // factor, 2, device=pts4.device, dtype=torch.float32)
    check(_lib.load().scf_reproject_down(ptr(pts4), ptr(k), ptr(rot), ptr(trs), float(invalid), ptr(flow), b, h, w, ptr(flow8),
                                         h // factor, w // factor, stream_ptr()), 'scf_reproject_down')
    return flow, flow8


def resize_bilinear_nchw(x: torch.Tensor, out_h: int, out_w: int, scale: float = 1.0, add: Optional[torch.Tensor] = None):
    """F.interpolate(mode='bilinear', align_corners=True) on NCHW, times ``scale``."""
    _req(x, 'x')
    b, c, h, w = x.shape
    out = torch.empty(b, c, out_h, out_w, device=x.device, dtype=torch.float32)
    check(_lib.load().scf_resize_bilinear(ptr(x), ptr(add), c * h * w, h * w, w, 1, h, w, ptr(out), c * out_h * out_w,
                                          out_h * out_w, out_w, 1, out_h, out_w, b, c, scale, stream_ptr()),
          'scf_resize_bilinear')
    return out


# ----------------------------------------------------------------------------------------------------------------------
# tcgen05 split-bf16 path.  A "split tensor" is a bf16 tensor [2, B, H, W, C]: plane 0 = hi, plane 1 = lo.
# ----------------------------------------------------------------------------------------------------------------------
def _req_split(t: torch.Tensor, name: str):
    _req(t, name, torch.bfloat16)
    if t.dim() != 5 or t.shape[0] != 2:
        raise ValueError(f'{name} must be a split-bf16 tensor [2,B,H,W,C]')
    return t


def split_nchw(x: torch.Tensor, out: Optional[torch.Tensor] = None, coff: int = 0, stride: Optional[int] = None,
               want_f32: bool = False):
    """NCHW fp32 -> split-bf16 NHWC planes [2,B,H,W,stride] (and optionally an fp32 NHWC copy)."""
    _req(x, 'x')
    b, c, h, w = x.shape
    if out is None:
        out = torch.zeros(2, b, h, w, stride or c, device=x.device, dtype=torch.bfloat16)
    f32 = torch.empty(b, h, w, c, device=x.device, dtype=torch.float32) if want_f32 else None
    check(_lib.load().scf_nchw_to_nhwc_split(ptr(x), ptr(out), out[0].numel(), out.shape[-1], coff, ptr(f32), c, b, c, h, w,
                                             stream_ptr()), 'scf_nchw_to_nhwc_split')
    return (out, f32) if want_f32 else out


def unsplit(t: torch.Tensor) -> torch.Tensor:
    """Debug helper: split-bf16 [2,B,H,W,C] -> fp32 NCHW (hi + lo)."""
    return (t[0].float() + t[1].float()).permute(0, 3, 1, 2).contiguous()


def pack_conv_weight_tc(weights: Sequence[torch.Tensor], cin_pad: Optional[int] = None, cout_pad: Optional[int] = None) -> torch.Tensor:
    """OIHW fp32 weights (merged along O) -> bf16 [2, taps, cout_pad, cin_pad]."""
    o_total = sum(int(w.shape[0]) for w in weights)
    _, i, kh, kw = weights[0].shape
    cin_pad = cin_pad or (i + 7) // 8 * 8
    cout_pad = cout_pad or (o_total + 15) // 16 * 16
    packed = torch.zeros(2, kh * kw, cout_pad, cin_pad, device=weights[0].device, dtype=torch.bfloat16)
    off = 0
    for w in weights:
        _req(w, 'weight')
        check(_lib.load().scf_pack_conv_weight_tc(ptr(w), ptr(packed), w.shape[0], i, kh, kw, cin_pad, cout_pad, off, stream_ptr()),
              'scf_pack_conv_weight_tc')
        off += int(w.shape[0])
    return packed


def conv2d_tc(segs: Sequence, packed_w: torch.Tensor, bias: Optional[torch.Tensor], cout: int, kernel, act: str = 'none',
              out_f32: Optional[torch.Tensor] = None, out_f32_coff: int = 0, out_hl: Optional[torch.Tensor] = None,
              out_hl_coff: int = 0, epi: int = _lib.EPI_ACT, aux0=None, aux1=None, out2_hl=None, scale: float = 1.0,
              w_batched: bool = False, stride: int = 1, pre: Optional[torch.Tensor] = None, stride_xy=None,
              stats: Optional[torch.Tensor] = None, out_pad_writable: bool = False, aux0_hl: Optional[torch.Tensor] = None):
    """tcgen05 convolution. ``segs`` = [(split tensor [2,B,H,W,stride], coff, nch), ...].
    ``pre``: fp32 NHWC map added before the GRU gate non-linearity; ``stride_xy``: per-axis strides; ``stats``: fp32
    [tiles*4*2*cout] buffer receiving per-tile InstanceNorm partial sums (see include/scflow_b200.h)."""
    kh, kw = (kernel, kernel) if isinstance(kernel, int) else kernel
    d = TcConvDesc()
    for n, (t, coff, nch) in enumerate(segs):
        _req_split(t, f'seg{n}')
        d.seg[n] = TcSeg(t.data_ptr(), t[0].numel(), t.shape[-1], coff, nch)
    d.nseg = len(segs)
    _, b, h, w, _ = segs[0][0].shape
    d.B, d.H, d.W, d.kh, d.kw, d.stride = b, h, w, kh, kw, stride
    _req(packed_w, 'packed_w', torch.bfloat16)
    d.w = packed_w.data_ptr()
    d.cin_pad, d.cout_pad, d.cout, d.w_batched = packed_w.shape[-1], packed_w.shape[-2], cout, int(w_batched)
    if bias is not None and bias.numel() % 16:      # the epilogue reads bias in aligned groups of 16
        bias = torch.nn.functional.pad(bias, (0, 16 - bias.numel() % 16))
    d.bias = None if bias is None else bias.data_ptr()
    d.scale, d.epi, d.act = scale, epi, _lib.ACT[act]
    if out_f32 is not None:
        _req(out_f32, 'out_f32')
        d.out_f32, d.out_f32_stride, d.out_f32_coff = out_f32.data_ptr(), out_f32.shape[-1], out_f32_coff
    if out_hl is not None:
        _req_split(out_hl, 'out_hl')
        d.out_hl, d.out_hl_plane, d.out_hl_stride, d.out_hl_coff = out_hl.data_ptr(), out_hl[0].numel(), out_hl.shape[-1], out_hl_coff
    if aux0 is not None:
        d.aux0, d.aux0_stride = aux0.data_ptr(), aux0.shape[-1]
    if aux1 is not None:
        d.aux1, d.aux1_stride = aux1.data_ptr(), aux1.shape[-1]
    if out2_hl is not None:
        _req_split(out2_hl, 'out2_hl')
        d.out2_hl, d.out2_hl_plane, d.out2_hl_stride = out2_hl.data_ptr(), out2_hl[0].numel(), out2_hl.shape[-1]
    if pre is not None:
        _req(pre, 'pre')
        d.pre, d.pre_stride = pre.data_ptr(), pre.shape[-1]
    if stride_xy is not None:
        d.stride_x, d.stride_y = stride_xy
    if stats is not None:
        _req(stats, 'stats')
        d.stats = stats.data_ptr()
    d.out_pad_writable = int(out_pad_writable)
    if aux0_hl is not None:                    # residual as split-bf16 planes [2, B, H, W, C] (rolling-rows kernel only)
        _req_split(aux0_hl, 'aux0_hl')
        d.aux0_hl, d.aux0_hl_plane, d.aux0_hl_stride = aux0_hl.data_ptr(), aux0_hl[0].numel(), aux0_hl.shape[-1]
    check(_lib.load().scf_conv2d_tc(C.byref(d), stream_ptr()), 'scf_conv2d_tc')
    return out_f32 if out_f32 is not None else out_hl


def conv2d_tc_tiles(b: int, hout: int, wout: int):
    """(number of 128-pixel tiles, tiles per sample or 0) of scf_conv2d_tc for this output geometry."""
    per = C.c_int(0)
    n = _lib.load().scf_conv2d_tc_tiles(b, hout, wout, C.byref(per))
    return n, per.value


def make_tc_gru_zr_bench(h, cxt, mot, wz, wr, bias, z, rh):
    """bench.py helper: a closure launching the GRU z|r tensor-core convolution (the dominant kernel) exactly as
    scf_decoder_forward issues it: input [h | motion] (K = taps*256), the loop-invariant context contribution + bias
    (computed here once, as the decoder does once per forward) added in the epilogue.
    h/cxt/mot: fp32 NHWC [B,H,W,128]; wz/wr: OIHW [128,384,kh,kw]; returns (launch, kernel name, mma passes, K)."""
    hs = split_nchw(h.permute(0, 3, 1, 2).contiguous())
    cs = split_nchw(cxt.permute(0, 3, 1, 2).contiguous())
    ms = split_nchw(mot.permute(0, 3, 1, 2).contiguous())
    kernel = (int(wz.shape[2]), int(wz.shape[3]))
    hm = [torch.cat([w[:, :128], w[:, 256:]], 1).contiguous() for w in (wz, wr)]
    cx = [w[:, 128:256].contiguous() for w in (wz, wr)]
    pw = pack_conv_weight_tc(hm)
    pre = torch.empty(*h.shape[:3], 256, device=h.device, dtype=torch.float32)
    conv2d_tc([(cs, 0, 128)], pack_conv_weight_tc(cx), bias, 256, kernel, act='none', out_f32=pre)
    rhs = torch.zeros(2, *rh.shape, device=rh.device, dtype=torch.bfloat16)

    def launch():
        conv2d_tc([(hs, 0, 128), (ms, 0, 128)], pw, None, 256, kernel, act='sigmoid', out_f32=z,
                  epi=_lib.EPI_GRU_ZR, aux0=h, out2_hl=rhs, pre=pre)
    k = kernel[0] * kernel[1] * 256
    return launch, f'conv_tc_kernel (GRU z|r 1x5, tcgen05 split-bf16, N=256, K={k}; context term hoisted out of the loop)', 3, k


def lookup_conv(levels: Sequence[torch.Tensor], flow8_nhwc: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CorrLookup + corr_net[0] in one kernel (scf_lookup_conv): ``weight`` OIHW [256, 324, 1, 1]; returns the split-bf16
    activation [2, B, H8, W8, 256] (hi, lo planes) of relu(conv1x1(lookup))."""
    _req(flow8_nhwc, 'flow8')
    _req(weight, 'weight')
    b, h8, w8, _ = flow8_nhwc.shape
    lib = _lib.load()
    packed = torch.empty(lib.scf_lookup_conv_packed_bytes(), device=weight.device, dtype=torch.uint8)
    check(lib.scf_lookup_conv_pack(ptr(weight), ptr(packed), stream_ptr()), 'scf_lookup_conv_pack')
    out = torch.empty(2, b, h8, w8, 256, device=weight.device, dtype=torch.bfloat16)
    arr = (C.c_void_p * len(levels))(*[_req(t, 'level').data_ptr() for t in levels])
    check(lib.scf_lookup_conv(arr, ptr(flow8_nhwc), ptr(mask), ptr(packed), ptr(bias), ptr(out), out[0].numel(), 256, b, h8, w8, stream_ptr()),
          'scf_lookup_conv')
    return out


class GruPassFused:
    """One SepConvGRU pass as ONE kernel (scf_gru_pass_fused; raft_decoder.py:245-253).  Built from the reference's weights
    ``conv_z/r/q.{pass}.conv.weight`` [128,384,kh,kw] and biases; the context columns' contribution (loop invariant) is
    evaluated once per context map with ``precompute`` and then every ``__call__(h, motion)`` is one launch.
    h / cxt / motion: fp32 NHWC [B,H,W,128]."""

    def __init__(self, wz, wr, wq, bz, br, bq):
        self.kernel = (int(wz.shape[2]), int(wz.shape[3]))
        if self.kernel not in ((1, 5), (5, 1)):
            raise ValueError('GruPassFused: SepConvGRU kernels are 1x5 or 5x1')
        self.vertical = int(self.kernel == (5, 1))
        hm = lambda w: torch.cat([w[:, :128], w[:, 256:]], 1).contiguous()      # [h | motion] columns
        self.w_zr = pack_conv_weight_tc([hm(wz), hm(wr)])
        self.w_q = pack_conv_weight_tc([hm(wq)])
        self.w_czr = pack_conv_weight_tc([wz[:, 128:256].contiguous(), wr[:, 128:256].contiguous()])
        self.w_cq = pack_conv_weight_tc([wq[:, 128:256].contiguous()])
        self.b_zr, self.b_q = torch.cat([bz, br]).contiguous(), bq.contiguous()
        self.pre_zr = self.pre_q = None

    def precompute(self, cxt: torch.Tensor):
        cs = split_nchw(cxt.permute(0, 3, 1, 2).contiguous())
        self.pre_zr = torch.empty(*cxt.shape[:3], 256, device=cxt.device, dtype=torch.float32)
        self.pre_q = torch.empty(*cxt.shape[:3], 128, device=cxt.device, dtype=torch.float32)
        conv2d_tc([(cs, 0, 128)], self.w_czr, self.b_zr, 256, self.kernel, act='none', out_f32=self.pre_zr)
        conv2d_tc([(cs, 0, 128)], self.w_cq, self.b_q, 128, self.kernel, act='none', out_f32=self.pre_q)

    def __call__(self, h: torch.Tensor, motion: torch.Tensor):
        """Returns (h' fp32 NHWC [B,H,W,128], h' split-bf16 [2,B,H,W,128])."""
        _req(h, 'h')
        _req(motion, 'motion')
        b, hh, ww, _ = h.shape
        hs = split_nchw(h.permute(0, 3, 1, 2).contiguous())
        ms = split_nchw(motion.permute(0, 3, 1, 2).contiguous())
        out = torch.empty_like(h)
        out_hl = torch.empty(2, b, hh, ww, 128, device=h.device, dtype=torch.bfloat16)
        d = _lib.GruPassDesc()
        d.h_hl, d.h_plane, d.h_f32 = hs.data_ptr(), hs[0].numel(), h.data_ptr()
        d.m_hl, d.m_plane = ms.data_ptr(), ms[0].numel()
        d.w_zr, d.w_q = self.w_zr.data_ptr(), self.w_q.data_ptr()
        d.pre_zr, d.pre_q = self.pre_zr.data_ptr(), self.pre_q.data_ptr()
        d.out_f32, d.out_hl, d.out_plane = out.data_ptr(), out_hl.data_ptr(), out_hl[0].numel()
        d.B, d.H, d.W, d.vertical = b, hh, ww, self.vertical
        self._keep = (hs, ms)
        check(_lib.load().scf_gru_pass_fused(C.byref(d), stream_ptr()), 'scf_gru_pass_fused')
        return out, out_hl


# ----------------------------------------------------------------------------------------------------------------------
# input formatting next to the path
# ----------------------------------------------------------------------------------------------------------------------
def format_rendered(images: torch.Tensor, zbuf: torch.Tensor, mean: Sequence[float], std: Sequence[float]):
    """Renderer output -> network input (base_refiner.py:96-107): ``images`` [B,H,W,C>=3] in [0,1], ``zbuf`` [B,H,W,K];
    ``mean`` / ``std`` are the three fp32 values the reference subtracts / divides by (dataset values already / 255).
    Returns (rendered_images [B,3,H,W], rendered_depths [B,H,W], rendered_masks [B,H,W])."""
    _req(images, 'images')
    _req(zbuf, 'zbuf')
    if images.dim() != 4 or zbuf.dim() != 4 or images.shape[:3] != zbuf.shape[:3] or images.shape[-1] < 3:
        raise ValueError(f'format_rendered: images [B,H,W,C>=3] and zbuf [B,H,W,K] expected, got {tuple(images.shape)} {tuple(zbuf.shape)}')
    b, h, w, cin = images.shape
    out = torch.empty(b, 3, h, w, device=images.device, dtype=torch.float32)
    depth = torch.empty(b, h, w, device=images.device, dtype=torch.float32)
    mask = torch.empty(b, h, w, device=images.device, dtype=torch.float32)
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    check(_lib.load().scf_format_rendered(ptr(images), cin, ptr(zbuf), zbuf.shape[-1], m3, s3, ptr(out), ptr(depth), ptr(mask), b, h, w,
                                          stream_ptr()), 'scf_format_rendered')
    return out, depth, mask


def convex_upsample(x: torch.Tensor, mask: torch.Tensor, mul: float = 8.0) -> torch.Tensor:
    """RAFTDecoder._upsample with a predicted mask (raft_decoder.py:381-416; ``mul`` = 8) or RAFTDecoderMask.upsample_mask
    (raft_decoder_mask.py:143-162; ``mul`` = 1): x [B,C<=2,H,W], mask [B,576,H,W] -> [B,C,8H,8W] (softmax over the 3x3
    neighbourhood)."""
    _req(x, 'x')
    _req(mask, 'mask')
    b, c, h, w = x.shape
    if c not in (1, 2) or tuple(mask.shape) != (b, 576, h, w):
        raise ValueError(f'convex_upsample: x [B,1|2,H,W] and mask [B,576,H,W] expected, got {tuple(x.shape)} {tuple(mask.shape)}')
    out = torch.empty(b, c, 8 * h, 8 * w, device=x.device, dtype=torch.float32)
    check(_lib.load().scf_convex_upsample(ptr(x), ptr(mask), ptr(out), b, c, h, w, float(mul), stream_ptr()), 'scf_convex_upsample')
    return out
