"""ctypes binding of libscflow_sm100a.so (the C ABI declared in include/scflow_b200.h).

There is deliberately NO fallback: if the shared library is missing (and cannot be built because nvcc is absent)
or a call fails, a ScfError is raised.  The oracle under oracle/ is never imported from here.
"""
import ctypes as C
import os
import shutil
import threading

from . import _build

c_float_p = C.POINTER(C.c_float)
c_void_p = C.c_void_p


class ScfError(RuntimeError):
    pass


class ConvSeg(C.Structure):
    _fields_ = [('ptr', c_void_p), ('stride', C.c_int), ('coff', C.c_int), ('nch', C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [
        ('seg', ConvSeg * 3), ('nseg', C.c_int),
        ('B', C.c_int), ('Hi', C.c_int), ('Wi', C.c_int), ('Ho', C.c_int), ('Wo', C.c_int),
        ('kh', C.c_int), ('kw', C.c_int), ('sh', C.c_int), ('sw', C.c_int), ('ph', C.c_int), ('pw', C.c_int),
        ('w', c_void_p), ('w_batch_stride', C.c_longlong), ('ldw', C.c_int), ('cout', C.c_int),
        ('bias', c_void_p), ('scale', C.c_float), ('epi', C.c_int), ('act', C.c_int),
        ('out', c_void_p), ('out_stride', C.c_int), ('out_coff', C.c_int),
        ('aux0', c_void_p), ('aux0_stride', C.c_int),
        ('aux1', c_void_p), ('aux1_stride', C.c_int),
        ('out2', c_void_p), ('out2_stride', C.c_int),
        ('out_hl', c_void_p), ('out_hl_plane', C.c_longlong), ('out_hl_stride', C.c_int), ('out_hl_coff', C.c_int),
    ]


class TcSeg(C.Structure):
    _fields_ = [('ptr', c_void_p), ('plane_stride', C.c_longlong), ('stride', C.c_int), ('coff', C.c_int), ('nch', C.c_int)]


class TcConvDesc(C.Structure):
    _fields_ = [
        ('seg', TcSeg * 3), ('nseg', C.c_int),
        ('B', C.c_int), ('H', C.c_int), ('W', C.c_int), ('kh', C.c_int), ('kw', C.c_int), ('stride', C.c_int),
        ('w', c_void_p), ('cin_pad', C.c_int), ('cout_pad', C.c_int), ('cout', C.c_int), ('w_batched', C.c_int),
        ('bias', c_void_p), ('scale', C.c_float), ('epi', C.c_int), ('act', C.c_int),
        ('out_f32', c_void_p), ('out_f32_stride', C.c_int), ('out_f32_coff', C.c_int),
        ('out_hl', c_void_p), ('out_hl_plane', C.c_longlong), ('out_hl_stride', C.c_int), ('out_hl_coff', C.c_int),
        ('aux0', c_void_p), ('aux0_stride', C.c_int), ('aux1', c_void_p), ('aux1_stride', C.c_int),
        ('out2_hl', c_void_p), ('out2_hl_plane', C.c_longlong), ('out2_hl_stride', C.c_int),
        ('pre', c_void_p), ('pre_stride', C.c_int), ('stride_x', C.c_int), ('stride_y', C.c_int), ('w_plane_stride', C.c_longlong), ('stats', c_void_p),
        ('out_pad_writable', C.c_int), ('ksplit', C.c_int), ('split_stride', C.c_longlong),
        ('aux0_hl', c_void_p), ('aux0_hl_plane', C.c_longlong), ('aux0_hl_stride', C.c_int),
    ]


class DecoderCfg(C.Structure):
    _fields_ = [('num_levels', C.c_int), ('radius', C.c_int), ('num_class', C.c_int), ('rot_dim', C.c_int),
                ('mask_flow', C.c_int), ('mask_corr', C.c_int), ('pose_head', C.c_int), ('precision', C.c_int)]


class DecoderIO(C.Structure):
    _fields_ = [
        ('feat_render', c_void_p), ('feat_real', c_void_p), ('h_feat', c_void_p), ('cxt_feat', c_void_p),
        ('ref_rotation', c_void_p), ('ref_translation', c_void_p), ('depth', c_void_p), ('internel_k', c_void_p),
        ('label', c_void_p), ('init_flow', c_void_p), ('invalid_flow_num', C.c_float),
        ('flow_from_pose', c_void_p), ('flow_from_pred', c_void_p), ('rotation', c_void_p), ('translation', c_void_p),
        ('mask', c_void_p), ('delta_rotation', c_void_p), ('delta_translation', c_void_p), ('h_out', c_void_p),
        ('out_batch_total', C.c_int), ('out_batch_offset', C.c_int), ('native_inputs', C.c_int),
    ]


class EncoderOut(C.Structure):
    _fields_ = [
        ('hl0', c_void_p), ('plane0', C.c_longlong), ('stride0', C.c_int), ('f32_0', c_void_p), ('f32_stride0', C.c_int), ('act0', C.c_int),
        ('hl1', c_void_p), ('plane1', C.c_longlong), ('stride1', C.c_int), ('f32_1', c_void_p), ('f32_stride1', C.c_int), ('act1', C.c_int),
        ('split', C.c_int),
    ]


class GruPassDesc(C.Structure):
    """scf_gru_pass_desc (include/scflow_b200.h)."""
    _fields_ = [
        ('h_hl', c_void_p), ('h_plane', C.c_longlong), ('h_f32', c_void_p), ('m_hl', c_void_p), ('m_plane', C.c_longlong),
        ('w_zr', c_void_p), ('w_q', c_void_p), ('pre_zr', c_void_p), ('pre_q', c_void_p),
        ('out_f32', c_void_p), ('out_hl', c_void_p), ('out_plane', C.c_longlong), ('B', C.c_int), ('H', C.c_int), ('W', C.c_int),
        ('vertical', C.c_int),
    ]


class LossDesc(C.Structure):
    _fields_ = [
        ('flow_pred', c_void_p), ('mask_pred', c_void_p), ('rotation', c_void_p), ('translation', c_void_p),
        ('gt_flow', c_void_p), ('valid', c_void_p), ('gt_rotation', c_void_p), ('gt_translation', c_void_p),
        ('label', c_void_p), ('points', c_void_p), ('num_points', c_void_p), ('symmetric', c_void_p), ('diameter', c_void_p),
        ('iters', C.c_int), ('B', C.c_int), ('H', C.c_int), ('W', C.c_int), ('num_class', C.c_int), ('max_points', C.c_int),
        ('max_flow', C.c_float), ('gamma', C.c_float), ('w_flow', C.c_float), ('w_pose', C.c_float), ('w_mask', C.c_float),
        ('eps', C.c_float), ('scratch', c_void_p), ('scratch_bytes', C.c_size_t), ('out', c_void_p),
    ]


ACT = {'none': 0, None: 0, 'relu': 1, 'sigmoid': 2, 'tanh': 3}
EPI_ACT, EPI_GRU_ZR, EPI_GRU_Q = 0, 1, 2

# order of scf_decoder_weight in include/scflow_b200.h  ->  reference state-dict key (relative to the decoder)
DECODER_WEIGHT_KEYS = [
    'encoder.corr_net.0.conv.weight', 'encoder.corr_net.0.conv.bias',
    'encoder.corr_net.1.conv.weight', 'encoder.corr_net.1.conv.bias',
    'encoder.flow_net.0.conv.weight', 'encoder.flow_net.0.conv.bias',
    'encoder.flow_net.1.conv.weight', 'encoder.flow_net.1.conv.bias',
    'encoder.out_net.0.conv.weight', 'encoder.out_net.0.conv.bias',
    'gru.conv_z.0.conv.weight', 'gru.conv_z.0.conv.bias', 'gru.conv_r.0.conv.weight', 'gru.conv_r.0.conv.bias',
    'gru.conv_q.0.conv.weight', 'gru.conv_q.0.conv.bias',
    'gru.conv_z.1.conv.weight', 'gru.conv_z.1.conv.bias', 'gru.conv_r.1.conv.weight', 'gru.conv_r.1.conv.bias',
    'gru.conv_q.1.conv.weight', 'gru.conv_q.1.conv.bias',
    'flow_pred.layers.0.conv.weight', 'flow_pred.layers.0.conv.bias',
    'flow_pred.predict_layer.weight', 'flow_pred.predict_layer.bias',
    'mask_pred.layers.0.conv.weight', 'mask_pred.layers.0.conv.bias',
    'mask_pred.predict_layer.weight', 'mask_pred.predict_layer.bias',
    'delta_flow_encoder.0.conv.weight', 'delta_flow_encoder.0.conv.bias',
    'delta_flow_encoder.1.conv.weight', 'delta_flow_encoder.1.conv.bias',
    'mask_encoder.0.conv.weight', 'mask_encoder.0.conv.bias',
    'mask_encoder.1.conv.weight', 'mask_encoder.1.conv.bias',
    'pose_pred.conv_layers.0.conv.weight', 'pose_pred.conv_layers.0.gn.weight', 'pose_pred.conv_layers.0.gn.bias',
    'pose_pred.conv_layers.1.conv.weight', 'pose_pred.conv_layers.1.gn.weight', 'pose_pred.conv_layers.1.gn.bias',
    'pose_pred.conv_layers.2.conv.weight', 'pose_pred.conv_layers.2.gn.weight', 'pose_pred.conv_layers.2.gn.bias',
    'pose_pred.fc_layers.0.0.weight', 'pose_pred.fc_layers.0.0.bias',
    'pose_pred.fc_layers.1.0.weight', 'pose_pred.fc_layers.1.0.bias',
    'pose_pred.rotation_pred.weight', 'pose_pred.rotation_pred.bias',
    'pose_pred.translation_pred.weight', 'pose_pred.translation_pred.bias',
]
SCF_W_COUNT = len(DECODER_WEIGHT_KEYS)

# RAFTEncoder conv units in the order of SCF_ENC_UNITS (include/scflow_b200.h): (conv module path, following norm path)
ENCODER_UNITS = [('conv1', '{n}1')]
for _stage in (1, 2, 3):
    for _blk in (0, 1):
        _q = f'res_layer{_stage}.{_blk}.'
        ENCODER_UNITS.append((_q + 'conv1', _q + '{n}1'))
        ENCODER_UNITS.append((_q + 'conv2', _q + '{n}2'))
        if _blk == 0 and _stage > 1:
            ENCODER_UNITS.append((_q + 'downsample.0', _q + 'downsample.1'))
ENCODER_UNITS.append(('conv2', None))
ENC_NORM_IN, ENC_NORM_BN = 0, 1

_SIGNATURES = {
    'scf_abi_version': (C.c_int, []),
    'scf_struct_size': (C.c_int, [C.c_int]),
    'scf_convex_upsample': (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_void_p]),
    'scf_format_rendered': (C.c_int, [c_void_p, C.c_int, c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), c_void_p, c_void_p,
                                      c_void_p, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_last_error': (C.c_char_p, []),
    'scf_device_supported': (C.c_int, []),
    'scf_launch_counter': (C.c_longlong, []),
    'scf_nchw_to_nhwc': (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_nhwc_to_nchw': (C.c_int, [c_void_p, C.c_int, C.c_int, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_pack_conv_weight': (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_conv2d': (C.c_int, [C.POINTER(ConvDesc), c_void_p]),
    'scf_conv2d_tc': (C.c_int, [C.POINTER(TcConvDesc), c_void_p]),
    'scf_conv2d_tc_tiles': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    'scf_pack_conv_weight_tc': (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_nchw_to_nhwc_split': (C.c_int, [c_void_p, c_void_p, C.c_longlong, C.c_int, C.c_int, c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, c_void_p]),
    'scf_split_copy': (C.c_int, [c_void_p, C.c_int, C.c_int, c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_longlong, C.c_int, c_void_p]),
    'scf_corr_lookup_split': (C.c_int, [C.POINTER(c_void_p), C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, C.c_longlong, C.c_int,
                                        C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_corr_build_scratch_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    'scf_corr_build': (C.c_int, [c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(c_void_p),
                                 c_void_p, C.c_int, c_void_p]),
    'scf_corr_lookup': (C.c_int, [C.POINTER(c_void_p), C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_corr_lookup_taps': (C.c_int, [C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_group_norm_relu': (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_void_p]),
    'scf_group_norm_relu_split': (C.c_int, [c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, c_void_p,
                                            C.c_longlong, c_void_p]),
    'scf_linear': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_linear_tc': (C.c_int, [c_void_p, C.c_int, C.c_longlong, c_void_p, C.c_int, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_pose_project': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_pose_update': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, c_void_p]),
    'scf_unproject': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_reproject': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_float, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_reproject_down': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_float, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p,
                                     C.c_int, C.c_int, c_void_p]),
    'scf_resize_bilinear': (C.c_int, [c_void_p, c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                                      C.c_int, c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_float, c_void_p]),
    'scf_filter_flow_by_mask': (C.c_int, [c_void_p, c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, c_void_p]),
    'scf_refiner_loss_scratch_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'scf_refiner_loss': (C.c_int, [C.POINTER(LossDesc), c_void_p]),
    'scf_gru_pass_fused': (C.c_int, [C.POINTER(GruPassDesc), c_void_p]),
    'scf_lookup_conv_packed_bytes': (C.c_size_t, []),
    'scf_lookup_conv_pack': (C.c_int, [c_void_p, c_void_p, c_void_p]),
    'scf_lookup_conv': (C.c_int, [C.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                  C.c_int, c_void_p]),
    'scf_clip_adamw': (C.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, C.c_longlong, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int, C.c_float, C.c_float, c_void_p, C.c_int, c_void_p, c_void_p]),
    'scf_encoder_packed_bytes': (C.c_size_t, []),
    'scf_encoder_workspace_bytes': (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    'scf_encoder_pack': (C.c_int, [C.c_int, C.POINTER(c_void_p), c_void_p, c_void_p]),
    'scf_encoder_forward': (C.c_int, [C.c_int, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, c_void_p, c_void_p, C.c_size_t,
                                      c_void_p]),
    'scf_encoder_forward_ex': (C.c_int, [C.c_int, c_void_p, c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(EncoderOut), c_void_p,
                                         C.c_size_t, c_void_p]),
    'scf_decoder_workspace_slots': (C.c_int, [C.POINTER(DecoderCfg), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    'scf_decoder_packed_bytes': (C.c_size_t, [C.POINTER(DecoderCfg)]),
    'scf_decoder_workspace_bytes': (C.c_size_t, [C.POINTER(DecoderCfg), C.c_int, C.c_int, C.c_int]),
    'scf_decoder_pack': (C.c_int, [C.POINTER(DecoderCfg), C.POINTER(c_void_p), c_void_p, c_void_p]),
    'scf_decoder_forward': (C.c_int, [C.POINTER(DecoderCfg), c_void_p, C.POINTER(DecoderIO), C.c_int, C.c_int, C.c_int,
                                      C.c_int, c_void_p, C.c_size_t, c_void_p]),
    'scf_decoder_launch_count': (C.c_int, [C.POINTER(DecoderCfg), C.c_int]),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)
# ctypes mirrors in the order of scf_struct_size(which)
STRUCT_MIRRORS = (ConvDesc, TcConvDesc, DecoderCfg, DecoderIO, EncoderOut, LossDesc, GruPassDesc)

_lib = None
_lock = threading.Lock()


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building in-tree first if nvcc is present and the library is missing/stale). Raises ScfError otherwise."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        have_nvcc = shutil.which('nvcc') is not None or os.path.exists('/usr/local/cuda/bin/nvcc')
        if have_nvcc and os.environ.get('SCFLOW_NO_AUTOBUILD') != '1':
            try:
                if _build.needs_build():
                    _build.build()
            except Exception as e:  # stale-but-present library is still usable; missing one is fatal below
                if not os.path.exists(path):
                    raise ScfError(f'could not build {path}: {e}') from e
        if not os.path.exists(path):
            raise ScfError(f'{path} not found: run `python -c "import __graft_entry__ as g; g.build()"` (needs nvcc). '
                           'scflow_b200 has no CPU or PyTorch fallback.')
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)     # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.scf_abi_version() != 1:
            raise ScfError('libscflow_sm100a.so ABI version mismatch')
        for which, cls in enumerate(STRUCT_MIRRORS):
            if lib.scf_struct_size(which) != C.sizeof(cls):
                raise ScfError(f'libscflow_sm100a.so was built with a different {cls.__name__} layout '
                               f'({lib.scf_struct_size(which)} bytes, binding {C.sizeof(cls)}): rebuild the library')
        _lib = lib
    return _lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().scf_last_error().decode(errors='replace')
        raise ScfError(f'{what or "scflow_b200"} failed (rc={rc}): {msg}')


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """torch's current stream on ``device`` (default: the current device).  Callers that take tensors pass the tensors'
    device and run the C call under ``torch.cuda.device(device)`` so that kernels, tensor maps and the library's side streams
    are issued on the device that owns the memory."""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
