"""CPU: host-side boundary - registries/config, constructor kwargs, state-dict keys, C-ABI exports, loud failures."""
import ctypes
import os
import re

import pytest
import torch

import scflow_b200 as S
from scflow_b200 import _lib
from oracle import scflow_oracle as O
from tests.util import ROOT, scflow_model_cfg


def test_registries_hold_reference_names():
    assert S.REFINERS.get('SCFlowRefiner') is S.SCFlowRefiner
    assert S.DECODERS.get('SCFlowDecoder') is S.SCFlowDecoder
    assert S.HEAD.get('MultiClassPoseHead') is S.MultiClassPoseHead
    assert S.HEAD.get('SingleClassPoseHead') is S.SingleClassPoseHead
    assert S.ENCODERS.get('RAFTEncoder') is S.RAFTEncoder
    with pytest.raises(KeyError):
        S.build_decoder(dict(type='NoSuchDecoder'))
    with pytest.raises(KeyError):
        S.build_from_cfg(dict(foo=1), S.DECODERS)


def test_refiner_builds_from_reference_config_and_keys_match():
    model = S.build_refiner(scflow_model_cfg(iters=8))
    assert model.test_iter_num == 8 and model.decoder.iters == 8
    assert model.real_encoder is model.render_encoder          # seperate_encoder=False
    want = set(O.make_model_weights(0).keys())
    have = {k for k in model.state_dict().keys() if not k.endswith('num_batches_tracked')}
    assert have == want, (sorted(want - have)[:5], sorted(have - want)[:5])
    sd = O.make_model_weights(0)
    for k, v in model.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), k
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.endswith('num_batches_tracked') for m in missing)
    n_dec = sum(p.numel() for p in model.decoder.parameters())
    assert n_dec == 6041630                                      # decoder parameter count


def test_decoder_rejects_unshipped_configurations():
    cfg = scflow_model_cfg()['decoder']
    for key, val in (('gru_type', 'Conv'), ('net_type', 'Small'), ('depth_transform', 'linear'), ('act_cfg', None)):
        bad = dict(cfg)
        bad[key] = val
        with pytest.raises((NotImplementedError, AssertionError)):
            S.build_decoder(bad)
    with pytest.raises(NotImplementedError):
        S.CorrLookup(radius=4, align_corners=False)


def test_pose_head_init_is_identity_delta():
    head = S.MultiClassPoseHead(num_class=21, in_channels=224, net_type='Basic', rotation_mode='ortho6d',
                                norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), act_cfg=dict(type='ReLU'))
    assert float(head.rotation_pred.weight.abs().sum()) == 0. and float(head.translation_pred.bias.abs().sum()) == 0.
    assert head.rotation_pred.bias.view(21, 6)[3].tolist() == [1., 0., 0., 0., 1., 0.]
    assert head.conv_layers[0].conv.bias is None              # bias='auto' with norm -> no conv bias
    assert head.fc_in_features == 2048


def test_no_cpu_fallback():
    dec = S.build_decoder(scflow_model_cfg(iters=1)['decoder']).eval()
    f = O.make_features(0, 1)
    s = O.make_scene(0, 1)
    with torch.no_grad(), pytest.raises(RuntimeError, match='CUDA'):
        dec(f['feat_render'], f['feat_real'], f['h_feat'], f['cxt_feat'], s['ref_rotation'], s['ref_translation'], s['depth'],
            s['internel_k'], label=s['label'], init_flow=torch.zeros(1, 2, 256, 256), invalid_flow_num=0.)
    with pytest.raises(RuntimeError, match='CUDA'):
        S.ops.corr_build(f['feat_render'], f['feat_real'])
    with pytest.raises(RuntimeError, match='CUDA'):
        S.CorrLookup(4)([torch.zeros(1024, 1, 32, 32)], torch.zeros(1, 2, 32, 32))


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only machine and exports exactly what include/scflow_b200.h declares."""
    header = open(os.path.join(ROOT, 'include', 'scflow_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(scf_[a-z0-9_]+)\s*\(', header))
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.scf_abi_version() == 1
    # host-only entry points are callable without a GPU
    cfg = _lib.DecoderCfg(4, 4, 21, 6, 0, 0, 1, 0)
    assert lib.scf_decoder_packed_bytes(ctypes.byref(cfg)) > 6041630 * 4
    ws32 = lib.scf_decoder_workspace_bytes(ctypes.byref(cfg), 32, 256, 256)
    assert 32 * 5.57e6 < ws32 < 2e9
    assert lib.scf_decoder_launch_count(ctypes.byref(cfg), 8) > 8 * 20
    bad = _lib.DecoderCfg(4, 4, 21, 4, 0, 0, 1, 0)            # quaternion head is not implemented -> 0 bytes + message
    assert lib.scf_decoder_packed_bytes(ctypes.byref(bad)) == 0
    assert b'ortho6d' in lib.scf_last_error()
    keys = _lib.DECODER_WEIGHT_KEYS
    assert len(keys) == _lib.SCF_W_COUNT == 55 and len(set(keys)) == len(keys)
    enum_body = re.search(r'enum scf_decoder_weight \{(.*?)SCF_W_COUNT', header, flags=re.S).group(1)
    assert len(re.findall(r'SCF_W_[A-Z0-9_]+', enum_body)) == len(keys)


def test_ctypes_mirrors_match_the_compiled_struct_layouts():
    """The ctypes mirrors in scflow_b200/_lib.py have the size the library was compiled with (ABI drift guard)."""
    lib = _lib.load()
    mirrors = list(_lib.STRUCT_MIRRORS)
    assert len(mirrors) >= 7
    for which, cls in enumerate(mirrors):
        assert lib.scf_struct_size(which) == ctypes.sizeof(cls), cls.__name__
    assert lib.scf_struct_size(len(mirrors)) == -1


def test_config_fromfile_with_base(tmp_path):
    (tmp_path / 'base.py').write_text("a = 1\nmodel = dict(type='X', k=dict(p=1, q=2))\n")
    (tmp_path / 'child.py').write_text("_base_ = './base.py'\nmodel = dict(k=dict(q=3), z=5)\nb = 'x'\n")
    cfg = S.Config.fromfile(str(tmp_path / 'child.py'))
    assert cfg.a == 1 and cfg.b == 'x'
    assert cfg.model.type == 'X' and cfg.model.k.p == 1 and cfg.model.k.q == 3 and cfg.model['z'] == 5


def test_shard_helpers():
    from scflow_b200 import dist as D
    for n, w in ((32, 8), (7, 4), (3, 8), (256, 8)):
        spans = [D.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    data = dict(labels=torch.tensor([9, 1, 2, 3, 4]), depth=torch.arange(5.)[:, None, None].repeat(1, 2, 2), tag='x')
    s1 = D.shard_batch(data, 1, 2)
    # per-sample labels stay the shard's own (they select meshes / symmetry in the loss and are returned per image); the pose
    # head's class selector - the GLOBAL label[0], pose_head.py:209-210 - travels separately
    assert s1['labels'].tolist() == [3, 4] and s1['pose_head_label'].tolist() == [9]
    assert s1['depth'][:, 0, 0].tolist() == [3., 4.] and s1['tag'] == 'x'
    assert 'pose_head_label' not in D.shard_batch(data, 1, 2, keep_global_label0=False)


def test_host_only_layout_queries():
    """Workspace slots (encoder -> loop hand-over), tiling and loss-scratch queries are host-only and consistent."""
    lib = _lib.load()
    cfg = _lib.DecoderCfg(4, 4, 21, 6, 0, 0, 1, 1)
    total = lib.scf_decoder_workspace_bytes(ctypes.byref(cfg), 32, 256, 256)
    slots = (ctypes.c_size_t * 4)()
    assert lib.scf_decoder_workspace_slots(ctypes.byref(cfg), 32, 256, 256, slots) == 0
    feat, sh, hf32, scxt = [int(v) for v in slots]
    bp = 32 * 1024
    sizes = [2 * 2 * bp * 256 * 2, 2 * bp * 128 * 2, bp * 128 * 4, 2 * bp * 128 * 2]
    spans = sorted(zip([feat, sh, hf32, scxt], sizes))
    for (o, n), (o2, _) in zip(spans, spans[1:]):
        assert o % 256 == 0 and o + n <= o2, 'slots must be aligned and must not overlap'
    assert spans[-1][0] + spans[-1][1] <= total
    cfg0 = _lib.DecoderCfg(4, 4, 21, 6, 0, 0, 1, 0)             # the fp32 path has no native slots
    assert lib.scf_decoder_workspace_slots(ctypes.byref(cfg0), 32, 256, 256, slots) != 0
    per = ctypes.c_int(0)
    assert lib.scf_conv2d_tc_tiles(32, 32, 32, ctypes.byref(per)) == 256 and per.value == 8
    assert lib.scf_conv2d_tc_tiles(2, 8, 8, ctypes.byref(per)) in (1, 2) and per.value == 0      # a 128-pixel tile spans both samples (the count is an upper bound over the tilings)
    assert lib.scf_refiner_loss_scratch_bytes(8, 32) >= 8 * 296 * 3 * 8 + 8 * 32 * 4
    assert lib.scf_refiner_loss_scratch_bytes(0, 32) == 0


def test_encoder_never_falls_back_silently():
    """eval(): the encoder is ONE C-ABI call or an error that says why - no silent switch to stock PyTorch kernels
    (north_star: no CPU fallback).  train() keeps the differentiable nn.Module graph."""
    import pytest
    enc = S.build_encoder(dict(scflow_model_cfg()['encoder']))
    x = torch.rand(1, 3, 32, 32)
    enc.eval()
    with pytest.raises(RuntimeError, match='autograd is enabled'):
        enc(x)                                   # grad-enabled eval call
    with torch.no_grad(), pytest.raises(RuntimeError, match='not a CUDA tensor'):
        enc(x)                                   # CPU tensor
    enc.train()
    y = enc(x)                                   # training graph: plain nn.Module ops with autograd
    assert y.shape == (1, 256, 4, 4) and y.requires_grad
