"""CPU oracle pieces of the RAFT baseline decoders.  TEST INFRASTRUCTURE ONLY.
Parity pin: `convex_upsample` was checked against the reference's own RAFTDecoder._upsample
(models/decoder/raft_decoder.py:381-416) run through oracle/ref_shim.py (oracle/make_golden_raft.py; fixture
tests/golden/convex_upsample_b2_6x9.npz)."""
import torch
import torch.nn.functional as F


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor, scale: int = 8, grid_side: int = 3) -> torch.Tensor:
    """raft_decoder.py:403-416 (num_levels = 4 => scale 8; radius = 4 => grid_size 9)."""
    n, _, h, w = flow.shape
    g = grid_side * grid_side
    m = torch.softmax(mask.view(n, 1, g, scale, scale, h, w), dim=2)
    up = F.unfold(scale * flow, [grid_side, grid_side], padding=1).view(n, 2, g, 1, 1, h, w)
    up = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, scale * h, scale * w)


def make_upsample_case(seed: int, b: int, h: int, w: int):
    g = torch.Generator().manual_seed(seed)
    return 3. * torch.randn(b, 2, h, w, generator=g), 2. * torch.randn(b, 576, h, w, generator=g)


# ----------------------------------------------------------------------------------------------------------------------
# RAFTDecoder (models/decoder/raft_decoder.py:296-457), composed from the blocks of scflow_oracle.py (each of which is
# pinned against the reference's module) plus convex_upsample above.  Pinned as a whole against the reference's own
# RAFTDecoder by oracle/make_golden_raft.py (fixture tests/golden/raft_decoder_b2_16x16_it3.npz).
# ----------------------------------------------------------------------------------------------------------------------
def make_raft_decoder_weights(seed: int):
    """Seeded state dict with the reference RAFTDecoder's key names and shapes (Basic net, radius 4: 576 mask channels)."""
    import math
    from . import scflow_oracle as O
    g = torch.Generator().manual_seed(seed + 7919)
    sd = {}
    for name, shp in O._DECODER_SHAPES:
        if name.startswith(('delta_flow_encoder', 'mask_encoder')):
            continue
        if name == 'mask_pred.predict_layer':
            shp = (576, 256, 1, 1)
        sd[name + '.weight'] = torch.randn(*shp, generator=g) / math.sqrt(shp[1] * shp[2] * shp[3])
        sd[name + '.bias'] = 0.05 * torch.randn(shp[0], generator=g)
    return sd


def raft_decoder_forward(sd, feat1, feat2, flow, h_feat, cxt_feat, iters: int, radius: int = 4, num_levels: int = 4):
    """raft_decoder.py:432-457."""
    from . import scflow_oracle as O
    pyramid = O.correlation_pyramid(feat1, feat2, num_levels)
    preds = []
    for _ in range(iters):
        corr = O.corr_lookup(pyramid, flow, radius)
        motion = O.motion_encoder(sd, corr, flow)
        h_feat = O.sepconv_gru(sd, h_feat, torch.cat([cxt_feat, motion], dim=1))
        flow = flow + O.xhead(sd, 'flow_pred.', h_feat, 'flow')
        mask = .25 * O.xhead(sd, 'mask_pred.', h_feat, 'mask')
        preds.append(convex_upsample(flow, mask))
    return preds


def make_raft_inputs(seed: int, b: int, h: int, w: int):
    from . import scflow_oracle as O
    f = O.make_features(seed, b, h, w)
    g = torch.Generator().manual_seed(seed + 31)
    return f['feat_render'], f['feat_real'], 0.5 * torch.randn(b, 2, h, w, generator=g), f['h_feat'], f['cxt_feat']


# ----------------------------------------------------------------------------------------------------------------------
# RAFTDecoderMask (models/decoder/raft_decoder_mask.py:21-208): RAFTDecoder + occlusion head, pinned against the reference's
# own module by oracle/make_golden_raft.py (fixture tests/golden/raft_decoder_mask_b2_16x16_it2.npz).
# ----------------------------------------------------------------------------------------------------------------------
def convex_upsample_mask(occlusion: torch.Tensor, mask: torch.Tensor, scale: int = 8, grid_side: int = 3) -> torch.Tensor:
    """raft_decoder_mask.py:154-162."""
    n, _, h, w = occlusion.shape
    g = grid_side * grid_side
    m = torch.softmax(mask.view(n, 1, g, scale, scale, h, w), dim=2)
    up = F.unfold(occlusion, [grid_side, grid_side], padding=1).view(n, 1, g, 1, 1, h, w)
    up = torch.sum(up * m, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 1, scale * h, scale * w)


def make_raft_decoder_mask_weights(seed: int):
    import math
    sd = make_raft_decoder_weights(seed)
    g = torch.Generator().manual_seed(seed + 104743)
    for name, shp in (('occlusion_pred.layers.0.conv', (256, 128, 3, 3)), ('occlusion_pred.predict_layer', (1, 256, 1, 1))):
        sd[name + '.weight'] = torch.randn(*shp, generator=g) / math.sqrt(shp[1] * shp[2] * shp[3])
        sd[name + '.bias'] = 0.05 * torch.randn(shp[0], generator=g)
    return sd


def raft_decoder_mask_forward(sd, feat1, feat2, flow, h_feat, cxt_feat, iters: int, radius: int = 4, num_levels: int = 4):
    """raft_decoder_mask.py:180-208."""
    from . import scflow_oracle as O
    pyramid = O.correlation_pyramid(feat1, feat2, num_levels)
    flows, occs = [], []
    for _ in range(iters):
        corr = O.corr_lookup(pyramid, flow, radius)
        motion = O.motion_encoder(sd, corr, flow)
        h_feat = O.sepconv_gru(sd, h_feat, torch.cat([cxt_feat, motion], dim=1))
        flow = flow + O.xhead(sd, 'flow_pred.', h_feat, 'flow')
        occlusion = torch.sigmoid(O.xhead(sd, 'occlusion_pred.', h_feat, 'mask'))
        mask = .25 * O.xhead(sd, 'mask_pred.', h_feat, 'mask')
        flows.append(convex_upsample(flow, mask))
        occs.append(convex_upsample_mask(occlusion, mask))
    return flows, occs
