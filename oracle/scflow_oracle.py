"""CPU oracle: a functional fp32 restatement of SCFlow's pose-refinement hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file.  The product package
(``scflow_b200``) never does; it fails loudly if its CUDA library is missing.

Parity pin: the reference ships no tests / golden vectors, so this restatement is pinned
against the reference's OWN code executed unmodified in the build container through
``oracle/ref_shim.py``; the fixtures in ``tests/golden/*.npz`` were produced by that run
(``oracle/make_golden.py``) and are what the CPU test-suite re-checks on every machine.

Everything is a pure function of a flat state dict that uses the reference's parameter names
(``encoder.corr_net.0.conv.weight`` ... ``pose_pred.rotation_pred.bias``), so a reference checkpoint's
``decoder.*`` sub-dict can be passed directly.  All paths below are relative to /root/reference.
"""
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def _conv(sd: SD, name: str, x: Tensor, padding, stride=1, act: str = 'none', bias: bool = True) -> Tensor:
    """ConvModule without norm: conv -> activation (mmcv ConvModule order; models/decoder/raft_decoder.py:140-148)."""
    b = sd.get(name + '.bias') if bias else None
    y = F.conv2d(x, sd[name + '.weight'], b, stride=stride, padding=padding)
    if act == 'relu':
        y = torch.relu(y)
    elif act == 'sigmoid':
        y = torch.sigmoid(y)
    elif act == 'tanh':
        y = torch.tanh(y)
    else:
        assert act == 'none'
    return y


# ----------------------------------------------------------------------------------------------
# correlation pyramid  (models/decoder/raft_decoder.py:35-58)
# ----------------------------------------------------------------------------------------------
def correlation_pyramid(feat_render: Tensor, feat_real: Tensor, num_levels: int = 4) -> List[Tensor]:
    """corr[n*P+q, 0, y, x] = <feat_render[n,:,q], feat_real[n,:,y,x]> / sqrt(C); then floor 2x2 mean pools."""
    n, c, h, w = feat_render.shape
    a = feat_render.reshape(n, c, h * w).transpose(1, 2)          # [n, P(query), C]
    b = feat_real.reshape(n, c, h * w)                            # [n, C, P(key)]
    vol = torch.bmm(a, b) / math.sqrt(float(c))                   # raft_decoder.py:48-52
    vol = vol.reshape(n * h * w, 1, h, w)
    levels = [vol]
    for _ in range(num_levels - 1):
        levels.append(F.avg_pool2d(levels[-1], kernel_size=2, stride=2))   # raft_decoder.py:54-56
    return levels


# ----------------------------------------------------------------------------------------------
# correlation lookup  (models/utils/corr_lookup.py:102-136)
# ----------------------------------------------------------------------------------------------
def lookup_taps(flow: Tensor, level: int, wl: int, hl: int, radius: int = 4):
    """Integer index work of the lookup, replayed with the reference's exact fp32 operation order.

    Returns (x0, y0, fx, fy): int32 top-left neighbour index and fp32 fractional offsets, each of shape
    [B, H, W, 2r+1, 2r+1] indexed [.., a, b] where the output channel is ``a*(2r+1)+b`` and the sample
    position is (cx + (a-r), cy + (b-r))  -- x-major window, corr_lookup.py:118-128.

    fp32 sequence (corr_lookup.py:115,127-128 then :64-65, then ATen grid_sampler un-normalise with
    align_corners=True):   c = (x + flow) / 2**level + d;  g = c*2/(W_l-1) - 1;  i = ((g+1)*0.5)*(W_l-1)
    """
    b, _, h, w = flow.shape
    r = radius
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, w)
    ys = torch.arange(h, dtype=torch.float32).view(1, h, 1)
    cx = ((xs + flow[:, 0]) / float(2 ** level)).view(b, h, w, 1, 1)
    cy = ((ys + flow[:, 1]) / float(2 ** level)).view(b, h, w, 1, 1)
    d = torch.arange(-r, r + 1, dtype=torch.float32)
    px = cx + d.view(1, 1, 1, -1, 1)          # first window axis moves x
    py = cy + d.view(1, 1, 1, 1, -1)          # second window axis moves y
    px, py = torch.broadcast_tensors(px, py)
    gx = px * 2.0 / float(max(wl - 1, 1)) - 1.0
    gy = py * 2.0 / float(max(hl - 1, 1)) - 1.0
    ix = ((gx + 1.0) * 0.5) * float(wl - 1)
    iy = ((gy + 1.0) * 0.5) * float(hl - 1)
    x0f, y0f = torch.floor(ix), torch.floor(iy)
    return x0f.to(torch.int32), y0f.to(torch.int32), ix - x0f, iy - y0f


def corr_lookup_explicit(pyramid: Sequence[Tensor], flow: Tensor, radius: int = 4) -> Tensor:
    """Gather-based lookup (no grid_sample): bilinear, zeros padding, align_corners=True.

    Values agree with :func:`corr_lookup` to ~1e-6 (weight rounding differs between ATen's CPU and CUDA
    grid samplers too); neighbour indices / in-bounds masks are the bit-exact part.
    """
    b, _, h, w = flow.shape
    k = 2 * radius + 1
    out = []
    for lvl, vol in enumerate(pyramid):
        hl, wl = vol.shape[-2:]
        x0, y0, fx, fy = lookup_taps(flow, lvl, wl, hl, radius)
        v = vol.reshape(b, h, w, hl * wl)
        acc = torch.zeros(b, h, w, k, k, dtype=torch.float32)
        for dy, dx, wgt in ((0, 0, (1 - fx) * (1 - fy)), (0, 1, fx * (1 - fy)),
                            (1, 0, (1 - fx) * fy), (1, 1, fx * fy)):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi < wl) & (yi >= 0) & (yi < hl)
            lin = (yi.clamp(0, hl - 1) * wl + xi.clamp(0, wl - 1)).long()
            g = torch.gather(v, 3, lin.reshape(b, h, w, k * k)).reshape(b, h, w, k, k)
            acc = acc + torch.where(ok, g * wgt, torch.zeros_like(g))
        out.append(acc.reshape(b, h, w, k * k))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous()


def corr_lookup(pyramid: Sequence[Tensor], flow: Tensor, radius: int = 4) -> Tensor:
    """The reference's own formulation (grid_sample); output [B, L*(2r+1)^2, H, W] fp32."""
    b, _, h, w = flow.shape
    k = 2 * radius + 1
    dev = flow.device                                                       # (CPU in the tests; bench.py's eager-GPU comparator passes CUDA tensors)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing='ij')
    base = torch.stack([xs, ys], dim=0).float()[None] + flow               # corr_lookup.py:113-115
    base = base.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
    d = torch.linspace(-radius, radius, k, device=dev)
    first, second = torch.meshgrid(d, d, indexing='ij')
    delta = torch.stack([first, second], dim=-1).view(1, k, k, 2)          # x-major: corr_lookup.py:118-123
    outs = []
    for lvl, vol in enumerate(pyramid):
        hl, wl = vol.shape[-2:]
        coords = base / 2 ** lvl + delta                                    # corr_lookup.py:127-128
        gx = coords[..., 0] * 2. / max(wl - 1, 1) - 1.                      # corr_lookup.py:64-65
        gy = coords[..., 1] * 2. / max(hl - 1, 1) - 1.
        smp = F.grid_sample(vol, torch.stack([gx, gy], dim=-1), mode='bilinear', padding_mode='zeros',
                            align_corners=True)
        outs.append(smp.view(b, h, w, k * k))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# ----------------------------------------------------------------------------------------------
# motion encoder / GRU / heads   (models/decoder/raft_decoder.py:152-166, 235-253, 292-294)
# ----------------------------------------------------------------------------------------------
def motion_encoder(sd: SD, corr: Tensor, flow: Tensor, p: str = 'encoder.') -> Tensor:
    c = _conv(sd, p + 'corr_net.0.conv', corr, 0, act='relu')      # 1x1 324->256
    c = _conv(sd, p + 'corr_net.1.conv', c, 1, act='relu')         # 3x3 256->192
    f = _conv(sd, p + 'flow_net.0.conv', flow, 3, act='relu')      # 7x7 2->128
    f = _conv(sd, p + 'flow_net.1.conv', f, 1, act='relu')         # 3x3 128->64
    o = _conv(sd, p + 'out_net.0.conv', torch.cat([c, f], 1), 1, act='relu')   # 3x3 256->126
    return torch.cat([o, flow], dim=1)                              # 128 channels


def sepconv_gru(sd: SD, h: Tensor, x: Tensor, p: str = 'gru.') -> Tensor:
    """Two GRU passes: 1x5 (pad (0,2)) then 5x1 (pad (2,0)); concat order is [h, x]."""
    for i, pad in enumerate(((0, 2), (2, 0))):
        hx = torch.cat([h, x], dim=1)
        z = _conv(sd, f'{p}conv_z.{i}.conv', hx, pad, act='sigmoid')
        r = _conv(sd, f'{p}conv_r.{i}.conv', hx, pad, act='sigmoid')
        q = _conv(sd, f'{p}conv_q.{i}.conv', torch.cat([r * h, x], dim=1), pad, act='tanh')
        h = (1 - z) * h + z * q
    return h


def xhead(sd: SD, p: str, h: Tensor, kind: str) -> Tensor:
    y = _conv(sd, p + 'layers.0.conv', h, 1, act='relu')           # 3x3 128->256, default ConvModule ReLU
    return _conv(sd, p + 'predict_layer', y, 1 if kind == 'flow' else 0)


# ----------------------------------------------------------------------------------------------
# pose head  (models/head/pose_head.py:201-211 ; SingleClassPoseHead :98-104)
# ----------------------------------------------------------------------------------------------
def pose_head(sd: SD, x: Tensor, label: Tensor, num_class: int = 21, rot_dim: int = 6,
              p: str = 'pose_pred.', multi_class: bool = True, num_groups: int = 32) -> Tuple[Tensor, Tensor]:
    for i in range(3):
        x = F.conv2d(x, sd[f'{p}conv_layers.{i}.conv.weight'], None, stride=2, padding=1)   # bias='auto' -> none
        x = F.group_norm(x, num_groups, sd[f'{p}conv_layers.{i}.gn.weight'], sd[f'{p}conv_layers.{i}.gn.bias'], 1e-5)
        x = torch.relu(x)
    x = x.flatten(1)
    for i in range(2):
        x = torch.relu(F.linear(x, sd[f'{p}fc_layers.{i}.0.weight'], sd[f'{p}fc_layers.{i}.0.bias']))
    rot = F.linear(x, sd[p + 'rotation_pred.weight'], sd[p + 'rotation_pred.bias'])
    trs = F.linear(x, sd[p + 'translation_pred.weight'], sd[p + 'translation_pred.bias'])
    if not multi_class:
        return rot, trs
    # quirk kept on purpose: index_select(dim=1, label)[:, 0] picks label[0]'s class for EVERY row
    cls = int(label[0])
    rot = rot.view(-1, num_class, rot_dim)[:, cls]
    trs = trs.view(-1, num_class, 3)[:, cls]
    return rot, trs


# ----------------------------------------------------------------------------------------------
# geometry  (models/utils/pose.py)
# ----------------------------------------------------------------------------------------------
def ortho6d_to_matrix(o6: Tensor) -> Tensor:
    """pose.py:153-169: columns (x, y, z), x = a1/|a1|, z = (x × a2)/|.|, y = z × x."""
    x = F.normalize(o6[:, 0:3], p=2, dim=1)
    z = F.normalize(torch.cross(x, o6[:, 3:6], dim=1), p=2, dim=1)
    y = torch.cross(z, x, dim=1)
    return torch.stack([x, y, z], dim=2)


def update_pose(d_rot: Tensor, d_trs: Tensor, rot: Tensor, trs: Tensor, weight: float = 10.,
                detach_depth_for_xy: bool = False) -> Tuple[Tensor, Tensor]:
    """pose.py:124-149 with depth_transform='exp': R' = dR R; tz' = tz/exp(dz); txy' = tz'(dxy/10 + txy/tz).
    ``detach_depth_for_xy`` (pose.py:143-145; True in the shipped config) changes gradients only."""
    r_new = torch.bmm(ortho6d_to_matrix(d_rot), rot)
    tz = trs[:, 2] / torch.exp(d_trs[:, 2])
    tzx = tz.detach() if detach_depth_for_xy else tz
    tx = tzx * (d_trs[:, 0] / weight + trs[:, 0] / trs[:, 2])
    ty = tzx * (d_trs[:, 1] / weight + trs[:, 1] / trs[:, 2])
    return r_new, torch.stack([tx, ty, tz], dim=-1)


def unproject_dense(depth: Tensor, k: Tensor, rot: Tensor, trs: Tensor) -> Tensor:
    """Dense form of pose.py:26-41,44-64: object-frame point for EVERY pixel, [B,3,H,W] (garbage where depth<=0).

    X_cam = K^-1 (x d, y d, d)^T ; X_obj = R^-1 (X_cam - t) with torch.inverse as in the reference.
    """
    b, h, w = depth.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=depth.device),
                            torch.arange(w, dtype=torch.float32, device=depth.device), indexing='ij')
    homo = torch.stack([xs, ys, torch.ones_like(xs)], dim=0)[None] * depth[:, None]     # [B,3,H,W]
    cam = torch.bmm(torch.inverse(k), homo.reshape(b, 3, -1))
    obj = torch.bmm(torch.inverse(rot), cam - trs[:, :, None])
    return obj.reshape(b, 3, h, w)


def reproject_dense(points_obj: Tensor, depth: Tensor, k: Tensor, rot: Tensor, trs: Tensor,
                    invalid: float = 0.) -> Tensor:
    """Dense form of pose.py:66-88: flow = proj(K(R X + t)) - pixel at depth>0, ``invalid`` elsewhere."""
    b, _, h, w = points_obj.shape
    u = torch.bmm(k, torch.bmm(rot, points_obj.reshape(b, 3, -1)) + trs[:, :, None]).reshape(b, 3, h, w)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=depth.device),
                            torch.arange(w, dtype=torch.float32, device=depth.device), indexing='ij')
    fx = u[:, 0] / u[:, 2] - xs
    fy = u[:, 1] / u[:, 2] - ys
    fg = depth > 0
    inv = torch.full_like(fx, invalid)
    return torch.stack([torch.where(fg, fx, inv), torch.where(fg, fy, inv)], dim=1)


def resize_bilinear_ac(x: Tensor, out_h: int, out_w: int) -> Tensor:
    """F.interpolate(mode='bilinear', align_corners=True) (scflow_decoder.py:196-197, 223-227)."""
    return F.interpolate(x, size=(out_h, out_w), mode='bilinear', align_corners=True)


# ----------------------------------------------------------------------------------------------
# the decoder loop  (models/decoder/scflow_decoder.py:150-251)
# ----------------------------------------------------------------------------------------------
def decoder_forward(sd: SD, feat_render: Tensor, feat_real: Tensor, h_feat: Tensor, cxt_feat: Tensor,
                    ref_rotation: Tensor, ref_translation: Tensor, depth: Tensor, internel_k: Tensor,
                    label: Tensor, init_flow: Tensor, invalid_flow_num: float = 0., iters: int = 8,
                    num_levels: int = 4, radius: int = 4, num_class: int = 21, identity_pose_head: bool = False,
                    trace: dict = None, detach: bool = False):
    """Returns the reference's 7 lists. ``identity_pose_head`` = config-3 mode (stock head cannot run at
    480x640): delta pose is the zero-init head's output (identity) in oracle and candidate."""
    scale = 2 ** (num_levels - 1)
    b, hh, ww = depth.shape
    h8, w8 = hh // scale, ww // scale
    pyramid = correlation_pyramid(feat_render, feat_real, num_levels)
    pts = unproject_dense(depth, internel_k, ref_rotation, ref_translation)
    rot, trs = ref_rotation, ref_translation
    flow = init_flow
    outs = ([], [], [], [], [], [], [])
    for it in range(iters):
        if detach:        # detach_flow / detach_pose / detach_depth_for_xy = True in the shipped config (scflow.py:57-60; scflow_decoder.py:192-195, 232-233):
            flow, rot, trs = flow.detach(), rot.detach(), trs.detach()            # changes gradients only, never forward values
        flow8 = (1.0 / scale) * resize_bilinear_ac(flow, h8, w8)                  # :196-197
        corr = corr_lookup(pyramid, flow8, radius)                                 # :198
        motion = motion_encoder(sd, corr, flow8)                                   # :206
        h_feat = sepconv_gru(sd, h_feat, torch.cat([cxt_feat, motion], dim=1))     # :207-208
        d_flow = xhead(sd, 'flow_pred.', h_feat, 'flow')                           # :210
        mask = torch.sigmoid(xhead(sd, 'mask_pred.', h_feat, 'mask'))              # :212-213
        df = _conv(sd, 'delta_flow_encoder.0.conv', d_flow, 3, act='relu')         # :216
        df = _conv(sd, 'delta_flow_encoder.1.conv', df, 1, act='relu')
        mf = _conv(sd, 'mask_encoder.0.conv', mask, 1, act='relu')                 # :217
        mf = _conv(sd, 'mask_encoder.1.conv', mf, 1, act='relu')
        if identity_pose_head:
            d_rot = torch.tensor([1., 0., 0., 0., 1., 0.], device=depth.device).repeat(b, 1)
            d_trs = torch.zeros(b, 3, device=depth.device)
        else:
            d_rot, d_trs = pose_head(sd, torch.cat([h_feat, df, mf], dim=1), label, num_class)   # :218-219
        flow_pred = scale * resize_bilinear_ac(flow8 + d_flow, hh, ww)             # :222-224
        mask_up = resize_bilinear_ac(mask, hh, ww)                                 # :226-227
        rot, trs = update_pose(d_rot, d_trs, rot, trs, detach_depth_for_xy=detach)  # :230-236
        flow = reproject_dense(pts, depth, internel_k, rot, trs, invalid_flow_num)  # :239-243
        if trace is not None:
            trace.setdefault('flow8', []).append(flow8)
            trace.setdefault('corr', []).append(corr)
            trace.setdefault('motion', []).append(motion)
            trace.setdefault('h', []).append(h_feat)
            trace.setdefault('d_flow', []).append(d_flow)
            trace.setdefault('mask8', []).append(mask)
        for lst, v in zip(outs, (flow, flow_pred, rot, trs, mask_up, d_rot, d_trs)):
            lst.append(v)
    return outs


# ----------------------------------------------------------------------------------------------
# RAFT 'Basic' encoder  (models/encoder/raft_encoder.py:286-314, models/backbone/resnet.py:14-94,678-773)
# ----------------------------------------------------------------------------------------------
def _norm(sd: SD, name: str, x: Tensor, kind: str) -> Tensor:
    if kind == 'IN':
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'], sd[name + '.weight'],
                        sd[name + '.bias'], training=False, eps=1e-5)


def raft_encoder(sd: SD, x: Tensor, norm: str = 'IN', p: str = '') -> Tensor:
    n = 'in' if norm == 'IN' else 'bn'
    x = torch.relu(_norm(sd, f'{p}{n}1', F.conv2d(x, sd[p + 'conv1.weight'], sd[p + 'conv1.bias'], stride=2, padding=3), norm))
    for stage, stride in ((1, 1), (2, 2), (3, 2)):
        for blk in range(2):
            q = f'{p}res_layer{stage}.{blk}.'
            s = stride if blk == 0 else 1
            y = F.conv2d(x, sd[q + 'conv1.weight'], sd[q + 'conv1.bias'], stride=s, padding=1)
            y = torch.relu(_norm(sd, q + n + '1', y, norm))
            y = F.conv2d(y, sd[q + 'conv2.weight'], sd[q + 'conv2.bias'], padding=1)
            y = _norm(sd, q + n + '2', y, norm)
            if (q + 'downsample.0.weight') in sd:
                x = F.conv2d(x, sd[q + 'downsample.0.weight'], sd[q + 'downsample.0.bias'], stride=s)
                x = _norm(sd, q + 'downsample.1', x, norm)
            x = torch.relu(y + x)
    return F.conv2d(x, sd[p + 'conv2.weight'], sd[p + 'conv2.bias'])


def get_pose(sd_model: SD, render_images: Tensor, real_images: Tensor, ref_rotation: Tensor,
             ref_translation: Tensor, depth: Tensor, internel_k: Tensor, label: Tensor, iters: int = 8,
             identity_pose_head: bool = False):
    """models/refiner/scflow_refiner.py:112-142 (+ extract_feat :88-110). ``sd_model`` uses the refiner's
    key names: ``real_encoder.* / render_encoder.*`` (shared), ``context.*``, ``decoder.*``."""
    enc = {k[len('render_encoder.'):]: v for k, v in sd_model.items() if k.startswith('render_encoder.')}
    ctx = {k[len('context.'):]: v for k, v in sd_model.items() if k.startswith('context.')}
    dec = {k[len('decoder.'):]: v for k, v in sd_model.items() if k.startswith('decoder.')}
    feat_real = raft_encoder(enc, real_images, 'IN')
    feat_render = raft_encoder(enc, render_images, 'IN')
    c = raft_encoder(ctx, render_images, 'BN')
    h_feat, cxt = torch.tanh(c[:, :128]), torch.relu(c[:, 128:])
    n, _, hh, ww = real_images.shape
    init_flow = torch.zeros(n, 2, hh, ww, device=render_images.device)
    return decoder_forward(dec, feat_render, feat_real, h_feat, cxt, ref_rotation, ref_translation, depth,
                           internel_k, label, init_flow, 0., iters=iters, identity_pose_head=identity_pose_head)


# ----------------------------------------------------------------------------------------------
# synthetic scenes + weights (SURVEY.md §8d generator; shared by tests and bench so inputs are identical)
# ----------------------------------------------------------------------------------------------
def random_rotation(gen: torch.Generator, n: int) -> Tensor:
    q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=gen))
    q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2))[:, None, :]
    det = torch.det(q)
    q[:, :, 2] = q[:, :, 2] * det[:, None]
    return q.contiguous()


def axis_angle_matrix(axis: Tensor, angle: Tensor) -> Tensor:
    axis = F.normalize(axis, dim=1)
    x, y, z = axis.unbind(1)
    zero = torch.zeros_like(x)
    kx = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], dim=1).view(-1, 3, 3)
    s, c = torch.sin(angle)[:, None, None], torch.cos(angle)[:, None, None]
    return torch.eye(3)[None] + s * kx + (1 - c) * torch.bmm(kx, kx)


def make_scene(seed: int, batch: int, height: int = 256, width: int = 256, num_class: int = 21):
    """Seeded synthetic inputs: intrinsics, reference pose, analytic ellipsoid z-buffer (~35-45 % fg), labels,
    and procedural real/rendered images in [0,1]. Units are millimetres like YCB-V."""
    g = torch.Generator().manual_seed(seed)
    f = 500. + 1500. * torch.rand(batch, generator=g)
    k = torch.zeros(batch, 3, 3)
    k[:, 0, 0] = f
    k[:, 1, 1] = f
    k[:, 0, 2] = width / 2.
    k[:, 1, 2] = height / 2.
    k[:, 2, 2] = 1.
    rot = random_rotation(g, batch)
    trs = torch.stack([20. * torch.randn(batch, generator=g), 20. * torch.randn(batch, generator=g),
                       600. + 600. * torch.rand(batch, generator=g)], dim=1)
    # ellipsoid semi-axes chosen so the silhouette covers ~40 % of the crop
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32), indexing='ij')
    rays = torch.stack([(xs[None] - k[:, 0, 2, None, None]) / f[:, None, None],
                        (ys[None] - k[:, 1, 2, None, None]) / f[:, None, None],
                        torch.ones(batch, height, width)], dim=1)                   # camera rays, [B,3,H,W]
    radius_px = 0.36 * min(height, width)
    base = radius_px * trs[:, 2] / f                                               # mm
    axes = base[:, None] * (0.75 + 0.5 * torch.rand(batch, 3, generator=g))
    # ray / ellipsoid intersection in the object frame: X_obj = R^T (s*ray - t)
    d_o = torch.einsum('bji,bjhw->bihw', rot, rays) / axes[:, :, None, None]
    o_o = -(torch.einsum('bji,bj->bi', rot, trs) / axes)[:, :, None, None]
    a = (d_o * d_o).sum(1)
    bq = 2 * (d_o * o_o).sum(1)
    c = (o_o * o_o).sum(1) - 1
    disc = bq * bq - 4 * a * c
    s = (-bq - torch.sqrt(disc.clamp_min(0))) / (2 * a)
    depth = torch.where(disc > 0, s, torch.zeros_like(s)).contiguous()             # z-buffer (ray z-component is 1)
    label = torch.randint(0, num_class, (batch,), generator=g)
    # ground-truth pose = reference jittered (angle N(0,15deg), xy N(0,15mm), z N(0,50mm)); used for the real image
    ang = torch.randn(batch, generator=g) * math.radians(15.)
    rot_gt = torch.bmm(axis_angle_matrix(torch.randn(batch, 3, generator=g), ang), rot)
    trs_gt = trs + torch.stack([15. * torch.randn(batch, generator=g), 15. * torch.randn(batch, generator=g),
                                50. * torch.randn(batch, generator=g)], dim=1)

    def texture(r_, t_):
        # procedural texture of the object-frame surface point under pose (r_, t_), using the same ellipsoid
        d2 = torch.einsum('bji,bjhw->bihw', r_, rays) / axes[:, :, None, None]
        o2 = -(torch.einsum('bji,bj->bi', r_, t_) / axes)[:, :, None, None]
        a2, b2, c2 = (d2 * d2).sum(1), 2 * (d2 * o2).sum(1), (o2 * o2).sum(1) - 1
        disc2 = b2 * b2 - 4 * a2 * c2
        s2 = (-b2 - torch.sqrt(disc2.clamp_min(0))) / (2 * a2)
        pt = o2 + d2 * s2[:, None]                                                  # unit-sphere coords
        tex = torch.stack([0.5 + 0.5 * torch.sin(6.0 * pt[:, 0] + 2.0 * pt[:, 1]),
                           0.5 + 0.5 * torch.sin(5.0 * pt[:, 1] - 3.0 * pt[:, 2]),
                           0.5 + 0.5 * torch.cos(4.0 * pt[:, 2] + 1.5 * pt[:, 0])], dim=1)
        return torch.where((disc2 > 0)[:, None], tex, torch.full_like(tex, 0.5))

    render = (texture(rot, trs) + 0.02 * torch.randn(batch, 3, height, width, generator=g)).clamp(0, 1)
    real = (texture(rot_gt, trs_gt) + 0.02 * torch.randn(batch, 3, height, width, generator=g)).clamp(0, 1)
    return dict(internel_k=k, ref_rotation=rot, ref_translation=trs, depth=depth, label=label,
                render_images=render.contiguous(), real_images=real.contiguous(), gt_rotation=rot_gt, gt_translation=trs_gt)


def make_features(seed: int, batch: int, h8: int = 32, w8: int = 32, channels: int = 256):
    """Feature-level synthetic inputs: feat ~ N(0,1), h = tanh(N), cxt = relu(N)."""
    g = torch.Generator().manual_seed(seed + 7919)
    return dict(feat_render=torch.randn(batch, channels, h8, w8, generator=g),
                feat_real=torch.randn(batch, channels, h8, w8, generator=g),
                h_feat=torch.tanh(torch.randn(batch, 128, h8, w8, generator=g)),
                cxt_feat=torch.relu(torch.randn(batch, 128, h8, w8, generator=g)))


_DECODER_SHAPES = [
    ('encoder.corr_net.0.conv', (256, 324, 1, 1)), ('encoder.corr_net.1.conv', (192, 256, 3, 3)),
    ('encoder.flow_net.0.conv', (128, 2, 7, 7)), ('encoder.flow_net.1.conv', (64, 128, 3, 3)),
    ('encoder.out_net.0.conv', (126, 256, 3, 3)),
    ('gru.conv_z.0.conv', (128, 384, 1, 5)), ('gru.conv_r.0.conv', (128, 384, 1, 5)), ('gru.conv_q.0.conv', (128, 384, 1, 5)),
    ('gru.conv_z.1.conv', (128, 384, 5, 1)), ('gru.conv_r.1.conv', (128, 384, 5, 1)), ('gru.conv_q.1.conv', (128, 384, 5, 1)),
    ('flow_pred.layers.0.conv', (256, 128, 3, 3)), ('flow_pred.predict_layer', (2, 256, 3, 3)),
    ('mask_pred.layers.0.conv', (256, 128, 3, 3)), ('mask_pred.predict_layer', (1, 256, 1, 1)),
    ('delta_flow_encoder.0.conv', (128, 2, 7, 7)), ('delta_flow_encoder.1.conv', (64, 128, 3, 3)),
    ('mask_encoder.0.conv', (64, 1, 3, 3)), ('mask_encoder.1.conv', (32, 64, 3, 3)),
]


def make_decoder_weights(seed: int, num_class: int = 21) -> SD:
    """Seeded decoder state dict with the reference's key names and shapes. Conv/FC weights are
    fan-in-scaled normals (keeps activations O(1) over 8+ iterations); the pose head's output layers get
    N(0, 0.02) weights so that the predicted delta pose is NOT the identity (stock init is zero,
    pose_head.py:187-198, which would make the pose branch degenerate - SURVEY §8d)."""
    g = torch.Generator().manual_seed(seed + 104729)
    sd = {}

    def normal(shape, std):
        return torch.randn(*shape, generator=g) * std

    for name, shp in _DECODER_SHAPES:
        fan_in = shp[1] * shp[2] * shp[3]
        sd[name + '.weight'] = normal(shp, 1.0 / math.sqrt(fan_in))
        sd[name + '.bias'] = normal((shp[0],), 0.05)
    cin = 224
    for i in range(3):
        sd[f'pose_pred.conv_layers.{i}.conv.weight'] = normal((128, cin, 3, 3), 1.0 / math.sqrt(cin * 9))
        sd[f'pose_pred.conv_layers.{i}.gn.weight'] = 1.0 + normal((128,), 0.1)
        sd[f'pose_pred.conv_layers.{i}.gn.bias'] = normal((128,), 0.1)
        cin = 128
    sd['pose_pred.fc_layers.0.0.weight'] = normal((1024, 2048), 1.0 / math.sqrt(2048))
    sd['pose_pred.fc_layers.0.0.bias'] = normal((1024,), 0.05)
    sd['pose_pred.fc_layers.1.0.weight'] = normal((256, 1024), 1.0 / math.sqrt(1024))
    sd['pose_pred.fc_layers.1.0.bias'] = normal((256,), 0.05)
    sd['pose_pred.rotation_pred.weight'] = normal((6 * num_class, 256), 0.02 / 16)
    sd['pose_pred.rotation_pred.bias'] = torch.tensor([1., 0., 0., 0., 1., 0.] * num_class)
    sd['pose_pred.translation_pred.weight'] = normal((3 * num_class, 256), 0.02 / 16)
    sd['pose_pred.translation_pred.bias'] = torch.zeros(3 * num_class)
    return sd


def make_encoder_weights(seed: int, norm: str) -> SD:
    """Seeded RAFTEncoder('Basic') state dict (keys of raft_encoder.py / resnet.py BasicBlock)."""
    g = torch.Generator().manual_seed(seed + (15485863 if norm == 'IN' else 32452843))
    sd = {}
    n = 'in' if norm == 'IN' else 'bn'

    def conv(name, co, ci, k):
        sd[name + '.weight'] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(2.0 / (ci * k * k))
        sd[name + '.bias'] = torch.randn(co, generator=g) * 0.02

    def bn(name, c):
        if norm == 'BN':
            sd[name + '.weight'] = 1.0 + 0.1 * torch.randn(c, generator=g)
            sd[name + '.bias'] = 0.1 * torch.randn(c, generator=g)
            sd[name + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
            sd[name + '.running_var'] = 1.0 + 0.2 * torch.rand(c, generator=g)

    conv('conv1', 64, 3, 7)
    bn(n + '1', 64)
    cin = 64
    for stage, (planes, stride) in enumerate(((64, 1), (96, 2), (128, 2)), start=1):
        for blk in range(2):
            q = f'res_layer{stage}.{blk}.'
            conv(q + 'conv1', planes, cin if blk == 0 else planes, 3)
            bn(q + n + '1', planes)
            conv(q + 'conv2', planes, planes, 3)
            bn(q + n + '2', planes)
            if blk == 0 and (stride != 1 or cin != planes):
                conv(q + 'downsample.0', planes, cin, 1)
                bn(q + 'downsample.1', planes)
        cin = planes
    conv('conv2', 256, 128, 1)
    return sd


def make_model_weights(seed: int, num_class: int = 21) -> SD:
    """Full refiner state dict: shared IN encoder under both real_/render_encoder, BN context, decoder."""
    sd = {}
    enc = make_encoder_weights(seed, 'IN')
    for k, v in enc.items():
        sd['real_encoder.' + k] = v
        sd['render_encoder.' + k] = v
    for k, v in make_encoder_weights(seed, 'BN').items():
        sd['context.' + k] = v
    for k, v in make_decoder_weights(seed, num_class).items():
        sd['decoder.' + k] = v
    return sd
