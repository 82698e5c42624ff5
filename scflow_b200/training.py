"""Training step of the refinement path (BASELINE config 5): differentiable forward + loss,
backward, gradient all-reduce over NCCL, fused global-norm clip + AdamW.

Reference: ``SCFlowRefiner.loss`` (models/refiner/scflow_refiner.py:184-258), ``BaseRefiner.train_step``
(base_refiner.py:325-336), DDP (train.py:127-135, ``broadcast_buffers=False``), AdamW + ``grad_clip(max_norm=10)``
(configs/refine_models/scflow.py:117-125).

What runs where
* **forward**: the gradient-free parts are the path's own CUDA kernels - correlation volume + pyramid (one tcgen05 kernel),
  pyramid lookup, 2D-3D lift, pose-induced flow, GT flow and its mask filter.  The lookup coordinates, the flow fed to the next
  iteration and the running pose are detached by the reference (scflow_decoder.py:192-195, 232-233; config scflow.py:57-60), so
  the gradient reaches the encoders only through the sampled correlation VALUES and flows through time only via the GRU state.
* **backward**: composed from torch operators: the two custom autograd Functions below supply the adjoints
  of the native forward kernels (pyramid: un-pool + two batched GEMMs; lookup: bilinear scatter into the level gradients), the
  convolutions / norms / activations in between are torch's own differentiable ops on the modules' parameters.
* **collective**: the only one the path has - the gradient all-reduce (8,178,206 fp32 = 32.7 MB).  Gradients live in ONE flat
  buffer; buckets of it are all-reduced (NCCL, SUM) on a communication stream as soon as backward has produced them
  (post-accumulate-grad hooks, buckets in reverse parameter order as DDP does), overlapping the rest of backward.
* **optimizer**: ``scf_clip_adamw`` - global gradient norm (deterministic two-stage reduction), clip coefficient, 1/world
  averaging and the AdamW update in two launches over the flat buffers.
"""
import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, ops


# ----------------------------------------------------------------------------------------------------------------------
# custom autograd Functions: native forward, torch-composed backward
# ----------------------------------------------------------------------------------------------------------------------
class _CorrPyramidFn(torch.autograd.Function):
    """CorrelationPyramid.forward (raft_decoder.py:35-58): levels[l] = [B*P, 1, H_l, W_l]."""

    @staticmethod
    def forward(ctx, feat_render, feat_real, num_levels):
        ctx.save_for_backward(feat_render, feat_real)
        ctx.num_levels = num_levels
        with torch.no_grad():
            levels = ops.corr_build(feat_render.detach().contiguous().float(), feat_real.detach().contiguous().float(), num_levels,
                                    precision=ops.PRECISION_BF16X3)
        return tuple(levels)

    @staticmethod
    def backward(ctx, *grads):
        f1, f2 = ctx.saved_tensors
        b, c, h, w = f1.shape
        p = h * w
        # adjoint of the successive floor 2x2 mean pools, coarsest first: G_{l-1} += unpool(G_l) / 4
        total = None
        hs = [(h >> l, w >> l) for l in range(ctx.num_levels)]
        for l in range(ctx.num_levels - 1, -1, -1):
            g = grads[l]
            if total is not None:
                up = total.repeat_interleave(2, dim=-2).repeat_interleave(2, dim=-1) * 0.25
                hl, wl = hs[l]
                if up.shape[-2] != hl or up.shape[-1] != wl:          # odd sizes: the dropped last row / column gets no gradient
                    up = F.pad(up, (0, wl - up.shape[-1], 0, hl - up.shape[-2]))
                total = up if g is None else g + up
            elif g is not None:
                total = g
        if total is None:
            return None, None, None
        gmat = total.reshape(b, p, p) * (1.0 / math.sqrt(c))          # [b, query q, key k]
        d1 = torch.bmm(f2.reshape(b, c, p), gmat.transpose(1, 2))     # dL/df_render[b, c, q] = sum_k G[q, k] f_real[c, k]
        d2 = torch.bmm(f1.reshape(b, c, p), gmat)                     # dL/df_real[b, c, k]   = sum_q G[q, k] f_render[c, q]
        return d1.reshape(b, c, h, w), d2.reshape(b, c, h, w), None


def lookup_torch(levels: Sequence[torch.Tensor], flow8: torch.Tensor, radius: int) -> torch.Tensor:
    """CorrLookup.forward in torch operators (corr_lookup.py:102-136): used for the adjoint of the native lookup."""
    b, _, h, w = flow8.shape
    dev = flow8.device
    k = 2 * radius + 1
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing='ij')
    base = (torch.stack([xs, ys], dim=0).float()[None] + flow8).permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
    d = torch.linspace(-radius, radius, k, device=dev)
    first, second = torch.meshgrid(d, d, indexing='ij')
    delta = torch.stack([first, second], dim=-1).view(1, k, k, 2)          # x-major window (corr_lookup.py:118-128)
    outs = []
    for lvl, vol in enumerate(levels):
        hl, wl = vol.shape[-2:]
        coords = base / 2 ** lvl + delta
        gx = coords[..., 0] * 2. / max(wl - 1, 1) - 1.
        gy = coords[..., 1] * 2. / max(hl - 1, 1) - 1.
        smp = F.grid_sample(vol, torch.stack([gx, gy], dim=-1), mode='bilinear', padding_mode='zeros', align_corners=True)
        outs.append(smp.view(b, h, w, k * k))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous()


class _CorrLookupFn(torch.autograd.Function):
    """CorrLookup.forward: native gather forward; backward scatters the output gradient into the level gradients with the same
    bilinear weights (the coordinates carry no gradient: the flow is detached by the reference, scflow_decoder.py:192-193)."""

    @staticmethod
    def forward(ctx, flow8, radius, *levels):
        ctx.radius = radius
        ctx.save_for_backward(flow8, *levels)
        with torch.no_grad():
            out = ops.corr_lookup_nhwc([l.detach() for l in levels], flow8.detach().permute(0, 2, 3, 1).contiguous(), radius)
        return ops.nhwc_to_nchw(out, channels=len(levels) * (2 * radius + 1) ** 2)

    @staticmethod
    def backward(ctx, grad_out):
        flow8, *levels = ctx.saved_tensors
        with torch.enable_grad():
            leaves = [l.detach().requires_grad_(True) for l in levels]
            out = lookup_torch(leaves, flow8.detach(), ctx.radius)
            grads = torch.autograd.grad(out, leaves, grad_out.contiguous(), allow_unused=True)
        return (None, None) + tuple(grads)


# ----------------------------------------------------------------------------------------------------------------------
# differentiable decoder loop (scflow_decoder.py:150-251) on the modules' own parameters
# ----------------------------------------------------------------------------------------------------------------------
_ACTS = {'none': lambda x: x, 'relu': F.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}


# cuDNN's tensor-core convolutions are NHWC kernels: with NCHW activations every convolution (forward, dgrad, wgrad) is wrapped in
# layout-conversion launches (~1500 per step at config 5).  Activations therefore enter the convolutions channels-last and stay so
# downstream when SCFLOW_TRAIN_CHANNELS_LAST=1 (measured 122.6 -> 113.7 ms per config-5 step).  Off by default: the different cuDNN
# kernels change the summation order, and the gradient-parity test against the CPU oracle (tests/test_train.py) is pinned to NCHW.
_CHANNELS_LAST = os.environ.get('SCFLOW_TRAIN_CHANNELS_LAST', '0') != '0'


def _cl(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous(memory_format=torch.channels_last) if _CHANNELS_LAST and x.dim() == 4 else x


def _cm(m, x: torch.Tensor) -> torch.Tensor:
    """ConvModule (conv -> optional GroupNorm -> activation) as differentiable torch operators."""
    y = F.conv2d(_cl(x), m.conv.weight, m.conv.bias, m.stride, m.padding)
    if m.with_norm:
        y = F.group_norm(y, m.num_groups, m.gn.weight, m.gn.bias, m.norm_eps)
    return _ACTS[m.act](y)


def ortho6d_to_matrix(o6: torch.Tensor) -> torch.Tensor:
    """get_rotation_matrix_from_ortho6d (pose.py:153-169): columns (x, y, z)."""
    x = F.normalize(o6[:, 0:3], dim=-1)
    z = F.normalize(torch.cross(x, o6[:, 3:6], dim=-1), dim=-1)
    y = torch.cross(z, x, dim=-1)
    return torch.stack([x, y, z], dim=-1)


def pose_from_delta(d_rot, d_trs, rot, trs, detach_depth_for_xy: bool = False, weight: float = 10.):
    """get_pose_from_delta_pose (pose.py:124-148), rotation_mode='ortho6d', depth_transform='exp'."""
    r_new = torch.bmm(ortho6d_to_matrix(d_rot), rot)
    tz = trs[:, 2] / torch.exp(d_trs[:, 2])
    tz_xy = tz.detach() if detach_depth_for_xy else tz
    tx = tz_xy * (d_trs[:, 0] / weight + trs[:, 0] / trs[:, 2])
    ty = tz_xy * (d_trs[:, 1] / weight + trs[:, 1] / trs[:, 2])
    return r_new, torch.stack([tx, ty, tz], dim=-1)


def decoder_forward_train(dec, feat_render, feat_real, h_feat, cxt_feat, ref_rotation, ref_translation, depth, internel_k, label,
                          init_flow, invalid_flow_num: float = 0.):
    """SCFlowDecoder.forward with an autograd graph (training).  Returns the reference's 7 lists."""
    if dec.mask_corr or dec.mask_flow:
        raise NotImplementedError('training implements the shipped configuration (mask_corr = mask_flow = False)')
    scale = 2 ** (dec.num_levels - 1)
    b, hh, ww = depth.shape
    h8, w8 = hh // scale, ww // scale
    debug_torch = os.environ.get('SCFLOW_TRAIN_DEBUG_TORCH', '0') != '0'      # parity debugging only: pyramid / lookup in torch ops
    if debug_torch:
        bb, cc, h8_, w8_ = feat_render.shape
        vol = torch.bmm(feat_render.reshape(bb, cc, -1).transpose(1, 2), feat_real.reshape(bb, cc, -1)) / math.sqrt(cc)
        levels = [vol.reshape(bb * h8_ * w8_, 1, h8_, w8_)]
        for _ in range(dec.num_levels - 1):
            levels.append(F.avg_pool2d(levels[-1], 2, 2))
    else:
        levels = _CorrPyramidFn.apply(feat_render, feat_real, dec.num_levels)
    with torch.no_grad():
        k = internel_k.detach().float().contiguous()
        pts4 = ops.unproject(depth.detach().float().contiguous(), k, ref_rotation.detach().float().contiguous(),
                             ref_translation.detach().float().contiguous())
    rot, trs = ref_rotation, ref_translation
    flow = init_flow
    enc, gru, head = dec.encoder, dec.gru, dec.pose_pred
    outs = ([], [], [], [], [], [], [])
    for _ in range(int(dec.iters)):
        if dec.detach_flow:
            flow = flow.detach()
        flow8 = (1.0 / scale) * F.interpolate(flow, size=(h8, w8), mode='bilinear', align_corners=True)
        corr = lookup_torch(levels, flow8, dec.radius) if debug_torch else _CorrLookupFn.apply(flow8, dec.radius, *levels)
        # motion encoder (raft_decoder.py:152-166)
        c = _cm(enc.corr_net[1], _cm(enc.corr_net[0], corr))
        f = _cm(enc.flow_net[1], _cm(enc.flow_net[0], flow8))
        motion = torch.cat([_cm(enc.out_net[0], torch.cat([c, f], dim=1)), flow8], dim=1)
        # SepConvGRU (raft_decoder.py:235-253)
        x = torch.cat([cxt_feat, motion], dim=1)
        for cz, cr, cq in zip(gru.conv_z, gru.conv_r, gru.conv_q):
            hx = torch.cat([h_feat, x], dim=1)
            z, r = _cm(cz, hx), _cm(cr, hx)
            q = _cm(cq, torch.cat([r * h_feat, x], dim=1))
            h_feat = (1 - z) * h_feat + z * q
        d_flow = _xhead(dec.flow_pred, h_feat)
        mask = torch.sigmoid(_xhead(dec.mask_pred, h_feat))
        df = _cm(dec.delta_flow_encoder[1], _cm(dec.delta_flow_encoder[0], d_flow))
        mf = _cm(dec.mask_encoder[1], _cm(dec.mask_encoder[0], mask))
        d_rot, d_trs = _pose_head(head, torch.cat([h_feat, df, mf], dim=1), label)
        flow_pred = scale * F.interpolate(flow8 + d_flow, size=(hh, ww), mode='bilinear', align_corners=True)
        mask_up = F.interpolate(mask, size=(hh, ww), mode='bilinear', align_corners=True)
        rot, trs = pose_from_delta(d_rot, d_trs, rot.detach() if dec.detach_pose else rot, trs.detach() if dec.detach_pose else trs,
                                   dec.detach_depth_for_xy)
        with torch.no_grad():         # pose-induced flow: feeds the next iteration detached and is not part of the loss
            flow = ops.reproject(pts4, k, rot.detach().float().contiguous(), trs.detach().float().contiguous(), float(invalid_flow_num))
        if not dec.detach_flow:
            raise NotImplementedError('detach_flow=False (gradient through the pose-induced flow) is not implemented')
        for lst, v in zip(outs, (flow, flow_pred, rot, trs, mask_up, d_rot, d_trs)):
            lst.append(v)
    return outs


def _xhead(xh, h):
    """XHead.forward (raft_decoder.py:292-294): hidden ConvModules, then the plain predict conv."""
    y = h
    for layer in xh.layers:
        y = _cm(layer, y)
    p = xh.predict_layer
    return F.conv2d(_cl(y), p.weight, p.bias, p.stride, p.padding)


def _pose_head(head, x, label):
    """Multi/SingleClassPoseHead.forward (pose_head.py:201-211), incl. the label[0] class selection."""
    y = x
    for layer in head.conv_layers:
        y = _cm(layer, y)
    y = y.flatten(1)
    for fc in head.fc_layers:
        y = F.relu(F.linear(y, fc[0].weight, fc[0].bias))
    rot = F.linear(y, head.rotation_pred.weight, head.rotation_pred.bias)
    trs = F.linear(y, head.translation_pred.weight, head.translation_pred.bias)
    nc = getattr(head, 'num_class', None)
    if nc:
        b = x.shape[0]
        cls = int(label.reshape(-1)[0])                      # index_select(dim=1, label)[:, 0]: label[0] for every row
        rot = rot.view(b, nc, -1)[:, cls]
        trs = trs.view(b, nc, 3)[:, cls]
    return rot, trs


# ----------------------------------------------------------------------------------------------------------------------
# differentiable loss (scflow_refiner.py:204-258 with the shipped loss configuration)
# ----------------------------------------------------------------------------------------------------------------------
def point_matching_loss(pred_r, pred_t, gt_r, gt_t, labels, packed, weight: float):
    """DisentanglePointMatchingLoss.forward (point_matching_loss.py:160-218; loss_type='l1', disentangle_z=True), batched
    over the samples: ``packed`` = (points [C, maxP, 3], num_points [C], symmetric [C], diameter [C])."""
    pts_all, npts_all, sym_all, diam_all = packed
    pts = pts_all[labels]                                     # [B, maxP, 3]
    n = npts_all[labels].to(pred_r.dtype)                     # [B]
    valid = (torch.arange(pts.shape[1], device=pts.device)[None] < npts_all[labels][:, None]).to(pred_r.dtype)      # [B, maxP]
    gt_rot = torch.bmm(pts, gt_r.transpose(1, 2))             # R p for every point
    gt_rt = gt_rot + gt_t[:, None]
    pred_rot = torch.bmm(pts, pred_r.transpose(1, 2)) + gt_t[:, None]
    sym = sym_all[labels].bool()
    if bool(sym.any()):
        # symmetric objects: every ground-truth point is matched with its nearest predicted point (pytorch3d knn_points, K=1)
        idx_s = torch.nonzero(sym).flatten()
        with torch.no_grad():
            d = torch.cdist(gt_rt[idx_s], pred_rot[idx_s])                                     # [S, maxP(gt), maxP(pred)]
            d = d.masked_fill(valid[idx_s][:, None, :] == 0, float('inf'))
            nn_idx = d.argmin(dim=-1)
        matched = torch.gather(pred_rot[idx_s], 1, nn_idx[..., None].expand(-1, -1, 3))
        pred_rot = pred_rot.index_copy(0, idx_s, matched)
    l_rot = ((pred_rot - gt_rt).abs().sum(-1) * valid).sum(-1) / n
    tz = torch.cat([gt_t[:, :2], pred_t[:, 2:3]], dim=1)
    txy = torch.cat([pred_t[:, :2], gt_t[:, 2:3]], dim=1)
    l_depth = (((gt_rot + tz[:, None]) - gt_rt).abs().sum(-1) * valid).sum(-1) / n
    l_xy = (((gt_rot + txy[:, None]) - gt_rt).abs().sum(-1) * valid).sum(-1) / n
    per_sample = (l_depth + l_xy + l_rot) / diam_all[labels]
    return weight * per_sample.sum() / pred_r.shape[0]


def refiner_loss_train(outs, gt_flow, rendered_masks, gt_rotations, gt_translations, labels, pose_f, flow_f, mask_f, max_flow: float):
    """(loss, dict of detached per-term tensors).  ``gt_flow`` is already filtered (no gradient)."""
    _, flow_pred, seq_rot, seq_trs, seq_mask = outs[0], outs[1], outs[2], outs[3], outs[4]
    n = len(flow_pred)
    packed = pose_f.loss_func.packed(gt_flow.device)
    mag = torch.sum(gt_flow ** 2, dim=1).sqrt()
    v = ((rendered_masks >= 0.5) & (mag < flow_f.loss_func.max_flow)).to(gt_flow.dtype)
    occ = (torch.sum(gt_flow, dim=1) < max_flow).to(torch.float32)
    pose_l, flow_l, mask_l = [], [], []
    for i in range(n):
        pose_l.append(point_matching_loss(seq_rot[i], seq_trs[i], gt_rotations, gt_translations, labels, packed, pose_f.loss_func.loss_weight))
        flow_l.append(flow_f.loss_func.loss_weight * (v[:, None] * (flow_pred[i] - gt_flow).abs()).sum() / (v.sum() + flow_f.loss_func.eps))
        mask_l.append(mask_f.loss_func.loss_weight * torch.mean(torch.abs(seq_mask[i].squeeze(1) - occ)))

    def seq(ls, gamma):
        return sum(gamma ** (n - i - 1) * l for i, l in enumerate(ls))
    loss_pose, loss_flow, loss_mask = seq(pose_l, pose_f.gamma), seq(flow_l, flow_f.gamma), seq(mask_l, mask_f.gamma)
    loss = loss_pose + loss_flow + loss_mask
    terms = dict(loss=loss.detach(), loss_pose=loss_pose.detach(), loss_flow=loss_flow.detach(), loss_mask=loss_mask.detach(),
                 seq_pose=torch.stack([l.detach() for l in pose_l]), seq_flow=torch.stack([l.detach() for l in flow_l]),
                 seq_mask=torch.stack([l.detach() for l in mask_l]))
    return loss, terms


# ----------------------------------------------------------------------------------------------------------------------
# optimizer step: flat buffers, bucketed all-reduce overlapped with backward, fused clip + AdamW
# ----------------------------------------------------------------------------------------------------------------------
class Trainer:
    """DDP-equivalent data-parallel training of a ``SCFlowRefiner`` (one process per GPU).

    * parameters and gradients are views of two flat fp32 buffers (so the collective and the optimizer see contiguous memory);
    * every parameter's post-accumulate-grad hook counts down its bucket; a complete bucket is all-reduced (SUM) on a
      communication stream while backward continues - the reference's DDP semantics (train.py:127-135), averaging folded into the
      optimizer kernel;
    * ``scf_clip_adamw``: global gradient norm -> clip coefficient (max_norm, as mmcv's OptimizerHook / clip_grad_norm_) ->
      AdamW (torch.optim.AdamW semantics, configs/refine_models/scflow.py:117-124) in two launches."""

    def __init__(self, model, lr: float = 4e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-4,
                 max_norm: Optional[float] = 10., bucket_mb: float = 8., process_group=None, lr_schedule=None):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.lr_schedule = lr_schedule
        seen, params = set(), []
        for p in model.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise RuntimeError('Trainer: the model has no trainable parameters')
        # (flat buffers, buckets and the all-reduce are plumbing and work on any device - the gloo tests use that; the optimizer
        # step is a CUDA kernel and refuses anything else)
        self.params = params
        dev = params[0].device
        self.device = dev
        sizes = [p.numel() for p in params]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + (s + 3) // 4 * 4)             # 16 B aligned slots
        self.total = offs[-1]
        self.flat_p = torch.zeros(self.total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(self.total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(self.total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.total, device=dev, dtype=torch.float32)
        for p, o in zip(params, offs):
            self.flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.flat_p[o:o + p.numel()].view_as(p)
            p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        self.step_count = 0
        self.scratch = torch.zeros(4096, device=dev, dtype=torch.float32)
        self.stats = torch.zeros(4, device=dev, dtype=torch.float32)       # grad norm, clip coefficient
        # buckets: contiguous ranges of the flat gradient, filled in reverse parameter order (the order backward produces them)
        cap = max(1, int(bucket_mb * (1 << 20) / 4))
        self.buckets, cur_hi, cur_n, members = [], self.total, 0, []
        self.bucket_of = {}
        for i in range(len(params) - 1, -1, -1):
            members.append(i)
            cur_n += sizes[i]
            if cur_n >= cap or i == 0:
                self.buckets.append(dict(lo=offs[i], hi=cur_hi, members=list(members), pending=0, work=None))
                for m in members:
                    self.bucket_of[m] = len(self.buckets) - 1
                cur_hi, cur_n, members = offs[i], 0, []
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and dev.type == 'cuda') else None
        self._hooks = []
        if self.world > 1:
            for i, p in enumerate(params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(i)))
            # identical initial parameters on every rank (DDP broadcasts rank 0's)
            dist.broadcast(self.flat_p, src=0, group=process_group)

    def _make_hook(self, index):
        def hook(_):
            bk = self.buckets[self.bucket_of[index]]
            bk['pending'] -= 1
            if bk['pending'] == 0:
                self._launch(bk)
        return hook

    def _launch(self, bk):
        view = self.flat_g[bk['lo']:bk['hi']]
        if self.comm_stream is None:
            bk['work'] = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))       # the bucket's gradients have been written
        with torch.cuda.stream(self.comm_stream):
            bk['work'] = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def zero_grad(self):
        self.flat_g.zero_()
        for p in self.params:            # autograd accumulates into these views
            if p.grad is None or p.grad.data_ptr() < self.flat_g.data_ptr() or p.grad.data_ptr() >= self.flat_g.data_ptr() + 4 * self.total:
                raise RuntimeError('Trainer: a parameter\'s .grad was replaced; gradients must stay views of the flat buffer')
        for bk in self.buckets:
            bk['pending'], bk['work'] = len(bk['members']), None

    def backward(self, loss: torch.Tensor):
        loss.backward()
        if self.world > 1:
            for bk in self.buckets:      # parameters that received no gradient this step never fired their hook
                if bk['work'] is None:
                    self._launch(bk)
            for bk in self.buckets:
                bk['work'].wait()
            if self.comm_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)

    def optimizer_step(self):
        if self.device.type != 'cuda':
            raise RuntimeError('Trainer.optimizer_step: scf_clip_adamw is a CUDA kernel (scflow_b200 has no CPU path)')
        self.step_count += 1
        lr = self.lr_schedule(self.step_count) if self.lr_schedule is not None else self.lr
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.scf_clip_adamw(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                          C.c_longlong(self.total), float(lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                          float(self.weight_decay), int(self.step_count),
                                          float(self.max_norm if self.max_norm is not None else -1.), float(1.0 / self.world),
                                          _lib.ptr(self.scratch), int(self.scratch.numel()), _lib.ptr(self.stats),
                                          _lib.stream_ptr(self.device)), 'scf_clip_adamw')
        return lr

    def train_step(self, data: Dict) -> Dict:
        """One optimisation step on an already formatted batch (the keys ``format_data_train_sup`` produces)."""
        self.model.train()
        self.zero_grad()
        outputs = self.model.train_step(data, None)
        self.backward(outputs['loss'])
        lr = self.optimizer_step()
        outputs['log_vars']['lr'] = lr
        return outputs

    def grad_norm(self) -> float:
        """Global gradient norm of the last step (after averaging over ranks, before clipping) - one device->host read."""
        return float(self.stats[0])

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def one_cycle_lr(max_lr: float, total_steps: int, pct_start: float = 0.05, div_factor: float = 25., final_div_factor: float = 1e4):
    """mmcv's OneCycleLrUpdaterHook with anneal_strategy='linear' as a function of the step (configs/refine_models/scflow.py:126-131)."""
    initial, final = max_lr / div_factor, max_lr / div_factor / final_div_factor
    up = max(1, int(pct_start * total_steps))

    def lr(step: int) -> float:
        s = min(step, total_steps)
        if s <= up:
            return initial + (max_lr - initial) * s / up
        return max_lr + (final - max_lr) * (s - up) / max(1, total_steps - up)
    return lr
