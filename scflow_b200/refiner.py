"""SCFlowRefiner registered under the reference's name (models/refiner/scflow_refiner.py, base_refiner.py).

Scope: ``extract_feat``, ``get_pose`` and ``forward_single_pass`` - the iteration owner
of the hot path.  The renderer (pytorch3d), the losses and the OpenCV pose re-mapping are outside the replaced
path; the refiner accepts their config keys so the reference's config dicts build unchanged, takes an injected
``renderer`` callable, and raises a clear error where an out-of-scope component would be needed.
"""
import os
from typing import Callable, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from .builder import REFINERS, build_decoder, build_encoder, build_loss
from .cnn import BaseModule


@REFINERS.register_module()
class SCFlowRefiner(BaseModule):
    def __init__(self, seperate_encoder: bool, cxt_channels: int, h_channels: int, cxt_encoder: dict, encoder: dict,
                 decoder: dict, renderer: Optional[Union[dict, Callable]] = None, pose_loss_cfg: Optional[dict] = None,
                 flow_loss_cfg: Optional[dict] = None, mask_loss_cfg: Optional[dict] = None, max_flow: float = 400,
                 render_augmentations: list = None, filter_invalid_flow: bool = True, freeze_encoder: bool = False,
                 freeze_bn: bool = False, train_cfg: Optional[dict] = None, test_cfg: Optional[dict] = None,
                 init_cfg: Optional[Union[list, dict]] = None) -> None:
        super().__init__(init_cfg)
        self.seperate_encoder = seperate_encoder
        if seperate_encoder:
            self.render_encoder = build_encoder(encoder)
            self.real_encoder = build_encoder(encoder)
        else:   # one module under two names, as base_refiner.py:37-39 (state dict carries both prefixes)
            enc = build_encoder(encoder)
            self.render_encoder = enc
            self.real_encoder = enc
        self.decoder = build_decoder(decoder)
        self.context = build_encoder(cxt_encoder)
        self.renderer = renderer if callable(renderer) else None
        self.renderer_cfg = renderer if isinstance(renderer, dict) else None
        self.h_channels, self.cxt_channels = h_channels, cxt_channels
        assert self.h_channels == self.decoder.h_channels
        assert self.cxt_channels == self.decoder.cxt_channels
        assert self.h_channels + self.cxt_channels == self.context.out_channels
        self.max_flow = max_flow
        self.train_cfg = train_cfg or {}
        self.test_cfg = test_cfg or {}
        self.loss_cfgs = dict(pose=pose_loss_cfg, flow=flow_loss_cfg, mask=mask_loss_cfg)
        # built lazily (the point-matching loss reads the model point clouds from mesh_path)
        self._loss_funcs = None
        self.native_feature_path = True   # inference: encoders write the loop's inputs directly (no NCHW round trip)
        self._zero_flow = None
        self.overlap_encoders = os.environ.get('SCFLOW_ENC_OVERLAP', '1') != '0'
        self._side_stream = None
        self._graphs = {}
        self._last_key = None
        self.filter_invalid_flow = filter_invalid_flow
        self.test_by_flow = self.test_cfg.get('by_flow', False)
        self.test_iter_num = self.test_cfg.get('iters') if 'iters' in self.test_cfg else self.decoder.iters
        if freeze_bn:
            self.freeze_bn()
        if freeze_encoder:
            self.freeze_encoder()

    def freeze_encoder(self):
        for enc in (self.real_encoder, self.render_encoder):
            for m in enc.modules():
                m.eval()
                for p in m.parameters(recurse=False):
                    p.requires_grad = False

    def freeze_bn(self) -> None:
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def extract_feat(self, render_images, real_images) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        """scflow_refiner.py:88-110."""
        if self.real_encoder is self.render_encoder and not self.training and real_images.shape == render_images.shape:
            # shared weights (seperate_encoder=False): one pass over both image sets (InstanceNorm is per sample)
            both = self.real_encoder(torch.cat([real_images, render_images], dim=0))
            real_feat, render_feat = both[:real_images.shape[0]], both[real_images.shape[0]:]
        else:
            real_feat = self.real_encoder(real_images)
            render_feat = self.render_encoder(render_images)
        cxt_feat = self.context(render_images)
        h_feat, cxt_feat = torch.split(cxt_feat, [self.h_channels, self.cxt_channels], dim=1)
        return render_feat, real_feat, torch.tanh(h_feat), torch.relu(cxt_feat)

    def _native_plan(self, render_images, real_images, depth):
        """(b, h, w) when the fused inference path applies to these inputs, else None (the generic path is used then)."""
        enc, ctx, dec = self.real_encoder, self.context, self.decoder
        if (self.training or torch.is_grad_enabled() or enc is not self.render_encoder or real_images.shape != render_images.shape
                or not (enc._native_ok(real_images) and ctx._native_ok(render_images)) or not hasattr(dec, 'native_slots')):
            return None
        b, _, h, w = real_images.shape
        if tuple(depth.shape) != (b, h, w) or dec.native_slots(b, h, w, real_images.device) is None:
            return None
        return b, h, w

    def _native_eager(self, images2, ref_rotation, ref_translation, depth, internel_k, label, init_flow):
        """Enqueues the whole step on the current stream: the three encoder passes write the loop's inputs (pixel-major
        split-bf16 feature maps, tanh'ed hidden state, relu'ed context) straight into the decoder workspace - no NCHW tensors
        between encoder and loop - then the loop runs.  ``images2`` = [real ; rendered] images stacked along the batch."""
        from . import _lib
        enc, ctx, dec = self.real_encoder, self.context, self.decoder
        b2, _, h, w = images2.shape
        b = b2 // 2
        _, ws, (s_feat, s_h, s_hf32, s_cxt) = dec.native_slots(b, h, w, images2.device)
        p8 = (h // 8) * (w // 8)
        base = ws.data_ptr()
        ex = _lib.EncoderOut()
        ex.split = 256
        ex.hl0, ex.plane0, ex.stride0, ex.act0 = base + s_feat, 2 * b * p8 * 256, 256, _lib.ACT['none']
        ex2 = _lib.EncoderOut()
        ex2.split = 128
        ex2.hl0, ex2.plane0, ex2.stride0, ex2.act0 = base + s_h, b * p8 * 128, 128, _lib.ACT['tanh']
        ex2.f32_0, ex2.f32_stride0 = base + s_hf32, 128
        ex2.hl1, ex2.plane1, ex2.stride1, ex2.act1 = base + s_cxt, b * p8 * 128, 128, _lib.ACT['relu']
        render_images = images2[b:]
        if self.overlap_encoders:
            # the context encoder is independent of the feature encoder: run it on a side stream so that its tensor-core
            # convolutions fill the feature encoder's HBM-bound InstanceNorm passes (and vice versa)
            cur = torch.cuda.current_stream(images2.device)
            if self._side_stream is None or self._side_stream.device != images2.device:
                self._side_stream = torch.cuda.Stream(device=images2.device)
            side = self._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                ctx._forward_native(render_images, ex2)
            enc._forward_native(images2, ex)            # samples [0,b) real, [b,2b) rendered
            cur.wait_stream(side)
        else:
            enc._forward_native(images2, ex)
            ctx._forward_native(render_images, ex2)
        return dec.forward_prepared(ref_rotation, ref_translation, depth, internel_k, label, init_flow, 0., allow_graph=False)

    def _get_pose_native(self, render_images, real_images, ref_rotation, ref_translation, depth, internel_k, label, init_flow):
        """Inference fast path (see ``_native_eager``).  With ``decoder.use_cuda_graph`` the WHOLE step - encoders and loop -
        is captured once per shape and replayed; the outputs are then views of static buffers that the next call overwrites.
        Returns None when the configuration does not allow the fused path."""
        plan = self._native_plan(render_images, real_images, depth)
        if plan is None:
            return None
        b, h, w = plan
        dev = real_images.device
        dec = self.decoder
        ins = dict(ref_rotation=ref_rotation, ref_translation=ref_translation, depth=depth, internel_k=internel_k, init_flow=init_flow)
        for k, t in ins.items():
            if not t.is_cuda:
                raise RuntimeError(f'SCFlowRefiner: {k} must be a CUDA tensor (scflow_b200 has no CPU path)')
            ins[k] = t.detach().contiguous().float()
        if label is None:
            label = torch.zeros(1, dtype=torch.int64, device=dev)
        ins['label'] = label.detach().to(torch.int64).reshape(-1)[:1].contiguous()       # only label[0] is read (pose_head.py:209-210)
        real_images, render_images = real_images.detach().float(), render_images.detach().float()
        with torch.cuda.device(dev):
            if not getattr(dec, 'use_cuda_graph', False) or torch.cuda.is_current_stream_capturing():
                return self._native_eager(torch.cat([real_images, render_images], dim=0), **ins)
            _, ws, _ = dec.native_slots(b, h, w, dev)
            key = (b, h, w, int(dec.iters), str(dev), ws.data_ptr(), int(dec.identity_pose_head), self.overlap_encoders)
            entry = self._graphs.get(key)
            if entry is None:
                # Cache miss: this call's result comes from ONE eager run (also the warm-up that loads modules and sets function
                # attributes); the capture itself enqueues nothing.  A shape is captured only when it repeats - test-time batches
                # with a different patch count every call would otherwise pay a capture per call for nothing.
                static = {k: v.clone() for k, v in ins.items()}
                static['images2'] = torch.cat([real_images, render_images], dim=0)
                outs = self._native_eager(**static)
                seen, self._last_key = self._last_key == key, key
                if not seen:
                    return outs
                torch.cuda.current_stream(dev).synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    replay_outs = self._native_eager(**static)
                self._graphs = {key: (graph, static, replay_outs)}      # keep only the latest shape
                return outs                                             # (the eager run's; replay_outs are filled by replays)
            graph, static, outs = entry
            static['images2'][:b].copy_(real_images, non_blocking=True)
            static['images2'][b:].copy_(render_images, non_blocking=True)
            for k, v in ins.items():
                static[k].copy_(v, non_blocking=True)
            graph.replay()
            return outs

    def get_pose(self, render_images, real_images, ref_rotation, ref_translation, depth, internel_k, label,
                 init_flow=None, pose_head_label=None):
        """scflow_refiner.py:112-142: encoders -> decoder loop; returns the decoder's 7 lists.  ``pose_head_label`` (not in the
        reference): class selector of the pose head when this batch is a shard of a larger one (the reference's head uses the
        GLOBAL ``label[0]`` for every row, pose_head.py:209-210; see scflow_b200/dist.py)."""
        if pose_head_label is not None:
            label = pose_head_label.reshape(-1)[:1]
        if self.native_feature_path:
            if init_flow is None:
                n, _, h, w = real_images.shape
                key = (n, h, w, str(real_images.device))
                if self._zero_flow is None or self._zero_flow[0] != key:
                    self._zero_flow = (key, torch.zeros((n, 2, h, w), device=real_images.device, dtype=torch.float32))
                zero = self._zero_flow[1]
            outs = self._get_pose_native(render_images, real_images, ref_rotation, ref_translation, depth, internel_k, label,
                                         zero if init_flow is None else init_flow)
            if outs is not None:
                return outs
        feat_render, feat_real, h_feat, cxt_feat = self.extract_feat(render_images, real_images)
        if init_flow is None:
            n, _, h, w = real_images.shape
            init_flow = feat_render.new_zeros((n, 2, h, w), dtype=torch.float32)
        return self.decoder(feat_render, feat_real, h_feat, cxt_feat, ref_rotation, ref_translation, depth, internel_k,
                            init_flow=init_flow, label=label, invalid_flow_num=0.)

    def forward_single_pass(self, data: Dict, data_batch: Optional[Dict] = None, return_loss: bool = False):
        """scflow_refiner.py:146-179.  The re-mapping to the original image resolution is the identity for the shipped
        'adapt_intrinsic' pipelines; the cv2.solvePnP modes raise (see below)."""
        labels = data['labels']
        per_img_patch_num = data.get('per_img_patch_num', [len(labels)])
        iters = self.decoder.iters
        self.decoder.iters = self.test_iter_num
        try:
            outs = self.get_pose(data['rendered_images'], data['real_images'], data['ref_rotations'], data['ref_translations'],
                                 data['rendered_depths'], data['internel_k'], labels, pose_head_label=data.get('pose_head_label'))
        finally:
            self.decoder.iters = iters
        seq_rotations, seq_translations = outs[2], outs[3]
        # remap_pose_to_origin_resoluaion (models/utils/pose.py:264-309): with the shipped pipelines (RemapPose(keep_intrinsic=
        # False) without dst_k => 'adapt_intrinsic', datasets/pipelines/geometry_transform.py:35-45) the poses are returned
        # unchanged (:276-279); the other two modes re-solve the pose with cv2.solvePnP on the host, which is not built
        for meta in (data_batch or {}).get('img_metas', []):
            mode = meta.get('geometry_transform_mode', 'adapt_intrinsic') if isinstance(meta, dict) else 'adapt_intrinsic'
            if mode != 'adapt_intrinsic':
                raise NotImplementedError(f"geometry_transform_mode '{mode}' needs the OpenCV PnP re-mapping (pose.py:280-305), "
                                          "which is outside this package; the shipped configs use 'adapt_intrinsic'")
        return dict(
            rotations=torch.split(seq_rotations[-1], per_img_patch_num),
            translations=torch.split(seq_translations[-1], per_img_patch_num),
            labels=torch.split(labels, per_img_patch_num),
            scores=torch.split(torch.ones_like(labels, dtype=torch.float32), per_img_patch_num),
        )

    def forward(self, data_batch, return_loss=False):
        """base_refiner.py:338-343. Needs an injected ``renderer`` callable producing rendered images/depths for the
        reference poses; the pytorch3d rasteriser itself is not part of this package."""
        if 'rendered_images' in data_batch:       # already formatted (bench / tests / external renderer)
            return self.forward_single_pass(data_batch)
        if self.renderer is None:
            raise NotImplementedError('SCFlowRefiner.forward(data_batch) needs a renderer: assign a callable to '
                                      '`model.renderer` or pass a pre-formatted dict with `rendered_images` / `rendered_depths`.')
        return self.forward_single_pass(self.format_data_test(data_batch), data_batch)

    def set_renderer(self, renderer: Callable):
        """Plug in the renderer (reference interface: ``renderer(rotations, translations, internel_k, labels)`` returning
        ``{'images': [N,H,W,4], 'fragments': obj with .zbuf [N,H,W,K]}``, models/utils/renderer.py)."""
        self.renderer = renderer

    def format_data_train_sup(self, data_batch: Dict) -> Dict:
        """base_refiner.py:136-191 (render augmentations: none, as in every shipped config)."""
        from . import formatting
        return formatting.format_data_train_sup(data_batch, self.renderer, getattr(self, 'render_augmentation', None))

    def format_data_test(self, data_batch: Dict) -> Dict:
        """base_refiner.py:79-133: flatten the per-image patch lists, render the reference poses, format the render
        (one CUDA pass, scflow_b200/formatting.py)."""
        from . import formatting
        return formatting.format_data_test(data_batch, self.renderer)

    def train_step(self, data_batch, optimizer=None, **kwargs):
        """base_refiner.py:325-336: the loss WITH its autograd graph plus the logging dict; backward / all-reduce / optimizer
        are the caller's (mmcv's OptimizerHook in the reference; ``scflow_b200.training.Trainer`` here)."""
        loss, log_imgs, log_vars, _, _ = self.loss(data_batch)
        n = len(data_batch['img_metas']) if 'img_metas' in data_batch else int(data_batch['labels'].shape[0])
        return dict(loss=loss, log_vars=log_vars, log_imgs=log_imgs, num_samples=n)

    def loss_functions(self):
        """(pose, flow, mask) SequenceLoss modules built from the reference's config keys (scflow_refiner.py:61-63)."""
        if self._loss_funcs is None:
            if any(v is None for v in self.loss_cfgs.values()):
                raise RuntimeError('SCFlowRefiner: pose_loss_cfg / flow_loss_cfg / mask_loss_cfg were not given')
            from . import loss as _loss  # noqa: F401  (registers the loss classes)
            self._loss_funcs = tuple(build_loss(self.loss_cfgs[k]) for k in ('pose', 'flow', 'mask'))
        return self._loss_funcs

    def loss(self, data: Dict):
        """scflow_refiner.py:184-258 for a collated training batch (formatted through ``format_data_train_sup`` when a renderer is
        plugged in) or an already formatted one: keys ``gt_rotations, gt_translations, ref_rotations, ref_translations,
        real_images, rendered_images, rendered_depths, rendered_masks, gt_masks, internel_k, labels``.

        * under ``torch.no_grad()``: forward value only, three fused kernels (``scf_refiner_loss``); returns
          ``(loss, log_vars, seq_rotations, seq_translations)``;
        * with autograd enabled (training): the differentiable graph (``scflow_b200/training.py``); returns the reference's
          ``(loss, log_imgs, log_vars, seq_rotations, seq_translations)`` (``log_imgs`` is empty: visualisation is outside this
          package)."""
        from . import loss as L
        from . import ops
        if 'rendered_images' not in data:         # a collated training batch: format it first (scflow_refiner.py:186)
            data = self.format_data_train_sup(data)
        pose_f, flow_f, mask_f = self.loss_functions()
        train = torch.is_grad_enabled()
        outs = self.get_pose(data['rendered_images'], data['real_images'], data['ref_rotations'], data['ref_translations'],
                             data['rendered_depths'], data['internel_k'], data['labels'], pose_head_label=data.get('pose_head_label'))
        flow_from_pose, flow_from_pred, seq_rot, seq_trs, seq_masks = outs[0], outs[1], outs[2], outs[3], outs[4]
        with torch.no_grad():
            depth = data['rendered_depths'].float().contiguous()
            k = data['internel_k'].float().contiguous()
            # GT flow (models/utils/pose.py:92-121): lift with the reference pose, project with the ground-truth pose
            pts4 = ops.unproject(depth, k, data['ref_rotations'].float().contiguous(), data['ref_translations'].float().contiguous())
            gt_flow = ops.reproject(pts4, k, data['gt_rotations'].float().contiguous(), data['gt_translations'].float().contiguous(),
                                    float(self.max_flow))
            if self.filter_invalid_flow:
                gt_flow = L.filter_flow_by_mask(gt_flow, data['gt_masks'].float().contiguous(), float(self.max_flow))
        if train:
            from . import training
            loss, terms = training.refiner_loss_train(outs, gt_flow, data['rendered_masks'].float(), data['gt_rotations'].float(),
                                                      data['gt_translations'].float(), data['labels'], pose_f, flow_f, mask_f,
                                                      float(self.max_flow))
            iters = len(flow_from_pred)
            vals = torch.cat([terms['loss'].reshape(1), terms['loss_pose'].reshape(1), terms['loss_flow'].reshape(1),
                              terms['loss_mask'].reshape(1), terms['seq_pose'], terms['seq_flow'], terms['seq_mask']]).tolist()
        else:
            out, iters = L.refiner_loss(flow_from_pred, seq_masks, seq_rot, seq_trs, gt_flow, data['rendered_masks'], data['gt_rotations'],
                                        data['gt_translations'], data['labels'], pose_f, flow_f, mask_f)
            loss = out[0]
            vals = out.tolist()        # one device->host read for the whole log (the reference does 3*iters + 4 .item() calls)
        log_vars = {}
        for name in ('add', 'rot', 'trans'):
            if f'init_{name}_error_mean' in data and name == 'add':
                log_vars['init_add_mean'] = float(data['init_add_error_mean'])
                log_vars['init_add_std'] = float(data['init_add_error_std'])
        for i in range(iters):
            log_vars[f'seq_{i}_pose_loss'] = vals[4 + i]
            log_vars[f'seq_{i}_flow_loss'] = vals[4 + iters + i]
            log_vars[f'seq_{i}_mask_loss'] = vals[4 + 2 * iters + i]
        log_vars.update(loss_mask=vals[3], loss_flow=vals[2], loss_pose=vals[1], loss=vals[0])
        if train:
            return loss, {}, log_vars, seq_rot, seq_trs
        return loss, log_vars, seq_rot, seq_trs
