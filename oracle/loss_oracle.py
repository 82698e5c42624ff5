"""CPU oracle of SCFlow's training loss (forward): GT flow, flow filtering, sequence losses, point-matching loss.

TEST INFRASTRUCTURE ONLY (same rule as scflow_oracle.py): imported by tests/ and oracle/make_golden_loss.py, never by
the product package.  Parity pin: every function below was checked against the reference's own code run through
oracle/ref_shim.py (oracle/make_golden_loss.py; fixtures in tests/golden/loss_*.npz).  Third-party arithmetic: the
reference calls pytorch3d.ops.knn_points (fork YangHai-1218/pytorch3d, README.md:29, unpinned, not installable here) for
symmetric objects only; its published semantics (K=1 nearest neighbour, squared L2) are restated in `_nearest`.
All paths are relative to /root/reference.
"""
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

from . import scflow_oracle as O

Tensor = torch.Tensor


def gt_flow_from_poses(ref_rot: Tensor, ref_trs: Tensor, gt_rot: Tensor, gt_trs: Tensor, depth: Tensor, k: Tensor,
                       invalid: float = 400.) -> Tensor:
    """get_flow_from_delta_pose_and_depth (models/utils/pose.py:92-121): lift the rendered depth with the reference pose,
    project with the ground-truth pose; pixels without depth hold `invalid`."""
    pts = O.unproject_dense(depth, k, ref_rot, ref_trs)
    return O.reproject_dense(pts, depth, k, gt_rot, gt_trs, invalid)


def filter_flow_by_mask(flow: Tensor, gt_mask: Tensor, invalid: float = 400.) -> Tensor:
    """models/utils/flow.py:6-26 with coords_grid (models/utils/warp.py:9-28): the flow is invalid where it already is, or
    where it points outside the target mask (bilinear sample of the mask, zeros padding, align_corners=False, < 0.9)."""
    b, _, h, w = flow.shape
    not_valid = (flow[:, 0] >= invalid) & (flow[:, 1] >= invalid)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing='ij')
    grid = torch.stack([xs, ys], dim=0).float()[None].repeat(b, 1, 1, 1) + flow
    gx = grid[:, 0] * 2. / max(w - 1, 1) - 1.
    gy = grid[:, 1] * 2. / max(h - 1, 1) - 1.
    m = F.grid_sample(gt_mask[:, None].to(flow.dtype), torch.stack([gx, gy], dim=-1), mode='bilinear', padding_mode='zeros',
                      align_corners=False)
    bad = (m < 0.9) | not_valid[:, None]
    return torch.where(bad.expand_as(flow), torch.full_like(flow, invalid), flow)


def raft_loss(pred: Tensor, gt_flow: Tensor, valid: Tensor, weight: float = 1., max_flow: float = 400., eps: float = 1e-10) -> Tensor:
    """RAFTLoss.forward (models/loss/sequence_loss.py:17-25)."""
    mag = torch.sum(gt_flow ** 2, dim=1).sqrt()
    v = ((valid >= 0.5) & (mag < max_flow)).to(gt_flow)
    return weight * (v[:, None] * (pred - gt_flow).abs()).sum() / (v.sum() + eps)


def l1_loss(pred_mask: Tensor, gt_mask: Tensor, weight: float = 1.) -> Tensor:
    """L1Loss.forward (models/loss/sequence_loss.py:36-38); `valid` is ignored by the reference."""
    return weight * torch.mean(torch.abs(pred_mask - gt_mask))


def sequence(losses: Sequence[Tensor], gamma: float = 0.8) -> Tensor:
    """SequenceLoss.forward (models/loss/sequence_loss.py:60-82): sum_i gamma^(n-i-1) loss_i."""
    n = len(losses)
    total = 0.
    for i, l in enumerate(losses):
        total = total + gamma ** (n - i - 1) * l
    return total


def _nearest(query: Tensor, points: Tensor) -> Tensor:
    """K=1 nearest neighbour of every query point among `points` (squared L2; pytorch3d knn_points semantics)."""
    d = ((query[:, None, :] - points[None, :, :]) ** 2).sum(-1)
    return d.argmin(dim=1)


def disentangle_pm_loss(pred_r: Tensor, pred_t: Tensor, gt_r: Tensor, gt_t: Tensor, labels: Tensor, points_list: List[Tensor],
                        symmetric: Sequence[bool], diameters: Sequence[float], weight: float = 1., loss_type: int = 1,
                        disentangle_z: bool = True, scale_depth_factor: float = 1.) -> Tensor:
    """DisentanglePointMatchingLoss.forward (models/loss/point_matching_loss.py:160-218), scale_xy = scale_depth = False
    (shipped config).  points_list[i]: the model points of sample i; symmetric[c], diameters[c] per class c."""
    b = pred_r.shape[0]
    sp, sg = pred_t.clone(), gt_t.clone()
    sp[..., -1] = pred_t[..., -1] * scale_depth_factor
    sg[..., -1] = gt_t[..., -1] * scale_depth_factor
    loss = 0.
    for i in range(b):
        pts = points_list[i]
        c = int(labels[i])
        gt_rot = (gt_r[i] @ pts.t()).t()
        gt_rt = gt_rot + sg[i][None]
        pred_rot = (pred_r[i] @ pts.t()).t() + sg[i][None]
        if symmetric[c]:
            pred_rot = pred_rot[_nearest(gt_rt, pred_rot)]
        l_rot = torch.mean(torch.linalg.norm(pred_rot - gt_rt, dim=-1, ord=loss_type))
        if disentangle_z:
            tz = sg[i].clone(); tz[-1] = sp[i, -1]
            l_depth = torch.mean(torch.linalg.norm(gt_rot + tz[None] - gt_rt, dim=-1, ord=loss_type))
            txy = sp[i].clone(); txy[-1] = sg[i, -1]
            l_xy = torch.mean(torch.linalg.norm(gt_rot + txy[None] - gt_rt, dim=-1, ord=loss_type))
            l_trans = l_depth + l_xy
        else:
            l_trans = torch.mean(torch.linalg.norm(gt_rot + sp[i][None] - gt_rt, dim=-1, ord=loss_type))
        loss = loss + (l_trans + l_rot) / diameters[c]
    return weight * loss / b


def refiner_loss(seq_flow_pred: Sequence[Tensor], seq_rot: Sequence[Tensor], seq_trs: Sequence[Tensor], seq_mask: Sequence[Tensor],
                 ref_rot: Tensor, ref_trs: Tensor, gt_rot: Tensor, gt_trs: Tensor, depth: Tensor, k: Tensor, rendered_mask: Tensor,
                 gt_mask: Tensor, labels: Tensor, points_list: List[Tensor], symmetric: Sequence[bool], diameters: Sequence[float],
                 max_flow: float = 400., gamma: float = 0.8, w_flow: float = 0.1, w_pose: float = 10., w_mask: float = 10.,
                 filter_invalid_flow: bool = True) -> Dict[str, Tensor]:
    """SCFlowRefiner.loss after get_pose (models/refiner/scflow_refiner.py:204-258) with the shipped loss configuration
    (configs/refine_models/scflow.py:75-104)."""
    gt_flow = gt_flow_from_poses(ref_rot, ref_trs, gt_rot, gt_trs, depth, k, max_flow)
    if filter_invalid_flow:
        gt_flow = filter_flow_by_mask(gt_flow, gt_mask, max_flow)
    pose_l = [disentangle_pm_loss(r, t, gt_rot, gt_trs, labels, points_list, symmetric, diameters, w_pose) for r, t in zip(seq_rot, seq_trs)]
    flow_l = [raft_loss(f, gt_flow, rendered_mask, w_flow, max_flow) for f in seq_flow_pred]
    occ = (torch.sum(gt_flow, dim=1) < max_flow).to(torch.float32)
    mask_l = [l1_loss(m.squeeze(1), occ, w_mask) for m in seq_mask]
    loss_pose, loss_flow, loss_mask = sequence(pose_l, gamma), sequence(flow_l, gamma), sequence(mask_l, gamma)
    return dict(loss=loss_pose + loss_flow + loss_mask, loss_pose=loss_pose, loss_flow=loss_flow, loss_mask=loss_mask,
                seq_pose=torch.stack(pose_l), seq_flow=torch.stack(flow_l), seq_mask=torch.stack(mask_l), gt_flow=gt_flow)


def make_loss_case(seed: int, batch: int, iters: int, height: int = 256, width: int = 256, num_class: int = 21, num_points: int = 200):
    """Seeded synthetic inputs of the loss: a scene (oracle make_scene), jittered ground-truth poses, per-class random
    model points (stand-in for models_eval/*.ply), a sequence of predictions around the ground truth."""
    scene = O.make_scene(seed, batch, height, width, num_class)
    g = torch.Generator().manual_seed(seed + 77)
    ang = torch.randn(batch, generator=g) * 0.15
    axis = F.normalize(torch.randn(batch, 3, generator=g), dim=-1)
    gt_rot = O.axis_angle_matrix(axis, ang) @ scene['ref_rotation']
    gt_trs = scene['ref_translation'] + torch.cat([15. * torch.randn(batch, 2, generator=g), 50. * torch.randn(batch, 1, generator=g)], 1)
    meshes = [(torch.rand(num_points + 7 * c, 3, generator=g) - 0.5) * (100. + 7. * c) for c in range(num_class)]
    symmetric = [c in (12, 15, 18, 19, 20) for c in range(num_class)]      # classes 13,16,19,20,21 of YCB-V (1-based)
    diameters = [float(100. + 9. * c) for c in range(num_class)]
    gt_flow = gt_flow_from_poses(scene['ref_rotation'], scene['ref_translation'], gt_rot, gt_trs, scene['depth'], scene['internel_k'])
    rendered_mask = (scene['depth'] > 0).float()
    # target-image mask: the rendered silhouette shifted by the mean GT flow of each sample, eroded at the border
    gt_mask = torch.zeros_like(rendered_mask)
    for i in range(batch):
        fg = scene['depth'][i] > 0
        dx = int(round(float(gt_flow[i, 0][fg].mean()))) if fg.any() else 0
        dy = int(round(float(gt_flow[i, 1][fg].mean()))) if fg.any() else 0
        gt_mask[i] = torch.roll(rendered_mask[i], shifts=(dy, dx), dims=(0, 1))
    seq_rot, seq_trs, seq_flow, seq_mask = [], [], [], []
    for it in range(iters):
        s = 0.5 ** it
        a = torch.randn(batch, generator=g) * 0.1 * s
        ax = F.normalize(torch.randn(batch, 3, generator=g), dim=-1)
        seq_rot.append(O.axis_angle_matrix(ax, a) @ gt_rot)
        seq_trs.append(gt_trs + s * torch.cat([5. * torch.randn(batch, 2, generator=g), 20. * torch.randn(batch, 1, generator=g)], 1))
        seq_flow.append(torch.where(gt_flow >= 400., torch.zeros_like(gt_flow), gt_flow) + s * 2. * torch.randn(batch, 2, height, width, generator=g))
        seq_mask.append(torch.sigmoid(4. * (rendered_mask[:, None] - 0.5) + s * torch.randn(batch, 1, height, width, generator=g)))
    return dict(scene=scene, gt_rot=gt_rot, gt_trs=gt_trs, meshes=meshes, symmetric=symmetric, diameters=diameters,
                rendered_mask=rendered_mask, gt_mask=gt_mask, seq_rot=seq_rot, seq_trs=seq_trs, seq_flow=seq_flow, seq_mask=seq_mask)
