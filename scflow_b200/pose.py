"""Pose / flow geometry with the reference's function names (models/utils/pose.py), on the CUDA kernels."""
from typing import List, Sequence, Tuple

import torch

from . import ops


def get_pose_from_delta_pose(rotation_delta, translation_delta, rotation_src, translation_src, weight=10.,
                             depth_transform='exp', detach_depth_for_xy=False):
    """pose.py:124-149. Only the configuration SCFlow ships is implemented: ortho6d delta rotation (n,6),
    depth_transform='exp', weight=10 (detach_depth_for_xy only affects gradients)."""
    if rotation_delta.size(1) != 6:
        raise NotImplementedError('only rotation_mode="ortho6d" is implemented (configs/refine_models/scflow.py:68)')
    if depth_transform != 'exp' or float(weight) != 10.:
        raise NotImplementedError('only depth_transform="exp", weight=10 is implemented')
    return ops.pose_update(rotation_delta.contiguous(), translation_delta.contiguous(), rotation_src.contiguous(),
                           translation_src.contiguous())


def unproject_dense(depth, internel_k, rotation, translation) -> torch.Tensor:
    """Dense replacement of per-sample cal_3d_2d_corr (pose.py:44-64): [B,H,W,4] = (X_obj, valid) for every pixel,
    no nonzero() compaction and therefore no host synchronisation."""
    return ops.unproject(depth.contiguous(), internel_k.contiguous(), rotation.contiguous(), translation.contiguous())


def get_flow_from_delta_pose_dense(rotation_dst, translation_dst, k, points4, invalid_num=400.) -> torch.Tensor:
    """Dense replacement of get_flow_from_delta_pose_and_points (pose.py:66-88)."""
    return ops.reproject(points4, k.contiguous(), rotation_dst.contiguous(), translation_dst.contiguous(), invalid_num)


def cal_3d_2d_corr(depth, internel_k, rotation, translation, occlusion=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference-compatible list form for ONE sample: (points_2d [N,2] xy float, points_3d [N,3]) in row-major
    foreground order. Compaction uses torch.nonzero (host sync) exactly like the reference; the decoder loop itself
    uses the dense form."""
    pts4 = unproject_dense(depth[None], internel_k[None], rotation[None], translation[None])[0]
    mask = depth > 0
    if occlusion is not None:
        mask = mask & occlusion
    ys, xs = torch.nonzero(mask, as_tuple=True)
    return torch.stack([xs, ys], dim=-1).float(), pts4[ys, xs, :3]


def get_flow_from_delta_pose_and_points(rotation_dst, translation_dst, k, points_2d_list: Sequence[torch.Tensor],
                                        points_3d_list: Sequence[torch.Tensor], height: int, width: int,
                                        invalid_num: float = 400.) -> torch.Tensor:
    """Reference-compatible list form (pose.py:66-88): scatters the lists into the dense point map, then one kernel."""
    n = len(rotation_dst)
    pts4 = torch.zeros(n, height, width, 4, device=rotation_dst.device, dtype=torch.float32)
    for i in range(n):
        xy = points_2d_list[i].long()
        pts4[i, xy[:, 1], xy[:, 0], :3] = points_3d_list[i]
        pts4[i, xy[:, 1], xy[:, 0], 3] = 1.
    return get_flow_from_delta_pose_dense(rotation_dst, translation_dst, k, pts4, invalid_num)
