"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed over NCCL/NVLink).

The refinement path shards by independent samples: every crop's correlation volume, GRU state and
pose are private, so inference needs NO data-path collective - only a final gather of the refined poses
(12 floats per crop).  The reference's equivalents are ``MMDistributedDataParallel`` + ``collect_results_gpu``
(test.py:121-126, tools/eval.py:185-215).

One reference quirk couples samples: ``MultiClassPoseHead`` selects the class row of ``label[0]`` for the whole batch
(pose_head.py:209-210).  ``shard_batch`` therefore adds a ``pose_head_label`` entry holding the GLOBAL first label; the
refiner / decoder use it as the pose head's class selector, which makes an N-rank run reproduce the 1-rank result exactly.
The per-sample ``labels`` are left untouched (they pick meshes, diameters and symmetry flags in the loss and are returned
per image by ``forward_single_pass``).
"""
import os
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun contract).
    Returns (rank, world_size, local_rank). No-op for single-process runs."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples for `rank`; the first n % world ranks get one extra sample."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f'bad rank/world: {rank}/{world}')
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_BATCH_KEYS = ('real_images', 'rendered_images', 'render_images', 'ref_rotations', 'ref_translations', 'ref_rotation',
               'ref_translation', 'rendered_depths', 'depth', 'internel_k', 'labels', 'label', 'feat_render', 'feat_real',
               'h_feat', 'cxt_feat', 'init_flow')


def shard_batch(data: Dict[str, torch.Tensor], rank: int, world: int, keep_global_label0: bool = True) -> Dict[str, torch.Tensor]:
    """Slice every per-sample tensor of `data` to this rank's shard. Non-tensor entries are passed through.  With
    ``keep_global_label0`` the shard carries ``pose_head_label`` = the global batch's first label (1-element tensor)."""
    label_key = 'labels' if 'labels' in data else ('label' if 'label' in data else None)
    n = None
    for k in _BATCH_KEYS:
        if k in data and isinstance(data[k], torch.Tensor):
            n = data[k].shape[0]
            break
    if n is None:
        raise ValueError('shard_batch: no per-sample tensor found')
    lo, hi = shard_range(n, rank, world)
    out = {}
    for k, v in data.items():
        if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[lo:hi].contiguous()
        else:
            out[k] = v
    if keep_global_label0 and label_key is not None:
        out['pose_head_label'] = data[label_key][:1].clone()
    return out


def gather_poses(rotation: torch.Tensor, translation: torch.Tensor, n_total: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the refined poses of every shard (shards may differ by one sample) -> ([n_total,3,3], [n_total,3])
    in global sample order on every rank.  12 floats per sample: control-plane traffic, not a data-path collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rotation, translation
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = (n_total + world - 1) // world
    packed = torch.zeros(cap, 12, device=rotation.device, dtype=torch.float32)
    n_local = rotation.shape[0]
    packed[:n_local, :9] = rotation.reshape(n_local, 9)
    packed[:n_local, 9:] = translation
    bufs = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed, group=group)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        rows.append(bufs[r][:hi - lo])
    full = torch.cat(rows, dim=0)
    return full[:, :9].reshape(-1, 3, 3), full[:, 9:]


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Device-side max-reduce of a timing (bench.py: the step time is the slowest rank's)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu'))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
