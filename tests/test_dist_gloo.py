"""CPU, world_size=2 over gloo: batch sharding + pose gather reproduce the single-process result (incl. the
label[0] quirk). The per-shard compute is the oracle here (no GPU in this tier); the sharding/gather code is the
product's (scflow_b200.dist)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import scflow_oracle as O
from scflow_b200 import dist as D

B, ITERS, SEED = 3, 1, 5


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    scene = O.make_scene(SEED, B)
    feats = O.make_features(SEED, B)
    data = dict(scene)
    data.update(feats)
    data['label'] = torch.tensor([7, 3, 11])
    return data


def _run(data, sd):
    n = data['depth'].shape[0]
    with torch.no_grad():
        outs = O.decoder_forward(sd, data['feat_render'], data['feat_real'], data['h_feat'], data['cxt_feat'],
                                 data['ref_rotation'], data['ref_translation'], data['depth'], data['internel_k'],
                                 data.get('pose_head_label', data['label']), torch.zeros(n, 2, 256, 256), 0., iters=ITERS)
    return outs[2][-1], outs[3][-1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w, _ = D.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    data = _inputs()
    sd = O.make_decoder_weights(SEED)
    shard = D.shard_batch(data, rank, world)
    rot, trs = _run(shard, sd)
    rot_all, trs_all = D.gather_poses(rot, trs, B)
    t = D.max_over_ranks(float(rank + 1), device='cpu')
    if rank == 0:
        q.put((rot_all, trs_all, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    rot_all, trs_all, tmax = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    rot_ref, trs_ref = _run(_inputs(), O.make_decoder_weights(SEED))
    assert rot_all.shape == (B, 3, 3) and trs_all.shape == (B, 3)
    assert float((rot_all - rot_ref).abs().max()) < 1e-6
    assert float((trs_all - trs_ref).abs().max()) < 1e-3
    assert tmax == 2.0
