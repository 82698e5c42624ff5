#!/usr/bin/env python
"""bench.py - image-pairs/s of SCFlow pose refinement at 256x256, 8 GRU iterations (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32] [--iters 8]

A "step" is one pass of the hot path (SCFlowRefiner.get_pose: 3 encoder passes + correlation build + `iters`
refinement iterations) over one batch of B synthetic 256x256 rendered/real crop pairs (BASELINE config 2: B=32, 8
iterations, inference).  For N>1 (torchrun, one rank per GPU) every rank processes its own B crops (weak scaling,
no data-path collective; poses are all-gathered at the end of each step).

Prints ONE JSON line (rank 0).  `value` = pairs/s with inputs resident in HBM; `e2e` = the same through the public
API with pinned-host inputs and a device->host read of the refined poses inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'image-pairs/sec at 256x256, 8 GRU iters'
# DRAM bytes of one GRU z|r launch (B=32) from the committed ncu --set full capture (profiles/r01i_summary.md): 85.6 MB read
# + 10.2 MB written; the algorithmic minimum is 117 MB when nothing is L2-resident (inputs 83.9 MB, outputs 33.5 MB)
ZR_TRAFFIC_BYTES = 95.8e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=32, help='crop pairs per GPU per step (BASELINE config 2: 32)')
    ap.add_argument('--iters', type=int, default=8)
    ap.add_argument('--precision', type=int, default=int(os.environ.get('SCFLOW_PRECISION', '1')),
                    help='0 = fp32 CUDA-core convolutions, 1 = tcgen05 split-bf16 (fp32-accurate)')
    ap.add_argument('--no-graph', action='store_true', help='do not replay the decoder loop as a CUDA graph')
    ap.add_argument('--cpu-sample', type=int, default=4, help='crop pairs in the CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--seed', type=int, default=0)
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p['hbm_gbs'], bf16_burst=p['bf16_tflops'], bf16_sustained=p['bf16_tflops_sustained'], source='measured')
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '25',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_reference_time(batch: int, iters: int, steps: int, warmup: int, seed: int):
    """The reference's CPU implementation of the path (oracle port of SCFlowRefiner.get_pose) on all host cores."""
    from oracle import scflow_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scene = O.make_scene(seed, batch)
    sd = O.make_model_weights(seed)

    def step():
        with torch.no_grad():
            O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                       scene['depth'], scene['internel_k'], scene['label'], iters=iters)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = args.cpu_sample
    pairs_s, ms, threads = cpu_reference_time(sample, args.iters, args.steps, args.warmup, args.seed)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pairs_s, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'YCB-V-like 256x256 crop pairs, batch={args.batch}, {args.iters} iters, inference (BASELINE config 2)',
                   'sample': f'{sample} pairs per step on the host CPU'},
        'cpu_baseline': {'value': pairs_s, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{sample} crop pairs x {args.iters} iters per step, {args.steps} steps (oracle port of the '
                                   'reference get_pose; /root/reference is not present on the GPU box)'},
        'e2e': {'value': pairs_s, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import scflow_b200 as S
    from scflow_b200 import _lib, dist as D
    from oracle import scflow_oracle as O       # synthetic scene generator + cpu_baseline leg only
    from tests.util import scflow_model_cfg

    rank, world, local_rank = D.init_from_env()
    assert torch.cuda.is_available(), 'bench.py --impl ours needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    torch.backends.cudnn.allow_tf32 = False       # encoders (stock cuDNN, not replaced) stay fp32 like the parity tests
    torch.backends.cuda.matmul.allow_tf32 = False
    b, iters = args.batch, args.iters
    peaks = measured_peaks()

    model = S.build_refiner(scflow_model_cfg(iters=iters, precision=args.precision, use_cuda_graph=not args.no_graph))
    model.load_state_dict(O.make_model_weights(args.seed), strict=False)
    model = model.to(dev).eval()
    scene = O.make_scene(args.seed + 1000 * rank, b)
    keys = ('render_images', 'real_images', 'ref_rotation', 'ref_translation', 'depth', 'internel_k', 'label')
    host = {k: scene[k].contiguous().pin_memory() for k in keys}
    resident = {k: host[k].to(dev) for k in keys}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    out_host = torch.empty(b, 12, dtype=torch.float32).pin_memory()

    def step(inp):
        with torch.no_grad():
            outs = model.get_pose(inp['render_images'], inp['real_images'], inp['ref_rotation'], inp['ref_translation'],
                                  inp['depth'], inp['internel_k'], inp['label'])
        rot, trs = outs[2][-1], outs[3][-1]
        if world > 1:       # the job's result: refined poses of every shard on every rank (12 floats per crop)
            rot, trs = D.gather_poses(rot, trs, b * world)
        return rot, trs

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize(dev)

    lib = _lib.load()

    def timed(fn, steps):
        """Sum of per-step CUDA-event times on the launching stream; L2 flushed (untimed) before every step."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for s, e in evs:
            flush.zero_()
            s.record()
            fn()
            e.record()
        barrier()
        return sum(s.elapsed_time(e) for s, e in evs)

    # ---- warm-up (also counts our kernel launches per step: graph replays bypass the counter)
    model.decoder.use_cuda_graph = False
    c0 = lib.scf_launch_counter()
    step(resident)
    launches_per_step = lib.scf_launch_counter() - c0
    model.decoder.use_cuda_graph = not args.no_graph
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize(dev)

    # ---- value: inputs resident in HBM
    with ClockSampler(local_rank) as clocks:
        ms_total = timed(lambda: step(resident), args.steps)
    ms_total = D.max_over_ranks(ms_total, dev)
    ms_step = ms_total / args.steps
    value = world * b / (ms_step * 1e-3)

    # ---- e2e: pinned host inputs -> H2D -> get_pose -> D2H of the refined poses, all inside ONE timed region of K steps.
    # Every step copies its own inputs from pinned host memory and reads its own poses back; like any serving loop the
    # copy of step i+1 runs on a copy stream while step i computes (two device input buffers), L2 is flushed every step.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    bufs = [{k: torch.empty_like(resident[k]) for k in keys} for _ in range(2)]

    def e2e_run(steps):
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record(main_stream)
        copy_stream.wait_stream(main_stream)

        def upload(i):
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(done[i % 2])          # the step that last used this buffer has finished
                for k in keys:
                    bufs[i % 2][k].copy_(host[k], non_blocking=True)
                ready[i % 2].record(copy_stream)
        upload(0)
        for i in range(steps):
            if i + 1 < steps:
                upload(i + 1)
            flush.zero_()
            main_stream.wait_event(ready[i % 2])
            rot, trs = step(bufs[i % 2])
            lo = rank * b if world > 1 else 0
            out_host[:, :9].copy_(rot[lo:lo + b].reshape(b, 9), non_blocking=True)
            out_host[:, 9:].copy_(trs[lo:lo + b], non_blocking=True)
            done[i % 2].record(main_stream)
        t1.record(main_stream)
        barrier()
        return t0.elapsed_time(t1)
    e2e_run(2)
    e2e_ms = D.max_over_ranks(e2e_run(args.steps), dev) / args.steps
    h2d = sum(host[k].numel() * host[k].element_size() for k in keys)
    d2h = out_host.numel() * 4

    # ---- breakdown: encoders / decoder / per-iteration (device events, same hygiene)
    with torch.no_grad():
        feats = model.extract_feat(resident['render_images'], resident['real_images'])
    init_flow = torch.zeros(b, 2, 256, 256, device=dev)

    def dec_only():
        with torch.no_grad():
            model.decoder(*feats, resident['ref_rotation'], resident['ref_translation'], resident['depth'],
                          resident['internel_k'], label=resident['label'], init_flow=init_flow, invalid_flow_num=0.)

    def enc_only():
        with torch.no_grad():
            model.extract_feat(resident['render_images'], resident['real_images'])
    nsub = max(3, min(args.steps, 10))
    dec_only(); enc_only()
    dec_ms = timed(dec_only, nsub) / nsub
    enc_ms = timed(enc_only, nsub) / nsub
    model.decoder.iters = iters // 2
    dec_only()
    dec_half_ms = timed(dec_only, nsub) / nsub
    model.decoder.iters = iters
    per_iter_ms = (dec_ms - dec_half_ms) / (iters - iters // 2)

    # ---- roofline of the dominant kernel: the GRU z|r convolution (N=256, K=5*384=1920; 2 launches / iteration)
    roof = dominant_kernel_roofline(S, args, b, dev, flush, peaks)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            pairs_s, cms, threads = cpu_reference_time(args.cpu_sample, iters, 2, 1, args.seed)
            cpu = {'value': pairs_s, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                   'sample': f'{args.cpu_sample} crop pairs x {iters} iters, 1 warm-up + 2 timed reps of the oracle port '
                             f'(get_pose incl. encoders) on {threads} host threads'}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32' if args.precision == 0 else 'bf16x3 (split-bf16 tcgen05 MMAs, fp32 accumulate; parity-tested to EPE < 1e-3 px vs the fp32 CPU reference)',
            'data': 'synthetic',
            'config': {'workload': f'YCB-V-like 256x256 crop pairs, batch={b} per GPU, {iters} iters, inference (BASELINE config 2); '
                                   'step = get_pose (3 RAFT encoder passes + corr build + refinement loop), all on scflow_b200 kernels',
                       'l2': 'L2 flushed (256 MB write) before every timed step', 'cuda_graph': not args.no_graph,
                       'e2e': 'one timed region of K steps; per step: pinned-host -> device copy of that step\'s inputs (copy stream, '
                              'double-buffered so it overlaps the previous step\'s compute), L2 flush, get_pose, device -> pinned-host '
                              'read of the poses',
                       'precision': args.precision, 'parallelism': f'batch-sharded x{world}, no data-path collective'},
            'e2e': {'value': world * b / (e2e_ms * 1e-3), 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_ms},
            'gpu_launches': int(launches_per_step * args.steps),
            'clocks': clocks.summary(),
            'roofline': roof,
            'cpu_baseline': cpu,
            'breakdown': {'note': 'module-by-module on the generic API path (NCHW feature maps between encoder and decoder, no '
                                  'encoder overlap); the step itself uses the fused path, so encoders + decoder > ms_per_step',
                          'encoders_ms': enc_ms, 'decoder_ms': dec_ms, 'per_iter_ms': per_iter_ms,
                          'decoder_pairs_per_s': world * b / (dec_ms * 1e-3), 'launches_per_step': int(launches_per_step)},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return line


def dominant_kernel_roofline(S, args, b, dev, flush, peaks):
    """Times the GRU z|r convolution alone (CUDA events on its stream, L2 flushed between launches)."""
    from scflow_b200 import _lib
    g = torch.Generator().manual_seed(1)
    h = torch.tanh(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
    cxt = torch.relu(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
    mot = torch.randn(b, 32, 32, 128, generator=g).to(dev)
    wz = (torch.randn(128, 384, 1, 5, generator=g) * 0.02).to(dev)
    wr = (torch.randn(128, 384, 1, 5, generator=g) * 0.02).to(dev)
    bias = torch.zeros(256, device=dev)
    z, rh = torch.empty_like(h), torch.empty_like(h)
    if args.precision == 0 or not hasattr(S.ops, 'conv2d_tc'):
        packed = S.ops.pack_conv_weight([wz, wr])

        def launch():
            S.ops.conv2d_nhwc([(h, 0, 128), (cxt, 0, 128), (mot, 0, 128)], packed, bias, 256, (1, 5), 1, (0, 2), act='sigmoid',
                              out=z, epi=_lib.EPI_GRU_ZR, aux0=h, out2=rh)
        name, passes, kdim = 'conv_f32_kernel<4> (GRU z|r 1x5, fp32 CUDA cores)', 1, 1920
    else:
        launch, name, passes, kdim = S.ops.make_tc_gru_zr_bench(h, cxt, mot, wz, wr, bias, z, rh)
    for _ in range(3):
        launch()
    reps = 10
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in evs:
        flush.zero_()
        s.record()
        launch()
        e.record()
    torch.cuda.synchronize(dev)
    ms = sum(s.elapsed_time(e) for s, e in evs) / reps
    flops = 2.0 * b * 1024 * 256 * kdim              # FLOPs this launch performs: 2*M*N*K (the context columns, 1/3 of the
    #                                                  reference's K = 1920, are evaluated once per forward, not per iteration)
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks['bf16_burst']
    return {'kernel': name, 'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
            'traffic': ZR_TRAFFIC_BYTES if passes == 3 and b == 32 else None, 'traffic_source': 'ncu --set full, profiles/r01i_summary.md '
            '(dram__bytes_read.sum + dram__bytes_write.sum of one launch at B=32; bytes)', 'ms_per_launch': ms, 'algorithmic_flops_per_launch': flops, 'mma_passes': passes,
            'peak_source': f"{peaks['source']} bf16 dense burst (kernel timed alone)"}


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
