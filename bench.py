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
    ap.add_argument('--cpu-sample', type=int, default=32, help='crop pairs per CPU step (config 2: 32; shrunk automatically to fit the time budget)')
    ap.add_argument('--no-extras', action='store_true', help='skip the eager-GPU comparator and the config 1/3/4 lines')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--mode', default='infer', choices=['infer', 'train'],
                    help="'train': BASELINE config 5 - one optimisation step (forward + loss + backward + gradient all-reduce + clip + AdamW)")
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

    def __init__(self, index: int, keep_busy=None):
        # keep_busy: callable running one more untimed step of the same workload; used when the timed region was too short for
        # nvidia-smi (process start-up ~0.1-0.5 s on a busy 8-GPU box) to deliver a sample while the GPU was under that load
        self.index, self.rows, self.proc, self.keep_busy = index, [], None, keep_busy

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
            t_end = time.time() + 3.0
            while self.keep_busy is not None and len(self.rows) < 3 and time.time() < t_end:
                self.keep_busy()                      # untimed: only keeps the same load on the GPU until samples arrive
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
def cpu_reference_time(batch: int, iters: int, steps: int, warmup: int, seed: int, budget_s: float = 0.0):
    """The reference's CPU implementation of the path (oracle port of SCFlowRefiner.get_pose, incl. the three encoder passes) on
    all host cores.  With ``budget_s`` the per-step sample shrinks from ``batch`` crop pairs until warm-up + timed steps fit the
    budget (judged from one probe step).  Returns (pairs/s, ms per step, threads, pairs per step)."""
    from oracle import scflow_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.make_model_weights(seed)

    def make_step(n):
        scene = O.make_scene(seed, n)

        def step():
            with torch.no_grad():
                O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                           scene['depth'], scene['internel_k'], scene['label'], iters=iters)
        return step

    step = make_step(batch)
    t0 = time.perf_counter()
    step()                                           # probe (also the first warm-up step)
    probe = time.perf_counter() - t0
    if budget_s > 0 and probe * (steps + warmup) > budget_s and batch > 1:
        batch = max(1, int(batch * budget_s / (probe * (steps + warmup))))
        step = make_step(batch)
        step()
    for _ in range(max(warmup - 1, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, dt * 1e3, torch.get_num_threads(), batch


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    pairs_s, ms, threads, sample = cpu_reference_time(args.cpu_sample, args.iters, args.steps, args.warmup, args.seed, budget_s=200.)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pairs_s, 'unit': 'pairs/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'YCB-V-like 256x256 crop pairs, batch={args.batch}, {args.iters} iters, inference (BASELINE config 2)',
                   'sample': f'{sample} pairs per step on the host CPU' + (' (the full config-2 batch)' if sample == args.batch else
                                                                             ' (reduced to keep the run within a few minutes)')},
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
    def busy():            # the step WITHOUT the pose all-gather: it may run on one rank alone (no collective)
        with torch.no_grad():
            model.get_pose(resident['render_images'], resident['real_images'], resident['ref_rotation'], resident['ref_translation'],
                           resident['depth'], resident['internel_k'], resident['label'])
        torch.cuda.synchronize(dev)
    with ClockSampler(local_rank, keep_busy=busy) as clocks:
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

    def warmed(fn):            # three untimed calls: lazy initialisation, allocator growth and the graph capture stay outside
        for _ in range(3):
            fn()
        return timed(fn, nsub) / nsub
    dec_ms = warmed(dec_only)
    enc_ms = warmed(enc_only)
    model.decoder.iters = iters // 2
    dec_half_ms = warmed(dec_only)
    model.decoder.iters = iters
    per_iter_ms = (dec_ms - dec_half_ms) / (iters - iters // 2)

    # ---- roofline of the dominant kernel: the one-kernel SepConvGRU pass (2 launches / iteration, the largest single kernel)
    roof = dominant_kernel_roofline(S, args, b, dev, flush, peaks)
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        del model
        torch.cuda.empty_cache()
        extras = extra_lines(S, O, args, dev, flush)

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            pairs_s, cms, threads, sample = cpu_reference_time(args.cpu_sample, iters, 2, 1, args.seed, budget_s=45.)
            cpu = {'value': pairs_s, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                   'sample': f'{sample} crop pairs x {iters} iters (config-2 shape), 1 warm-up + 2 timed reps of the oracle port '
                             f'(get_pose incl. the three encoder passes) on {threads} host threads', 'ms_per_step': cms}
            if not args.no_extras:      # BASELINE config 1: one 256x256 pair, 4 iterations, the reference's own CPU-runnable case
                p1, ms1, _, _ = cpu_reference_time(1, 4, 5, 2, args.seed)
                cpu['config1'] = {'value': p1, 'unit': 'pairs/s', 'ms_per_step': ms1, 'sample': '1 pair x 4 iters, 2 warm-up + 5 timed reps'}
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
            'kernel_profiles': committed_kernel_profiles(peaks),
            'cpu_baseline': cpu,
            'extra': extras,
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


def committed_kernel_profiles(peaks):
    """The other kernels north_star names, from the COMMITTED `ncu --set full` summary (profiles/ncu_r02_kernels.csv; one launch each at
    B = 32 / 64 images, cold cache).  Static evidence copied into the line for convenience - not measured by this run."""
    import csv
    path = os.path.join(ROOT, 'profiles', 'ncu_r02_kernels.csv')
    algo = {   # kernel-name prefix -> (bound, algorithmic bytes or FLOPs per launch, what)
        'corr_pyramid_kernel': ('hbm', 32 * (2.10 + 5.57) * 1e6, 'correlation volume + 3 pooled levels, B=32: 7.67 MB per sample'),
        'void corr_lookup_l4r4_lean_kernel': ('hbm', 32 * (1.147 + 1.34) * 1e6, 'pyramid lookup, B=32: 1.147 MB read + 1.34 MB written per sample'),
        'void conv_rows_kernel': ('tensor', 2.0 * 64 * 16384 * 64 * 576, '3x3 64->64 encoder layer, 64 images of 128x128'),
        'void conv_stem_rows_kernel': ('hbm', 64 * 16384 * (2 * 2 * 64 + 64 * 4.0), '7x7/2 stem over the x-folded split-bf16 input (2 rows x 2 planes x 64 B per output pixel) -> fp32, 64 images'),
        'void gru_pass_kernel': ('tensor', 2.0 * 32 * 1024 * 384 * 1280, 'one SepConvGRU pass, B=32'),
    }
    try:
        out = []
        with open(path) as f:
            for r in csv.DictReader(f):
                key = next((k for k in algo if r['kernel'].startswith(k)), None)
                if key is None:
                    continue
                bound, work, what = algo[key]
                t = float(r['time_ns']) * 1e-9
                peak = peaks['hbm_gbs'] * 1e9 if bound == 'hbm' else peaks['bf16_burst'] * 1e12
                out.append({'kernel': r['kernel'], 'what': what, 'bound': bound, 'time_us_under_ncu': t * 1e6,
                            'achieved': work / t / (1e9 if bound == 'hbm' else 1e12), 'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s',
                            'frac': work / t / peak, 'tensor_pct_active': float(r['tensor_pct_active'] or 0),
                            'dram_bytes': float(r['dram_read_bytes'] or 0) + float(r['dram_write_bytes'] or 0), 'source': 'profiles/ncu_r02_kernels.csv'})
        return out or None
    except Exception:
        return None


def kernel_traffic(name: str):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` summary (profiles/r02_kernel_traffic.json: written by
    tools/ncu_summary.py from the .ncu-rep of the command it names); None if absent."""
    path = os.path.join(ROOT, 'profiles', 'r02_kernel_traffic.json')
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    e = t.get(name)
    return (e['dram_read_bytes'] + e['dram_write_bytes'], e.get('source')) if e else (None, None)


def dominant_kernel_roofline(S, args, b, dev, flush, peaks):
    """Times the one-kernel SepConvGRU pass alone (CUDA events on its stream, L2 flushed between launches).  Algorithmic FLOPs per
    launch = 2 * (B*1024 px) * 384 gate channels (z, r, q) * 1280 (5 taps x [h | motion] 256): the context columns are evaluated
    once per forward, not per iteration."""
    from scflow_b200 import _lib
    import ctypes as C
    g = torch.Generator().manual_seed(1)
    h = torch.tanh(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
    cxt = torch.relu(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
    mot = torch.randn(b, 32, 32, 128, generator=g).to(dev)
    if args.precision == 0:
        return None
    ws = [(torch.randn(128, 384, 1, 5, generator=g) * 0.02).to(dev) for _ in range(3)]
    bs = [torch.zeros(128, device=dev) for _ in range(3)]
    op = S.ops.GruPassFused(*ws, *bs)
    op.precompute(cxt)
    hs = S.ops.split_nchw(h.permute(0, 3, 1, 2).contiguous())
    ms_ = S.ops.split_nchw(mot.permute(0, 3, 1, 2).contiguous())
    out = torch.empty_like(h)
    out_hl = torch.empty(2, b, 32, 32, 128, device=dev, dtype=torch.bfloat16)
    d = _lib.GruPassDesc()
    d.h_hl, d.h_plane, d.h_f32 = hs.data_ptr(), hs[0].numel(), h.data_ptr()
    d.m_hl, d.m_plane = ms_.data_ptr(), ms_[0].numel()
    d.w_zr, d.w_q = op.w_zr.data_ptr(), op.w_q.data_ptr()
    d.pre_zr, d.pre_q = op.pre_zr.data_ptr(), op.pre_q.data_ptr()
    d.out_f32, d.out_hl, d.out_plane = out.data_ptr(), out_hl.data_ptr(), out_hl[0].numel()
    d.B, d.H, d.W, d.vertical = b, 32, 32, 0
    lib = _lib.load()

    def launch():
        _lib.check(lib.scf_gru_pass_fused(C.byref(d), _lib.stream_ptr()), 'scf_gru_pass_fused')
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
    flops = 2.0 * b * 1024 * 384 * 1280
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks['bf16_burst']
    traffic, src = kernel_traffic('gru_pass_kernel') if b == 32 else (None, None)
    return {'kernel': 'gru_pass_kernel<horizontal> (one SepConvGRU pass: z | r | q over [h | motion], 1x5 taps, tcgen05 split-bf16, M=128 N=256)',
            'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
            'traffic_source': src, 'ms_per_launch': ms, 'algorithmic_flops_per_launch': flops, 'mma_passes': 3,
            'launches_per_step': 2 * args.iters,
            'peak_source': f"{peaks['source']} bf16 dense burst (kernel timed alone); 3 MMAs per algorithmic MAC => ceiling 1/3"}


def extra_lines(S, O, args, dev, flush):
    """Comparators and the other BASELINE configurations, measured in the same run (rank 0, N=1):
      * gpu_eager_reference: the reference's arithmetic (oracle port of get_pose: torch conv2d / grid_sample / matmul, i.e.
        cuDNN / cuBLAS eager kernels) on THIS GPU at the config-2 shape, with TF32 off (fp32, the parity setting) and on (torch's
        default for convolutions) - the existing-Blackwell number the hand-written path has to beat;
      * config 3 (480x640, B=8, identity pose head), a config-4 shard (B=16 of 64 over 4 GPUs, 12 iterations) and config 1
        (B=1, 4 iterations) through scflow_b200."""
    from tests.util import scflow_model_cfg
    out = {}

    def timed(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for s, e in evs:
            flush.zero_()
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize(dev)
        return sum(s.elapsed_time(e) for s, e in evs) / reps

    # ---- eager PyTorch on the GPU
    b, iters = args.batch, args.iters
    sd = {k: v.to(dev) for k, v in O.make_model_weights(args.seed).items()}
    scene = {k: v.to(dev) for k, v in O.make_scene(args.seed, b).items()}

    def eager():
        with torch.no_grad():
            O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                       scene['depth'], scene['internel_k'], scene['label'], iters=iters)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, conv_tf32, mm_tf32 in (('fp32', False, False), ('tf32', True, True)):
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = conv_tf32, mm_tf32
            ms = timed(eager, reps=3, warm=2)
            out[f'gpu_eager_reference_{name}'] = {'ms_per_step': ms, 'value': b / (ms * 1e-3), 'unit': 'pairs/s',
                                                  'what': f'oracle port of get_pose under torch eager on the GPU (cuDNN / cuBLAS), B={b}, '
                                                          f'{iters} iters, cudnn.allow_tf32={conv_tf32}, matmul.allow_tf32={mm_tf32}'}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    del sd, scene
    torch.cuda.empty_cache()

    # ---- the other configurations through scflow_b200
    def ours(bb, hh, ww, it, identity):
        model = S.build_refiner(scflow_model_cfg(iters=it, precision=args.precision, use_cuda_graph=not args.no_graph))
        model.load_state_dict(O.make_model_weights(args.seed), strict=False)
        model = model.to(dev).eval()
        model.decoder.identity_pose_head = identity
        sc = {k: v.to(dev) for k, v in O.make_scene(args.seed, bb, hh, ww).items()}

        def step():
            with torch.no_grad():
                model.get_pose(sc['render_images'], sc['real_images'], sc['ref_rotation'], sc['ref_translation'], sc['depth'],
                               sc['internel_k'], sc['label'])
        ms = timed(step)
        return {'ms_per_step': ms, 'value': bb / (ms * 1e-3), 'unit': 'pairs/s'}
    out['config1_b1_it4'] = dict(ours(1, 256, 256, 4, False), what='BASELINE config 1 shape: one 256x256 pair, 4 iterations')
    out['config3_480x640_b8_it8'] = dict(ours(8, 480, 640, 8, True),
                                         what='BASELINE config 3: 480x640 full frames, B=8, 8 iterations, identity pose head (the stock '
                                              'head cannot run off 256x256, as in the reference)')
    out['config4_shard_b16_it12'] = dict(ours(16, 256, 256, 12, False),
                                         what='one GPU\'s shard of BASELINE config 4: B=64 over 4 GPUs = 16 crops, 12 iterations incl. pose head')
    return out


def run_train(args):
    """BASELINE config 5: YCB-V-like training step, 32 crops per GPU (256 at 8 GPUs), 8 iterations.  One step = forward (native
    kernels for the gradient-free parts, torch ops elsewhere) + loss + backward + bucketed NCCL all-reduce of the 32.7 MB gradient
    (overlapped with backward) + fused global-norm clip + AdamW.  Weak scaling; the collective is the all-reduce."""
    import scflow_b200 as S
    from scflow_b200 import _lib, dist as D
    from scflow_b200.training import Trainer
    from oracle import scflow_oracle as O
    from oracle import loss_oracle as L           # synthetic training-batch generator only
    from tests.util import scflow_model_cfg

    rank, world, local_rank = D.init_from_env()
    assert torch.cuda.is_available(), 'bench.py --mode train needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    b, iters = args.batch, args.iters
    c = L.make_loss_case(args.seed + 1000 * rank, b, iters)
    sym = {f'cls_{k + 1}': 1 for k, s_ in enumerate(c['symmetric']) if s_}
    cfg = scflow_model_cfg(iters=iters, precision=1)
    cfg.update(pose_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(
                   type='DisentanglePointMatchingLoss', symmetry_types=sym, mesh_diameter=c['diameters'], loss_type='l1',
                   disentangle_z=True, loss_weight=10.)),
               flow_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='RAFTLoss', loss_weight=.1, max_flow=400.)),
               mask_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='L1Loss', loss_weight=10.)))
    model = S.build_refiner(cfg)
    model.load_state_dict(O.make_model_weights(args.seed), strict=False)
    model = model.to(dev).train()
    model.loss_functions()[0].loss_func.set_meshes(c['meshes'])
    sc = c['scene']
    host = dict(gt_rotations=c['gt_rot'], gt_translations=c['gt_trs'], ref_rotations=sc['ref_rotation'], ref_translations=sc['ref_translation'],
                real_images=sc['real_images'], rendered_images=sc['render_images'], rendered_depths=sc['depth'],
                rendered_masks=c['rendered_mask'], gt_masks=c['gt_mask'], internel_k=sc['internel_k'], labels=sc['label'])
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    trainer = Trainer(model, lr=4e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=10.)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize(dev)

    loss_host = torch.zeros(1).pin_memory()

    def step(data):
        out = trainer.train_step(data)
        return out['loss']

    c0 = lib.scf_launch_counter()
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize(dev)
    launches = (lib.scf_launch_counter() - c0) / max(args.warmup, 3)

    def timed(e2e):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for s_, e_ in evs:
            flush.zero_()
            s_.record()
            data = {k: v.to(dev, non_blocking=True) for k, v in host.items()} if e2e else resident
            loss = step(data)
            if e2e:
                loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            e_.record()
        barrier()
        return sum(s_.elapsed_time(e_) for s_, e_ in evs) / args.steps
    with ClockSampler(local_rank) as clocks:
        ms = D.max_over_ranks(timed(False), dev)
    e2e_ms = D.max_over_ranks(timed(True), dev)
    grad_norm = trainer.grad_norm()
    if rank == 0:
        nparam = sum(p.numel() for p in trainer.params)
        line = {
            'metric': 'training crop-pairs/sec at 256x256, 8 GRU iters (forward + loss + backward + gradient all-reduce + clip + AdamW)',
            'value': world * b / (ms * 1e-3), 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (native kernels: split-bf16 tcgen05 / fp32; torch backward with the framework defaults, cuDNN TF32 allowed as in '
                     'the reference\'s training run)',
            'data': 'synthetic',
            'config': {'workload': f'YCB-V-like training step (BASELINE config 5): batch={b} per GPU ({b * world} global), {iters} iters, '
                                   'AdamW lr 4e-4 wd 1e-4, grad clip 10', 'l2': 'L2 flushed (256 MB write) before every timed step',
                       'parallelism': f'data parallel x{world}: bucketed NCCL all-reduce (SUM) of {nparam * 4 / 1e6:.1f} MB of gradients, '
                                      'overlapped with backward', 'mode': 'train'},
            'e2e': {'value': world * b / (e2e_ms * 1e-3), 'unit': 'pairs/s', 'ms_per_step': e2e_ms,
                    'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in host.values()), 'd2h_bytes_per_step': 4},
            'gpu_launches': int(launches * args.steps),
            'clocks': clocks.summary(),
            'train': {'parameters': nparam, 'grad_norm_last': grad_norm, 'buckets': len(trainer.buckets)},
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    elif args.mode == 'train':
        run_train(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
