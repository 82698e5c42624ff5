"""Times the single-GPU shapes of BASELINE.json's configs 2-4 (CUDA events, L2 flushed, inputs resident).
config 3 runs the decoder with the identity pose head (the stock head cannot run off 256x256 - SURVEY.md §7 item 5);
config 4 is the per-GPU shard (B=16 of 64 over 4 GPUs, 12 iterations through test_cfg.iters)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S
from oracle import scflow_oracle as O
from tests.util import scflow_model_cfg

dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, b, h, w, iters, ident in [('config 2: 256x256 B=32 8 iters', 32, 256, 256, 8, False),
                                    ('config 3: 480x640 B=8 8 iters (identity pose head)', 8, 480, 640, 8, True),
                                    ('config 4 shard: 256x256 B=16 12 iters', 16, 256, 256, 12, False)]:
    cfg = scflow_model_cfg(iters=8, precision=1, use_cuda_graph=True)
    cfg['test_cfg'] = dict(iters=iters)
    model = S.build_refiner(cfg)
    model.load_state_dict(O.make_model_weights(0), strict=False)
    model = model.to(dev).eval()
    model.decoder.identity_pose_head = ident
    scene = {k: v.to(dev) for k, v in O.make_scene(0, b, h, w).items()}
    data = dict(rendered_images=scene['render_images'], real_images=scene['real_images'], ref_rotations=scene['ref_rotation'],
                ref_translations=scene['ref_translation'], rendered_depths=scene['depth'], internel_k=scene['internel_k'],
                labels=scene['label'])

    def step():
        with torch.no_grad():
            return model.forward_single_pass(data)
    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s, e in evs:
        flush.zero_(); s.record(); step(); e.record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / 5
    r = out['rotations'][0]
    print(f'{name}: {ms:.2f} ms/step = {b / ms * 1e3:.0f} pairs/s  (finite outputs: {bool(torch.isfinite(r).all())})')
    del model, scene, data
    torch.cuda.empty_cache()
