"""One get_pose step (BASELINE config 2) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--precision 1] [--decoder-only]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scflow_b200 as S  # noqa: E402
from oracle import scflow_oracle as O  # noqa: E402
from tests.util import scflow_model_cfg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--precision', type=int, default=1)
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--iters', type=int, default=8)
ap.add_argument('--decoder-only', action='store_true')
args = ap.parse_args()

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda', 0)
model = S.build_refiner(scflow_model_cfg(iters=args.iters, precision=args.precision, use_cuda_graph=False))
model.load_state_dict(O.make_model_weights(0), strict=False)
model = model.to(dev).eval()
scene = {k: v.to(dev) for k, v in O.make_scene(0, args.batch).items()}


def step():
    with torch.no_grad():
        if args.decoder_only:
            model.decoder(*feats, scene['ref_rotation'], scene['ref_translation'], scene['depth'], scene['internel_k'],
                          label=scene['label'], init_flow=init_flow, invalid_flow_num=0.)
        else:
            model.get_pose(scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                           scene['depth'], scene['internel_k'], scene['label'])


with torch.no_grad():
    feats = model.extract_feat(scene['render_images'], scene['real_images'])
init_flow = torch.zeros(args.batch, 2, 256, 256, device=dev)
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled one step')
