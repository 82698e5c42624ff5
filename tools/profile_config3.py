"""One forward_single_pass of BASELINE config 3 (480x640, B=8, 8 iterations, identity pose head) between
cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S
from oracle import scflow_oracle as O
from tests.util import scflow_model_cfg

dev = torch.device('cuda', 0)
b, h, w = 8, 480, 640
model = S.build_refiner(scflow_model_cfg(iters=8, precision=1, use_cuda_graph=False))
model.load_state_dict(O.make_model_weights(0), strict=False)
model = model.to(dev).eval()
model.decoder.identity_pose_head = True
scene = {k: v.to(dev) for k, v in O.make_scene(0, b, h, w).items()}
data = dict(rendered_images=scene['render_images'], real_images=scene['real_images'], ref_rotations=scene['ref_rotation'],
            ref_translations=scene['ref_translation'], rendered_depths=scene['depth'], internel_k=scene['internel_k'], labels=scene['label'])
with torch.no_grad():
    for _ in range(2):
        model.forward_single_pass(data)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model.forward_single_pass(data)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('profiled config 3')
