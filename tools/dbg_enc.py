import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scflow_b200 as S
from oracle import scflow_oracle as O
from tests.util import scflow_model_cfg
dev = torch.device('cuda', 0)
model = S.build_refiner(scflow_model_cfg(iters=8, precision=1, use_cuda_graph=True))
model.load_state_dict(O.make_model_weights(0), strict=False)
model = model.to(dev).eval()
scene = {k: v.to(dev) for k, v in O.make_scene(0, 32).items()}
def enc_only():
    with torch.no_grad():
        return model.extract_feat(scene['render_images'], scene['real_images'])
for i in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); enc_only(); e.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f'enc_only {i}: gpu {s.elapsed_time(e):.2f} ms, cpu enqueue {(t1 - t0) * 1e3:.2f} ms')
