"""Where does the config-5 training step spend its time?  Wall clock vs summed GPU kernel time (torch.profiler) and the top kernels.
    python tools/profile_train.py [batch]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scflow_b200 as S  # noqa: E402
from scflow_b200.training import Trainer  # noqa: E402
from oracle import scflow_oracle as O  # noqa: E402
from oracle import loss_oracle as L  # noqa: E402
from tests.util import scflow_model_cfg  # noqa: E402

b, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 8
dev = torch.device('cuda', 0)
c = L.make_loss_case(0, b, iters)
sym = {f'cls_{k + 1}': 1 for k, s_ in enumerate(c['symmetric']) if s_}
cfg = scflow_model_cfg(iters=iters, precision=1)
cfg.update(pose_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(
               type='DisentanglePointMatchingLoss', symmetry_types=sym, mesh_diameter=c['diameters'], loss_type='l1',
               disentangle_z=True, loss_weight=10.)),
           flow_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='RAFTLoss', loss_weight=.1, max_flow=400.)),
           mask_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='L1Loss', loss_weight=10.)))
model = S.build_refiner(cfg)
model.load_state_dict(O.make_model_weights(0), strict=False)
model = model.to(dev).train()
model.loss_functions()[0].loss_func.set_meshes(c['meshes'])
sc = c['scene']
data = dict(gt_rotations=c['gt_rot'], gt_translations=c['gt_trs'], ref_rotations=sc['ref_rotation'], ref_translations=sc['ref_translation'],
            real_images=sc['real_images'], rendered_images=sc['render_images'], rendered_depths=sc['depth'],
            rendered_masks=c['rendered_mask'], gt_masks=c['gt_mask'], internel_k=sc['internel_k'], labels=sc['label'])
data = {k: v.to(dev) for k, v in data.items()}
trainer = Trainer(model, lr=4e-4, weight_decay=1e-4, max_norm=10.)
for _ in range(3):
    trainer.train_step(data)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    trainer.train_step(data)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3 * 1e3
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    trainer.train_step(data)
    torch.cuda.synchronize()
ev = prof.key_averages()
from torch.autograd import DeviceType
kern = [e for e in prof.events() if e.device_type == DeviceType.CUDA]          # GPU-side records only (kernels, memcpy, memset)
cuda_total = sum(e.device_time for e in kern) / 1e3
nk = len(kern)
print(f'train step B={b}: wall {wall:.1f} ms, summed GPU kernel time {cuda_total:.1f} ms, {nk} GPU launches')
for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:14]:
    print(f'  {e.self_device_time_total / 1e3:8.2f} ms  x{e.count:5d}  {e.key[:100]}')
