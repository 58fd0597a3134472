"""Which calls of one training step make the host wait for the GPU?  (torch.cuda.set_sync_debug_mode("warn") on the
fourth step of `TrainStep` fed from a pinned host batch; prints every synchronising torch call with its stack line.)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adsorbdiff_b200 import PaiNN, synthetic as S, train as T

dev = torch.device("cuda:0")
net = PaiNN(None, 0, 1, so3_denoising=True).to(dev)
net.load_state_dict(S.random_state_dict(0), strict=True)
optim = dict(optimizer="AdamW", optimizer_params=dict(weight_decay=0.001), lr_initial=1e-4, clip_grad_norm=100, ema_decay=0.999,
             denoising_pos_params=dict(ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55))
step = T.TrainStep(net, optim, T.IGSO3Tables(dev), generator=torch.Generator(device=dev).manual_seed(0))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
hosts = []
for k in range(4):
    h = S.collate([S.make_system(k * B + i) for i in range(B)])
    for name, v in list(h.__dict__.items()):
        if isinstance(v, torch.Tensor):
            setattr(h, name, v.pin_memory())
    hosts.append(h)
for i in range(4):
    step(hosts[i % 4])
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    step(hosts[0])
torch.cuda.set_sync_debug_mode("default")
import traceback
for x in w:
    print(f"{x.filename}:{x.lineno}: {str(x.message)[:100]}")
print(len(w), "synchronising calls")
# host time of a step when nothing waits: enqueue 10 steps, time the host side only
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    step(hosts[i % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 10:.2f} ms/step, wall {1e3 * (t2 - t0) / 10:.2f} ms/step")

import cProfile, pstats
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for i in range(5):
    step(hosts[i % 4])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
