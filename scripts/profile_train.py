"""Kernel-time table of one training step (torch profiler; CUDA time per kernel name)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from adsorbdiff_b200 import PaiNN, synthetic as S, train as T
import bench

dev = torch.device("cuda:0")
net = PaiNN(None, 0, 1, so3_denoising=True).to(dev)
net.load_state_dict(S.random_state_dict(0), strict=True)
step = T.TrainStep(net, bench.TRAIN_OPTIM, T.IGSO3Tables(dev))
B = int(os.environ.get("TRAIN_B", "48"))
hosts = [S.collate([S.make_system(k * B + i) for i in range(B)]).to(dev) for k in range(2)]
for i in range(3):
    step(hosts[i % 2].clone())
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        step(hosts[i % 2].clone())
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))

# host enqueue time of one step (no sync inside: TrainStep reads nothing back)
import time
torch.cuda.synchronize()
ts = []
for i in range(6):
    t0 = time.perf_counter()
    loss = step(hosts[i % 2].clone())
    ts.append(time.perf_counter() - t0)
torch.cuda.synchronize()
print("host enqueue ms per step:", [f"{1e3 * t:.1f}" for t in ts])
