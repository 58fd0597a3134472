"""Soak: 200 training steps on changing batches (sizes 16..64 systems): loss finite, memory flat, no status bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adsorbdiff_b200 import PaiNN, synthetic as S, train as T
import bench

dev = torch.device("cuda:0")
net = PaiNN(None, 0, 1, so3_denoising=True, scale_file={f"upd_out_scalar_scale_{i}": v for i, v in enumerate(S.SHIPPED_SCALE_FACTORS)}).to(dev)
net.load_state_dict(S.random_state_dict(0), strict=False)
optim = dict(bench.TRAIN_OPTIM, status_every=20)
step = T.TrainStep(net, optim, T.IGSO3Tables(dev))
mem = []
for i in range(200):
    B = 16 + (i * 7) % 49
    batch = S.collate([S.make_system((i * 64 + k) % 5000) for k in range(B)]).to(dev)
    loss = float(step(batch))
    assert loss == loss and abs(loss) < 1e9, (i, loss)
    if i % 20 == 0:
        mem.append(torch.cuda.memory_allocated() / 2**20)
        print(i, B, f"loss {loss:.4g}", f"alloc {mem[-1]:.0f} MiB", f"peak {torch.cuda.max_memory_allocated() / 2**20:.0f} MiB")
step.check_gemm_status()
assert abs(mem[8] - mem[1]) < 16 and abs(mem[9] - mem[2]) < 16, mem   # same batch size 140 steps apart: no growth
print("soak ok")
