"""Eager (no CUDA graph) denoising steps for ncu: `ncu ... python scripts/profile_step.py [systems] [steps]`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from adsorbdiff_b200 import _cabi
if os.environ.get("ADK_LIB"):
    _cabi._LIB_PATH = os.environ["ADK_LIB"]
from adsorbdiff_b200 import PaiNN, synthetic as S
from adsorbdiff_b200.denoiser import schedule_table

systems = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
model = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
model.load_state_dict(S.random_state_dict(0), strict=True)
if os.environ.get("ADK_GEMM"):
    model.gemm = os.environ["ADK_GEMM"]
base = [S.make_system(i) for i in range(min(64, systems))]
batch = S.collate([base[i % len(base)] for i in range(systems)]).to(dev)
plan, z, pos = model._prepare(batch)
tags = batch.tags.to(torch.int32).contiguous()
fixed = batch.fixed.to(torch.int32).contiguous()
torch.manual_seed(0)
noise = torch.rand(systems, 3).to(dev)
_cabi.call("adk_init_placement", dev, _cabi.ptr(pos), _cabi.ptr(plan.cell_f32), _cabi.ptr(plan.atom_off), _cabi.ptr(tags), _cabi.ptr(noise), systems)
params = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55)
sched = schedule_table(params, dev)
step = torch.zeros(1, dtype=torch.int32, device=dev)
max_upd = torch.zeros(systems, dtype=torch.float32, device=dev)
for _ in range(steps):
    model._run(plan, z, pos)
    _cabi.call("adk_se3_step", dev, _cabi.ptr(pos), _cabi.ptr(plan.cell_f32), _cabi.ptr(plan.atom_off), _cabi.ptr(tags), _cabi.ptr(fixed),
               _cabi.ptr(plan.out[0]), _cabi.ptr(plan.out[1]), _cabi.ptr(sched), _cabi.ptr(step), systems, None, _cabi.ptr(max_upd), None)
torch.cuda.synchronize()
model.check_status(plan)
print("ok", systems, steps)
