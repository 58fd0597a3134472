"""Bring-up check of the tcgen05 message kernel: t5 vs the mma kernel (per-layer traces) and timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adsorbdiff_b200 import _cabi
if os.environ.get("ADK_LIB"):
    _cabi._LIB_PATH = os.environ["ADK_LIB"]
from adsorbdiff_b200 import PaiNN, synthetic as S

dev = torch.device("cuda:0")
m = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
m.load_state_dict(S.random_state_dict(0), strict=True)
m.forward_graph = False
nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 32
base = [S.make_system(i) for i in range(min(64, nsys))]
b = S.collate([base[i % len(base)] for i in range(nsys)]).to(dev)
m.t5_min_ctas = 0
res = {}
for eng in (() if os.environ.get("T5_TIME_ONLY") else ("mma", "t5")):
    m.msg = eng
    tr = {}
    o = m(b, trace=tr)
    torch.cuda.synchronize()
    res[eng] = (o, tr)
rel = lambda a, r: float((a.double() - r.double()).abs().max() / r.double().abs().max())
for k in (res["mma"][1] if res else ()):
    print(f"{k:10s} t5 vs mma: {rel(res['t5'][1][k], res['mma'][1][k]):.3e}")
for i in (range(2) if res else ()):
    print("out", i, rel(res["t5"][0][i], res["mma"][0][i]))
if len(sys.argv) > 2:
    from oracle import painn_oracle as O
    hb = S.collate(base[:2])
    sd = S.random_state_dict(0)
    tr64 = {}
    o64 = O.painn_forward(sd, hb.atomic_numbers, hb.pos.numpy(), hb.cell.numpy(), hb.natoms, trace=tr64, dtype=torch.float64)
    n2 = int(hb.pos.shape[0])
    for eng in ("mma", "t5"):
        worst = max(rel(res[eng][1][k][:n2].cpu(), tr64[k]) for k in tr64 if k in res[eng][1])
        print(eng, "worst feature vs fp64:", f"{worst:.3e}", "outs:", [f"{rel(res[eng][0][i][:n2].cpu(), o64[i]):.3e}" for i in range(2)])
for eng in ("mma", "t5"):
    m.msg = eng
    p, z, pos = m._prepare(b)
    for _ in range(2):
        m._run(p, z, pos)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m._run(p, z, pos)
    e1.record()
    torch.cuda.synchronize()
    print(eng, "forward ms:", e0.elapsed_time(e1) / 5)
