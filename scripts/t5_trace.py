"""Pipeline trace of the tcgen05 message kernel (debug build with -DT5_TRACE): per-tile clock stamps of one CTA."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from adsorbdiff_b200 import _cabi

lib = os.environ.get("ADK_LIB") or os.path.join(ROOT, "adsorbdiff_b200", "lib", "libadsorbdiff_b200_trace.so")
_cabi._LIB_PATH = lib
from adsorbdiff_b200 import PaiNN, synthetic as S

dev = torch.device("cuda:0")
m = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
m.load_state_dict(S.random_state_dict(0), strict=True)
m.forward_graph = False
nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = [S.make_system(i) for i in range(min(64, nsys))]
b = S.collate([base[i % len(base)] for i in range(nsys)]).to(dev)
m(b); m(b)
torch.cuda.synchronize()
L = _cabi.load()
n = 512 * 16
buf = (ctypes.c_longlong * n)()
L.adk_message_t5_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
L.adk_message_t5_trace(buf, n)
t = np.array(buf[:], dtype=np.int64).reshape(512, 16)
nt = int((t[:511, 0] > 0).sum())
t0 = t[:nt, :12][t[:nt, :12] > 0].min()
names = ["g.top", "g.taps", "g.dE", "g.bE", "g.bF", "m.dE", "m.bF", "m.done", "e.top", "e.rdy", "e.body", "e.rel"]
print("tile " + " ".join(f"{x:>7s}" for x in names))
for i in range(min(nt, 48)):
    print(f"{i:4d} " + " ".join(f"{(t[i, k] - t0) if t[i, k] > 0 else -1:7d}" for k in range(12)))
c = t[511, :9]
print("CTA stamps (cycles from entry): setup %d | A staged %d  A mma-issue done %d  A end %d | B staged %d  B mma-issue done %d  B end %d | exit %d  (ntiles %d)"
      % (c[1] - c[0], c[2] - c[0], c[3] - c[0], c[4] - c[0], c[5] - c[0], c[6] - c[0], c[7] - c[0], c[8] - c[0], nt))
d = np.diff(t[:nt, 7])
print("MMA done-to-done period: mean", d[4:].mean(), "tiles", nt)
