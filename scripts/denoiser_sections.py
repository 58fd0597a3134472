"""Where the wall time of a 20-step `Denoiser.run()` from a pinned host batch goes (1024 systems)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adsorbdiff_b200 import PaiNN, Denoiser, synthetic as S, _cabi
import bench

dev = torch.device("cuda:0")
m = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
m.load_state_dict(S.random_state_dict(0), strict=True)
host = S.collate(bench.global_systems(1024))
for k, v in list(host.__dict__.items()):
    if isinstance(v, torch.Tensor):
        setattr(host, k, v.pin_memory())
params = dict(bench.SAMPLER_PARAMS, num_steps=20)
def run_once():
    t = [time.perf_counter()]
    def mark():
        torch.cuda.synchronize(); t.append(time.perf_counter())
    b = host.to(dev, non_blocking=True); mark()                       # 1 h2d
    den = Denoiser(b, m, params, device=dev, init_noise=torch.rand(1024, 3)); mark()   # 2 ctor
    orig_prepare = m._prepare
    def prep(batch):
        r = orig_prepare(batch); mark(); return r                     # 3 plan
    m._prepare = prep
    orig_cap = _cabi.capture_graph
    def cap(fn, device):
        mark()                                                         # 4 placement + eager first step + status
        g = orig_cap(fn, device); mark(); return g                     # 5 capture
    import adsorbdiff_b200.denoiser as D
    D._cabi.capture_graph = cap
    den.run(); mark()                                                  # 6 replays + tail
    m._prepare = orig_prepare; D._cabi.capture_graph = orig_cap
    out = b.pos.to("cpu"); mark()                                      # 7 d2h
    return [1e3 * (t[i + 1] - t[i]) for i in range(len(t) - 1)]
if os.environ.get("WITH_LOOPS"):   # what bench.py holds while it times e2e: two resident step loops
    from adsorbdiff_b200 import partition as PT
    noise_all = PT.initial_noise(1024, seed=1234)
    loop = bench.StepLoop(m, host.clone().to(dev), noise_all, pruned=False); loop.run(3)
    ploop = bench.StepLoop(m, host.clone().to(dev), noise_all, pruned=True); ploop.run(3)
    del ploop
    torch.cuda.synchronize()
    print("resident GiB", torch.cuda.memory_allocated() / 2**30, "reserved", torch.cuda.memory_reserved() / 2**30)
for it in range(3):
    r = run_once()
    print("ms [h2d, ctor, (calibration plan,) plan, placement + eager first step, capture, 19 replays + tail, d2h]:",
          [round(x, 2) for x in r], "total", round(sum(r), 1))
