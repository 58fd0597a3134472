"""Wall clock of a whole `Denoiser.run()` (setup, eager first step, graph capture, replays) at 1024 systems."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adsorbdiff_b200 import PaiNN, Denoiser, synthetic as S
dev = torch.device("cuda:0")
m = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
m.load_state_dict(S.random_state_dict(0, score_scale=S.SAMPLER_SCORE_SCALE), strict=True)
base = [S.make_system(i) for i in range(64)]
for early in (False, True):
    b = S.collate([base[i % 64] for i in range(1024)]).to(dev)
    params = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55, early_stop=early)
    torch.manual_seed(0)
    den = Denoiser(b, m, params, device=dev)
    torch.cuda.synchronize(); t0 = time.time()
    den.run()
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"Denoiser.run 1024 systems x {den.steps_run} steps (early_stop={early}): {dt:.2f} s -> {1024*den.steps_run/dt:.0f} system*steps/s wall")
