"""A few training steps at 48 systems for ncu: `ncu ... python scripts/train_steps.py [steps]`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adsorbdiff_b200 import PaiNN, synthetic as S, train as T
import bench

dev = torch.device("cuda:0")
net = PaiNN(None, 0, 1, so3_denoising=True).to(dev)
net.load_state_dict(S.random_state_dict(0), strict=True)
step = T.TrainStep(net, bench.TRAIN_OPTIM, T.IGSO3Tables(dev))
B = 48
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    print(float(step(S.collate([S.make_system(i * B + k) for k in range(B)]).to(dev))))
