import sys, os, time
sys.path.insert(0, "/root/repo")
import torch
from adsorbdiff_b200 import PaiNN, synthetic as S, train as T
import bench
dev = torch.device("cuda:0")
net = PaiNN(None, 0, 1, so3_denoising=True).to(dev)
net.load_state_dict(S.random_state_dict(0), strict=True)
step = T.TrainStep(net, bench.TRAIN_OPTIM, T.IGSO3Tables(dev))
B = 48
hosts = [S.collate([S.make_system(k * B + i) for i in range(B)]) for k in range(3)]
for h in hosts:
    for name, v in list(h.__dict__.items()):
        if isinstance(v, torch.Tensor):
            setattr(h, name, v.pin_memory())
for i in range(3):
    float(step(hosts[i % 3].to(dev, non_blocking=True)))
def t(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / n, r
ms, b = t(lambda: hosts[0].to(dev, non_blocking=True)); print("h2d batch", ms)
ms, nb = t(lambda: T.tr_so3_schedule(hosts[0].to(dev, non_blocking=True), bench.TRAIN_OPTIM["denoising_pos_params"], step.tables)); print("h2d + noising", ms)
net.train()
def plan_only():
    nb2 = hosts[1].to(dev, non_blocking=True)
    return net._prepare(nb2)
ms, _ = t(plan_only); print("h2d + plan/_prepare", ms)
def fwd():
    return T.forward_train(net, nb, check=False)
ms, out = t(fwd); print("forward", ms)
def fwd_bwd():
    o = T.forward_train(net, nb, check=False)
    l = T.denoising_loss(o, nb, step.tables)
    net.zero_grad(set_to_none=True)
    l.backward()
ms, _ = t(fwd_bwd); print("forward+loss+backward", ms)
def full():
    return float(step(hosts[2].to(dev, non_blocking=True)))
ms, _ = t(full); print("full step incl loss read", ms)
def opt():
    torch.nn.utils.clip_grad_norm_(step.params, max_norm=100.0, foreach=True); step.optimizer.step()
    torch._foreach_lerp_(step.shadow, [q.detach() for q in step.params], 1e-3)
ms, _ = t(opt); print("clip+adamw+ema", ms)
