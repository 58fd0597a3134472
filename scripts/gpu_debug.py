"""Stage-by-stage diagnostics on a GPU box (prints, never asserts): used while bringing kernels up."""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from adsorbdiff_b200 import PaiNN, Denoiser, _cabi, synthetic as S
from oracle import painn_oracle as O
from tests.cases import CASES

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0))
sd = S.random_state_dict(0)
model = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
model.load_state_dict(sd, strict=True)


def stage(name, fn):
    try:
        fn()
    except Exception:
        print(f"[{name}] EXCEPTION")
        traceback.print_exc()


def graph_check():
    for name, (make, pbc) in CASES.items():
        if name == "empty":
            continue
        from adsorbdiff_b200 import painn
        painn._PBC_STICKY[:] = [True, True, True]
        b = make()
        o = O.generate_graph_values(b.pos.numpy(), b.cell.numpy(), b.natoms, pbc=pbc or (True, True, True))
        bd = b.clone()
        if pbc is not None:
            bd.pbc = torch.tensor([pbc] * b.num_graphs)
        ei, neigh, d, rv, _ = model.generate_graph_values(bd.to(dev))
        ei = ei.cpu().numpy()
        same_shape = ei.shape == o["edge_index"].shape
        eq = same_shape and np.array_equal(ei, o["edge_index"])
        print(f"[graph:{name}] E cuda={ei.shape[1]} oracle={o['edge_index'].shape[1]} equal={eq} "
              f"neigh cuda={neigh.cpu().tolist()} oracle={o['neighbors'].tolist()}")
        if same_shape and not eq:
            bad = np.nonzero((ei != o["edge_index"]).any(0))[0]
            print("   first mismatches", bad[:5], ei[:, bad[:5]].T.tolist(), o["edge_index"][:, bad[:5]].T.tolist())
        if eq:
            print("   d maxrel", float(np.abs(d.cpu().numpy() - o["dist"]).max() / o["dist"].max()),
                  "unit maxabs", float(np.abs(rv.cpu().numpy() - o["unit_vec"]).max()))
        plan = model._plan_cache
        deg = plan.row_deg.cpu().numpy()
        print("   in-degree min/mean/max", deg.min(), deg.mean(), deg.max(), "sum", deg.sum())


def forward_check():
    for name in ("jit2", "mixed", "gas"):
        make, pbc = CASES[name]
        b = make()
        tr_o, tr_c = {}, {}
        o1, o2 = O.painn_forward(sd, b.atomic_numbers, b.pos.numpy(), b.cell.numpy(), b.natoms, trace=tr_o)
        f1, f2 = model(b.clone().to(dev), trace=tr_c)
        for k in tr_c:
            ref = tr_o[k]
            print(f"[fwd:{name}] {k:10s} err/max {float((tr_c[k].cpu() - ref).abs().max() / ref.abs().max()):.3e}")
        for got, ref, nm in ((f1, o1, "f1"), (f2, o2, "f2")):
            print(f"[fwd:{name}] {nm} err/max {float((got.cpu() - ref).abs().max() / ref.abs().max()):.3e} "
                  f"max|ref| {float(ref.abs().max()):.4g}")


def sampler_check():
    import ast
    from tests.cases import sampler_batch
    g = np.load(os.path.join(ROOT, "tests/golden/sampler.npz"))
    params = ast.literal_eval(str(g["params"]))
    params["early_stop"] = False
    for use_graph in (False, True):
        b = sampler_batch().to(dev)
        torch.manual_seed(1234)
        den = Denoiser(b, model, params, device="cuda:0", use_cuda_graph=use_graph)
        import pathlib
        den.traj_dir = pathlib.Path("/tmp/adk_traj")
        den.traj_names = b.sid
        den.run()
        fr = den.frames.cpu().numpy()
        print(f"[sampler graph={use_graph}] per-step max|dpos|",
              ["%.2e" % float(np.abs(fr[t] - g["traj"][t]).max()) for t in range(g["traj"].shape[0])])


def timing():
    for B in (1, 64, 256, 1024):
        b = S.make_placements(0, B).to(dev)
        model(b)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            model(b)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"[time] forward B={B}: {dt*1e3:.2f} ms  -> {B/dt:.0f} systems/s (eager, incl. launch overhead)")


stage("graph", graph_check)
stage("forward", forward_check)
stage("sampler", sampler_check)
stage("timing", timing)
