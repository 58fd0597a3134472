"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden
The reference cannot travel to the GPU box, so its outputs on fixed synthetic inputs are
frozen here; `tests/test_oracle_golden.py` then pins `oracle/painn_oracle.py` to them and
the `-m gpu` tests pin the CUDA path to both.

Inputs are never stored: every case is regenerated from `adsorbdiff_b200.synthetic` with the
seeds recorded in the fixture (`case` string), weights from `random_state_dict(seed=0)`.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adsorbdiff_b200 import synthetic as S  # noqa: E402
from oracle import painn_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
PROBE_ATOMS = 6  # rows of (x, vec) kept per checkpoint to keep fixtures small

CHECKPOINT_CASES = ("jit2", "mixed")  # cases that keep per-layer (x, vec) probes
CHECKPOINT_LAYERS = (0, 2, 5)

SAMPLER_PARAMS = dict(num_steps=8, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55)


def cases():
    """name -> (batch, pbc or None).  Keep in sync with tests/cases.py."""
    from tests.cases import CASES

    return CASES


def probe_rows(n):
    return np.unique(np.linspace(0, n - 1, PROBE_ATOMS).astype(np.int64))


class _FakeTrainer:
    """The 6 lines of `DenoisingTrainer` the sampler touches (sde_denoising_trainer.py:539-553)."""

    def __init__(self, model):
        self.model = model
        self._unwrapped_model = model

    @torch.no_grad()
    def predict_denoising(self, batch, per_image=False, disable_tqdm=True):
        p1, p2 = self.model(batch)
        return {"positions": p1.detach(), "positions_free": p2.detach()}


def main():
    ns = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    sd = S.random_state_dict(0)
    model = ns.PaiNN(None, 0, 1, scale_file=ns.scale_file, so3_denoising=True).eval()
    model.load_state_dict(sd, strict=True)

    for name, (make, pbc) in cases().items():
        b = make()
        if pbc is not None:
            b.pbc = torch.tensor([pbc] * b.num_graphs)
        # the reference mutates a *mutable default* pbc list (utils.py:561,568-572): reset it
        ns.utils.radius_graph_pbc.__defaults__[-1][:] = [True, True, True]
        rec = {}
        # the reference formats data.id/sid/fid into its zero-neighbour error (:372-375)
        b.id = b.fid = torch.arange(b.num_graphs)
        sid_names, b.sid = b.sid, torch.arange(b.num_graphs)
        try:
            with torch.no_grad():
                ei, neigh, d, rv, _ = model.generate_graph_values(b.clone())
                f1, f2 = model(b.clone())
        except ValueError as e:
            np.savez_compressed(os.path.join(OUT, f"{name}.npz"), raises=np.array(str(e)[:40]))
            print(name, "raises ValueError")
            continue
        # layer checkpoints through forward hooks on the unmodified modules
        feats = {}
        hooks = []
        state = {}

        def mk(tag):
            def hook(mod, inp, out):
                feats[tag] = (inp, out)
            return hook

        for l in range(model.num_layers):
            hooks.append(model.message_layers[l].register_forward_hook(mk(f"msg{l}")))
            hooks.append(model.update_layers[l].register_forward_hook(mk(f"upd{l}")))
        with torch.no_grad():
            model(b.clone())
        for h in hooks:
            h.remove()
        rows = probe_rows(b.pos.shape[0])
        for l in (CHECKPOINT_LAYERS if name in CHECKPOINT_CASES else ()):
            (x_in, vec_in, *_), (dx, dvec) = feats[f"msg{l}"]
            x = (x_in + dx) * model.inv_sqrt_2
            vec = vec_in + dvec
            rec[f"msg{l}.x"] = x[rows].numpy()
            rec[f"msg{l}.vec"] = vec[rows].numpy()
            (x_in, vec_in), (dx, dvec) = feats[f"upd{l}"]
            sc = getattr(model, f"upd_out_scalar_scale_{l}")
            rec[f"upd{l}.x"] = sc(x_in + dx)[rows].numpy()
            rec[f"upd{l}.vec"] = (vec_in + dvec)[rows].numpy()
        # how far the stock (unstable-sort) reference is from the canonical stable-tie semantics
        g = O.generate_graph_values(b.pos.numpy(), b.cell.numpy(), b.natoms, pbc=pbc or (True, True, True))
        same = tuple(ei.shape) == g["edge_index"].shape and bool((ei.numpy() == g["edge_index"]).all())
        rec.update(
            case=np.array(name), rows=rows,
            edge_index=ei.numpy().astype(np.int32), neighbors=neigh.numpy(),
            dist=d.numpy(), unit_vec=rv.numpy(),
            forces=f1.numpy(), forces2=f2.numpy(),
            n_ties=np.array(g["n_ties"]), stable_equal=np.array(same),
        )
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
        print(f"{name}: N={b.pos.shape[0]} E={ei.shape[1]} ties={g['n_ties']} stock==stable:{same} "
              f"|f1|max={float(f1.abs().max()):.4g}")

    sampler_goldens(ns, model)
    training_goldens(ns)
    com_schedule_goldens(ns)


def run_reference_sampler(ns, model, b, params, seed):
    """Drive the UNMODIFIED `Denoiser.reverse_sde_sampling_rot` on CPU.  Returns (init noise [B,3],
    SDE noise [steps,2,B,3] or None, trajectory [steps,N,3], steps run).  The reference draws its random numbers
    from the global CPU generator (:215 torch.rand, :274-289 torch.normal; the model forward draws none), so the
    draws are reproduced by replaying the same call sequence after the same manual_seed."""
    import ase.io  # the inert shim

    ns.utils.radius_graph_pbc.__defaults__[-1][:] = [True, True, True]
    B, steps = b.num_graphs, params["num_steps"]
    torch.manual_seed(seed)
    noise = torch.rand(B, 3)
    sde = None
    if not params.get("ode", True):
        sde = torch.stack([torch.stack([torch.normal(mean=0, std=1, size=(B, 3)),
                                        torch.normal(mean=0, std=1, size=(B, 3))]) for _ in range(steps)])
    calc = ns.DiffTorchCalc(_FakeTrainer(model))
    den = ns.Denoiser(b, calc, dict(params), device="cpu", traj_dir=None, traj_names=b.sid)
    den.trajectories = [ase.io.Trajectory() for _ in b.sid]
    torch.manual_seed(seed)
    den.reverse_sde_sampling_rot()
    n_run = len(den.trajectories[0].frames)
    traj = [np.concatenate([tr.frames[t].kw["positions"] for tr in den.trajectories], 0) for t in range(n_run)]
    return noise, sde, np.stack(traj).astype(np.float32), n_run


def sampler_goldens(ns, model, only=None):
    from tests.cases import sampler_batch, sampler100_batch

    tamed = S.random_state_dict(0, score_scale=S.SAMPLER_SCORE_SCALE)
    raw = S.random_state_dict(0)
    full = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55)
    jobs = {
        # name: (batch factory, weights, params, seed)
        # tamed output scale: see synthetic.random_state_dict(score_scale=...)
        "sampler": (sampler_batch, tamed, dict(SAMPLER_PARAMS), 1234),
        # BASELINE config #1: ONE system, the shipped 100-step schedule (configs/denoising/painn_so3.yml:74-87)
        "sampler100": (sampler100_batch, tamed, dict(full), 4321),
        # the same with untamed random-init weights: a chaotic map, used step by step (teacher forcing)
        "sampler100_raw": (sampler100_batch, raw, dict(full), 4321),
        # SDE branch (ode=False): translation + rotation noise injected every step
        "sampler_sde": (sampler_batch, tamed, dict(SAMPLER_PARAMS, num_steps=6, ode=False), 99),
    }
    for name, (make, sd, params, seed) in jobs.items():
        if only and name not in only:
            continue
        b = make()
        model.load_state_dict(sd, strict=True)
        noise, sde, traj, n_run = run_reference_sampler(ns, model, b, params, seed)
        rec = dict(noise=noise.numpy(), traj=traj, final=b.pos.numpy(), params=np.array(repr(params)),
                   steps_run=np.array(n_run))
        if sde is not None:
            rec["sde_noise"] = sde.numpy()
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
        print(f"{name}: steps {n_run}, final pos checksum {float(b.pos.double().sum()):.6f}")


# ---------------------------------------------------------------------------------------------------------------
# training side: IGSO(3) tables, noising schedule, loss
# ---------------------------------------------------------------------------------------------------------------
TRAIN_PARAMS = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55,
                    free_std_low=0.01, free_std_high=0.1)
TABLE_ROWS = (0, 137, 500, 863, 999)   # eps rows of the IGSO(3) tables kept in the fixture


def _reference_functions(names):
    """Compile the named top-level functions / methods of the reference's trainer module from the text where it lies
    (the module itself cannot be imported: it pulls in lmdb, wandb, the dataset and registry stack).  Nothing is
    copied into the repo; the functions run once, here, to make the fixtures."""
    import ast

    import torch_scatter  # the shim
    rot_utils = sys.modules["adsorbdiff.utils.rot_utils"]
    path = os.path.join(ref_import.REF_ROOT, "adsorbdiff", "trainers", "sde_denoising_trainer.py")
    src = open(path).read()
    tree = ast.parse(src)
    env = dict(torch=torch, np=np, scatter=torch_scatter.scatter, rot_utils=rot_utils)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), env)
            found[node.name] = env[node.name]
    assert set(found) == set(names), set(names) - set(found)
    return found, rot_utils


def training_goldens(ns):
    import time

    fns, rot_utils = _reference_functions(["pbc_correction", "tr_so3_schedule", "_compute_loss"])
    # the reference's own series code (rot_utils.py:151-188), all 1000 eps: this is what its import does on a
    # machine that has no cache yet (:196-215)
    t0 = time.time()
    eps_array = 10 ** np.linspace(np.log10(rot_utils.MIN_EPS), np.log10(rot_utils.MAX_EPS), rot_utils.N_EPS)
    omegas = np.linspace(0, np.pi, rot_utils.X_N + 1)[1:]
    exp_vals = np.asarray([rot_utils._expansion(omegas, e) for e in eps_array])
    pdf_vals = np.asarray([rot_utils._density(e, omegas, marginal=True) for e in exp_vals])
    cdf_vals = np.asarray([p.cumsum() / rot_utils.X_N * np.pi for p in pdf_vals])
    score_norms = np.asarray([rot_utils._score(exp_vals[i], omegas, eps_array[i]) for i in range(len(eps_array))])
    exp_score_norms = np.sqrt(np.sum(score_norms**2 * pdf_vals, axis=1) / np.sum(pdf_vals, axis=1) / np.pi)
    rot_utils._omegas_array, rot_utils._cdf_vals = omegas, cdf_vals
    rot_utils._score_norms, rot_utils._exp_score_norms = score_norms, exp_score_norms
    print(f"reference IGSO(3) tables: {time.time() - t0:.0f} s")
    rows = np.array(TABLE_ROWS)
    rec = dict(rows=rows, cdf=cdf_vals[rows], score_norms=score_norms[rows], exp_score_norms=exp_score_norms)
    # lookups through the reference's own sample / score_vec / score_norm
    probe_eps = np.array([0.01, 0.0123, 0.05, 0.3, 0.77, 1.55, 2.0, 3.0])
    np.random.seed(7)
    draws_u, samp, svec = [], [], []
    state = np.random.get_state()
    for e in probe_eps:
        samp.append(rot_utils.sample(e))
    np.random.set_state(state)
    for e in probe_eps:
        draws_u.append(np.random.rand())
    vecs = np.random.RandomState(3).randn(len(probe_eps), 3) * np.array([0.01, 0.1, 0.5, 1, 1.5, 2, 2.5, 3.0])[:, None] / 2
    for e, v in zip(probe_eps, vecs):
        svec.append(rot_utils.score_vec(eps=e, vec=v))
    rec.update(probe_eps=probe_eps, probe_u=np.array(draws_u), probe_sample=np.array(samp), probe_vecs=vecs,
               probe_score_vec=np.array(svec), probe_score_norm=rot_utils.score_norm(torch.tensor(probe_eps)).numpy())
    np.savez_compressed(os.path.join(OUT, "igso3.npz"), **rec)

    # noising schedule on the jit2 batch (2 systems) and a 3-system mixed batch
    from tests.cases import CASES
    for case, seed in (("jit2", 11), ("mixed", 12)):
        b = CASES[case][0]()
        pos0 = b.pos.clone()
        B = b.num_graphs
        torch.manual_seed(seed)
        t = torch.rand(B)
        normal = torch.zeros(B, 3).normal_()
        np.random.seed(seed)
        axis, u = [], []
        for _ in range(B):
            axis.append(np.random.randn(3))
            u.append(np.random.rand())
        torch.manual_seed(seed)
        np.random.seed(seed)
        nb = fns["tr_so3_schedule"](b, dict(TRAIN_PARAMS))
        assert torch.equal(pos0[b.tags != 2], nb.pos[b.tags != 2])
        # loss on seeded stand-ins for the two heads (differentiated)
        g = torch.Generator().manual_seed(seed + 100)
        out1 = torch.randn(b.pos.shape[0], 3, generator=g).requires_grad_()
        out2 = torch.randn(b.pos.shape[0], 3, generator=g).requires_grad_()
        fake = type("T", (), {})()
        fake.config = {"optim": {}, "model_attributes": {"so3_denoising": True}}
        fake.device = "cpu"
        out = {"positions": out1 * 1.0, "positions_free": out2 * 1.0}
        loss = fns["_compute_loss"](fake, out, nb)
        loss.backward()
        np.savez_compressed(
            os.path.join(OUT, f"train_{case}.npz"), seed=np.array(seed), t=t.numpy(), normal=normal.numpy(),
            axis=np.array(axis), u=np.array(u), pos=nb.pos.numpy(), tr_sigma=nb.tr_sigma.numpy(),
            rot_sigma=nb.rot_sigma.numpy(), rot_score=nb.rot_score.numpy(), tr_score=nb.tr_score.numpy(),
            ads_center_noise_vec=nb.ads_center_noise_vec.numpy(), out1=out1.detach().numpy(),
            out2=out2.detach().numpy(), loss=loss.detach().numpy(), g_out1=out1.grad.numpy(), g_out2=out2.grad.numpy())
        print(f"train_{case}: loss {float(loss):.6f}")


def com_schedule_goldens(ns):
    """`ads_COM_gaussian_schedule` (sde_denoising_trainer.py:138-177, the translation-only model's noising) run from the
    reference's text on two seeded batches."""
    fns, _ = _reference_functions(["ads_COM_gaussian_schedule"])
    from tests.cases import CASES
    for case, seed in (("jit2", 21), ("mixed", 22)):
        b = CASES[case][0]()
        B = b.num_graphs
        torch.manual_seed(seed)
        t = torch.rand(B)
        normal = torch.zeros(B, 3).normal_()
        torch.manual_seed(seed)
        nb = fns["ads_COM_gaussian_schedule"](b, dict(TRAIN_PARAMS))
        np.savez_compressed(os.path.join(OUT, f"train_com_{case}.npz"), seed=np.array(seed), t=t.numpy(),
                            normal=normal.numpy(), pos=nb.pos.numpy(), tr_sigma=nb.tr_sigma.numpy(),
                            tr_score=nb.tr_score.numpy(), ads_center_noise_vec=nb.ads_center_noise_vec.numpy())
        print(f"train_com_{case}: ok")


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--sampler-only", nargs="*", default=None,
                    help="regenerate only the named sampler fixtures (e.g. sampler100 sampler_sde)")
    ap.add_argument("--training-only", action="store_true", help="regenerate only igso3.npz / train_*.npz")
    ap.add_argument("--com-only", action="store_true", help="regenerate only train_com_*.npz")
    a = ap.parse_args()
    if a.com_only:
        com_schedule_goldens(ref_import.load())
    elif a.training_only:
        training_goldens(ref_import.load())
    elif a.sampler_only is not None:
        ns_ = ref_import.load()
        m_ = ns_.PaiNN(None, 0, 1, scale_file=ns_.scale_file, so3_denoising=True).eval()
        sampler_goldens(ns_, m_, only=a.sampler_only or None)
    else:
        main()
