"""Training-side host logic against fixtures made by the reference's own functions
(oracle/gen_golden.py::training_goldens: rot_utils._expansion/_score/sample/score_vec/score_norm,
sde_denoising_trainer.tr_so3_schedule/_compute_loss executed from the reference tree) and against the oracle
restatement.  Everything here is torch on the CPU: the kernels are not involved."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adsorbdiff_b200 import train as T
from oracle import train_oracle as TO
from tests.cases import CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PARAMS = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55,
              free_std_low=0.01, free_std_high=0.1)


@pytest.fixture(scope="module")
def tables():
    return T.IGSO3Tables("cpu")


@pytest.fixture(scope="module")
def igso3():
    return np.load(os.path.join(GOLDEN, "igso3.npz"))


def _close(a, b, rtol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
    assert err <= rtol, f"{what}: relative-to-max error {err:.3e} > {rtol}"


def test_tables_match_reference_series(tables, igso3):
    rows = igso3["rows"]
    for r, cdf, sn in zip(rows, igso3["cdf"], igso3["score_norms"]):
        _close(tables.cdf[r].numpy(), cdf, 1e-10, f"cdf row {r}")
        # score = d(series)/series: where the density has decayed to rounding noise of the series (narrow eps, large
        # omega) both sides are noise divided by noise; compare where the density is above 1e-6 of its peak
        pdf = np.diff(cdf, prepend=0.0)
        live = pdf > 1e-6 * pdf.max()
        assert live.sum() >= 20
        np.testing.assert_allclose(tables.score_norms[r].numpy()[live], sn[live], rtol=1e-6, err_msg=f"score_norms row {r}")
    np.testing.assert_allclose(tables.exp_score_norms.numpy(), igso3["exp_score_norms"], rtol=1e-9)


def test_oracle_series_match_reference_rows(igso3):
    rows = igso3["rows"][[0, 2, 4]]
    ref = TO.igso3_tables(rows=rows)
    for k, r in enumerate([0, 2, 4]):
        np.testing.assert_allclose(ref["cdf"][k], igso3["cdf"][r], rtol=1e-12, atol=1e-300)
        pdf = np.diff(igso3["cdf"][r], prepend=0.0)
        live = pdf > 1e-6 * pdf.max()
        np.testing.assert_allclose(ref["score_norms"][k][live], igso3["score_norms"][r][live], rtol=1e-12)
        np.testing.assert_allclose(ref["exp_score_norms"][k], igso3["exp_score_norms"][rows[k]], rtol=1e-12)


def test_lookups_match_reference(tables, igso3):
    eps = torch.tensor(igso3["probe_eps"])
    B = eps.shape[0]
    got = tables._interp(torch.tensor(igso3["probe_u"]), tables.cdf[tables.eps_index(eps)], tables.omegas[None].expand(B, -1))
    np.testing.assert_allclose(got.numpy(), igso3["probe_sample"], rtol=1e-9, atol=1e-12)
    sv = tables.score_vec(eps, torch.tensor(igso3["probe_vecs"]))
    np.testing.assert_allclose(sv.numpy(), igso3["probe_score_vec"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(tables.score_norm(eps).numpy(), igso3["probe_score_norm"], rtol=1e-6)


def test_small_tables_match_oracle_loops():
    kw = dict(min_eps=0.2, max_eps=1.5, n_eps=6, x_n=40, L=300)
    ref = TO.igso3_tables(**kw)
    got = T.IGSO3Tables("cpu", **kw)
    _close(got.cdf.numpy(), ref["cdf"], 1e-11, "cdf")
    pdf = np.diff(ref["cdf"], axis=1, prepend=0.0)
    live = pdf > 1e-6 * pdf.max(axis=1, keepdims=True)
    np.testing.assert_allclose(got.score_norms.numpy()[live], ref["score_norms"][live], rtol=1e-7)
    np.testing.assert_allclose(got.exp_score_norms.numpy(), ref["exp_score_norms"], rtol=1e-8)
    for e in (0.02, 0.2, 0.9, 1.5, 4.0):
        assert int(got.eps_index(torch.tensor([e]))[0]) == int(TO.eps_index(ref, e))
        v = np.array([0.3, -0.2, 0.5])
        np.testing.assert_allclose(got.score_vec(torch.tensor([e]), torch.tensor(v)[None])[0].numpy(),
                                   TO.score_vec(ref, e, v), rtol=1e-9)


def test_pbc_correction_matches_oracle():
    g = torch.Generator().manual_seed(0)
    cell = torch.eye(3)[None].repeat(4, 1, 1) * 9 + torch.randn(4, 3, 3, generator=g)
    v = torch.randn(4, 3, generator=g) * 20
    got = T.pbc_correction(v, cell).numpy()
    np.testing.assert_allclose(got, TO.pbc_correction(v.numpy(), cell.numpy()), atol=2e-5)


@pytest.mark.parametrize("case", ["jit2", "mixed"])
def test_noising_and_loss_match_reference(tables, case):
    ref = np.load(os.path.join(GOLDEN, f"train_{case}.npz"))
    b = CASES[case][0]()
    pos0 = b.pos.clone()
    draws = dict(t=torch.tensor(ref["t"]), normal=torch.tensor(ref["normal"]), axis=torch.tensor(ref["axis"]),
                 u=torch.tensor(ref["u"]))
    nb = T.tr_so3_schedule(b, PARAMS, tables, draws=draws)
    assert torch.equal(nb.pos[b.tags != 2], pos0[b.tags != 2])   # only the adsorbate moves
    np.testing.assert_allclose(nb.tr_sigma.numpy(), ref["tr_sigma"], rtol=1e-6)
    np.testing.assert_allclose(nb.rot_sigma.numpy(), ref["rot_sigma"], rtol=1e-6)
    np.testing.assert_allclose(nb.ads_center_noise_vec.numpy(), ref["ads_center_noise_vec"], atol=2e-5)
    np.testing.assert_allclose(nb.tr_score.numpy(), ref["tr_score"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(nb.rot_score.numpy(), ref["rot_score"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nb.pos.numpy(), ref["pos"], atol=2e-5)
    # loss and its gradient w.r.t. the two heads, on the reference's noised batch
    for k in ("tr_sigma", "rot_sigma", "rot_score", "tr_score"):
        setattr(nb, k, torch.tensor(ref[k]))
    out1 = torch.tensor(ref["out1"]).requires_grad_()
    out2 = torch.tensor(ref["out2"]).requires_grad_()
    loss = T.denoising_loss((out1, out2), nb, tables)
    loss.backward()
    np.testing.assert_allclose(float(loss), float(ref["loss"]), rtol=1e-5)
    np.testing.assert_allclose(out1.grad.numpy(), ref["g_out1"], rtol=1e-4, atol=1e-7 * np.abs(ref["g_out1"]).max())
    np.testing.assert_allclose(out2.grad.numpy(), ref["g_out2"], rtol=1e-4, atol=1e-7 * np.abs(ref["g_out2"]).max())
    # the oracle restatement of _compute_loss agrees with the reference too
    o1, o2 = torch.tensor(ref["out1"]), torch.tensor(ref["out2"])
    lo = TO.compute_loss(o1, o2, nb.tags, nb.batch, nb.tr_sigma, nb.rot_sigma, nb.tr_score, nb.rot_score,
                         torch.tensor(TO.score_norm(dict(exp_score_norms=tables.exp_score_norms.numpy(), min_eps=0.01, max_eps=2.0, n_eps=1000), nb.rot_sigma.numpy())))
    np.testing.assert_allclose(float(lo), float(ref["loss"]), rtol=1e-5)


@pytest.mark.parametrize("case", ["jit2", "mixed"])
def test_com_only_noising_matches_reference(case):
    """`ads_COM_gaussian_schedule` (the so3_denoising=False variant) with the reference's draws injected."""
    ref = np.load(os.path.join(GOLDEN, f"train_com_{case}.npz"))
    b = CASES[case][0]()
    pos0 = b.pos.clone()
    nb = T.ads_com_gaussian_schedule(b, PARAMS, draws=dict(t=torch.tensor(ref["t"]), normal=torch.tensor(ref["normal"])))
    assert torch.equal(nb.pos[b.tags != 2], pos0[b.tags != 2])
    np.testing.assert_allclose(nb.tr_sigma.numpy(), ref["tr_sigma"], rtol=1e-6)
    np.testing.assert_allclose(nb.ads_center_noise_vec.numpy(), ref["ads_center_noise_vec"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(nb.tr_score.numpy(), ref["tr_score"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nb.pos.numpy(), ref["pos"], atol=2e-5)


def test_host_found_adsorbate_rows_give_the_same_noising_and_loss(tables):
    """`TrainStep.to_device` finds the adsorbate rows on the host copy and hands them over as index tensors
    (`_ads_idx`, `_ads_seg`), so that noising and loss index with them instead of a boolean mask (whose size the host
    would have to wait for).  Same draws => the same noised batch and the same loss, bit for bit."""
    outs = []
    for with_rows in (False, True):
        b = CASES["mixed"][0]()
        if with_rows:
            b._ads_idx = torch.nonzero(b.tags == 2).flatten()
            b._ads_seg = b.batch.index_select(0, b._ads_idx)
        idx, seg = T._adsorbate_rows(b)
        assert torch.equal(idx, torch.nonzero(b.tags == 2).flatten()) and torch.equal(seg, b.batch[b.tags == 2])
        nb = T.tr_so3_schedule(b, PARAMS, tables, generator=torch.Generator().manual_seed(11))
        g = torch.Generator().manual_seed(12)
        out = (torch.randn(nb.pos.shape[0], 3, generator=g), torch.randn(nb.pos.shape[0], 3, generator=g))
        outs.append((nb.pos.clone(), nb.tr_score.clone(), nb.rot_score.clone(), T.denoising_loss(out, nb, tables)))
    for a, c in zip(*outs):
        assert torch.equal(a, c)


def test_rbf_weight_scale_is_a_power_of_two_with_headroom():
    """Prescale of the message weights for the tcgen05 forward of the training step: s * max|w| in (1024, 2048]."""
    class Net:
        pass
    net = Net()
    keep = []   # (the table is keyed on the weight's storage: keep the tensors alive, as parameters are)
    for k, mag in enumerate((3e-4, 0.07, 1.0, 55.0)):
        w = torch.randn(1536, 128, generator=torch.Generator().manual_seed(k)) * mag
        keep.append(w)
        s = T.rbf_weight_scale(net, w)
        assert math.log2(s) == round(math.log2(s)) and 1024 < s * float(w.abs().max()) <= 2048
        assert T.rbf_weight_scale(net, w) == s
    w = torch.full((4, 4), 0.01)
    s0 = T.rbf_weight_scale(net, w)
    w.mul_(64)                                   # the weight grew: the cached scale stays until a refresh
    assert T.rbf_weight_scale(net, w) == s0 and T.rbf_weight_scale(net, w, refresh=True) == s0 / 64
    assert T.rbf_weight_scale(net, torch.zeros(3, 3)) == 1.0


def test_schedule_statistics(tables):
    """Size-independent properties of the random path: sigma ranges, xy-only translation, rigid adsorbate."""
    b = CASES["mixed"][0]()
    pos0 = b.pos.clone()
    g = torch.Generator().manual_seed(5)
    nb = T.tr_so3_schedule(b, PARAMS, tables, generator=g)
    assert (nb.tr_sigma >= 0.1).all() and (nb.tr_sigma <= 10).all()
    assert (nb.rot_sigma >= 0.01).all() and (nb.rot_sigma <= 1.55).all()
    assert (nb.ads_center_noise_vec[:, 2] == 0).all()
    ads = b.tags == 2
    for s in range(b.num_graphs):
        m = ads & (b.batch == s)
        d0 = torch.cdist(pos0[m], pos0[m])
        d1 = torch.cdist(nb.pos[m], nb.pos[m])
        assert torch.allclose(d0, d1, atol=1e-4)   # rotation + translation: internal distances unchanged
        shift = nb.pos[m].mean(0) - pos0[m].mean(0)
        assert torch.allclose(shift, nb.ads_center_noise_vec[s] + torch.tensor([0, 0, 1.0]), atol=1e-4)


def _ar_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(rank)
    grads = [torch.randn(3, 5, generator=g), torch.randn(7, generator=g), torch.randn(2, 2, 2, generator=g)]
    T.allreduce_mean_(grads)
    if rank == 0:
        torch.save(grads, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_ar_worker, args=(2, 29741, out), nprocs=2, join=True)
    got = torch.load(out)
    want = []
    for shape in ((3, 5), (7,), (2, 2, 2)):
        want.append(None)
    gens = [torch.Generator().manual_seed(r) for r in range(2)]
    per_rank = [[torch.randn(3, 5, generator=g), torch.randn(7, generator=g), torch.randn(2, 2, 2, generator=g)] for g in gens]
    for i in range(3):
        assert torch.allclose(got[i], (per_rank[0][i] + per_rank[1][i]) / 2)


def test_operand_space_formula_covers_the_library():
    """`_TcWorkspace.operand_space` sizes the scratch in Python; it must never be below what the C side carves."""
    from adsorbdiff_b200 import _cabi

    lib = _cabi.load()
    for M, K, N in [(1, 64, 64), (77, 1024, 512), (3700, 512, 1536), (11111, 512, 1024), (128, 256, 256), (129, 2048, 2048)]:
        mp, np_, kr = T._pad(M, 128), T._pad(N, 128), T._pad(M, 64)
        need = 4 * max(mp * K + N * K, mp * N + K * N + np_ * kr + K * kr) + 2048
        assert need >= lib.adk_linear_train_ws_bytes(M, K, N)


def _sync_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Fake:   # the two attributes params_in_sync reads
        world = 2
        params = [torch.nn.Parameter(torch.arange(6.0).reshape(2, 3)), torch.nn.Parameter(torch.ones(4))]

    same = T.TrainStep.params_in_sync(Fake())
    Fake.params[1].data[2] += 1e-6 * rank          # rank 1 drifts by one element
    differ = T.TrainStep.params_in_sync(Fake())
    if rank == 0:
        torch.save((same, differ), out)
    dist.barrier()
    dist.destroy_process_group()


def test_params_in_sync_detects_a_drifting_rank(tmp_path):
    out = str(tmp_path / "s.pt")
    mp.spawn(_sync_worker, args=(2, 29743, out), nprocs=2, join=True)
    same, differ = torch.load(out)
    assert same is True and differ is False
