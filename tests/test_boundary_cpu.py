"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the module mirrors the reference's constructor / state dict, and the product
path refuses to run without CUDA (no fallback)."""
import os
import re

import pytest
import torch

from adsorbdiff_b200 import PaiNN, _cabi, synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "adsorbdiff_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t)\s+(adk_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.adk_abi_version() == _cabi.ABI_VERSION


def test_neighbor_staging_capacity_query():
    lib = _cabi.load()
    assert lib.adk_neighbors_smem_bytes(90, 75, 50) > 0
    assert lib.adk_neighbors_smem_bytes(5000, 75, 50) < 0
    assert lib.adk_neighbors_smem_bytes(90, 5000, 50) < 0
    assert 0 < lib.adk_message_mma_smem_bytes(128, 90) <= 227 * 1024
    assert lib.adk_message_mma_smem_bytes(128, 400) < 0 and lib.adk_message_mma_smem_bytes(100, 90) < 0


def test_state_dict_matches_reference_layout(weights):
    m = PaiNN(None, 0, 1, so3_denoising=True)
    spec = dict(S.state_dict_spec())
    sd = m.state_dict()
    assert set(sd) == set(spec)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert m.num_params == 21451888  # SURVEY.md 8b (probed on the reference)
    m.load_state_dict(weights, strict=True)
    cond = PaiNN(None, 0, 1, so3_denoising=True, energy_encoding="scalar")
    assert cond.num_params == 21451888 + 263680
    assert "atom_emb.embeddings.weight" in m.no_weight_decay()


def test_scale_file_dict_and_unfitted_default():
    m = PaiNN(None, 0, 1, scale_file={"upd_out_scalar_scale_0": 1.25})
    assert float(m.upd_out_scalar_scale_0.scale_factor) == 1.25
    assert float(m.upd_out_scalar_scale_1.scale_factor) == 0.0 and not m.upd_out_scalar_scale_1.fitted


def test_no_cpu_fallback(weights):
    m = PaiNN(None, 0, 1, so3_denoising=True).eval()
    b = S.make_batch(1)
    with pytest.raises(_cabi.AdkError):
        m(b)  # CPU tensors: must fail loudly, not fall back


def test_training_forward_is_refused():
    m = PaiNN(None, 0, 1, so3_denoising=True).train()
    with torch.enable_grad(), pytest.raises(NotImplementedError):
        m(S.make_batch(1))


def test_schedule_table_matches_oracle():
    from adsorbdiff_b200.denoiser import schedule_table
    from oracle import painn_oracle as O

    params = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55)
    tab = schedule_table(params, "cpu")
    for t in (0, 1, 50, 99):
        tr_g, rot_g, dt = O.schedule(t, params)
        assert tab[t, 0].item() == float(0.5 * tr_g**2 * dt)
        assert tab[t, 1].item() == float(dt)
        assert tab[t, 2].item() == float((rot_g**2).float())
