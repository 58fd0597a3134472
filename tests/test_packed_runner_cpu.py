"""Host logic of the packed input format and of the batch driver (no GPU: the sampler is a stand-in)."""
import os

import numpy as np
import pytest
import torch

from adsorbdiff_b200 import packed as PK
from adsorbdiff_b200 import runner as R
from adsorbdiff_b200 import synthetic as S


def _systems(n):
    return [S.make_system(100 + i, adsorbate=("CH3" if i % 3 == 0 else None)) for i in range(n)]


def test_pack_save_load_collate_round_trip(tmp_path):
    sysl = _systems(5)
    for i, s in enumerate(sysl):
        s["sid"] = f"s{i}"
    p = PK.PackedSystems.from_data_list(sysl)
    assert len(p) == 5 and p.natoms.tolist() == [len(s["pos"]) for s in sysl]
    p.save(tmp_path / "pk")
    q = PK.PackedSystems.load(tmp_path / "pk")           # memory-mapped
    assert isinstance(q.pos, np.memmap) and q.sid == [f"s{i}" for i in range(5)]
    b = q.collate([3, 1], pin=False)
    ref = S.collate([sysl[3], sysl[1]], sids=["s3", "s1"])
    for k in ("pos", "cell", "atomic_numbers", "tags", "fixed", "natoms", "batch"):
        assert torch.equal(getattr(b, k), getattr(ref, k)), k
    assert b.sid == ["s3", "s1"]
    # placements: every system repeated back to back, distinct trajectory names
    c = q.collate([2], placements=3, pin=False)
    n = int(q.natoms[2])
    assert c.natoms.tolist() == [n, n, n] and c.sid == ["s2_p0", "s2_p1", "s2_p2"]
    assert torch.equal(c.pos[:n], c.pos[2 * n:]) and c.batch.tolist() == [0] * n + [1] * n + [2] * n


def test_float_attributes_and_tensor_inputs_are_accepted():
    """atoms_to_graphs.py:147,153 writes atomic_numbers / tags as float32 tensors; cell as [1,3,3]."""
    s = S.make_system(7)
    d = dict(pos=torch.from_numpy(s["pos"]), cell=torch.from_numpy(s["cell"])[None], sid=torch.tensor([42]),
             atomic_numbers=torch.from_numpy(s["atomic_numbers"]).float(), tags=torch.from_numpy(s["tags"]).float(),
             fixed=torch.from_numpy(s["fixed"]).float())
    p = PK.PackedSystems.from_data_list([d])
    assert p.atomic_numbers.dtype == np.int64 and p.sid == ["42"] and p.cell.shape == (1, 3, 3)
    assert np.array_equal(p.atomic_numbers, s["atomic_numbers"])


def test_loader_shards_cover_every_system_once_and_prefetch_matches():
    p = PK.PackedSystems.from_data_list(_systems(11))
    seen = []
    for rank in range(3):
        ld = PK.PackedLoader(p, systems_per_batch=2, rank=rank, world_size=3)
        plain = [b.sid for b in PK.PackedLoader(p, 2, rank=rank, world_size=3, prefetch=False)]
        got = [b.sid for b in ld]
        assert got == plain and len(got) == len(ld)
        seen += [s for b in got for s in b]
    assert sorted(seen, key=int) == [str(i) for i in range(11)]


def test_from_lmdb_explains_missing_packages(tmp_path):
    try:
        import lmdb  # noqa: F401
        import torch_geometric  # noqa: F401
        pytest.skip("lmdb and torch_geometric are installed here")
    except ImportError:
        with pytest.raises(ImportError, match="lmdb"):
            PK.PackedSystems.from_lmdb(tmp_path / "x.lmdb")


class _Ema:
    def __init__(self):
        self.log = []

    def store(self): self.log.append("store")
    def copy_to(self): self.log.append("copy_to")
    def restore(self): self.log.append("restore")
    def __bool__(self): return True


class _Trainer:
    def __init__(self):
        self.ema = _Ema()
        self._unwrapped_model = torch.nn.Linear(1, 1)

    def predict_denoising(self, *a, **k):
        raise AssertionError("stand-in sampler never calls the model")


def test_batch_driver_skips_finished_batches_swaps_ema_once_and_merges_like_the_reference(tmp_path):
    p = PK.PackedSystems.from_data_list(_systems(6))
    loader = PK.PackedLoader(p, systems_per_batch=2, prefetch=False)
    traj = tmp_path / "traj"
    traj.mkdir()
    for sid in ("2", "3"):                       # batch 1 already has its trajectories: resume skips it
        (traj / f"{sid}.traj").write_bytes(b"x")
    trainer, calls = _Trainer(), []

    def fake_diffuse(batch, model, denoising_pos_params, traj_dir, save_full_traj, device, transform):
        assert model.ema is None                 # the job-level swap is not repeated per batch
        calls.append(list(batch.sid))
        batch.pos = batch.pos + 1.0
        return batch

    out = R.run_diffusion_batches(trainer, loader, dict(num_steps=1), device="cpu", traj_dir=traj,
                                  results_dir=tmp_path / "res", diffuse=fake_diffuse)
    assert calls == [["0", "1"], ["4", "5"]]
    assert trainer.ema.log == ["store", "copy_to", "restore"]
    f = np.load(tmp_path / "res" / "relaxed_positions.npz")
    assert f["ids"].tolist() == ["0", "1", "4", "5"] == out["ids"].tolist()
    nat = p.natoms
    assert f["chunk_idx"].tolist() == np.cumsum(nat[[0, 1, 4, 5]])[:-1].tolist()
    parts = np.split(f["pos"], f["chunk_idx"])   # how the reference's consumers read it back
    o = p.offsets
    for part, i in zip(parts, (0, 1, 4, 5)):
        np.testing.assert_array_equal(part, p.pos[o[i]:o[i + 1]] + 1.0)


def test_atoms_to_batch_uses_the_reference_field_conventions():
    s = S.make_system(3)

    class FixAtoms:
        def __init__(self, idx): self.idx = idx
        def get_indices(self): return self.idx

    class Atoms:   # the accessors of ase.Atoms the converter touches
        constraints = [FixAtoms(np.nonzero(s["fixed"])[0])]
        def get_positions(self): return s["pos"].astype(np.float64)
        def get_cell(self): return s["cell"].astype(np.float64)
        def get_atomic_numbers(self): return s["atomic_numbers"]
        def get_tags(self): return s["tags"]

    b = R.atoms_to_batch(Atoms())
    assert b.atomic_numbers.dtype == torch.float32 and b.tags.dtype == torch.float32   # atoms_to_graphs.py:147,153
    assert torch.equal(b.fixed.long(), torch.from_numpy(s["fixed"])) and b.cell.shape == (1, 3, 3)
    assert b.natoms.tolist() == [len(s["pos"])] and b.sid == ["0"]
