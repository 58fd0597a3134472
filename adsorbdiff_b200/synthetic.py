"""Synthetic OC20-Dense-shaped adsorbate+slab systems.

No dataset can be fetched in this environment, so the benchmark and the parity
tests run on generated systems with the shape of what `AdsorbateSlabConfig`
produces (reference: adsorbdiff/placement/adsorbate_slab_config.py:22-457) and
what `data_list_collater` hands to the model
(reference: adsorbdiff/datasets/lmdb_dataset.py:246-263): a flat attribute bag
with `pos, cell, atomic_numbers, natoms, tags, fixed, batch, sid`.

`SystemBatch` deliberately mirrors the small part of the PyG `Batch` interface
the hot path touches (attribute access, `in`, `.to(device)`), so the model and
sampler accept either a real PyG `Batch` or this bag.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SLAB_ELEMENTS = (29, 78, 28, 47, 46, 79, 45, 27)  # Cu Pt Ni Ag Pd Au Rh Co  (all <= 83)
_ADSORBATES = {
    "CO": ([6, 8], [[0.0, 0.0, 0.0], [0.0, 0.0, 1.15]]),
    "OH": ([8, 1], [[0.0, 0.0, 0.0], [0.0, 0.0, 0.97]]),
    # larger fragments for the 2-10 atom range of the north star
    "CH3": ([6, 1, 1, 1], [[0, 0, 0], [1.03, 0, 0.36], [-0.51, 0.89, 0.36], [-0.51, -0.89, 0.36]]),
    "CHOHCH3": (
        [6, 1, 8, 1, 6, 1, 1, 1],
        [[0, 0, 0], [0.0, 1.0, 0.4], [1.2, -0.4, 0.6], [1.9, 0.2, 0.8],
         [-1.3, -0.6, 0.7], [-1.2, -1.6, 1.1], [-2.1, -0.5, 0.0], [-1.6, 0.0, 1.6]],
    ),
}


class SystemBatch:
    """Attribute bag standing in for `torch_geometric.data.Batch` on the hot path."""

    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)

    def __contains__(self, key):
        return hasattr(self, key)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, non_blocking: bool = False):
        # PyG's Batch.to is in-place-and-return-self for the attributes; same here.
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device, non_blocking=non_blocking))
        return self

    def clone(self):
        out = SystemBatch()
        for k, v in self.__dict__.items():
            setattr(out, k, v.clone() if isinstance(v, torch.Tensor) else v)
        return out

    @property
    def num_graphs(self):
        return int(self.natoms.shape[0])

    _PER_ATOM = ("pos", "atomic_numbers", "tags", "fixed", "force")
    _PER_SYSTEM = ("cell", "natoms", "y", "pbc")

    def to_data_list(self):
        """Single-system batches, like PyG's `Batch.to_data_list` (used by ml_diffuse's split-and-retry)."""
        nat = [int(n) for n in self.natoms.tolist()]
        off = [0]
        for n in nat:
            off.append(off[-1] + n)
        out = []
        for i, n in enumerate(nat):
            f = {}
            for k, v in self.__dict__.items():
                if not isinstance(v, torch.Tensor):
                    continue
                if k in self._PER_ATOM and v.shape[0] == off[-1]:
                    f[k] = v[off[i]:off[i + 1]].clone()
                elif k in self._PER_SYSTEM and v.shape[0] == len(nat):
                    f[k] = v[i:i + 1].clone()
            f["batch"] = torch.zeros(n, dtype=torch.long, device=self.pos.device)
            f["sid"] = [self.sid[i]]
            out.append(SystemBatch(**f))
        return out

    @classmethod
    def from_data_list(cls, data_list):
        """Concatenate batches (single systems or finished multi-system batches) in order."""
        keys = [k for k, v in data_list[0].__dict__.items() if isinstance(v, torch.Tensor) and k != "batch"]
        f = {k: torch.cat([getattr(d, k) for d in data_list], 0) for k in keys
             if all(hasattr(d, k) for d in data_list)}
        f["batch"] = torch.repeat_interleave(torch.arange(f["natoms"].shape[0], device=f["pos"].device), f["natoms"])
        f["sid"] = [s for d in data_list for s in d.sid]
        return cls(**f)


def make_system(system_id: int, adsorbate: str | None = None, jitter: float = 0.05,
                size=(4, 4, 5), height: float = 2.0, skew: float = 0.0):
    """One slab+adsorbate system as numpy arrays.

    fcc(111)-like `size` lattice, nearest neighbour 2.8 A, layer spacing 2.3 A,
    ABC stacking; cell rows a=(nx*2.8,0,0), b=(ny*1.4, ny*2.425, 0), c=(skew,skew,32).
    Seed = 1000 + system_id (jitter, element, adsorbate site).
    """
    rng = np.random.RandomState(1000 + system_id)
    nx, ny, nl = size
    a1 = np.array([2.8, 0.0, 0.0])
    a2 = np.array([1.4, 2.425, 0.0])
    z0 = 8.0
    pos, tags, fixed = [], [], []
    for l in range(nl):
        shift = (l % 3) * (a1 + a2) / 3.0
        for iy in range(ny):
            for ix in range(nx):
                p = ix * a1 + iy * a2 + shift
                p[2] = z0 + 2.3 * l
                pos.append(p)
                top = l >= nl - 2
                tags.append(1 if top else 0)
                fixed.append(0 if top else 1)
    pos = np.array(pos)
    z_slab = SLAB_ELEMENTS[rng.randint(len(SLAB_ELEMENTS))]
    numbers = [z_slab] * len(pos)
    if jitter > 0:
        pos = pos + rng.normal(0.0, jitter, size=pos.shape)
    if adsorbate is None:
        adsorbate = "CO" if system_id % 2 == 0 else "OH"
    ads_z, ads_rel = _ADSORBATES[adsorbate]
    ads_rel = np.array(ads_rel, dtype=np.float64)
    site = rng.rand() * nx * a1 + rng.rand() * ny * a2
    site[2] = z0 + 2.3 * (nl - 1) + height
    ads_pos = site[None, :] + ads_rel
    pos = np.concatenate([pos, ads_pos], 0)
    numbers += list(ads_z)
    tags += [2] * len(ads_z)
    fixed += [0] * len(ads_z)
    cell = np.array([nx * a1, ny * a2, [skew, skew, 32.0]])
    return dict(
        pos=pos.astype(np.float32),
        cell=cell.astype(np.float32),
        atomic_numbers=np.array(numbers, dtype=np.int64),
        tags=np.array(tags, dtype=np.int64),
        fixed=np.array(fixed, dtype=np.int64),
    )


def collate(systems, sids=None, float_attrs: bool = False) -> SystemBatch:
    """Concatenate systems the way `Batch.from_data_list` does (flat atoms, `batch` vector).

    `float_attrs=True` reproduces the ASE front door, where atomic_numbers/tags
    arrive as float32 (reference: adsorbdiff/utils/atoms_to_graphs.py:147,153).
    """
    natoms = torch.tensor([len(s["pos"]) for s in systems], dtype=torch.long)
    cat = lambda k, dt: torch.from_numpy(np.concatenate([s[k] for s in systems], 0)).to(dt)
    idt = torch.float32 if float_attrs else torch.long
    b = SystemBatch(
        pos=cat("pos", torch.float32),
        cell=torch.from_numpy(np.stack([s["cell"] for s in systems], 0)).float(),
        atomic_numbers=cat("atomic_numbers", idt),
        tags=cat("tags", idt),
        fixed=cat("fixed", idt),
        natoms=natoms,
        batch=torch.repeat_interleave(torch.arange(len(systems)), natoms),
        sid=list(sids) if sids is not None else [f"sys{i}" for i in range(len(systems))],
    )
    return b


def make_batch(num_systems: int, first_id: int = 0, **kw) -> SystemBatch:
    systems = [make_system(first_id + i, **kw) for i in range(num_systems)]
    return collate(systems, sids=[f"sys{first_id + i}" for i in range(num_systems)])


def make_placements(system_id: int, num_placements: int, **kw) -> SystemBatch:
    """BASELINE config #2: one system replicated `num_placements` times; the sampler's
    `torch.rand(B,3)` rows then give each copy an independent initial placement."""
    s = make_system(system_id, **kw)
    return collate([s] * num_placements, sids=[f"sys{system_id}_p{i}" for i in range(num_placements)])


# Fitted values shipped in the reference's configs/scaling_factors/painn_nb6_scaling_factors.pt
# (keys upd_out_scalar_scale_{0..5}; read with torch.load in the build container).
SHIPPED_SCALE_FACTORS = (1.0364354848861694, 0.8951448202133179, 0.8934778571128845,
                         0.8899308443069458, 0.8886106610298157, 0.8822302222251892)


def state_dict_spec(hidden=512, num_layers=6, num_rbf=128, num_elements=83, so3_denoising=True):
    """(key, shape) list of the reference PaiNN state dict (reference:
    adsorbdiff/models/painn/painn_denoising.py:99-148), in registration order."""
    h = hidden
    spec = [("atom_radii", (101,)), ("atom_emb.embeddings.weight", (num_elements, h)),
            ("radial_basis.rbf.offset", (num_rbf,))]
    for i in range(num_layers):
        m, u = f"message_layers.{i}", f"update_layers.{i}"
        spec += [(m + ".x_proj.0.weight", (h, h)), (m + ".x_proj.0.bias", (h,)),
                 (m + ".x_proj.2.weight", (3 * h, h)), (m + ".x_proj.2.bias", (3 * h,)),
                 (m + ".rbf_proj.weight", (3 * h, num_rbf)), (m + ".rbf_proj.bias", (3 * h,)),
                 (m + ".x_layernorm.weight", (h,)), (m + ".x_layernorm.bias", (h,)),
                 (u + ".vec_proj.weight", (2 * h, h)),
                 (u + ".xvec_proj.0.weight", (h, 2 * h)), (u + ".xvec_proj.0.bias", (h,)),
                 (u + ".xvec_proj.2.weight", (3 * h, h)), (u + ".xvec_proj.2.bias", (3 * h,)),
                 (f"upd_out_scalar_scale_{i}.scale_factor", ())]
    spec += [("out_energy.0.weight", (h // 2, h)), ("out_energy.0.bias", (h // 2,)),
             ("out_energy.2.weight", (1, h // 2)), ("out_energy.2.bias", (1,))]
    heads = ["out_forces"] + (["out_forces2"] if so3_denoising else [])
    for hd in heads:
        for blk, (cin, cout) in enumerate(((h, h // 2), (h // 2, 1))):
            p = f"{hd}.output_network.{blk}"
            spec += [(p + ".vec1_proj.weight", (cin, cin)), (p + ".vec2_proj.weight", (cout, cin)),
                     (p + ".update_net.0.weight", (cin, 2 * cin)), (p + ".update_net.0.bias", (cin,)),
                     (p + ".update_net.2.weight", (2 * cout, cin)), (p + ".update_net.2.bias", (2 * cout,))]
    return spec


SAMPLER_SCORE_SCALE = 0.002  # see random_state_dict(score_scale=...)


def random_state_dict(seed=0, score_scale=1.0, **arch):
    """Random weights at the shipped architecture, keyed like the reference state dict.

    Independent of module construction order so the reference model, the oracle and the
    CUDA model can all be loaded with bit-identical values on any machine: each tensor is
    drawn from its own `torch.Generator` seeded by (seed, position in the spec).  Weights are
    Xavier-uniform like the reference's `reset_parameters`; biases / LayerNorm affine get small
    non-trivial values (the reference zero/one-initialises them, a trained checkpoint does not).

    `score_scale` multiplies the last linear of both output heads (the model output is linear in
    it).  Random-init scores are O(10), ~100x a trained model's at t=1, which turns the sampler
    (step = 57.6 * score at t=1) into a chaotic map that amplifies fp32 rounding noise by orders of
    magnitude per step; trajectory-parity tests use score_scale=SAMPLER_SCORE_SCALE so that the
    comparison measures the implementation, not the Lyapunov exponent of a random network.
    """
    sd = {}
    num_rbf = arch.get("num_rbf", 128)
    for idx, (key, shape) in enumerate(state_dict_spec(**arch)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if key == "atom_radii":
            t = torch.rand(shape, generator=g) + 0.5
        elif key == "radial_basis.rbf.offset":
            t = torch.linspace(0.0, 1.0, num_rbf)
        elif key.endswith("scale_factor"):
            layer = int(key.split(".")[0].rsplit("_", 1)[1])
            t = torch.tensor(SHIPPED_SCALE_FACTORS[layer % len(SHIPPED_SCALE_FACTORS)])
        elif key == "atom_emb.embeddings.weight":
            t = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(3.0)
        elif key.endswith("x_layernorm.weight"):
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif key.endswith(".bias"):
            t = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        else:
            fan_out, fan_in = shape
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        if score_scale != 1.0 and ".output_network.1.update_net.2." in key:
            t = t * score_scale
        sd[key] = t.float()
    return sd
