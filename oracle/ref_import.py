"""Import the UNMODIFIED reference (`/root/reference/adsorbdiff`) in this container.

ORACLE / TEST INFRASTRUCTURE ONLY.  Nothing under `adsorbdiff_b200/` may import
this module.  It exists so that `oracle/gen_golden.py` can run the reference's
own PaiNN forward and sampler loop on CPU and freeze their outputs as golden
vectors under `tests/golden/` (the reference tree does not exist on the GPU box).

Recipe:
  1. register an empty module object named `adsorbdiff` (and `adsorbdiff.relaxation`)
     whose `__path__` points into the read-only tree, so the package `__init__`
     files (which pull in ase / lmdb / e3nn) are skipped while every submodule
     still comes from the reference sources;
  2. put `oracle/ref_shims/` on `sys.path` for the absent third-party packages
     (torch_scatter, torch_geometric, matplotlib, ase);
  3. make the four IGSO(3) table loads in `utils/rot_utils.py:189-194` succeed with
     dummy arrays (sampling only uses `axis_angle_to_matrix`).
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def _find_root() -> str:
    """Where the unmodified reference package lives: $ADSORBDIFF_REF, the mounted tree of the build container, or
    the git-ignored `pip install --target baseline/_ref` copy that travels to the GPU box with the snapshot
    (made by `__graft_entry__.build()`; never committed)."""
    for cand in (os.environ.get("ADSORBDIFF_REF"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "adsorbdiff", "models", "painn")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "adsorbdiff", "models", "painn"))


def scale_file():
    """The shipped fitted scale factors: the reference's `.pt` when the source tree is there, else the same six
    values as a dict (the pip-installed package carries no `configs/`; `load_scales_compat` takes a dict too)."""
    path = os.path.join(REF_ROOT, "configs/scaling_factors/painn_nb6_scaling_factors.pt")
    if os.path.exists(path):
        return path
    from adsorbdiff_b200.synthetic import SHIPPED_SCALE_FACTORS

    return {f"upd_out_scalar_scale_{i}": float(v) for i, v in enumerate(SHIPPED_SCALE_FACTORS)}


def _stub_package(name: str, path: str) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod


def load():
    """Returns a namespace with the reference symbols on the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    pkg = os.path.join(REF_ROOT, "adsorbdiff")
    _stub_package("adsorbdiff", pkg)
    _stub_package("adsorbdiff.relaxation", os.path.join(pkg, "relaxation"))
    _stub_package("adsorbdiff.relaxation.diffusers", os.path.join(pkg, "relaxation", "diffusers"))

    import numpy as np

    real_exists, real_load = os.path.exists, np.load
    so3_dir = "/home/jovyan/shared-scratch/adeesh/denoising/so3_precompute/"

    def fake_exists(p):
        return True if str(p).startswith(so3_dir) else real_exists(p)

    def fake_load(p, *a, **k):
        if str(p).startswith(so3_dir):
            return np.zeros((2, 2))
        return real_load(p, *a, **k)

    os.path.exists, np.load = fake_exists, fake_load
    try:
        import adsorbdiff.utils.rot_utils as rot_utils
    finally:
        os.path.exists, np.load = real_exists, real_load

    import adsorbdiff.models.painn.painn_denoising as painn_denoising
    import adsorbdiff.relaxation.diffusers.denoising_torch as denoising_torch
    import adsorbdiff.utils.utils as utils

    ns = types.SimpleNamespace(
        PaiNN=painn_denoising.PaiNN,
        painn_denoising=painn_denoising,
        repeat_blocks=painn_denoising.repeat_blocks,
        radius_graph_pbc=utils.radius_graph_pbc,
        get_pbc_distances=utils.get_pbc_distances,
        get_max_neighbors_mask=utils.get_max_neighbors_mask,
        utils=utils,
        Denoiser=denoising_torch.Denoiser,
        DiffTorchCalc=denoising_torch.DiffTorchCalc,
        axis_angle_to_matrix=rot_utils.axis_angle_to_matrix,
        scale_file=scale_file(),
        root=REF_ROOT,
    )
    return ns
