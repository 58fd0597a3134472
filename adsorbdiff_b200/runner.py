"""Caller side of the sampler: the batch driver of `DenoisingTrainer.run_relaxations`
and the single-structure front door of `AdsorbDiffCalculator.run_diffusion`.

Reference: adsorbdiff/trainers/sde_denoising_trainer.py:750-951 (loop over `relax_loader`, skip batches whose
trajectories exist, `ml_diffuse`, per-rank `relaxed_pos_{rank}.npz`, barrier, rank-0 merge into
`relaxed_positions.npz`), adsorbdiff/utils/utils.py:968-973 (`check_traj_files`),
adsorbdiff/relaxation/calculator.py:180-210 (`run_diffusion`).

What is different, on purpose:
  * the EMA weight swap is made once around the whole job (the reference does it here too, :763-765, and then again
    inside every `predict_denoising` call);
  * batches come from `PackedLoader` (pinned, prefetched) and go to the device with a non-blocking copy, so the GPU
    never waits for collation; many placements of one system are one batch (the reference re-launches the job with
    another `--seed` per placement, run.py:44-55);
  * the ranks' results are merged with one `all_gather_object` instead of files + barrier; the merged file has the
    reference's name and keys (`ids`, `pos`, `chunk_idx`), duplicates removed the same way (np.unique on ids).
"""
from __future__ import annotations

import logging
import os
from pathlib import Path
from typing import Iterable, Optional

import numpy as np
import torch
import torch.distributed as dist

from .denoiser import ml_diffuse, unwrap_model
from .synthetic import SystemBatch


def check_traj_files(batch, traj_dir) -> bool:
    """reference: adsorbdiff/utils/utils.py:968-973 (the deferred writer may also leave `<sid>.npz`)."""
    if traj_dir is None:
        return False
    traj_dir = Path(traj_dir)
    return all((traj_dir / f"{i}.traj").exists() or (traj_dir / f"{i}.npz").exists() for i in batch.sid)


def merge_positions(ids, positions, natoms, results_dir=None, filename="relaxed_positions.npz"):
    """Gather every rank's (ids, positions, natoms) and, on rank 0, write the reference's result file
    (sde_denoising_trainer.py:862-909): ids de-duplicated with np.unique, `pos` concatenated in that order,
    `chunk_idx` = cumsum(natoms)[:-1].  Returns the merged dict on rank 0, None elsewhere."""
    payload = (list(ids), [np.asarray(p, dtype=np.float32) for p in positions], [int(n) for n in natoms])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        gathered = [None] * dist.get_world_size()
        dist.all_gather_object(gathered, payload)
        if dist.get_rank() != 0:
            return None
    else:
        gathered = [payload]
    all_ids = [i for g in gathered for i in g[0]]
    all_pos = [p for g in gathered for p in g[1]]
    all_nat = [n for g in gathered for n in g[2]]
    if not all_ids:
        return {"ids": np.array([]), "pos": np.zeros((0, 3), np.float32), "chunk_idx": np.array([], dtype=np.int64)}
    _, idx = np.unique(np.array(all_ids), return_index=True)
    merged = {"ids": np.array(all_ids)[idx], "pos": np.concatenate([all_pos[i] for i in idx]),
              "chunk_idx": np.cumsum(np.array(all_nat)[idx])[:-1]}
    if results_dir is not None:
        os.makedirs(results_dir, exist_ok=True)
        full = os.path.join(results_dir, filename)
        logging.info(f"Writing results to {full}")
        np.savez_compressed(full, **merged)
    return merged


@torch.no_grad()
def run_diffusion_batches(model, loader: Iterable, denoising_pos_params: dict, device="cuda:0", traj_dir=None,
                          save_full_traj: bool = True, results_dir=None, write_pos: bool = True,
                          num_batches: Optional[int] = None, diffuse=ml_diffuse):
    """The sampling job of one rank over its batches (`DenoisingTrainer.run_relaxations`).  `model` is the trainer
    (anything with `predict_denoising` / `_unwrapped_model` / optional `ema`) or an `adsorbdiff_b200.PaiNN`.
    Returns what `merge_positions` returns."""
    net = unwrap_model(model)
    net.eval()
    ema = getattr(model, "ema", None)
    if ema:
        ema.store()
        ema.copy_to()
        model_for_loop = _NoEma(model)   # the swap above covers the whole job: no second swap per batch
    else:
        model_for_loop = model
    ids, positions, natoms = [], [], []
    try:
        for i, batch in enumerate(loader):
            if num_batches is not None and i >= num_batches:
                break
            if check_traj_files(batch, traj_dir):   # resume: every trajectory of this batch is already there
                logging.info(f"Skipping batch: {list(batch.sid)}")
                continue
            dbatch = batch.to(device, non_blocking=True)
            relaxed = diffuse(batch=dbatch, model=model_for_loop, denoising_pos_params=denoising_pos_params,
                              traj_dir=traj_dir, save_full_traj=save_full_traj, device=device, transform=None)
            if write_pos:
                nat = [int(n) for n in relaxed.natoms.tolist()]
                pos = relaxed.pos.detach().to("cpu", torch.float32)
                positions += [p.numpy() for p in torch.split(pos, nat)]
                natoms += nat
                ids += [str(s) for s in relaxed.sid]
    finally:
        if ema:
            ema.restore()
    return merge_positions(ids, positions, natoms, results_dir) if write_pos else None


class _NoEma:
    """View of a trainer with `ema` hidden (the job-level swap has been made already)."""

    def __init__(self, trainer):
        self._t = trainer

    ema = None

    def __getattr__(self, name):
        return getattr(self._t, name)


def atoms_to_batch(atoms, sid="0") -> SystemBatch:
    """One ASE `Atoms`(-like) object -> a one-system batch with the fields `AtomsToGraphs.convert` +
    `data_list_collater([...], otf_graph=True)` produce (atoms_to_graphs.py:131-198): float32 `atomic_numbers` and
    `tags` (the reference's quirk, :147,153), `fixed` from a FixAtoms-style constraint.  Only duck-typed accessors
    are used, so ASE itself is not required."""
    pos = np.asarray(atoms.get_positions(), dtype=np.float32)
    n = pos.shape[0]
    fixed = np.zeros(n, dtype=np.float32)
    for c in getattr(atoms, "constraints", []) or []:
        idx = c.get_indices() if hasattr(c, "get_indices") else getattr(c, "index", [])
        fixed[np.asarray(idx, dtype=np.int64)] = 1.0
    tags = np.asarray(atoms.get_tags(), dtype=np.float32)
    cell = np.asarray(atoms.get_cell(), dtype=np.float32).reshape(1, 3, 3)
    return SystemBatch(pos=torch.from_numpy(pos), cell=torch.from_numpy(cell),
                       atomic_numbers=torch.from_numpy(np.asarray(atoms.get_atomic_numbers(), dtype=np.float32)),
                       tags=torch.from_numpy(tags), fixed=torch.from_numpy(fixed),
                       natoms=torch.tensor([n], dtype=torch.long), batch=torch.zeros(n, dtype=torch.long), sid=[sid])


@torch.no_grad()
def run_diffusion(atoms, model, denoising_pos_params: dict, trajectory=None, device="cuda:0", placements: int = 1,
                  save_full_traj: bool = True):
    """`AdsorbDiffCalculator.run_diffusion` (calculator.py:180-210): sample a placement for one structure and return
    its final positions [n, 3] (numpy).  With `placements > 1` that many independent placements of the same structure
    run as one batch and the result is [placements, n, 3]."""
    one = atoms_to_batch(atoms)
    batch = one if placements == 1 else type(one).from_data_list([one.clone() for _ in range(placements)])
    if placements > 1:
        batch.sid = [f"0_p{i}" for i in range(placements)]
    ema = getattr(model, "ema", None)
    net = unwrap_model(model)
    net.eval()
    if ema:
        ema.store()
        ema.copy_to()
    try:
        out = ml_diffuse(batch=batch.to(device), model=_NoEma(model) if ema else model,
                         denoising_pos_params=denoising_pos_params, traj_dir=trajectory, save_full_traj=save_full_traj,
                         device=device, transform=None)
    finally:
        if ema:
            ema.restore()
    pos = out.pos.detach().cpu().numpy()
    return pos if placements == 1 else pos.reshape(placements, -1, 3)
