"""Multi-GPU partition of independent adsorbate+slab systems and the final result gather.

Systems never interact, so sampling shards by system with no collective on
the data path: every rank runs its own CUDA-graphed loop on a contiguous block of systems.  The
single collective is the gather of final positions, which replaces the reference's per-rank
`.npz` files + barrier + rank-0 merge (reference: adsorbdiff/trainers/sde_denoising_trainer.py:862-909).
Partitioning mirrors the intent of the reference's atom-count balancing
(reference: adsorbdiff/datasets/data_parallel.py:32-48) but keeps blocks contiguous so that the
gathered tensor is already in input order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def contiguous_partition(natoms: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """[start, end) system ranges per rank, contiguous, balanced by atom count (prefix-sum split)."""
    n = len(natoms)
    total = float(sum(natoms))
    bounds = [0]
    acc = 0.0
    r = 1
    for i, a in enumerate(natoms):
        acc += a
        while r < world_size and acc >= total * r / world_size - 1e-9 and len(bounds) < world_size:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world_size:
        bounds.append(n)
    bounds.append(n)
    return [(bounds[k], max(bounds[k], bounds[k + 1])) for k in range(world_size)]


def initial_noise(num_systems_total: int, seed: int) -> torch.Tensor:
    """Every rank draws the FULL torch.rand(B_total, 3) and slices its rows, so a multi-process run
    uses the same initial placements as a single-process run (reference draw: denoising_torch.py:215)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(num_systems_total, 3, generator=g)


def gather_positions(local_pos: torch.Tensor, atoms_per_rank: Sequence[int]) -> torch.Tensor:
    """All-gather of the ranks' final positions into input order: [sum(atoms), 3].
    NCCL on GPU tensors, gloo on CPU tensors (tests)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_pos
    world = dist.get_world_size()
    assert len(atoms_per_rank) == world and local_pos.shape[0] == atoms_per_rank[dist.get_rank()]
    m = max(atoms_per_rank)
    padded = local_pos.new_zeros(m, 3)
    padded[: local_pos.shape[0]] = local_pos
    out = local_pos.new_empty(world * m, 3)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * m: r * m + atoms_per_rank[r]] for r in range(world)], 0)
