"""World-size-2 gloo test of the multi-GPU host logic (partition, shared noise draw, result gather).
The per-rank compute is replaced by a deterministic stand-in: this test is about the plumbing."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adsorbdiff_b200 import partition as P
from adsorbdiff_b200 import synthetic as S


def test_contiguous_partition_balances_atoms():
    nat = [82] * 10 + [60] * 6 + [90] * 4
    for w in (1, 2, 3, 4, 8):
        parts = P.contiguous_partition(nat, w)
        assert parts[0][0] == 0 and parts[-1][1] == len(nat)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        loads = [sum(nat[a:b]) for a, b in parts]
        assert max(loads) - min(loads) <= 2 * max(nat)
    tiny = P.contiguous_partition([5, 5], 4)  # more ranks than systems: some ranks get nothing
    assert tiny[0][0] == 0 and tiny[-1][1] == 2 and sum(b - a for a, b in tiny) == 2


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = S.make_batch(5)
    nat = batch.natoms.tolist()
    parts = P.contiguous_partition(nat, world)
    a, b = parts[rank]
    offs = [0]
    for x in nat:
        offs.append(offs[-1] + x)
    noise = P.initial_noise(len(nat), seed=7)[a:b]
    # stand-in for the sampler: shift every atom of a system by that system's noise row
    local = batch.pos[offs[a]:offs[b]].clone()
    for s in range(a, b):
        local[offs[s] - offs[a]: offs[s + 1] - offs[a]] += noise[s - a]
    atoms_per_rank = [offs[e] - offs[s] for s, e in parts]
    full = P.gather_positions(local, atoms_per_rank)
    if rank == 0:
        torch.save(full, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(tmp_path):
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, 29731, out), nprocs=2, join=True)
    full = torch.load(out)
    batch = S.make_batch(5)
    noise = P.initial_noise(5, seed=7)
    ref = batch.pos + noise[batch.batch]
    assert torch.equal(full, ref)
