"""Packed input format for the sampler.

The reference feeds the sampler from an LMDB of pickled PyG `Data` objects, unpickled one by one by DataLoader
workers and collated with `Batch.from_data_list` (reference: adsorbdiff/datasets/lmdb_dataset.py:124-162, 246-263;
fields written by scripts/create_lmdbs/pred_traj_to_lmdb.py:77-92).  Here the same fields live in flat arrays --
one `.npy` per field in a directory, memory-mappable -- and a batch is a handful of slices copied into PINNED host
tensors, so collation costs microseconds and the host-to-device copy of the next batch overlaps the GPU's work on
the current one (`PackedLoader`).  The batch that comes out has the attribute set `data_list_collater(...,
otf_graph=True)` produces (`pos, cell, atomic_numbers, natoms, tags, fixed, batch, sid`).

`from_lmdb` converts a reference LMDB; it needs the `lmdb` and `torch_geometric` packages (to unpickle `Data`), which
this image does not have -- it raises ImportError with that explanation when they are absent.
"""
from __future__ import annotations

import json
import os
import threading
from pathlib import Path
from queue import Queue
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from .partition import contiguous_partition
from .synthetic import SystemBatch

_FIELDS = ("pos", "atomic_numbers", "tags", "fixed")   # per atom
_DTYPES = dict(pos=np.float32, atomic_numbers=np.int64, tags=np.int64, fixed=np.int64, cell=np.float32)


def _get(obj, key):
    v = obj[key] if isinstance(obj, dict) else getattr(obj, key)
    return v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)


class PackedSystems:
    """Flat storage of S systems: per-atom arrays concatenated, `offsets[S + 1]` delimiting them."""

    def __init__(self, pos, atomic_numbers, tags, fixed, cell, offsets, sid):
        self.pos, self.atomic_numbers, self.tags, self.fixed = pos, atomic_numbers, tags, fixed
        self.cell, self.offsets, self.sid = cell, offsets, list(sid)
        assert len(self.sid) == len(self.offsets) - 1 == self.cell.shape[0]

    def __len__(self) -> int:
        return len(self.sid)

    @property
    def natoms(self) -> np.ndarray:
        return np.diff(self.offsets)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_data_list(cls, data_list: Iterable) -> "PackedSystems":
        """From `Data`-like objects or dicts with the reference's field names; `cell` may be [3,3] or [1,3,3],
        `atomic_numbers` / `tags` / `fixed` may be float (atoms_to_graphs.py:147,153 writes float32)."""
        cols = {k: [] for k in _FIELDS}
        cells, sids, offs = [], [], [0]
        for i, d in enumerate(data_list):
            p = _get(d, "pos").astype(np.float32).reshape(-1, 3)
            cols["pos"].append(p)
            for k in ("atomic_numbers", "tags", "fixed"):
                cols[k].append(np.rint(_get(d, k)).astype(np.int64).reshape(-1))
                if cols[k][-1].shape[0] != p.shape[0]:
                    raise ValueError(f"system {i}: `{k}` has {cols[k][-1].shape[0]} entries for {p.shape[0]} atoms")
            cells.append(_get(d, "cell").astype(np.float32).reshape(3, 3))
            try:
                s = _get(d, "sid")
                s = s.reshape(-1)[0] if isinstance(s, np.ndarray) and s.ndim else s
                sids.append(str(s.item() if hasattr(s, "item") else s))
            except (KeyError, AttributeError):
                sids.append(str(i))
            offs.append(offs[-1] + p.shape[0])
        return cls(np.concatenate(cols["pos"]), np.concatenate(cols["atomic_numbers"]), np.concatenate(cols["tags"]),
                   np.concatenate(cols["fixed"]), np.stack(cells), np.asarray(offs, dtype=np.int64), sids)

    @classmethod
    def from_lmdb(cls, path, limit: Optional[int] = None) -> "PackedSystems":
        """Convert a reference LMDB (keys "0", "1", ...; values pickled PyG Data; lmdb_dataset.py:124-162)."""
        try:
            import lmdb  # noqa: F401
            import torch_geometric  # noqa: F401  (the pickles reference its classes)
        except ImportError as e:
            raise ImportError("PackedSystems.from_lmdb needs the `lmdb` and `torch_geometric` packages to read the "
                              "reference's pickled-PyG LMDB; convert where they are installed and ship the packed "
                              f"directory instead ({e})") from e
        import pickle

        env = lmdb.open(str(path), subdir=False, readonly=True, lock=False, readahead=True, meminit=False, max_readers=1)
        with env.begin() as txn:
            n = txn.stat()["entries"]
            length = txn.get("length".encode("ascii"))
            n = pickle.loads(length) if length is not None else n
            if limit is not None:
                n = min(n, limit)
            data = [pickle.loads(txn.get(f"{i}".encode("ascii"))) for i in range(n)]
        env.close()
        return cls.from_data_list(data)

    # ------------------------------------------------------------------ storage
    def save(self, directory) -> None:
        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        for k in _FIELDS + ("cell", "offsets"):
            np.save(d / f"{k}.npy", getattr(self, k))
        (d / "meta.json").write_text(json.dumps({"format": "adsorbdiff_b200.packed/1", "systems": len(self), "sid": self.sid}))

    @classmethod
    def load(cls, directory, mmap: bool = True) -> "PackedSystems":
        d = Path(directory)
        meta = json.loads((d / "meta.json").read_text())
        if meta.get("format") != "adsorbdiff_b200.packed/1":
            raise ValueError(f"{d} is not a packed-systems directory")
        arr = {k: np.load(d / f"{k}.npy", mmap_mode="r" if mmap else None) for k in _FIELDS + ("cell", "offsets")}
        return cls(sid=meta["sid"], **arr)

    # ------------------------------------------------------------------ batches
    def collate(self, indices: Sequence[int], placements: int = 1, pin: bool = True) -> SystemBatch:
        """Batch of the given systems, each repeated `placements` times back to back (BASELINE config #2: the sampler's
        independent initial placements make the copies different).  Tensors are pinned when a CUDA runtime is there."""
        idx = [int(i) for i in indices for _ in range(placements)]
        nat = self.natoms[idx]
        total = int(nat.sum())
        pin = pin and torch.cuda.is_available()

        def buf(shape, dtype):
            t = torch.empty(shape, dtype=dtype)
            return t.pin_memory() if pin else t

        out = {"pos": buf((total, 3), torch.float32), "atomic_numbers": buf((total,), torch.int64),
               "tags": buf((total,), torch.int64), "fixed": buf((total,), torch.int64)}
        views = {k: v.numpy() for k, v in out.items()}
        o = 0
        for i, n in zip(idx, nat):
            a, b = int(self.offsets[i]), int(self.offsets[i + 1])
            for k in _FIELDS:
                views[k][o:o + n] = getattr(self, k)[a:b]
            o += int(n)
        cell = buf((len(idx), 3, 3), torch.float32)
        cell.numpy()[:] = self.cell[idx]
        natoms = torch.from_numpy(nat.astype(np.int64))
        sid = [self.sid[i] if placements == 1 else f"{self.sid[i]}_p{r}" for i in indices for r in range(placements)]
        return SystemBatch(batch=torch.repeat_interleave(torch.arange(len(idx)), natoms), natoms=natoms, cell=cell, sid=sid, **out)


class PackedLoader:
    """Batches of `systems_per_batch` systems (x `placements`) for this rank, collated into pinned memory by a
    background thread one batch ahead of the consumer.  Ranks take contiguous, atom-balanced shards of the dataset
    (the intent of the reference's BalancedBatchSampler, datasets/data_parallel.py:32-48,165-200, without its
    per-batch all_gather)."""

    def __init__(self, data: PackedSystems, systems_per_batch: int, placements: int = 1, rank: int = 0, world_size: int = 1,
                 prefetch: bool = True):
        self.data, self.spb, self.placements, self.prefetch = data, int(systems_per_batch), int(placements), prefetch
        a, b = contiguous_partition(data.natoms.tolist(), world_size)[rank]
        self.indices = list(range(a, b))

    def __len__(self) -> int:
        return (len(self.indices) + self.spb - 1) // self.spb

    def _batches(self) -> List[List[int]]:
        return [self.indices[i:i + self.spb] for i in range(0, len(self.indices), self.spb)]

    def __iter__(self):
        chunks = self._batches()
        if not self.prefetch:
            for c in chunks:
                yield self.data.collate(c, self.placements)
            return
        q: Queue = Queue(maxsize=2)

        def work():
            try:
                for c in chunks:
                    q.put(self.data.collate(c, self.placements))
                q.put(None)
            except BaseException as e:  # surfaced in the consumer
                q.put(e)

        t = threading.Thread(target=work, daemon=True)
        t.start()
        while True:
            item = q.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            yield item
        t.join()
