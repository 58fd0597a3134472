"""Named synthetic cases shared by oracle/gen_golden.py (which freezes the reference's outputs
for them) and the parity tests.  Inputs are regenerated from seeds, never stored."""
import numpy as np

from adsorbdiff_b200 import synthetic as S


def _gas(n_atoms, box=30.0):
    """A few atoms in a big box: every atom has < 50 neighbours (n_atoms=1: none at all)."""
    pos = np.array([[3.0, 3.0, 3.0], [4.1, 3.2, 3.1], [3.3, 4.4, 2.7], [9.0, 9.0, 9.0]], dtype=np.float32)[:n_atoms]
    z = np.array([6, 8, 1, 29])[:n_atoms]
    tags = np.array([2, 2, 2, 1])[:n_atoms]
    return dict(pos=pos, cell=(np.eye(3) * box).astype(np.float32), atomic_numbers=z.astype(np.int64),
                tags=tags.astype(np.int64), fixed=np.zeros(n_atoms, dtype=np.int64))


CASES = {
    # name: (factory, pbc override or None)
    "jit2": (lambda: S.make_batch(2), None),
    "mixed": (lambda: S.collate([S.make_system(3), S.make_system(4, adsorbate="CHOHCH3", size=(3, 3, 4)),
                                 S.make_system(5, adsorbate="CH3")]), None),
    "skew": (lambda: S.make_batch(2, first_id=7, skew=1.3), None),
    "ideal": (lambda: S.make_batch(1, first_id=9, jitter=0.0), None),
    "pbc_ttf": (lambda: S.make_batch(1, first_id=11), [True, True, False]),
    "tiny": (lambda: S.collate([S.make_system(13, size=(2, 2, 2)), S.make_system(14, size=(2, 3, 2))]), None),
    "gas": (lambda: S.collate([_gas(3), _gas(4)]), None),
    "empty": (lambda: S.collate([_gas(3), _gas(1)]), None),
}


def sampler_batch():
    return S.make_batch(2, first_id=20)


def sampler100_batch():
    """BASELINE config #1: a single adsorbate+slab system for the full 100-step schedule."""
    return S.make_batch(1, first_id=30)
