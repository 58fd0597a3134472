"""Inert stand-in for ASE so the reference sampler's `write()` runs
(oracle scaffolding only)."""
from . import io  # noqa: F401


class Atoms:
    def __init__(self, **kw):
        import numpy as np

        # real ASE copies its inputs; the reference hands in views of batch.pos that it
        # later overwrites in place, so the stand-in must copy too.
        self.kw = {k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}

    def set_calculator(self, calc):
        self.calc = calc
