"""CPU restatement of the training-side pieces of the reference (TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU legs -- never by the product path).

Follows, loop for loop, in numpy/torch on the CPU:
  * IGSO(3) tables and lookups      adsorbdiff/utils/rot_utils.py:9-10, 142-262
  * `pbc_correction`                adsorbdiff/trainers/sde_denoising_trainer.py:45-64
  * the noising of one system       adsorbdiff/trainers/sde_denoising_trainer.py:67-135 (given its random draws)
  * `_compute_loss`                 adsorbdiff/trainers/sde_denoising_trainer.py:675-728

Parity pin: `oracle/gen_golden.py::training_goldens` runs the reference's OWN functions in the build container --
`rot_utils._expansion/_density/_score/sample/score_vec/score_norm` from the imported module (all 1000 eps, 409 s),
and `pbc_correction`, `tr_so3_schedule`, `_compute_loss` compiled from the text of sde_denoising_trainer.py where it
lies (the module itself needs lmdb / wandb / the registry stack) -- and freezes their outputs as
tests/golden/igso3.npz and tests/golden/train_{jit2,mixed}.npz.  tests/test_train_cpu.py checks this restatement
(selected table rows, lookups, loss) and the product (`adsorbdiff_b200/train.py`) against those files.
"""
from __future__ import annotations

import numpy as np
import torch


# ---------------------------------------------------------------- rot_utils.py:142-215
def igso3_tables(min_eps=0.01, max_eps=2.0, n_eps=1000, x_n=2000, L=2000, rows=None):
    """`rows`: compute only these eps rows (the full table is 7 minutes of numpy loops); the returned arrays then have
    len(rows) rows and `eps_index` does not apply."""
    eps_array = 10 ** np.linspace(np.log10(min_eps), np.log10(max_eps), n_eps)
    if rows is not None:
        eps_array = eps_array[np.asarray(rows)]
    omegas = np.linspace(0, np.pi, x_n + 1)[1:]

    def expansion(omega, eps):            # :151-160
        p = 0
        for l in range(L):
            p += (2 * l + 1) * np.exp(-l * (l + 1) * eps**2) * np.sin(omega * (l + 1 / 2)) / np.sin(omega / 2)
        return p

    def score(exp, omega, eps):           # :174-188
        d_sigma = 0
        for l in range(L):
            hi = np.sin(omega * (l + 1 / 2))
            dhi = (l + 1 / 2) * np.cos(omega * (l + 1 / 2))
            lo = np.sin(omega / 2)
            dlo = 1 / 2 * np.cos(omega / 2)
            d_sigma += (2 * l + 1) * np.exp(-l * (l + 1) * eps**2) * (lo * dhi - hi * dlo) / lo**2
        return d_sigma / exp

    exp_vals = np.asarray([expansion(omegas, e) for e in eps_array])
    pdf_vals = np.asarray([e * (1 - np.cos(omegas)) / np.pi for e in exp_vals])        # `_density`, marginal (:163-171)
    cdf_vals = np.asarray([p.cumsum() / x_n * np.pi for p in pdf_vals])
    score_norms = np.asarray([score(exp_vals[i], omegas, eps_array[i]) for i in range(len(eps_array))])
    exp_score_norms = np.sqrt(np.sum(score_norms**2 * pdf_vals, axis=1) / np.sum(pdf_vals, axis=1) / np.pi)
    return dict(omegas=omegas, cdf=cdf_vals, score_norms=score_norms, exp_score_norms=exp_score_norms,
                min_eps=min_eps, max_eps=max_eps, n_eps=n_eps)


def eps_index(T, eps):                    # :219-224, 243-248, 256-261
    idx = (np.log10(eps) - np.log10(T["min_eps"])) / (np.log10(T["max_eps"]) - np.log10(T["min_eps"])) * T["n_eps"]
    return np.clip(np.around(idx).astype(int), a_min=0, a_max=T["n_eps"] - 1)


def sample_omega(T, eps, u):              # `sample` (:218-233) with its uniform draw `u` given
    return np.interp(u, T["cdf"][eps_index(T, eps)], T["omegas"])


def score_vec(T, eps, vec):               # :242-251
    om = np.linalg.norm(vec)
    return np.interp(om, T["omegas"], T["score_norms"][eps_index(T, eps)]) * vec / om


def score_norm(T, eps):                   # :254-262
    return T["exp_score_norms"][eps_index(T, np.asarray(eps))].astype(np.float32)


# ---------------------------------------------------------------- sde_denoising_trainer.py:45-64
def pbc_correction(noise_vec, cell):
    """noise_vec [B,3], cell [B,3,3] -> minimum-image vector per system."""
    out = np.zeros_like(noise_vec, dtype=np.float32)
    for b in range(noise_vec.shape[0]):
        frac = np.linalg.solve(cell[b].T.astype(np.float64), noise_vec[b].astype(np.float64))
        frac %= 1.0
        frac %= 1.0
        frac[frac > 0.5] -= 1
        out[b] = frac.astype(np.float32) @ cell[b].astype(np.float32)
    return out


# ---------------------------------------------------------------- sde_denoising_trainer.py:675-728
def compute_loss(out1, out2, tags, batch_idx, tr_sigma, rot_sigma, tr_score, rot_score, rot_score_norm):
    """torch (differentiable) restatement; out2 / rot_* may be None for the translation-only model."""
    B = int(batch_idx.max()) + 1
    ads = tags == 2

    def scatter_mean(v):
        s = torch.zeros(B, v.shape[1], dtype=v.dtype).index_add_(0, batch_idx[ads], v[ads])
        c = torch.zeros(B, dtype=v.dtype).index_add_(0, batch_idx[ads], torch.ones(int(ads.sum()), dtype=v.dtype))
        return s / c[:, None]

    pos = scatter_mean(out1) / tr_sigma
    pos = torch.cat([pos[:, :2], torch.zeros_like(pos[:, 2:])], dim=1)        # out["positions"][:, -1] = 0
    energy_mask = torch.ones(B, dtype=out1.dtype)
    loss = [((pos - tr_score) ** 2 * tr_sigma**2 * energy_mask.unsqueeze(-1)).mean()]
    if out2 is not None:
        free = scatter_mean(out2) / rot_sigma
        loss.append((((free - rot_score) / rot_score_norm) ** 2 * energy_mask.unsqueeze(-1)).mean())
    return sum(loss)
