"""CPU oracle: a restatement of the reference's PaiNN denoising hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module; it is
the checker, never the product.  `adsorbdiff_b200/` never imports it and fails
loudly when its CUDA library is missing.

Parity status: the reference ships no tests and no golden vectors for this path
except the seven `repeat_blocks` docstring examples
(reference: adsorbdiff/models/painn/painn_denoising.py:718-736).  This oracle is
therefore pinned by (a) those seven known answers and (b) outputs of the UNMODIFIED
reference run in the build container through `oracle/ref_import.py`, frozen under
`tests/golden/` by `oracle/gen_golden.py` (integer tensors must match exactly; float
tensors: the host-independent fp64 evaluation to 1e-5 of the frozen fp32 reference output,
the host-dependent fp32 evaluation to 5e-5 -- 2e-6 on the generating host).
See tests/test_oracle_golden.py.

Everything is plain numpy / torch-CPU with the arithmetic order written out where
the result is order sensitive (the d^2 used for the cutoff test and the top-k).

One deliberate, documented deviation ("canonical semantics"):
the reference's per-atom top-k uses `torch.sort` with `stable=False`
(utils/utils.py:806), so which of two candidates with *bit-equal* d^2 survives at
rank 50/51 is an accident of the sort implementation.  The oracle breaks such ties
by enumeration order (= stable sort).  On inputs without exact ties (any jittered
system) this is identical to the reference; `gen_golden.py` records the tie count.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32


# --------------------------------------------------------------------------------------
# repeat_blocks  (reference: painn_denoising.py:700-842) -- loop restatement
# --------------------------------------------------------------------------------------
def repeat_blocks(sizes, repeats, continuous_indexing=True, start_idx=0, block_inc=0, repeat_inc=0):
    """Index array that repeats blocks of consecutive indices.

    Written from the docstring contract of the reference function: block b has
    `sizes[b]` indices and is emitted `repeats[b]` times; the r-th repetition is
    shifted by r*repeat_inc; with continuous indexing block b starts where block
    b-1 ended (+block_inc), and blocks that are skipped (repeats 0) still advance it.
    """
    sizes = [int(s) for s in sizes]
    nb = len(sizes)
    rep = [int(repeats)] * nb if np.isscalar(repeats) else [int(r) for r in repeats]
    rinc = [int(repeat_inc)] * nb if np.isscalar(repeat_inc) else [int(r) for r in repeat_inc]
    binc = [int(block_inc)] * max(nb - 1, 0) if np.isscalar(block_inc) else [int(b) for b in block_inc]
    out = []
    base = start_idx
    for b in range(nb):
        for r in range(rep[b]):
            out.extend(base + r * rinc[b] + k for k in range(sizes[b]))
        if continuous_indexing:
            base += sizes[b]
        if b < nb - 1:
            base += binc[b]
    return np.asarray(out, dtype=np.int64)


# --------------------------------------------------------------------------------------
# neighbour search  (reference: utils/utils.py:556-730 radius_graph_pbc, 733-853 top-k)
# --------------------------------------------------------------------------------------
def cell_repeats(cell, radius, pbc=(True, True, True)):
    """rep_k = ceil(r * |a_{k+1} x a_{k+2}| / |V|), max over the batch  (utils.py:634-662)."""
    cell = np.asarray(cell, dtype=f32)
    t = torch.from_numpy(cell)
    cross23 = torch.cross(t[:, 1], t[:, 2], dim=-1)
    vol = torch.sum(t[:, 0] * cross23, dim=-1, keepdim=True)
    reps = []
    for k, (a, b) in enumerate(((1, 2), (2, 0), (0, 1))):
        if pbc[k]:
            cr = torch.cross(t[:, a], t[:, b], dim=-1)
            inv = torch.norm(cr / vol, p=2, dim=-1)
            reps.append(int(torch.ceil(radius * inv).max().item()))
        else:
            reps.append(0)
    return reps


def image_table(reps):
    """Integer images in `cartesian_prod` order: u1 slowest, u3 fastest (utils.py:665-669)."""
    r1, r2, r3 = reps
    u = [(a, b, c) for a in range(-r1, r1 + 1) for b in range(-r2, r2 + 1) for c in range(-r3, r3 + 1)]
    return np.asarray(u, dtype=f32)  # [C,3]


def image_offsets(cell_b, images):
    """Cartesian offset of each image, fp32, in the order the reference's K=3 `torch.bmm`
    (utils.py:680-681) evaluates on the build image: (a1*u1 + a3*u3) + a2*u2, each product
    rounded separately (no FMA).  Probed bit-equal over random cells, reps up to 5."""
    c = np.asarray(cell_b, dtype=f32)
    u = images
    off = np.empty((3, len(u)), dtype=f32)
    for x in range(3):
        t0 = (c[0, x] * u[:, 0]).astype(f32)
        t1 = (c[1, x] * u[:, 1]).astype(f32)
        t2 = (c[2, x] * u[:, 2]).astype(f32)
        off[x] = ((t0 + t2).astype(f32) + t1).astype(f32)
    return off  # [3,C]


def radius_graph_pbc(pos, cell, natoms, radius, max_nbrs, pbc=(True, True, True)):
    """Returns (edge_index[2,E] i64 (row0 = source j, row1 = target i), cell_offsets[E,3] f32,
    neighbors[B] i64, d2[E] f32, n_ties) with edges ordered by (i, j, image)."""
    pos = np.asarray(pos, dtype=f32)
    cell = np.asarray(cell, dtype=f32)
    natoms = [int(n) for n in natoms]
    reps = cell_repeats(cell, radius, pbc)
    images = image_table(reps)
    r2 = f32(radius * radius)
    src_l, tgt_l, off_l, d2_l, neigh = [], [], [], [], []
    start = 0
    n_ties = 0
    for b, n in enumerate(natoms):
        p = pos[start:start + n]
        off = image_offsets(cell[b], images)  # [3,C]
        # pos2 = pos_j + offset (utils.py:692); diff = pos_i - pos2; d2 = (dx^2+dy^2)+dz^2 (utils.py:695)
        p2 = (p[:, :, None] + off[None, :, :]).astype(f32)  # [j,3,C]
        diff = (p[:, None, :, None] - p2[None, :, :, :]).astype(f32)  # [i,j,3,C]
        sq = (diff * diff).astype(f32)
        d2 = ((sq[:, :, 0] + sq[:, :, 1]).astype(f32) + sq[:, :, 2]).astype(f32)  # [i,j,C]
        mask = (d2 <= r2) & (d2 > f32(0.0001))  # utils.py:699-702
        count = 0
        for i in range(n):
            jj, cc = np.nonzero(mask[i])  # row-major => (j, image) order
            dd = d2[i, jj, cc]
            if len(dd) > max_nbrs:
                order = np.argsort(dd, kind="stable")
                kth = dd[order[max_nbrs - 1]]
                if np.count_nonzero(dd == kth) > 1 and np.count_nonzero(dd <= kth) > max_nbrs:
                    n_ties += 1
                keep = np.sort(order[:max_nbrs])  # survivors keep enumeration order (utils.py:850-851)
                jj, cc, dd = jj[keep], cc[keep], dd[keep]
            src_l.append(jj + start)
            tgt_l.append(np.full(len(jj), i + start, dtype=np.int64))
            off_l.append(images[cc])
            d2_l.append(dd)
            count += len(jj)
        neigh.append(count)
        start += n
    edge_index = np.stack([np.concatenate(src_l), np.concatenate(tgt_l)]).astype(np.int64)
    return (edge_index, np.concatenate(off_l).astype(f32).reshape(-1, 3),
            np.asarray(neigh, dtype=np.int64), np.concatenate(d2_l).astype(f32), n_ties)


def get_pbc_distances(pos, edge_index, cell, cell_offsets, neighbors):
    """vec = pos[j] - pos[i] + u.cell ; d = |vec| ; drop d == 0 without touching `neighbors`
    (reference: utils/utils.py:513-553)."""
    pos = torch.as_tensor(pos, dtype=torch.float32)
    ei = torch.as_tensor(edge_index)
    row, col = ei[0], ei[1]
    vec = pos[row] - pos[col]
    cell_e = torch.repeat_interleave(torch.as_tensor(cell, dtype=torch.float32),
                                     torch.as_tensor(neighbors), dim=0)
    u = torch.as_tensor(cell_offsets, dtype=torch.float32)
    offs = (u[:, :, None] * cell_e).sum(1)
    vec = vec + offs
    d = vec.norm(dim=-1)
    nz = d != 0
    return ei[:, nz], d[nz], vec[nz], u[nz]


def symmetrize_edges(edge_index, cell_offsets, neighbors, d, unit_vec):
    """Non-symmetric branch of `symmetrize_edges` (reference: painn_denoising.py:262-327):
    keep j<i (or j==i with a lexicographically negative image), append the reversed copies,
    order per system = [kept..., reversed...] via `repeat_blocks`."""
    ei = np.asarray(edge_index)
    u = np.asarray(cell_offsets, dtype=f32)
    d = np.asarray(d, dtype=f32)
    rv = np.asarray(unit_vec, dtype=f32)
    earlier = (u[:, 0] < 0) | ((u[:, 0] == 0) & (u[:, 1] < 0)) | ((u[:, 0] == 0) & (u[:, 1] == 0) & (u[:, 2] < 0))
    mask = (ei[0] < ei[1]) | ((ei[0] == ei[1]) & earlier)
    kept = ei[:, mask]
    cat = np.concatenate([kept, kept[::-1]], axis=1)
    batch_edge = np.repeat(np.arange(len(neighbors)), neighbors)[mask]
    per_image = 2 * np.bincount(batch_edge, minlength=len(neighbors))
    reorder = repeat_blocks(per_image // 2, repeats=2, continuous_indexing=True, repeat_inc=kept.shape[1])
    ei_new = cat[:, reorder]
    u_new = np.concatenate([u[mask], -u[mask]])[reorder]
    d_new = np.concatenate([d[mask], d[mask]])[reorder]
    rv_new = np.concatenate([rv[mask], -rv[mask]])[reorder]
    return ei_new, u_new, per_image.astype(np.int64), d_new, rv_new


def generate_graph_values(pos, cell, natoms, radius=12.0, max_nbrs=50, pbc=(True, True, True)):
    """`PaiNN.generate_graph_values` (reference: painn_denoising.py:353-400) on raw arrays.
    Raises ValueError when a system has no neighbours, as the reference does (:370-375)."""
    ei, u, neigh, _, n_ties = radius_graph_pbc(pos, cell, natoms, radius, max_nbrs, pbc)
    ei_t, d, vec, u_t = get_pbc_distances(pos, ei, cell, u, neigh)
    d = d.clone()
    d[torch.isclose(d, torch.tensor(0.0), atol=1e-3)] = 1.0e-3
    unit = vec / d[:, None]
    if (neigh == 0).any():
        raise ValueError("An image has no neighbors")
    ei_s, u_s, neigh_s, d_s, unit_s = symmetrize_edges(ei_t.numpy(), u_t.numpy(), neigh, d.numpy(), unit.numpy())
    return dict(edge_index=ei_s, cell_offsets=u_s, neighbors=neigh_s, dist=d_s, unit_vec=unit_s,
                raw_edge_index=ei, raw_cell_offsets=u, raw_neighbors=neigh, n_ties=n_ties)


# --------------------------------------------------------------------------------------
# edge featurisation (reference: gemnet_oc/layers/radial_basis.py:18-43, 64-82, 235-244)
# --------------------------------------------------------------------------------------
def radial_basis(d, cutoff=12.0, num_rbf=128, exponent=5):
    d = torch.as_tensor(d)
    s = d * (1.0 / cutoff)
    p = float(exponent)
    a, b, c = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
    env = 1 + a * s**p + b * s ** (p + 1) + c * s ** (p + 2)
    env = torch.where(s < 1, env, torch.zeros_like(s))
    offset = torch.linspace(0.0, 1.0, num_rbf, dtype=torch.float32).to(d.dtype)
    coeff = -0.5 / ((1.0 - 0.0) / (num_rbf - 1)) ** 2
    return env[:, None] * torch.exp(coeff * (s[:, None] - offset[None, :]) ** 2)


def ssilu(x):
    """ScaledSiLU (reference: gemnet_oc/layers/base_layers.py:65-72)."""
    return F.silu(x) * (1 / 0.6)


# --------------------------------------------------------------------------------------
# PaiNN forward (reference: painn_denoising.py:402-481, 530-567, 601-623, 626-697)
# --------------------------------------------------------------------------------------
def _lin(x, P, name, bias=True):
    return F.linear(x, P[name + ".weight"], P[name + ".bias"] if bias else None)


def gated_block(P, prefix, x, v, out_channels):
    """GatedEquivariantBlock.forward (reference: painn_denoising.py:688-697)."""
    vec1 = torch.norm(_lin(v, P, prefix + ".vec1_proj", bias=False), dim=-2)
    vec2 = _lin(v, P, prefix + ".vec2_proj", bias=False)
    h = torch.cat([x, vec1], dim=-1)
    h = _lin(ssilu(_lin(h, P, prefix + ".update_net.0")), P, prefix + ".update_net.2")
    xo, g = torch.split(h, out_channels, dim=-1)
    return ssilu(xo), g.unsqueeze(1) * vec2


def painn_forward(P, atomic_numbers, pos, cell, natoms, pbc=(True, True, True), cutoff=12.0,
                  max_nbrs=50, num_layers=6, hidden=512, num_rbf=128, graph=None, trace=None,
                  dtype=torch.float32):
    """Returns (forces[N,3], forces2[N,3]).  `P` = reference state dict (key names of
    SURVEY.md section 8b).  `trace`, if a dict, receives per-layer (x, vec)."""
    P = {k: v.to(dtype) if v.is_floating_point() else v for k, v in P.items()}
    g = graph if graph is not None else generate_graph_values(pos, cell, natoms, cutoff, max_nbrs, pbc)
    ei = torch.as_tensor(g["edge_index"])
    d = torch.as_tensor(g["dist"]).to(dtype)
    rhat = torch.as_tensor(g["unit_vec"]).to(dtype)
    j, i = ei[0], ei[1]
    rbf = radial_basis(d, cutoff, num_rbf)
    z = torch.as_tensor(atomic_numbers).long()
    x = P["atom_emb.embeddings.weight"][z - 1]  # embedding_block.py:42
    n = x.shape[0]
    vec = torch.zeros(n, 3, hidden, dtype=dtype)
    inv_sqrt_2, inv_sqrt_3, inv_sqrt_h = 1 / math.sqrt(2.0), 1 / math.sqrt(3.0), 1 / math.sqrt(hidden)
    if trace is not None:
        trace["rbf"] = rbf
        trace["x_emb"] = x
    for l in range(num_layers):
        m = f"message_layers.{l}"
        xn = F.layer_norm(x, (hidden,), P[m + ".x_layernorm.weight"], P[m + ".x_layernorm.bias"])
        xh = _lin(ssilu(_lin(xn, P, m + ".x_proj.0")), P, m + ".x_proj.2")
        rbfh = _lin(rbf, P, m + ".rbf_proj")
        mx, xh2, xh3 = torch.split(xh[j] * rbfh, hidden, dim=-1)
        xh2 = xh2 * inv_sqrt_3
        mv = (vec[j] * xh2.unsqueeze(1) + xh3.unsqueeze(1) * rhat.unsqueeze(2)) * inv_sqrt_h
        dx = torch.zeros_like(x).index_add_(0, i, mx)
        dvec = torch.zeros_like(vec).index_add_(0, i, mv)
        x = (x + dx) * inv_sqrt_2
        vec = vec + dvec
        if trace is not None:
            trace[f"msg{l}.x"], trace[f"msg{l}.vec"] = x, vec
        u = f"update_layers.{l}"
        v1, v2 = torch.split(_lin(vec, P, u + ".vec_proj", bias=False), hidden, dim=-1)
        vec_dot = (v1 * v2).sum(dim=1) * inv_sqrt_h
        h = torch.cat([x, torch.sqrt(torch.sum(v2**2, dim=-2) + 1e-8)], dim=-1)
        h = _lin(ssilu(_lin(h, P, u + ".xvec_proj.0")), P, u + ".xvec_proj.2")
        a, bq, c = torch.split(h, hidden, dim=-1)
        dx = (a + bq * vec_dot) * inv_sqrt_2
        dvec = c.unsqueeze(1) * v1
        x = x + dx
        vec = vec + dvec
        sc = P[f"upd_out_scalar_scale_{l}.scale_factor"]
        if float(sc) != 0.0:  # ScaleFactor.forward multiplies only when fitted (scale_factor.py:166-167)
            x = x * sc
        if trace is not None:
            trace[f"upd{l}.x"], trace[f"upd{l}.vec"] = x, vec
    outs = []
    for head in ("out_forces", "out_forces2"):
        if head + ".output_network.0.vec1_proj.weight" not in P:
            continue
        hx, hv = gated_block(P, head + ".output_network.0", x, vec, hidden // 2)
        hx, hv = gated_block(P, head + ".output_network.1", hx, hv, 1)
        outs.append(hv.squeeze(-1))  # [N,3,1] -> [N,3]  (PaiNNOutput.forward :647-650)
    return tuple(outs)


# --------------------------------------------------------------------------------------
# sampler step (reference: relaxation/diffusers/denoising_torch.py:198-367; rot_utils.py:18-98)
# --------------------------------------------------------------------------------------
def axis_angle_to_matrix(aa):
    """rot_utils.py:50-98 for one rotation vector [3] (fp32)."""
    aa = torch.as_tensor(aa, dtype=torch.float32)
    angle = torch.norm(aa, p=2)
    half = 0.5 * angle
    if float(angle.abs()) < 1e-6:
        k = 0.5 - (angle * angle) / 48
    else:
        k = torch.sin(half) / angle
    q = torch.cat([torch.cos(half)[None], aa * k])
    r, i, j, kq = q
    two_s = 2.0 / (q * q).sum()
    return torch.stack([
        1 - two_s * (j * j + kq * kq), two_s * (i * j - kq * r), two_s * (i * kq + j * r),
        two_s * (i * j + kq * r), 1 - two_s * (i * i + kq * kq), two_s * (j * kq - i * r),
        two_s * (i * kq - j * r), two_s * (j * kq + i * r), 1 - two_s * (i * i + j * j)]).reshape(3, 3)


def schedule(t_idx, params):
    """Per-step scalars (denoising_torch.py:209-261): returns (tr_g [f32], rot_g [f64], dt [f32])."""
    num_steps = params["num_steps"]
    tr_schedule = torch.tensor(np.linspace(1, 0, num_steps + 1)[:-1], dtype=torch.float32)
    t = tr_schedule[t_idx]
    tr_sigma = params["ads_std_low"] ** (1 - t) * params["ads_std_high"] ** t
    rot_sigma = params["rot_std_low"] ** (1 - t) * params["rot_std_high"] ** t
    tr_g = tr_sigma * (2 * np.log(params["ads_std_high"] / params["ads_std_low"])) ** 0.5
    rot_g = 2 * rot_sigma * torch.sqrt(torch.tensor(np.log(params["rot_std_high"] / params["rot_std_low"])))
    dt = tr_schedule[t_idx] - tr_schedule[t_idx + 1] if t_idx < num_steps - 1 else tr_schedule[t_idx]
    return tr_g, rot_g, dt


def _ads_mean(vals, batch, ads_mask, nsys):
    out = torch.zeros(nsys, vals.shape[1], dtype=vals.dtype).index_add_(0, batch[ads_mask], vals[ads_mask])
    cnt = torch.zeros(nsys, dtype=vals.dtype).index_add_(0, batch[ads_mask], torch.ones(int(ads_mask.sum()), dtype=vals.dtype))
    return out / cnt.clamp(min=1)[:, None]


def init_placement(pos, cell, batch, tags, noise):
    """Random in-plane COM, keep z (denoising_torch.py:215-232).  `noise` = torch.rand(B,3)."""
    pos = pos.clone()
    ads = tags == 2
    nsys = cell.shape[0]
    com_noise = torch.einsum("bi,bij->bj", noise, cell.transpose(1, 2))
    init_com = _ads_mean(pos, batch, ads, nsys)
    com_noise[:, -1] = init_com[:, -1]
    rel = pos[ads] - init_com[batch][ads]
    pos[ads] = rel + com_noise[batch][ads]
    return pos


def se3_step(pos, cell, batch, tags, fixed, score_tr, score_rot, tr_g, rot_g, dt, tr_z=None, rot_z=None):
    """One reverse step (denoising_torch.py:263-353).  `tr_z`/`rot_z` None: the ODE branch (`ode=True`, :269-272);
    given ([B,3] standard-normal draws): the SDE branch (`ode=False`, :273-295), same torch expressions as the
    reference so that dtype promotion (rot_g is float64, everything else fp32) is the reference's.
    Returns (new_pos, delta_com)."""
    ads = tags == 2
    nsys = cell.shape[0]
    score_rot = score_rot.clone()
    score_rot[fixed == 1] = 0  # DiffTorchCalc.get_denoising_prediction :498
    npred = _ads_mean(score_tr, batch, ads, nsys)
    rpred = _ads_mean(score_rot, batch, ads, nsys)
    if tr_z is None:
        upd = 0.5 * tr_g**2 * dt * npred
        rotv = 0.5 * rpred * dt * rot_g**2
    else:
        sqrt_dt = torch.sqrt(dt)  # np.sqrt(dt) on a 0-dim fp32 tensor returns a 0-dim fp32 tensor
        upd = tr_g**2 * dt * npred + tr_g * sqrt_dt * tr_z
        rotv = rpred * dt * rot_g**2 + rot_g * sqrt_dt * rot_z
    com = _ads_mean(pos, batch, ads, nsys)
    upd[:, -1] = 0
    frac = torch.linalg.solve(cell, com + upd)
    frac %= 1
    frac %= 1
    upd = torch.einsum("bi,bij->bj", frac, cell.transpose(1, 2)) - com
    new_pos = pos.clone()
    ads_idx = torch.nonzero(ads).flatten()
    for b in range(nsys):
        sel = ads_idx[batch[ads_idx] == b]
        R = axis_angle_to_matrix(rotv[b]).float()
        new_pos[sel] = (pos[sel] - com[b]) @ R.T + upd[b] + com[b]
    return new_pos, upd


def sample(P, batch_fields, params, noise, num_steps=None, model_kw=None, record=None, sde_noise=None):
    """`Denoiser.reverse_sde_sampling_rot` with early stop disabled (fixed step count).
    `sde_noise` [steps,2,B,3] (tr_z, rot_z per step) runs the SDE branch (`ode=False`)."""
    model_kw = model_kw or {}
    pos = batch_fields["pos"].clone().float()
    cell, bvec = batch_fields["cell"].float(), batch_fields["batch"]
    tags, fixed = batch_fields["tags"].long(), batch_fields["fixed"].long()
    natoms, z = batch_fields["natoms"], batch_fields["atomic_numbers"]
    pos = init_placement(pos, cell, bvec, tags, noise)
    steps = params["num_steps"] if num_steps is None else num_steps
    for t in range(steps):
        tr_g, rot_g, dt = schedule(t, params)
        s_tr, s_rot = painn_forward(P, z, pos.numpy(), cell.numpy(), natoms, **model_kw)
        zs = (None, None) if sde_noise is None else (sde_noise[t, 0], sde_noise[t, 1])
        pos, _ = se3_step(pos, cell, bvec, tags, fixed, s_tr, s_rot, tr_g, rot_g, dt, *zs)
        if record is not None:
            record.append(pos.clone())
    return pos
