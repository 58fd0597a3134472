"""Training step of the score model.

Reference: adsorbdiff/trainers/sde_denoising_trainer.py -- `tr_so3_schedule` (:67-135, noising of a batch),
`_forward_denoising` (:539-553), `_compute_loss` (:675-728), the loop body of `train` (:409-447) and
`BaseTrainer._backward` (base_trainer.py:787-820: zero_grad, backward, clip, step, EMA); IGSO(3) tables
adsorbdiff/utils/rot_utils.py:142-264.

What runs where:
  * the radius graph is the sampler's own kernel (csrc/neighbors.cu) -- with `direct_forces` nothing
    differentiates through positions, so the graph is data;
  * the message op (edge featurisation + rbf projection + message + aggregation, the part of PaiNN that is not a
    plain GEMM) is a `torch.autograd.Function` over two hand-written kernels: forward csrc/message.cu (exact fp32),
    backward csrc/message_bwd.cu -- no per-edge tensor ever exists in HBM, where the reference's autograd keeps
    rbf [E,128], rbf_proj(rbf) [E,1536] and the messages [E,3,512] alive per layer (about 3.5 GB per layer at 48
    systems);
  * LayerNorm, the Linear layers (cuBLAS fp32), SiLU and the gated blocks are plain torch ops under autograd:
    library GEMMs, fp32 like the reference's default (no AMP);
  * the noising schedule is vectorised on the device (the reference loops over systems on the host with numpy
    RNG and scipy-free table lookups, :104-125); same distributions, different random stream.
"""
from __future__ import annotations

import copy
import math
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _cabi
from ._cabi import call, ptr

INV_SQRT_2 = 1.0 / math.sqrt(2.0)


# --------------------------------------------------------------------------------------------------------------
# message op under autograd
# --------------------------------------------------------------------------------------------------------------
def rbf_weight_scale(net, w: torch.Tensor, refresh: bool = False) -> float:
    """Power-of-two prescale of an `rbf_proj` weight for its fp16x2 planes: s * max|w| in [1024, 2048], 32x below the
    fp16 limit.  Measured on the host the first time a weight is seen (one device read) and again whenever `TrainStep`
    reads the status word (`status_every` steps): an optimiser step moves a weight by at most lr, so the headroom
    cannot be used up in between; should it ever be, the split sets the overflow bit and `check_gemm_status` raises."""
    tab = net.__dict__.setdefault("_train_rbf_scale", {})
    key = (w.data_ptr(), tuple(w.shape))   # (the table is dropped by `load_state_dict` and rebuilt by `TrainStep.__init__`)
    if refresh or key not in tab:
        amax = float(w.detach().abs().max())
        tab[key] = 2.0 ** math.floor(math.log2(SPLIT_TARGET / amax)) if amax > 0 and math.isfinite(amax) else 1.0
    return tab[key]


class MessageFn(torch.autograd.Function):
    """(x, vec, xh, W_rbf, b_rbf) -> ((x + dx) / sqrt(2), vec + dvec); PaiNNMessage + residual
    (painn_denoising.py:443-445, 534-567).  `vec` may be None (first layer)."""

    @staticmethod
    def forward(ctx, x, vec, xh, w, b, net, p):
        N, F_, R = p.N, net.hidden_channels, net.num_rbf
        x_out = x.detach().clone().contiguous()
        vec_in = vec.detach().contiguous() if vec is not None else None
        vec_out = torch.empty(N, 3, F_, dtype=torch.float32, device=x.device)
        xh_c, w_c, b_c = xh.detach().contiguous(), w.detach().contiguous(), b.detach().contiguous()
        if getattr(net, "train_msg", "t5") == "t5" and getattr(p, "t5_fits", False):
            # the sampler's tcgen05 kernel (csrc/message_t5.cu): fp16x2 planes of the weight at a power-of-two scale
            # measured on the weight itself (`rbf_weight_scale`), overflow reported through the training status word
            scale = rbf_weight_scale(net, w_c)
            planes = torch.empty(2 * w_c.numel(), dtype=torch.float16, device=x.device)
            status = _TcWorkspace.get(x.device).status
            call("adk_split_f16", p.device, ptr(w_c), R, 3 * F_, R, scale, ptr(planes), 3 * F_, ptr(status))
            call("adk_message_t5", p.device, ptr(p.atom_off), p.B, p.n_max, None, ptr(p.row_start), ptr(p.row_deg),
                 ptr(p.e_src), ptr(p.e_geo), ptr(xh_c), ptr(vec_in), ptr(planes), scale, ptr(b_c),
                 ptr(net.radial_basis.rbf.offset), F_, R, float(net.cutoff), net.radial_basis.exponent,
                 float(net.msg_t5_comp), ptr(x_out), ptr(vec_out), None, 0, 1.0, ptr(status))
        else:   # exact-fp32 row-tiled kernel (any system size, any num_rbf)
            call("adk_message", p.device, None, ptr(p.row_start), ptr(p.row_deg), ptr(p.e_src), ptr(p.e_geo), ptr(xh_c),
                 ptr(vec_in), ptr(w_c), ptr(b_c), ptr(net.radial_basis.rbf.offset), N, F_, R, float(net.cutoff),
                 net.radial_basis.exponent, ptr(x_out), ptr(vec_out))
        ctx.net, ctx.p, ctx.has_vec = net, p, vec is not None
        ctx.save_for_backward(xh_c, vec_in if vec_in is not None else xh_c.new_empty(0), w_c, b_c)
        return x_out, vec_out

    @staticmethod
    def backward(ctx, g_x, g_vec):
        net, p = ctx.net, ctx.p
        xh, vec_in, w, b = ctx.saved_tensors
        N, F_, R = p.N, net.hidden_channels, net.num_rbf
        dev = p.device
        g_dx = (g_x * INV_SQRT_2).contiguous()
        g_dvec = g_vec.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        d_xh = torch.empty(N, 3 * F_, **f32)
        d_vec = torch.empty(N, 3, F_, **f32)
        d_w = torch.empty(3 * F_, R, **f32)
        d_b = torch.empty(3 * F_, **f32)
        n_scratch = _cabi.load().adk_message_bwd_scratch_floats(N, F_, R, None)
        scratch = torch.empty(n_scratch, **f32)
        call("adk_message_bwd", dev, ptr(p.row_start), ptr(p.row_deg), ptr(p.e_src), ptr(p.e_geo), ptr(p.bwd_plan), ptr(xh),
             ptr(vec_in) if ctx.has_vec else None, ptr(w), ptr(b), ptr(net.radial_basis.rbf.offset), N, F_, R,
             float(net.cutoff), net.radial_basis.exponent, ptr(g_dx), ptr(g_dvec), ptr(d_xh), ptr(d_vec), ptr(d_w),
             ptr(d_b), ptr(scratch))
        grad_vec = (g_dvec + d_vec) if ctx.has_vec else None
        return g_dx, grad_vec, d_xh, d_w, d_b, None, None


# --------------------------------------------------------------------------------------------------------------
# nn.Linear forward / backward on the tcgen05 GEMM (csrc/linear_tc.cu) with device-side prescales (csrc/train_ops.cu)
# --------------------------------------------------------------------------------------------------------------
SPLIT_TARGET = 2048.0    # scaled magnitudes are placed in [1024, 2048]: 32x below the fp16 limit


def _pad(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _TcWorkspace:
    """Per-device scratch of the training GEMMs: the amax kernel's two counters and an overflow status word."""

    def __init__(self, device):
        self.device = device
        self.scratch = torch.zeros(2, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.planes = torch.empty(0, dtype=torch.uint8, device=device)
        self.rec_g = torch.empty(2, dtype=torch.float32, device=device)

    def operand_space(self, M: int, K: int, N: int) -> torch.Tensor:
        """Device scratch for the operand planes of one Linear pass (grown on demand; stream-ordered reuse)."""
        # = adk_linear_train_ws_bytes(M, K, N), evaluated here (this runs ~80 times per step)
        mp, np_, kr = _pad(M, 128), _pad(N, 128), _pad(M, 64)
        need = 4 * max(mp * K + N * K, mp * N + K * N + np_ * kr + K * kr) + 2048
        if self.planes.numel() < need:
            self.planes = torch.empty(int(need * 1.25), dtype=torch.uint8, device=self.device)
        return self.planes

    _by_device: dict = {}

    @classmethod
    def get(cls, device) -> "_TcWorkspace":
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        ws = cls._by_device.get(key)
        if ws is None:
            ws = cls._by_device[key] = cls(torch.device("cuda", key[1]))
        return ws


def _rec(t: torch.Tensor, ws: _TcWorkspace) -> torch.Tensor:
    """{s, 1/s}: the power-of-two prescale of `t`, computed on the device (no host sync)."""
    rec = torch.empty(2, dtype=torch.float32, device=t.device)
    call("adk_amax_scale", ws.device, ptr(t), t.numel(), SPLIT_TARGET, ptr(rec), ptr(ws.scratch))
    return rec


def _split(t: torch.Tensor, rec, ws, plane_rows: int) -> torch.Tensor:
    M, K = t.shape
    out = torch.empty(2 * plane_rows * K, dtype=torch.float16, device=t.device)
    call("adk_split_f16_dev", ws.device, ptr(t), K, M, K, ptr(rec), ptr(out), plane_rows, ptr(ws.status))
    return out


def _split_t(t: torch.Tensor, rec, ws, plane_rows: int, kp: int) -> torch.Tensor:
    M, C = t.shape
    out = torch.empty(2 * plane_rows * kp, dtype=torch.float16, device=t.device)
    call("adk_split_f16_t_dev", ws.device, ptr(t), C, M, C, ptr(rec), ptr(out), plane_rows, kp, ptr(ws.status))
    return out


def _gemm_nt(a, a_rows: int, M: int, w, N: int, K: int, bias, rec_a, rec_w, ws) -> torch.Tensor:
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    call("adk_linear_tc_dev", ws.device, ptr(a), a_rows, M, ptr(w), N, K, ptr(bias) if bias is not None else None,
         ptr(rec_a), ptr(rec_w), ptr(out), N, ptr(ws.status))
    return out


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (the kernels read float4)."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


class TcLinearFn(torch.autograd.Function):
    """y = x W^T + b with all three GEMMs (Y, dX = dY W, dW = dY^T X) on the tcgen05 fp16x2-split kernel; one C call
    per pass (`adk_linear_train_fwd` / `_bwd`: device amax -> power-of-two prescale -> split -> GEMM)."""

    @staticmethod
    def forward(ctx, x, w, b):
        ws = _TcWorkspace.get(x.device)
        x, w = _aligned(x.detach()), _aligned(w.detach())
        M, K = x.shape
        N = w.shape[0]
        recs = torch.empty(4, dtype=torch.float32, device=x.device)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        call("adk_linear_train_fwd", ws.device, ptr(x), ptr(w), ptr(_aligned(b.detach())) if b is not None else None, M, K, N,
             SPLIT_TARGET, ptr(recs), ptr(ws.operand_space(M, K, N)), ptr(ws.scratch), ptr(ws.status), ptr(y))
        ctx.save_for_backward(x, w, recs)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, w, recs = ctx.saved_tensors
        ws = _TcWorkspace.get(x.device)
        g = _aligned(g)
        M, K = x.shape
        N = w.shape[0]
        dx = torch.empty(M, K, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty(N, K, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[1] else None
        call("adk_linear_train_bwd", ws.device, ptr(g), ptr(x), ptr(w), M, K, N, SPLIT_TARGET, ptr(recs), ptr(ws.rec_g),
             ptr(ws.operand_space(M, K, N)), ptr(ws.scratch), ptr(ws.status), ptr(dx), ptr(dw))
        db = g.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def _tc_shape_ok(lin) -> bool:
    n, k = lin.weight.shape
    return k % 64 == 0 and n % 64 == 0 and n <= 2048 and k <= 2048


def _lin(net, lin, x):
    """`lin(x)`: on the tensor-core GEMM when `net.train_gemm == "tc"` and the shape allows, else torch (cuBLAS fp32)."""
    if getattr(net, "train_gemm", "tc") != "tc" or not _tc_shape_ok(lin):
        return lin(x)
    lead = x.shape[:-1]
    y = TcLinearFn.apply(x.reshape(-1, x.shape[-1]), lin.weight, lin.bias)
    return y.reshape(*lead, y.shape[-1])


class UpdatePrepFn(torch.autograd.Function):
    """(x, vp) -> (dot, cat): vec_dot and [x | |v2|] of PaiNNUpdate.forward (painn_denoising.py:602-613), one kernel
    each way instead of nine / fifteen eager ops."""

    @staticmethod
    def forward(ctx, x, vp):
        x, vp = x.detach().contiguous(), vp.detach().contiguous()
        N, F_ = x.shape
        dot = torch.empty(N, F_, dtype=torch.float32, device=x.device)
        cat = torch.empty(N, 2 * F_, dtype=torch.float32, device=x.device)
        call("adk_update_prep", x.device, ptr(x), ptr(vp), N, F_, ptr(dot), ptr(cat), None, 0, 1.0, None)
        ctx.save_for_backward(vp)
        return dot, cat

    @staticmethod
    def backward(ctx, g_dot, g_cat):
        (vp,) = ctx.saved_tensors
        N, F_ = g_dot.shape
        g_x = torch.empty(N, F_, dtype=torch.float32, device=vp.device)
        g_vp = torch.empty_like(vp)
        call("adk_update_prep_bwd", vp.device, ptr(vp), ptr(g_dot.contiguous()), ptr(g_cat.contiguous()), N, F_, ptr(g_x), ptr(g_vp))
        return g_x, g_vp


class UpdateGateFn(torch.autograd.Function):
    """(x, vec, vp, h, dot) -> (x', vec'): the residual / gating half of PaiNNUpdate.forward (:614-623) plus the
    fitted ScaleFactor (scale_factor.py:157-172)."""

    @staticmethod
    def forward(ctx, x, vec, vp, h, dot, scale):
        vp, h, dot = vp.detach().contiguous(), h.detach().contiguous(), dot.detach().contiguous()
        N, F_ = x.shape
        x_out, vec_out = x.detach().clone(), vec.detach().clone()
        call("adk_update_gate", x.device, ptr(h), ptr(dot), ptr(vp), ptr(scale), N, F_, ptr(x_out), ptr(vec_out), None, 0, 1.0, None)
        ctx.save_for_backward(vp, h, dot, scale)
        return x_out, vec_out

    @staticmethod
    def backward(ctx, g_xo, g_vo):
        vp, h, dot, scale = ctx.saved_tensors
        N, F_ = dot.shape
        f32 = dict(dtype=torch.float32, device=vp.device)
        g_x, g_h, g_dot, g_vp = torch.empty(N, F_, **f32), torch.empty(N, 3 * F_, **f32), torch.empty(N, F_, **f32), torch.empty_like(vp)
        g_vo = g_vo.contiguous()
        call("adk_update_gate_bwd", vp.device, ptr(h), ptr(dot), ptr(vp), ptr(scale), ptr(g_xo.contiguous()), ptr(g_vo), N, F_,
             ptr(g_x), ptr(g_h), ptr(g_dot), ptr(g_vp))
        return g_x, g_vo, g_vp, g_h, g_dot, None


def _ssilu(x):
    """ScaledSiLU (gemnet_oc/layers/base_layers.py:65-72)"""
    return F.silu(x) * (1.0 / 0.6)


def _mlp(net, seq, x):
    return _lin(net, seq[2], _ssilu(_lin(net, seq[0], x)))


def _gated_block(net, blk, x, v, out_channels: int):
    """GatedEquivariantBlock.forward (painn_denoising.py:688-697)"""
    vec1 = torch.norm(_lin(net, blk.vec1_proj, v), dim=-2)
    vec2 = _lin(net, blk.vec2_proj, v)
    h = _mlp(net, blk.update_net, torch.cat([x, vec1], dim=-1))
    xo, g = torch.split(h, out_channels, dim=-1)
    return _ssilu(xo), g.unsqueeze(1) * vec2


def _status_to_host(net, p) -> None:
    """Copy the status word to pinned host memory on a side stream, right behind the kernels that set it (neighbour
    search, element check): reading it later costs no drain of the main stream."""
    io = getattr(net, "_status_io", None)
    if io is None or io["device"] != p.device:
        io = net._status_io = {"device": p.device, "host": torch.zeros(1, dtype=torch.int32).pin_memory(),
                               "stream": torch.cuda.Stream(p.device), "event": torch.cuda.Event()}
    io["stream"].wait_stream(torch.cuda.current_stream(p.device))
    with torch.cuda.stream(io["stream"]):
        io["host"].copy_(p.status, non_blocking=True)
        io["event"].record(io["stream"])


def check_status_async(net, p) -> None:
    """Raise what `PaiNN.check_status` raises, from the side-stream copy made by the forward."""
    io = net._status_io
    io["event"].synchronize()
    if int(io["host"][0]):
        net.check_status(p)


def forward_train(net, data, check: bool = True):
    """The differentiable forward of `adsorbdiff_b200.PaiNN` (PaiNN.forward, painn_denoising.py:402-481).  Returns
    forces [N,3] (and forces2 with `so3_denoising`) attached to the autograd graph of the parameters.  `check` raises
    on the device status word (empty system, bad element, row overflow) before returning; it waits only for the
    neighbour search, not for the forward.  `TrainStep` defers it until the backward pass has been enqueued."""
    with torch.autocast(device_type="cuda", enabled=False):   # PaiNN trains in fp32 (no AMP for this model in the
        return _forward_train_fp32(net, data, check)           # reference, run.py:21-25); the kernels take fp32 only


def _forward_train_fp32(net, data, check: bool):
    p, z, pos = net._prepare(data)
    net._graph(p, pos)
    if torch.is_grad_enabled():
        # the weight-gradient pass walks the edges sorted by tap slot: planned once per graph, used by every layer
        if getattr(p, "bwd_plan", None) is None:
            n_int = _cabi.load().adk_message_bwd_plan_ints(p.N, p.e_src.numel())
            p.bwd_plan = torch.empty(n_int, dtype=torch.int32, device=p.device)
        call("adk_message_bwd_plan", p.device, ptr(p.row_start), ptr(p.row_deg), ptr(p.e_src), ptr(p.e_geo), p.N,
             net.num_rbf, float(net.cutoff), net.radial_basis.exponent, ptr(p.bwd_plan))
    F_ = net.hidden_channels
    ne = net.atom_emb.embeddings.weight.shape[0]
    bad = ((z < 1) | (z > ne)).any()    # the reference's embedding raises IndexError; torch's CUDA lookup would assert
    p.status.bitwise_or_(bad.to(torch.int32) * _cabi.STATUS_BAD_ELEMENT)
    _status_to_host(net, p)
    x = net.atom_emb.embeddings(z.clamp(1, ne) - 1)
    vec = None
    for l in range(net.num_layers):
        m, u = net.message_layers[l], net.update_layers[l]
        xh = _mlp(net, m.x_proj, m.x_layernorm(x))
        x, vec = MessageFn.apply(x, vec, xh, m.rbf_proj.weight, m.rbf_proj.bias, net, p)
        sc = getattr(net, "upd_out_scalar_scale_%d" % l).scale_factor
        vp = _lin(net, u.vec_proj, vec)
        if getattr(net, "train_fused_update", True):
            vec_dot, cat = UpdatePrepFn.apply(x, vp)
            h = _mlp(net, u.xvec_proj, cat)
            x, vec = UpdateGateFn.apply(x, vec, vp, h, vec_dot, sc)
        else:   # the same block in eager torch ops (kept for the parity test of the fused kernels)
            v1, v2 = torch.split(vp, F_, dim=-1)
            vec_dot = (v1 * v2).sum(dim=1) * (1.0 / math.sqrt(F_))
            h = _mlp(net, u.xvec_proj, torch.cat([x, torch.sqrt(torch.sum(v2 ** 2, dim=-2) + 1e-8)], dim=-1))
            a, bq, c = torch.split(h, F_, dim=-1)
            x = x + (a + bq * vec_dot) * INV_SQRT_2
            vec = vec + c.unsqueeze(1) * v1
            x = torch.where(sc != 0.0, x * sc, x)   # ScaleFactor.forward multiplies only when fitted (no host sync)
    outs = []
    for head in ([net.out_forces, net.out_forces2] if net.so3_denoising else [net.out_forces]):
        hx, hv = _gated_block(net, head.output_network[0], x, vec, F_ // 2)
        hx, hv = _gated_block(net, head.output_network[1], hx, hv, 1)
        outs.append(hv.squeeze(-1))
    net._train_plan = p
    if check:
        check_status_async(net, p)
    return outs[0] if not net.so3_denoising else tuple(outs)


# --------------------------------------------------------------------------------------------------------------
# IGSO(3) tables (rot_utils.py:9-10, 142-215)
# --------------------------------------------------------------------------------------------------------------
class IGSO3Tables:
    """cdf / score-norm tables of the isotropic Gaussian on SO(3), built once on `device` in float64 by the series
    the reference sums on the host (:151-160 `_expansion`, :174-188 `_score`, :205-215) -- about a second on a GPU
    at the reference's sizes (N_EPS=1000, X_N=2000, L=2000) against minutes of numpy.  Lookups reproduce the
    reference's index arithmetic, including `* N_EPS` (not N_EPS - 1) before rounding (:226-231)."""

    def __init__(self, device="cpu", min_eps=0.01, max_eps=2.0, n_eps=1000, x_n=2000, L=2000):
        self.min_eps, self.max_eps, self.n_eps, self.x_n = float(min_eps), float(max_eps), int(n_eps), int(x_n)
        # the series' terms decay like exp(-l^2 eps^2): below L ~ 6 / eps the narrowest table rows are not converged
        # (negative densities, NaN scores).  The reference's L = 2000 covers its min_eps = 0.01 (6 / eps = 600).
        if L * float(min_eps) < 6.0:
            raise ValueError(f"IGSO3Tables: L={L} terms do not converge the series at min_eps={min_eps} (need L >= {6.0 / float(min_eps):.0f})")
        dev = torch.device(device)
        f64 = dict(dtype=torch.float64, device=dev)
        eps = 10 ** torch.linspace(math.log10(min_eps), math.log10(max_eps), n_eps, **f64)
        om = torch.linspace(0, math.pi, x_n + 1, **f64)[1:]
        l = torch.arange(L, **f64)
        lo = torch.sin(om / 2)                                  # [X]
        dlo = 0.5 * torch.cos(om / 2)
        hi = torch.sin(om[:, None] * (l[None, :] + 0.5))        # [X, L]
        dhi = (l[None, :] + 0.5) * torch.cos(om[:, None] * (l[None, :] + 0.5))
        term_e = hi / lo[:, None]
        term_s = (lo[:, None] * dhi - hi * dlo[:, None]) / (lo[:, None] ** 2)
        wgt = (2 * l[None, :] + 1) * torch.exp(-l[None, :] * (l[None, :] + 1) * eps[:, None] ** 2)   # [n_eps, L]
        exp_vals, dsig = wgt @ term_e.T, wgt @ term_s.T                                                  # [n_eps, X]
        pdf = exp_vals * (1 - torch.cos(om))[None, :] / math.pi
        self.omegas = om
        self.cdf = pdf.cumsum(dim=1) / x_n * math.pi
        self.score_norms = dsig / exp_vals
        self.exp_score_norms = torch.sqrt((self.score_norms ** 2 * pdf).sum(1) / pdf.sum(1) / math.pi)

    def to(self, device):
        for k in ("omegas", "cdf", "score_norms", "exp_score_norms"):
            setattr(self, k, getattr(self, k).to(device))
        return self

    def eps_index(self, eps: torch.Tensor) -> torch.Tensor:
        idx = (torch.log10(eps.double()) - math.log10(self.min_eps)) / (math.log10(self.max_eps) - math.log10(self.min_eps)) * self.n_eps
        return torch.clamp(torch.round(idx).long(), 0, self.n_eps - 1)

    @staticmethod
    def _interp(x, xp, fp):
        """np.interp row-wise: x [B], xp [B, X] increasing, fp [B, X] (clamped at the ends)."""
        X = xp.shape[1]
        hi = torch.searchsorted(xp.contiguous(), x[:, None].contiguous(), right=True).squeeze(1)
        lo_i = torch.clamp(hi - 1, 0, X - 1)
        hi_i = torch.clamp(hi, 0, X - 1)
        x0, x1 = xp.gather(1, lo_i[:, None]).squeeze(1), xp.gather(1, hi_i[:, None]).squeeze(1)
        f0, f1 = fp.gather(1, lo_i[:, None]).squeeze(1), fp.gather(1, hi_i[:, None]).squeeze(1)
        w = torch.where(x1 > x0, (x - x0) / torch.where(x1 > x0, x1 - x0, torch.ones_like(x0)), torch.zeros_like(x0))
        out = f0 + w * (f1 - f0)
        out = torch.where(x <= xp[:, 0], fp[:, 0], out)
        return torch.where(x >= xp[:, -1], fp[:, -1], out)

    def sample_vec(self, eps: torch.Tensor, generator=None, axis=None, u=None) -> torch.Tensor:
        """`sample_vec` (:236-239) for a vector of eps: uniform axis, angle by inverse-cdf lookup.  float64 [B,3].
        `axis` [B,3] (unnormalised normal draws) and `u` [B] (uniform draws) may be supplied (tests)."""
        B, dev = eps.shape[0], self.cdf.device
        idx = self.eps_index(eps)
        if axis is None:
            axis = torch.randn(B, 3, dtype=torch.float64, device=dev, generator=generator)
        axis = axis / axis.norm(dim=1, keepdim=True)
        if u is None:
            u = torch.rand(B, dtype=torch.float64, device=dev, generator=generator)
        omega = self._interp(u, self.cdf[idx], self.omegas[None, :].expand(B, -1))
        return axis * omega[:, None]

    def score_vec(self, eps: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
        """`score_vec` (:242-251): interp(|vec|, omegas, score_norms[eps]) * vec / |vec|."""
        idx = self.eps_index(eps)
        om = vec.norm(dim=1)
        B = vec.shape[0]
        s = self._interp(om, self.omegas[None, :].expand(B, -1), self.score_norms[idx])
        return s[:, None] * vec / om[:, None]

    def score_norm(self, eps: torch.Tensor) -> torch.Tensor:
        """`score_norm` (:254-262), float32 with the shape of `eps`."""
        return self.exp_score_norms[self.eps_index(eps)].float()


def axis_angle_to_matrix(aa: torch.Tensor) -> torch.Tensor:
    """rot_utils.py:18-98 (axis-angle -> quaternion -> matrix) for [B,3]."""
    angle = aa.norm(dim=1, keepdim=True)
    half = 0.5 * angle
    small = angle.abs() < 1e-6
    k = torch.where(small, 0.5 - angle * angle / 48, torch.sin(half) / torch.where(small, torch.ones_like(angle), angle))
    q = torch.cat([torch.cos(half), aa * k], dim=1)
    r, i, j, kq = q.unbind(1)
    two_s = 2.0 / (q * q).sum(1)
    return torch.stack([
        1 - two_s * (j * j + kq * kq), two_s * (i * j - kq * r), two_s * (i * kq + j * r),
        two_s * (i * j + kq * r), 1 - two_s * (i * i + kq * kq), two_s * (j * kq - i * r),
        two_s * (i * kq - j * r), two_s * (j * kq + i * r), 1 - two_s * (i * i + j * j)], dim=1).reshape(-1, 3, 3)


def _segment_mean(values, seg, B):
    out = torch.zeros(B, values.shape[1], dtype=values.dtype, device=values.device).index_add_(0, seg, values)
    cnt = torch.zeros(B, dtype=values.dtype, device=values.device).index_add_(0, seg, torch.ones_like(seg, dtype=values.dtype))
    return out / cnt.clamp_min(1)[:, None]


_CONSTS = {}


def _const(values: tuple, device) -> torch.Tensor:
    """A small constant tensor per device, created once (torch.tensor(..., device=cuda) is a blocking copy)."""
    key = (values, str(device))
    if key not in _CONSTS:
        _CONSTS[key] = torch.tensor(values, device=device)
    return _CONSTS[key]


def _adsorbate_rows(batch):
    """(rows, system of each row) of the adsorbate atoms (tags == 2), in atom order.  A batch that `TrainStep` moved to
    the device itself carries them (`_ads_idx`, `_ads_seg`: found on the host copy, so nothing here waits for the GPU);
    otherwise they are found with a boolean mask, which costs a device -> host round trip like the reference's
    `batch.pos[batch.tags == 2]`."""
    idx = getattr(batch, "_ads_idx", None)
    if idx is not None and idx.device == batch.pos.device:
        return idx, batch._ads_seg
    idx = torch.nonzero(batch.tags == 2).flatten()
    return idx, batch.batch.index_select(0, idx)


@torch.no_grad()
def pbc_correction(noise_vec, cell):
    """Minimum-image wrap of a per-system vector (sde_denoising_trainer.py:45-64): fractional coordinates mod 1,
    mapped to (-0.5, 0.5]."""
    # (solve_ex: the same LU solve without `solve`'s host-side check of the info word, which waits for the GPU)
    frac = torch.linalg.solve_ex(cell.double().transpose(1, 2), noise_vec.double()[:, :, None], check_errors=False)[0].squeeze(2)
    frac = frac % 1.0
    frac = frac % 1.0
    frac = torch.where(frac > 0.5, frac - 1, frac)
    return torch.einsum("bi,bij->bj", frac.float(), cell.float())


@torch.no_grad()
def tr_so3_schedule(batch, params: dict, tables: IGSO3Tables, generator=None, draws: Optional[dict] = None):
    """Noise a batch for one training step (sde_denoising_trainer.py:67-135), all systems at once on the device:
    t ~ U(0,1) per system, sigma = low^(1-t) high^t, xy-translation of the adsorbate by N(0, sigma_tr^2) wrapped to
    the minimum image, a rotation about the adsorbate's centre drawn from IGSO(3)(sigma_rot), +1 A in z.  Sets
    `tr_sigma`, `rot_sigma`, `rot_score`, `tr_score`, `ads_center_noise_vec` and rewrites `batch.pos[tags == 2]`.
    `draws` = dict(t [B], normal [B,3], axis [B,3], u [B]) replaces the random draws (parity tests)."""
    dev = batch.pos.device
    B = int(batch.natoms.shape[0])
    d = draws or {}
    t = d["t"] if "t" in d else torch.rand(B, device=dev, generator=generator)
    tr_sigma = params["ads_std_low"] ** (1 - t) * params["ads_std_high"] ** t
    rot_sigma = params["rot_std_low"] ** (1 - t) * params["rot_std_high"] ** t
    ads, seg = _adsorbate_rows(batch)
    ads_pos = batch.pos.index_select(0, ads)
    center = _segment_mean(ads_pos, seg, B)
    noise = (d["normal"] if "normal" in d else torch.randn(B, 3, device=dev, generator=generator)) * tr_sigma[:, None]
    noise = pbc_correction(noise, batch.cell.reshape(B, 3, 3))
    noise[:, -1] = 0
    rot_update = tables.sample_vec(rot_sigma, generator, d.get("axis"), d.get("u"))   # float64 like the numpy original
    rot_mat = axis_angle_to_matrix(rot_update).float()
    rot_score = tables.score_vec(rot_sigma, rot_update).float()
    rel = ads_pos - center[seg]
    new_pos = torch.einsum("nj,nij->ni", rel, rot_mat[seg]) + noise[seg] + center[seg]   # rel @ R^T
    new_pos[:, -1] += 1
    batch.tr_sigma, batch.rot_sigma = tr_sigma[:, None], rot_sigma[:, None]
    batch.rot_score = rot_score
    batch.pos = batch.pos.index_copy(0, ads, new_pos)
    batch.ads_center_noise_vec = noise
    batch.tr_score = -noise / tr_sigma[:, None] ** 2
    return batch


@torch.no_grad()
def ads_com_gaussian_schedule(batch, params: dict, generator=None, draws: Optional[dict] = None):
    """Noising for the translation-only model (`so3_denoising=False`; sde_denoising_trainer.py:138-177): the adsorbate's
    centre is moved by N(0, sigma^2) in xy, wrapped into the cell (`solve(cell, c)`, `% 1` twice -- the reference's
    arithmetic, not the transposed solve of `pbc_correction`), lifted by 1 A, and EVERY adsorbate atom is placed at that
    point (the model only learns the centre of mass)."""
    dev = batch.pos.device
    B = int(batch.natoms.shape[0])
    d = draws or {}
    t = d["t"] if "t" in d else torch.rand(B, device=dev, generator=generator)
    tr_sigma = params["ads_std_low"] ** (1 - t) * params["ads_std_high"] ** t
    ads, seg = _adsorbate_rows(batch)
    center = _segment_mean(batch.pos.index_select(0, ads), seg, B)
    noise = (d["normal"] if "normal" in d else torch.randn(B, 3, device=dev, generator=generator)) * tr_sigma[:, None]
    noise[:, -1] = 0
    center = center + noise
    cell = batch.cell.reshape(B, 3, 3)
    frac = torch.linalg.solve_ex(cell, center[:, :, None], check_errors=False)[0].squeeze(2)
    frac = frac % 1
    frac = frac % 1
    center = torch.einsum("bi,bij->bj", frac, cell.transpose(1, 2))
    center[:, -1] += 1
    batch.pos = batch.pos.index_copy(0, ads, center[seg])
    batch.tr_sigma = tr_sigma[:, None]
    batch.ads_center_noise_vec = noise
    batch.tr_score = -noise / tr_sigma[:, None] ** 2
    return batch


def denoising_loss(out, batch, tables: Optional[IGSO3Tables], denoising_pos_coefficient: float = 1.0):
    """`DenoisingTrainer._compute_loss` (:675-728): adsorbate-mean of the two heads, divided by sigma; translation
    term weighted by sigma^2 with z zeroed, rotation term divided by the expected score norm; both `.mean()`s are
    over [B, 3].  (`denoising_pos_coefficient` is read and never applied by the reference, :680-682 -- same here.)"""
    so3 = isinstance(out, (tuple, list))
    tr_out = out[0] if so3 else out
    ads, seg = _adsorbate_rows(batch)
    B = int(batch.natoms.shape[0])
    tr = _segment_mean(tr_out.index_select(0, ads), seg, B) / batch.tr_sigma
    tr = tr * _const((1.0, 1.0, 0.0), tr.device)
    loss = ((tr - batch.tr_score) ** 2 * batch.tr_sigma ** 2).mean()
    if so3:
        rot = _segment_mean(out[1].index_select(0, ads), seg, B) / batch.rot_sigma
        norm = tables.score_norm(batch.rot_sigma)
        loss = loss + (((rot - batch.rot_score) / norm) ** 2).mean()
    return loss


# --------------------------------------------------------------------------------------------------------------
# one optimisation step
# --------------------------------------------------------------------------------------------------------------
def allreduce_mean_(tensors) -> None:
    """Average `tensors` over the ranks in place with ONE all-reduce of a flat buffer (what DistributedDataParallel
    does bucket by bucket for the gradients; the whole model is 42 MB, one bucket)."""
    world = dist.get_world_size()
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat)
    flat.div_(world)
    o = 0
    for t in tensors:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()


class ShadowEma:
    """The slice of `ExponentialMovingAverage` (modules/exponential_moving_average.py:20-130) the trainer and the sampler
    use: `update`, `store`, `copy_to`, `restore` -- the sampler swaps the shadow weights in for a run
    (sde_denoising_trainer.py:580-583, 650-651)."""

    def __init__(self, params, decay: float):
        self.params, self.decay = list(params), float(decay)
        self.shadow_params = [q.detach().clone() for q in self.params]
        self.collected_params = []

    def update(self) -> None:
        torch._foreach_lerp_(self.shadow_params, [q.detach() for q in self.params], 1.0 - self.decay)

    def store(self) -> None:
        self.collected_params = [q.detach().clone() for q in self.params]

    def copy_to(self) -> None:
        for s_, q in zip(self.shadow_params, self.params):
            q.data.copy_(s_)

    def restore(self) -> None:
        for c, q in zip(self.collected_params, self.params):
            q.data.copy_(c)
        self.collected_params = []


class TrainStep:
    """noise -> forward -> loss -> backward -> (all-reduce) -> clip -> AdamW -> EMA; the body of
    `DenoisingTrainer.train` (:409-447) + `_backward` (base_trainer.py:787-820) with the optimizer set up as
    `load_optimizer` does (:556-612: no weight decay on embeddings and biases).

    Data parallel: one process per GPU, each on its own batch; gradients are averaged with ONE all-reduce of a flat
    fp32 buffer (10.6 M parameters = 42 MB) over NCCL -- DistributedDataParallel's average (world-size mean) without
    its per-bucket hooks."""

    def __init__(self, net, optim: dict, tables: Optional[IGSO3Tables] = None, generator=None):
        self.net, self.optim, self.generator = net, optim, generator
        self.params = [q for q in net.parameters() if q.requires_grad]
        dev = self.params[0].device
        self.tables = tables if tables is not None else (IGSO3Tables(dev) if net.so3_denoising else None)
        wd = float(optim.get("optimizer_params", {}).get("weight_decay", 0))
        skip = set(net.no_weight_decay())
        no_decay = [q for n, q in net.named_parameters() if q.requires_grad and n in skip]
        decay = [q for n, q in net.named_parameters() if q.requires_grad and n not in skip]
        groups = [{"params": no_decay, "weight_decay": 0.0}, {"params": decay, "weight_decay": wd}]
        name = optim.get("optimizer", "AdamW")
        if name != "AdamW":
            raise NotImplementedError("the denoising configs train with AdamW")
        extra = {k: v for k, v in optim.get("optimizer_params", {}).items() if k != "weight_decay"}   # betas, eps, amsgrad ...
        self.optimizer = torch.optim.AdamW(groups, lr=float(optim.get("lr_initial", 1e-4)), fused=dev.type == "cuda", **extra)
        self.clip = optim.get("clip_grad_norm")
        self.ema_decay = optim.get("ema_decay")
        self.ema = ShadowEma(self.params, self.ema_decay) if self.ema_decay else None   # what `Denoiser` / `ml_diffuse` look for
        self.shadow = self.ema.shadow_params if self.ema else None
        self._unwrapped_model = net                                                      # trainer-like: unwrap_model(step)
        self.pos_params = optim.get("denoising_pos_params", {})
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.step_count = 0
        self.status_every = int(optim.get("status_every", 50))   # host read of the GEMMs' overflow word (one sync)
        net.__dict__.pop("_train_rbf_scale", None)                # message-weight prescales: measured again on first use
        # loss read-out without draining the stream: each step copies its loss to a pinned ring slot right after the
        # forward and records an event; `read_loss(lag)` waits for that event only
        self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory() if dev.type == "cuda" else None
        self._loss_events = [torch.cuda.Event() for _ in range(4)] if dev.type == "cuda" else None

    def to_device(self, batch):
        """A device copy of a HOST batch (what a DataLoader yields) for `__call__`; the caller's batch is left alone.
        The host side keeps what the launch plan needs as Python numbers -- atoms per system, cell, pbc -- and finds the
        adsorbate rows there, so the step that follows never waits for the GPU to learn them: `natoms.cpu()`, the image
        count's `.tolist()` and `pos[tags == 2]` each drain the stream when done on device tensors, which serialises
        host and device once per step (5 ms of a 32 ms step at 48 systems).  Pinned host tensors make the copies
        asynchronous."""
        dev = next(self.net.parameters()).device
        host = {k: getattr(batch, k, None) for k in ("natoms", "cell", "pbc")}
        ads_idx = torch.nonzero(batch.tags == 2).flatten()
        ads_seg = batch.batch.index_select(0, ads_idx)
        d = copy.copy(batch)
        d = d.to(dev, non_blocking=True)
        d._ads_idx, d._ads_seg = ads_idx.to(dev, non_blocking=True), ads_seg.to(dev, non_blocking=True)
        self.net._host_meta = {k: (getattr(d, k, None), v) for k, v in host.items()}
        return d

    def __call__(self, batch, noised: bool = False) -> torch.Tensor:
        net = self.net
        net.train()
        if batch.pos.device.type == "cpu" and next(net.parameters()).device.type == "cuda":
            batch = self.to_device(batch)
        if not noised:
            if net.so3_denoising:
                batch = tr_so3_schedule(batch, self.pos_params, self.tables, self.generator)
            else:
                batch = ads_com_gaussian_schedule(batch, self.pos_params, self.generator)
        out = forward_train(net, batch, check=False)
        loss = denoising_loss(out, batch, self.tables)
        if self._loss_host is not None:
            slot = self.step_count % 4
            self._loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            self._loss_events[slot].record()
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        check_status_async(net, net._train_plan)   # raises before the optimizer sees gradients of a malformed batch
        if self.world > 1:
            allreduce_mean_([q.grad for q in self.params if q.grad is not None])
        if self.clip:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=float(self.clip), foreach=True)
        self.optimizer.step()
        if self.ema is not None:   # ExponentialMovingAverage.update (exponential_moving_average.py:71-97)
            self.ema.update()
        self.step_count += 1
        if self.status_every and self.step_count % self.status_every == 0:
            self.check_gemm_status()
        return loss.detach()

    def read_loss(self, lag: int = 0) -> float:
        """Loss of the step made `lag` calls ago (0 = the last one, up to 3) as a Python float.  Waits for that step's
        forward only, not for the stream: logging the previous step's loss while the current step runs keeps the GPU fed
        (the reference's `loss.item()` right after `backward()` drains it every step)."""
        k = self.step_count - 1 - lag
        if k < 0 or lag > 3:
            raise ValueError("no such step")
        self._loss_events[k % 4].synchronize()
        return float(self._loss_host[k % 4])

    @torch.no_grad()
    def predict_denoising(self, batch, per_image: bool = False, disable_tqdm: bool = True) -> dict:
        """`DenoisingTrainer.predict_denoising` for one batch (sde_denoising_trainer.py:555-673, without its per-call EMA
        swap: the sampler swaps once per run): makes this object usable wherever the sampler takes a trainer."""
        self.net.eval()
        out = self.net(batch)
        if self.net.so3_denoising:
            return {"positions": out[0], "positions_free": out[1]}
        return {"positions": out}

    def check_gemm_status(self) -> None:
        """The tensor-core GEMMs pick their prescales from the operands' own maxima, so the fp16 range can only be
        left by an operand that already holds inf / nan; the status word says so (read every `status_every` steps)."""
        ws = _TcWorkspace.get(self.params[0].device)
        st = int(ws.status.item())
        for m in self.net.message_layers:   # (the stream is drained anyway: re-measure the message weights' prescales)
            rbf_weight_scale(self.net, m.rbf_proj.weight, refresh=True)
        if st:
            ws.status.zero_()
            raise _cabi.AdkOverflow("a GEMM operand of the training step was not finite (fp16x2 split overflow bit set): "
                                    "loss or gradients have diverged -- or an rbf_proj weight grew 32x within "
                                    f"{self.status_every} steps and left the prescale of the message forward "
                                    "(train.rbf_weight_scale; re-measured just now)")

    def params_in_sync(self) -> bool:
        """Data parallel sanity: every rank holds bit-identical parameters (they start equal and see the same averaged
        gradients).  One all-reduce of a checksum pair; call it outside timed regions."""
        if self.world == 1:
            return True
        flat = torch.cat([q.detach().reshape(-1) for q in self.params]).double()
        probe = torch.stack([flat.sum(), (flat * torch.arange(1, flat.numel() + 1, device=flat.device, dtype=torch.float64)).sum()])
        lo, hi = probe.clone(), probe.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))
