"""Reverse-diffusion sampler: the reference's `Denoiser` / `DiffTorchCalc` / `ml_diffuse`
interfaces on top of the fused kernels.

Reference: adsorbdiff/relaxation/diffusers/denoising_torch.py:18-84 (Denoiser), :198-367
(reverse_sde_sampling_rot), :486-511 (DiffTorchCalc); adsorbdiff/relaxation/ml_relaxation.py:98-168
(ml_diffuse).  Same constructor arguments, same `run()` contract: the *same* batch object comes
back with `pos` updated in place (fp32) and `y` / `force` set to zeros.

One denoising step = PaiNN forward (neighbour search .. output heads) + one SE(3) update kernel.
After the first step the whole step is replayed as a single CUDA graph: the per-step scalars live
in a device schedule table indexed by a device step counter, so replays need no host input.

Deviations from the reference, all opt-in or documented in DESIGN.md:
  * the EMA weight swap of `predict_denoising` (3 full-model copies per step,
    sde_denoising_trainer.py:580-583, 650-651) is hoisted out of the loop (once per run);
  * the reference checks `allclose(delta_com, 0)` on the host every step (:312-320); here the same
    test runs on the device after every step (`adk_early_stop`: counter, break before the tenth hit is
    applied) and the host only polls the result every `early_stop_every` steps (default 10) -- final
    positions and step count are the reference's; `denoising_pos_params["early_stop"] = False` disables it;
  * per-step ASE trajectory writing (:358-367) needs ASE, which is optional here: frames are
    collected on the device and written once at the end (`.traj` through ASE when importable,
    else `<sid>.npz`).
"""
from __future__ import annotations

import logging
from collections import deque
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import call, ptr
from .painn import PaiNN


def unwrap_model(model):
    """trainer -> DDP -> module, like the reference's `_unwrapped_model` (base_trainer.py)."""
    m = getattr(model, "_unwrapped_model", model)
    m = getattr(m, "module", m)
    return m


class DiffTorchCalc:
    """reference: denoising_torch.py:486-511"""

    def __init__(self, model, transform=None) -> None:
        self.model = model
        self.transform = transform

    def get_denoising_prediction(self, atoms, apply_constraint: bool = True):
        predictions = self.model.predict_denoising(atoms, per_image=False, disable_tqdm=True)
        positions = predictions["positions"]
        if "positions_free" in predictions and apply_constraint:
            positions_free = predictions["positions_free"]
            positions_free[atoms.fixed == 1] = 0
            return positions, positions_free
        return positions


def schedule_table(params: dict, device) -> torch.Tensor:
    """[num_steps][ADK_SCHED_COLS] per-step scalars evaluated with the reference's own torch/numpy expressions and
    dtypes (denoising_torch.py:209-261, 269-295): tr_g is fp32, rot_g is float64.
    Columns: 0.5*tr_g^2*dt | dt | fp32(rot_g^2) | tr_g^2*dt | tr_g*sqrt(dt) | fp32(rot_g*sqrt(dt))."""
    num_steps = params["num_steps"]
    lo, hi = params["ads_std_low"], params["ads_std_high"]
    rlo, rhi = params["rot_std_low"], params["rot_std_high"]
    tr_schedule = torch.tensor(np.linspace(1, 0, num_steps + 1)[:-1], dtype=torch.float32)
    rows = []
    for t_idx in range(num_steps):
        t = tr_schedule[t_idx]
        tr_sigma = lo ** (1 - t) * hi ** t
        rot_sigma = rlo ** (1 - t) * rhi ** t
        tr_g = tr_sigma * (2 * np.log(hi / lo)) ** 0.5
        rot_g = 2 * rot_sigma * torch.sqrt(torch.tensor(np.log(rhi / rlo)))
        dt = tr_schedule[t_idx] - tr_schedule[t_idx + 1] if t_idx < num_steps - 1 else tr_schedule[t_idx]
        sqrt_dt = torch.sqrt(dt)  # what np.sqrt(dt) evaluates to for a 0-dim fp32 tensor
        rows.append([float(0.5 * tr_g**2 * dt), float(dt), float((rot_g**2).to(torch.float32)),
                     float(tr_g**2 * dt), float(tr_g * sqrt_dt), float((rot_g * sqrt_dt).to(torch.float32))])
    out = torch.tensor(rows, dtype=torch.float32, device=device)
    assert out.shape[1] == _cabi.SCHED_COLS
    return out


def draw_sde_noise(num_steps: int, num_systems: int, device) -> torch.Tensor:
    """[num_steps][2][B][3]: the standard-normal draws of the SDE branch, made with the reference's calls in the
    reference's order (tr_z then rot_z every step, `torch.normal(mean=0, std=1, size=[B,3], device=device)`,
    denoising_torch.py:274-289) so that the same generator state gives the same trajectory.  They are drawn up
    front because the step runs as a CUDA-graph replay; an early-stopped run therefore advances the generator
    further than the reference would."""
    out = torch.empty(num_steps, 2, num_systems, 3, dtype=torch.float32, device=device)
    for t in range(num_steps):
        out[t, 0] = torch.normal(mean=0, std=1, size=(num_systems, 3), device=device)
        out[t, 1] = torch.normal(mean=0, std=1, size=(num_systems, 3), device=device)
    return out


class Denoiser:
    def __init__(
        self,
        batch,
        model,
        denoising_pos_params: dict,
        device: str = "cuda:0",
        save_full_traj: bool = True,
        traj_dir: Optional[Path] = None,
        traj_names=None,
        early_stop_batch: bool = False,
        logger=None,
        use_cuda_graph: bool = True,
        init_noise: Optional[torch.Tensor] = None,
        sde_noise: Optional[torch.Tensor] = None,
    ) -> None:
        """Same arguments as the reference `Denoiser` (denoising_torch.py:18-62) plus:
        `init_noise` [B,3]: the uniform draws of the initial placement (:215) when the caller has made them --
        a rank that holds rows [a, b) of a larger job passes rows [a, b) of the job-wide `torch.rand(B_total, 3)`
        (`partition.initial_noise`), so that an N-rank run reproduces the 1-rank run; default: drawn here from the
        CPU global generator exactly like the reference.
        `sde_noise` [num_steps,2,B,3]: the normal draws of the SDE branch (`draw_sde_noise`), same idea."""
        self.batch = batch
        self.model = model  # DiffTorchCalc(trainer) like the reference, or a PaiNN directly
        self.device = device
        self.save_full = save_full_traj
        self.traj_dir = traj_dir
        self.traj_names = traj_names
        self.early_stop_batch = early_stop_batch
        self.denoising_pos_params = denoising_pos_params
        self.use_cuda_graph = use_cuda_graph
        self.init_noise = init_noise
        self.sde_noise = sde_noise
        trainer = getattr(model, "model", model)
        self.trainer = trainer if hasattr(trainer, "predict_denoising") else None
        self.net = unwrap_model(trainer)
        if not isinstance(self.net, PaiNN):
            raise TypeError("adsorbdiff_b200.Denoiser drives adsorbdiff_b200.PaiNN; got "
                            f"{type(self.net).__name__} (use the reference Denoiser for other models)")
        self.otf_graph = self.net.otf_graph
        assert not self.traj_dir or (traj_dir and len(traj_names)), \
            "Trajectory names should be specified to save trajectories"
        self.steps_run = 0
        self.frames = None

    # ------------------------------------------------------------------
    def run(self):
        self.reverse_sde_sampling_rot()
        if self.traj_dir:
            self._write_trajectories()
        return self.batch

    @torch.no_grad()
    def reverse_sde_sampling_rot(self):
        params = self.denoising_pos_params
        if "ads_std_low" not in params:
            return
        ode = bool(params.get("ode", True))
        batch, net = self.batch, self.net
        dev = batch.pos.device
        if dev.type != "cuda":
            raise _cabi.AdkError("Denoiser needs the batch on a CUDA device (no CPU fallback)")
        num_steps = params["num_steps"]
        early_stop = params.get("early_stop", True)
        # the stop decision is taken on the device after every step (adk_early_stop: exact reference semantics);
        # the host only polls the flag, every `early_stop_every` steps, to leave the loop
        poll_every = max(1, int(params.get("early_stop_every", 10)))
        status_every = max(1, int(params.get("status_every", 25)))
        record = bool(self.traj_dir) or bool(params.get("keep_frames", False))

        ema = getattr(self.trainer, "ema", None) if self.trainer is not None else None
        if ema:
            ema.store()
            ema.copy_to()
        net.eval()
        try:
            # operand prescales of the tensor-core GEMMs: measured on the weights that will actually run (the EMA
            # shadow, if any, was just swapped in) and on a sample of this batch
            if net.auto_calibrate and (ema or net._needs_calibration()):
                net.calibrate(batch)
            plan, z, pos = net._prepare(batch)
            if pos.data_ptr() != batch.pos.data_ptr():
                batch.pos = pos  # fp32 contiguous working copy becomes the batch's positions
            tags = batch.tags.to(torch.int32).contiguous()
            fixed = batch.fixed.to(torch.int32).contiguous()
            B = plan.B
            # initial placement: the reference draws on the CPU global generator (:215)
            noise = self.init_noise if self.init_noise is not None else torch.rand(B, 3)
            if tuple(noise.shape) != (B, 3):
                raise ValueError(f"init_noise must be [{B}, 3], got {tuple(noise.shape)}")
            noise = noise.to(dev, torch.float32).contiguous()
            call("adk_init_placement", dev, ptr(pos), ptr(plan.cell_f32), ptr(plan.atom_off), ptr(tags),
                 ptr(noise), B)
            sched = schedule_table(params, dev)
            sde = None
            if not ode:
                sde = self.sde_noise if self.sde_noise is not None else draw_sde_noise(num_steps, B, dev)
                if tuple(sde.shape) != (num_steps, 2, B, 3):
                    raise ValueError(f"sde_noise must be [{num_steps}, 2, {B}, 3], got {tuple(sde.shape)}")
                sde = sde.to(dev, torch.float32).contiguous()
            step = torch.zeros(1, dtype=torch.int32, device=dev)
            max_upd = torch.zeros(B, dtype=torch.float32, device=dev)
            prev = torch.empty_like(pos) if early_stop else None
            stop = torch.zeros(4, dtype=torch.int32, device=dev) if early_stop else None
            if record:
                self.frames = torch.empty(num_steps, plan.N, 3, dtype=torch.float32, device=dev)

            # The SE(3) update averages both scores over the adsorbate atoms and reads nothing else
            # (reference :460-467, 322-338), and nothing after the last message layer mixes atoms: that layer's
            # message, its update block and the heads are evaluated for the adsorbate rows only
            # (denoising_pos_params["full_forward"] = True evaluates every atom like the reference).
            out_rows = None
            if not params.get("full_forward", False):
                flags = (tags == 2).to(torch.int32).contiguous()
                idx = torch.nonzero(flags, as_tuple=False).flatten().to(torch.int32).contiguous()
                if 0 < idx.numel() < plan.N:
                    out_rows = (idx, flags)

            def one_step(weights_ready=False):
                if early_stop:
                    prev.copy_(pos)
                net._run(plan, z, pos, weights_ready=weights_ready, out_rows=out_rows)
                call("adk_se3_step", dev, ptr(pos), ptr(plan.cell_f32), ptr(plan.atom_off), ptr(tags), ptr(fixed),
                     ptr(plan.out[0]), ptr(plan.out[1]), ptr(sched), ptr(step), B, ptr(sde), ptr(max_upd), ptr(stop))
                if early_stop:
                    call("adk_early_stop", dev, ptr(max_upd), B, 1e-3, ptr(step), ptr(stop), ptr(pos), ptr(prev),
                         pos.numel())

            graph = None
            t_idx = 0
            stopped = False
            while t_idx < num_steps:
                if graph is not None:
                    graph.replay()
                else:
                    one_step()  # first step eager: warms every kernel up outside capture
                    if self.use_cuda_graph and num_steps > 2:
                        net.check_status(plan)  # (a device synchronisation: one D2H read)
                        # Between the EMA swap-in above and the swap-out below nobody else writes the parameters,
                        # and the eager step just produced their fp16x2 planes: the replayed step does not redo it.
                        # (capture does not execute: pos / step / stop keep their values)
                        graph = _cabi.capture_graph(lambda: one_step(weights_ready=True), dev)
                t_idx += 1
                if record:
                    self.frames[t_idx - 1].copy_(pos)
                # a data-dependent failure (a system without neighbours, an operand leaving the fp16 range) is caught
                # after the first replayed step and then every `status_every` steps, not only at the end of the run
                if t_idx == 2 or t_idx % status_every == 0:
                    net.check_status(plan)
                if early_stop and (t_idx % poll_every == 0 or t_idx == num_steps):
                    st = stop.tolist()
                    if st[1]:
                        t_idx = st[2]  # steps applied before the break (:312-320)
                        stopped = True
                        break
            self.steps_run = t_idx
            self.stopped_early = stopped
            net.check_status(plan)
        finally:
            if ema:
                ema.restore()
        batch.y = torch.zeros(B, device=dev)
        batch.force = torch.zeros(pos.shape, device=dev)

    # ------------------------------------------------------------------
    def _write_trajectories(self) -> None:
        """Deferred trajectory output: one file per system, frames = steps (reference :64-82, 469-477)."""
        self.traj_dir.mkdir(exist_ok=True, parents=True)
        frames = self.frames[: self.steps_run].cpu().numpy()
        if not self.save_full:
            frames = frames[-1:]
        nat = self.batch.natoms.tolist()
        numbers = self.batch.atomic_numbers.cpu().numpy()
        tags = self.batch.tags.cpu().numpy()
        fixed = self.batch.fixed.cpu().numpy()
        cells = self.batch.cell.cpu().numpy()
        try:
            import ase
            import ase.io
            from ase.constraints import FixAtoms
            have_ase = hasattr(ase.io, "Trajectory") and hasattr(ase, "Atoms") and hasattr(ase.Atoms, "get_positions")
        except Exception:
            have_ase = False
        start = 0
        for i, (name, n) in enumerate(zip(self.traj_names, nat)):
            sl = slice(start, start + n)
            if have_ase:
                tmp = self.traj_dir / f"{name}.traj_tmp"
                traj = ase.io.Trajectory(tmp, mode="w")
                for f in frames:
                    traj.write(ase.Atoms(numbers=numbers[sl].astype(int).tolist(), positions=f[sl],
                                         tags=tags[sl].astype(int).tolist(), cell=cells[i],
                                         constraint=FixAtoms(mask=fixed[sl].astype(bool).tolist()),
                                         pbc=[True, True, True]))
                traj.close()
                tmp.rename(self.traj_dir / f"{name}.traj")
            else:
                np.savez(self.traj_dir / f"{name}.npz", positions=frames[:, sl], numbers=numbers[sl],
                         tags=tags[sl], fixed=fixed[sl], cell=cells[i])
            start += n


def _collate(data_list, like):
    """Re-collate a list of single systems (or finished batches) the way the reference does with
    `data_list_collater` / `Batch.from_data_list` (ml_relaxation.py:163-167): PyG when the batch is a PyG batch,
    else the batch class's own `from_data_list` (adsorbdiff_b200.synthetic.SystemBatch has one)."""
    cls = type(like)
    if hasattr(cls, "from_data_list") and not cls.__module__.startswith("torch_geometric"):
        return cls.from_data_list(data_list)
    from torch_geometric.data import Batch

    return Batch.from_data_list(data_list)


def ml_diffuse(batch, model, denoising_pos_params: dict, traj_dir, save_full_traj, device: str = "cuda:0",
               transform=None, early_stop_batch: bool = False, logger=None):
    """reference: adsorbdiff/relaxation/ml_relaxation.py:98-168.  Like the reference, ANY RuntimeError of a batch
    (out of memory first of all) frees the cache, splits the batch in two and retries (second half first, :163-165);
    a single-system batch re-raises."""
    batches = deque([batch])
    done = []
    while batches:
        b = batches.popleft()
        calc = DiffTorchCalc(model, transform)
        den = Denoiser(b, calc, denoising_pos_params, device=device, save_full_traj=save_full_traj,
                       traj_dir=Path(traj_dir) if traj_dir is not None else None, traj_names=b.sid,
                       early_stop_batch=early_stop_batch, logger=logger)
        e: Optional[RuntimeError] = None
        try:
            done.append(den.run())
        except RuntimeError as err:
            e = err
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        if e is not None:
            # recovery outside the except clause so that the failed attempt's tensors can be freed (:156-157)
            if not hasattr(b, "to_data_list"):
                raise e
            data_list = b.to_data_list()
            if len(data_list) == 1:
                raise e
            logging.info(f"Failed to relax batch with size: {len(data_list)}, splitting into two...")
            mid = len(data_list) // 2
            batches.appendleft(_collate(data_list[:mid], b))
            batches.appendleft(_collate(data_list[mid:], b))
    if len(done) == 1:
        return done[0]
    return _collate(done, batch)
