#!/usr/bin/env python
"""Benchmark of the PaiNN denoising hot path (BASELINE.json metric: system*steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--systems S] [--quick]

A "step" is one reverse-diffusion step over the job's adsorbate+slab systems: PaiNN forward (neighbour search,
6 message/update layers, two output heads) + the SE(3) update.

Workload = BASELINE config #3: S = 1024 synthetic 82-atom systems IN TOTAL, partitioned over the N ranks by
`adsorbdiff_b200.partition.contiguous_partition` (1024 / 512 / 256 / 128 systems per GPU at N = 1 / 2 / 4 / 8): strong
scaling.  No collective on the data path; the job's one collective is the final gather of positions.

Keys of the JSON line (rank 0 prints ONE line):
  value         whole-job system*steps/s, device timed (CUDA events, max over ranks), inputs resident in HBM, the
                FULL forward (every atom's scores, what `PaiNN.forward` returns) replayed as a CUDA graph
  sampler_step  the same for the step as `Denoiser` replays it (last layer + heads on the adsorbate rows only)
  e2e           the same metric through the public API: `Denoiser.run()` on a HOST-resident batch -- H2D of the batch,
                plan, eager first step, graph capture, K steps, D2H of the final positions all inside the timed region
  weak          (N > 1) the round-1 weak-scaling figure: 1024 systems on EVERY rank
  partition_check  (N > 1) the gathered positions of a short N-rank `Denoiser` run are bit-identical to rank 0 running
                all S systems alone
  roofline / kernels / cpu_baseline / reference_gpu / clocks: see DESIGN.md section 5

`--impl reference` times the UNMODIFIED reference (baseline/_ref, a git-ignored `pip install --target` of
/root/reference made by `__graft_entry__.build()`) through `oracle/ref_import.py` on the host cores -- the oracle port
when that copy is absent -- on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLER_PARAMS = dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55,
                      early_stop=False)
METRIC = "denoising_system_steps_per_sec"
UNIT = "system*steps/s"
DISTINCT = 64  # distinct synthetic systems, tiled to S (independent initial placements make every copy different)


# ----------------------------------------------------------------------------------------------
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def workload_name(systems):
    return (f"PaiNN reverse-diffusion sampling, {systems} synthetic OC20-Dense-shaped systems in total "
            "(80-atom slab + CO/OH, 82 atoms, ~4110 edges each), partitioned over the ranks; hidden 512 x 6 layers, "
            "128 RBF, cutoff 12 A, 50 nbrs")


def global_systems(total):
    from adsorbdiff_b200 import synthetic as S

    base = [S.make_system(i) for i in range(min(DISTINCT, total))]
    return [base[i % len(base)] for i in range(total)]


# ---------------------------------------------------------------------------------------------- CPU baselines
class _RefTrainer:
    """The lines of `DenoisingTrainer` the reference sampler touches (sde_denoising_trainer.py:539-553): the model
    call on `batch.to(device)` (PyG's `.to` is in place, so the batch migrates on the first step)."""

    def __init__(self, model, device, stamps=None):
        self.model = self._unwrapped_model = model
        self.device, self.stamps = device, stamps

    @torch.no_grad()
    def predict_denoising(self, batch, per_image=False, disable_tqdm=True):
        if self.stamps is not None:
            if self.device != "cpu":
                torch.cuda.synchronize()
            self.stamps.append(time.perf_counter())
        p1, p2 = self.model(batch.to(self.device))
        return {"positions": p1.detach(), "positions_free": p2.detach()}


def reference_rate(n_systems: int, steps: int, warmup: int, threads: int, device: str = "cpu"):
    """system*steps/s of the UNMODIFIED reference (`PaiNN` + `Denoiser.reverse_sde_sampling_rot`, imported through
    oracle/ref_import.py) on `threads` host threads, or with its torch ops on `device`.  Returns (rate, seconds)."""
    import logging

    from adsorbdiff_b200 import synthetic as S
    from oracle import ref_import

    logging.disable(logging.INFO)
    ns = ref_import.load()
    torch.set_num_threads(threads)
    model = ns.PaiNN(None, 0, 1, scale_file=ns.scale_file, so3_denoising=True).eval()
    model.load_state_dict(S.random_state_dict(0), strict=True)
    model = model.to(device)
    b = S.collate(global_systems(n_systems))
    b.id = b.fid = torch.arange(n_systems)
    sid_names, b.sid = b.sid, torch.arange(n_systems)
    ns.utils.radius_graph_pbc.__defaults__[-1][:] = [True, True, True]
    stamps = []
    params = dict(SAMPLER_PARAMS, num_steps=steps + warmup)
    params.pop("early_stop")
    den = ns.Denoiser(b, ns.DiffTorchCalc(_RefTrainer(model, device, stamps)), params, device=device, traj_dir=None,
                      traj_names=sid_names)
    import ase.io  # inert shim (oracle/ref_shims): the per-step trajectory write stays in the loop, the file does not

    den.trajectories = [ase.io.Trajectory() for _ in sid_names]
    torch.manual_seed(0)
    den.reverse_sde_sampling_rot()
    if device != "cpu":
        torch.cuda.synchronize()
    stamps.append(time.perf_counter())
    done = len(stamps) - 1  # steps actually run (the reference may stop early)
    w = min(warmup, max(done - 1, 0))
    dt = stamps[-1] - stamps[w]
    return n_systems * (done - w) / dt, dt


def port_rate(n_systems: int, steps: int, warmup: int, threads: int):
    """system*steps/s of the oracle port (CPU restatement of the reference) on `threads` host threads."""
    from adsorbdiff_b200 import synthetic as S
    from oracle import painn_oracle as O

    torch.set_num_threads(threads)
    sd = S.random_state_dict(0)
    b = S.collate(global_systems(n_systems))
    torch.manual_seed(0)
    pos = O.init_placement(b.pos.clone(), b.cell, b.batch, b.tags, torch.rand(n_systems, 3))
    t0 = None
    for t in range(steps + warmup):
        if t == warmup:
            t0 = time.perf_counter()
        tr_g, rot_g, dt = O.schedule(t % SAMPLER_PARAMS["num_steps"], SAMPLER_PARAMS)
        s_tr, s_rot = O.painn_forward(sd, b.atomic_numbers, pos.numpy(), b.cell.numpy(), b.natoms)
        pos, _ = O.se3_step(pos, b.cell, b.batch, b.tags, b.fixed, s_tr, s_rot, tr_g, rot_g, dt)
    el = time.perf_counter() - t0
    return n_systems * steps / el, el


def cpu_rate(n_systems, steps, warmup, threads):
    """(rate, seconds, kind): the unmodified reference when its copy is importable, else the port."""
    from oracle import ref_import

    if ref_import.available():
        try:
            r, dt = reference_rate(n_systems, steps, warmup, threads)
            return r, dt, "reference"
        except Exception as e:  # a broken copy must not take the bench line down
            print(f"[bench] unmodified reference failed ({type(e).__name__}: {e}); timing the port", file=sys.stderr)
    r, dt = port_rate(n_systems, steps, warmup, threads)
    return r, dt, "port"


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # size the sample so that (steps + warmup) reference steps end within ~2 minutes
    probe_rate, _, kind = cpu_rate(1, 1, 0, cores)
    total_steps = args.steps + args.warmup
    n_sys = max(1, min(16, int(probe_rate * 120.0 / total_steps)))
    value, elapsed, kind = cpu_rate(n_sys, args.steps, args.warmup, cores)
    what = ("UNMODIFIED reference PaiNN + Denoiser loop imported from baseline/_ref (absent third-party packages "
            "shimmed, oracle/ref_shims)" if kind == "reference" else "oracle port of the reference")
    sample = f"{n_sys} systems x {args.steps} steps of the same synthetic workload, fp32, {cores} threads: {what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.systems), "reference_sample_systems": n_sys},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- our arm
def se3_call(_cabi, call, dev, pos, plan, tags, fixed, s_tr, s_rot, sched, step, B, max_upd):
    call("adk_se3_step", dev, _cabi.ptr(pos), _cabi.ptr(plan.cell_f32), _cabi.ptr(plan.atom_off), _cabi.ptr(tags),
         _cabi.ptr(fixed), _cabi.ptr(s_tr), _cabi.ptr(s_rot), _cabi.ptr(sched), _cabi.ptr(step), B, None,
         _cabi.ptr(max_upd), None)


def kernel_breakdown(model, plan, z, pos, tags, fixed, sched, step, max_upd, reps=3):
    """Per-entry-point device time of one step, CUDA events on the launching stream."""
    from adsorbdiff_b200 import _cabi

    dev = plan.device
    stream = torch.cuda.current_stream(dev)
    records = []
    orig_call = _cabi.call

    def timed_call(name, device, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        orig_call(name, device, *a)
        e1.record(stream)
        records.append((name, a, e0, e1))

    import adsorbdiff_b200.painn as painn_mod

    agg = {}
    for r in range(reps):
        records.clear()
        painn_mod.call = timed_call
        try:
            model._run(plan, z, pos)
            se3_call(_cabi, timed_call, dev, pos, plan, tags, fixed, plan.out[0], plan.out[1], sched, step, plan.B, max_upd)
        finally:
            painn_mod.call = orig_call
        torch.cuda.synchronize(dev)
        step.zero_()
        if r == 0:
            continue  # first instrumented pass is warm-up
        for name, a, e0, e1 in records:
            key, flops = name, 0.0
            if name == "adk_linear":
                M, N, K = a[4], a[5], a[6]
                key, flops = f"adk_linear[{N}x{K}]", 2.0 * M * N * K
            elif name == "adk_linear_tc":
                M, N, K = a[2], a[4], a[5]
                key, flops = f"adk_linear_tc[{N}x{K}]", 2.0 * M * N * K
            d = agg.setdefault(key, {"ms": 0.0, "launches": 0, "flops": 0.0})
            d["ms"] += e0.elapsed_time(e1) / (reps - 1)
            d["launches"] += 1.0 / (reps - 1)
            d["flops"] += flops / (reps - 1)
    return agg


class StepLoop:
    """K denoising steps over one resident batch, replayed as a CUDA graph (what `value` / `sampler_step` time)."""

    def __init__(self, model, batch, noise, pruned: bool):
        from adsorbdiff_b200 import _cabi
        from adsorbdiff_b200.denoiser import schedule_table

        self._cabi = _cabi
        self.model, self.batch = model, batch
        dev = batch.pos.device
        self.dev = dev
        if model._needs_calibration():
            model.calibrate(batch)   # (what the public forward / Denoiser do on first use)
        self.plan, self.z, self.pos = model._prepare(batch)
        self.tags = batch.tags.to(torch.int32).contiguous()
        self.fixed = batch.fixed.to(torch.int32).contiguous()
        B = self.plan.B
        _cabi.call("adk_init_placement", dev, _cabi.ptr(self.pos), _cabi.ptr(self.plan.cell_f32),
                   _cabi.ptr(self.plan.atom_off), _cabi.ptr(self.tags), _cabi.ptr(noise.to(dev).contiguous()), B)
        self.sched = schedule_table(SAMPLER_PARAMS, dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.max_upd = torch.zeros(B, dtype=torch.float32, device=dev)
        self.out_rows = None
        if pruned:
            flags = (self.tags == 2).to(torch.int32).contiguous()
            self.out_rows = (torch.nonzero(flags).flatten().to(torch.int32).contiguous(), flags)
        l0 = _cabi.launch_count
        self.one_step(False)  # eager: warms every kernel up, produces the weight operand planes
        self.launches_per_step = _cabi.launch_count - l0
        torch.cuda.synchronize(dev)
        model.check_status(self.plan)
        # (pruned: the sampler owns the parameters for its run, the weight operand planes are prepared once)
        self.graph = _cabi.capture_graph(lambda: self.one_step(pruned), dev)
        self.done = 1

    def one_step(self, weights_ready):
        self.model._run(self.plan, self.z, self.pos, weights_ready=weights_ready, out_rows=self.out_rows)
        se3_call(self._cabi, self._cabi.call, self.dev, self.pos, self.plan, self.tags, self.fixed, self.plan.out[0],
                 self.plan.out[1], self.sched, self.step, self.plan.B, self.max_upd)

    def run(self, k):
        for _ in range(k):
            if self.done % SAMPLER_PARAMS["num_steps"] == 0:
                self.step.zero_()  # the schedule wraps every 100 steps (a new sampling run)
            self.graph.replay()
            self.done += 1


def run_ours(args):
    import torch.distributed as dist

    from adsorbdiff_b200 import Denoiser, PaiNN, partition as PT, synthetic as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = _peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the job: S systems in total, contiguous atom-balanced partition over the ranks (config #3) -------------
    S_total = args.systems
    systems = global_systems(S_total)
    natoms_all = [len(s["pos"]) for s in systems]
    parts = PT.contiguous_partition(natoms_all, world)
    a, b = parts[rank]
    if b <= a:
        raise SystemExit(f"rank {rank} got no systems: --systems {S_total} is too small for {world} ranks")
    S_loc = b - a
    atoms_per_rank = [sum(natoms_all[s:e]) for s, e in parts]
    noise_all = PT.initial_noise(S_total, seed=1234)  # every rank draws the job-wide rows and slices its own
    host = S.collate(systems[a:b], sids=[f"sys{i}" for i in range(a, b)])
    N = int(host.pos.shape[0])

    model = PaiNN(None, 0, 1, so3_denoising=True).to(dev).eval()
    model.load_state_dict(S.random_state_dict(0), strict=True)

    def timed(loop_run, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop_run(steps)
        if after is not None:
            after()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- value: K full-forward steps, device timed, max over ranks; then the job's one collective --------------
    loop = StepLoop(model, host.clone().to(dev), noise_all[a:b], pruned=False)
    plan = loop.plan
    loop.run(max(args.warmup, 3))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    gathered = {}

    def final_gather():
        gathered["pos"] = PT.gather_positions(loop.pos, atoms_per_rank)

    ms = timed(loop.run, args.steps, final_gather if world > 1 else None)
    clocks = sampler.stop() if sampler else None
    value = S_total * args.steps / (ms / 1e3)
    model.check_status(plan)
    E_loc = int(plan.row_deg.sum().item())

    # ---- sampler_step: the step exactly as adsorbdiff_b200.Denoiser replays it ----------------------------------
    # (weights' operand planes prepared once per run; last message layer, its update block and the heads on the
    # adsorbate rows only -- the only rows the SE(3) update reads; positions bit-identical, tests/test_gpu_parity.py)
    ploop = StepLoop(model, host.clone().to(dev), noise_all[a:b], pruned=True)
    ploop.run(max(args.warmup, 3))
    sampler_ms = timed(ploop.run, args.steps)
    model.check_status(ploop.plan)
    sampler_launches = ploop.launches_per_step
    del ploop

    # ---- e2e: Denoiser.run() on a HOST-resident batch ----------------------------------------------------------
    # timed region, per run of K steps: batch H2D (pinned), plan, initial placement, eager first step + graph capture,
    # K - 1 replays, final positions D2H (pinned).  Two flavours: the product default (adsorbate-rows tail) and
    # full_forward=True (every atom's scores each step, directly comparable with `value`).
    def pin_batch(h):
        out = h.clone()
        for k, v in list(out.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(out, k, v.pin_memory())
        return out

    h_out = torch.empty(N, 3, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.__dict__.values() if isinstance(v, torch.Tensor))

    def e2e_run(full_forward):
        params = dict(SAMPLER_PARAMS, num_steps=args.steps, full_forward=full_forward)
        hb = pin_batch(host)

        def go(_steps):
            ts = [time.perf_counter()]
            d = hb.to(dev, non_blocking=True)
            ts.append(time.perf_counter())
            Denoiser(d, model, params, device=str(dev), init_noise=noise_all[a:b]).run()
            ts.append(time.perf_counter())
            h_out.copy_(d.pos, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            ts.append(time.perf_counter())
            if os.environ.get("ADK_BENCH_DEBUG") and rank == 0:
                print("[bench] e2e run: to(dev) %.1f ms, Denoiser.run enqueue %.1f ms, drain + D2H %.1f ms"
                      % tuple(1e3 * (ts[i + 1] - ts[i]) for i in range(3)), file=sys.stderr)

        # two warm-up runs: a run builds a new launch plan while the model still holds the previous one, so the caching
        # allocator reaches its steady state (two plan generations) only after the second
        go(args.steps)
        hb = pin_batch(host)
        go(args.steps)
        hb = pin_batch(host)
        return timed(go, args.steps)

    e2e_ms = e2e_run(False)
    e2e_full_ms = e2e_run(True) if not args.quick else None

    # ---- weak scaling (N > 1 only): every rank holds all S systems, as round 1 measured -----------------------
    weak = None
    if world > 1 and not args.quick:
        wsys = global_systems(S_total)
        wloop = StepLoop(model, S.collate(wsys).to(dev), PT.initial_noise(S_total, seed=99 + rank), pruned=False)
        wloop.run(max(args.warmup, 3))
        wms = timed(wloop.run, args.steps)
        weak = {"value": world * S_total * args.steps / (wms / 1e3), "unit": UNIT, "systems_per_gpu": S_total,
                "ms_per_step": wms / args.steps}
        del wloop

    # ---- partition check (N > 1): N-rank Denoiser + gather == rank 0 running all S systems alone --------------
    partition_check = None
    if world > 1:
        cparams = dict(SAMPLER_PARAMS, num_steps=3)
        d = host.clone().to(dev)
        Denoiser(d, model, cparams, device=str(dev), init_noise=noise_all[a:b]).run()
        merged = PT.gather_positions(d.pos, atoms_per_rank)
        if rank == 0:
            full = S.collate(systems).to(dev)
            Denoiser(full, model, cparams, device=str(dev), init_noise=noise_all).run()
            partition_check = bool(torch.equal(merged, full.pos))
            if not partition_check:
                print(f"[bench] partition check FAILED: max |d| = {float((merged - full.pos).abs().max()):.3e}",
                      file=sys.stderr)
            del full
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel breakdown + roofline of the dominant kernel (rank 0, eager, CUDA events) ----
    loop.step.zero_()
    agg = kernel_breakdown(model, plan, loop.z, loop.pos, loop.tags, loop.fixed, loop.sched, loop.step, loop.max_upd)
    total_ms = sum(d["ms"] for d in agg.values())
    F, L = model.hidden_channels, model.num_layers
    msg_bytes_per_launch = N * 20480.0 * (F / 512.0) + E_loc * 24.0  # fused-design algorithmic bytes
    kernels = {}
    msg_keys = [k for k in agg if k.startswith("adk_message")]
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        ent = {"ms_per_step": round(d["ms"], 4), "share": round(d["ms"] / total_ms, 4), "launches": round(d["launches"])}
        if k in msg_keys:
            per = d["ms"] / max(d["launches"], 1)
            ent["hbm_gbs"] = round(msg_bytes_per_launch / (per * 1e-3) / 1e9, 1)
            ent["hbm_frac"] = round(ent["hbm_gbs"] / peaks["hbm"], 4)
        if d["flops"] > 0:
            ent["tflops"] = round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 2)
            if k.startswith("adk_linear_tc"):
                # fp32-equivalent product rate; the kernel issues 3 fp16 MMAs per product (fp16x2 split)
                ent["mma_tflops_fp16"] = round(3 * ent["tflops"], 1)
        kernels[k] = ent
    msg_ms = sum(agg[k]["ms"] for k in msg_keys)
    msg_launches = sum(agg[k]["launches"] for k in msg_keys)
    tc_ms = sum(d["ms"] for k, d in agg.items() if k.startswith("adk_linear_tc"))
    tc_flops = sum(d["flops"] for k, d in agg.items() if k.startswith("adk_linear_tc"))
    tensor_peak = peaks["bf16_sustained"] or peaks["bf16"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram bytes per launch from the committed ncu --set full capture, if sizes match
        tj = json.load(open(tpath))
        if tj.get("systems") == S_loc:
            traffic = tj
    if tc_ms >= msg_ms:
        ach = tc_flops / (tc_ms * 1e-3) / 1e12
        roofline = {"kernel": "linear_tc_kernel (node-wise GEMMs, tcgen05 fp16x2 split: 3 fp16 MMAs per fp32 product)",
                    "bound": "tensor", "achieved": round(ach, 1), "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": round(ach / tensor_peak, 4), "traffic": None,
                    "algorithmic_flops_per_step": tc_flops,
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "note": "achieved counts the ALGORITHMIC fp32 products (2MNK); the tensor pipe issues three fp16 "
                            "MMAs per product, so its own rate is 3x this"}
    else:
        per = msg_ms / max(msg_launches, 1)
        ach = msg_bytes_per_launch / (per * 1e-3) / 1e9
        roofline = {"kernel": msg_keys[0] + " (fused rbf + rbf_proj + message + CSR segmented reduction)", "bound": "hbm",
                    "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm"], 4),
                    "traffic": traffic["message_dram_bytes_per_launch"] if traffic else None,
                    "algorithmic_bytes_per_launch": msg_bytes_per_launch,
                    "peak_source": peaks["source"],
                    "note": "compute/shared-memory bound by design (per-edge tensors never reach HBM); see DESIGN.md 4"}
        # the same kernel against the tensor roofline: algorithmic rbf_proj FLOPs (2*R*3F per edge)
        rbf_flops = 2.0 * model.num_rbf * 3 * F * E_loc
        roofline["tensor_view"] = {"algorithmic_tflops": round(rbf_flops / (per * 1e-3) / 1e12, 1), "peak": tensor_peak,
                                   "frac": round(rbf_flops / (per * 1e-3) / 1e12 / tensor_peak, 4),
                                   "note": "dense-equivalent rbf_proj FLOPs (the kernel evaluates only the band of "
                                           "centres an edge tile touches, as fp16x2-split MMAs)"}
    roofline["tensor_kernel"] = {"name": "linear_tc_kernel", "ms_per_step": round(tc_ms, 3),
                                 "algorithmic_tflops": round(tc_flops / max(tc_ms, 1e-9) / 1e9, 1),
                                 "frac_of_peak": round(tc_flops / max(tc_ms, 1e-9) / 1e9 / tensor_peak, 4),
                                 "mma_tflops_fp16": round(3.0 * tc_flops / max(tc_ms, 1e-9) / 1e9, 1)}
    roofline["message_kernel"] = {"ms_per_step": round(msg_ms, 3), "launches": round(msg_launches)}

    # ---- reference baselines (rank 0) ----------------------------------------------------------------------
    cores = os.cpu_count() or 1
    # bounded sample of the same workload: 16 systems x 5 steps is 10-12 s of the reference on 16 cores (--quick: 1 s)
    cpu_sys, cpu_steps = (4, 2) if args.quick else (16, 5)
    cpu_val, cpu_dt, cpu_kind = cpu_rate(cpu_sys, cpu_steps, 1, cores)
    reference_gpu = None
    if world == 1 and not args.quick:
        from oracle import ref_import

        if ref_import.available():
            try:
                n_ref = 64
                r, dt_ = reference_rate(n_ref, 3, 1, cores, device=str(dev))
                reference_gpu = {"value": r, "unit": UNIT, "systems": n_ref,
                                 "what": "the UNMODIFIED reference's own torch ops on this B200 (device='cuda'; "
                                         "torch_scatter / PyG propagate stood in for by index_add_-based shims): "
                                         f"{n_ref} systems x 3 steps in {dt_:.2f} s"}
            except Exception as e:
                reference_gpu = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(S_total), "systems_total": S_total, "systems_per_gpu": S_loc,
                   "atoms_per_gpu": N, "edges_per_gpu": E_loc, "partition": "contiguous, atom balanced "
                   "(adsorbdiff_b200.partition.contiguous_partition); job-wide initial noise sliced per rank",
                   "l2": "inputs larger than L2 (per-step activation working set ~%.1f GB per GPU)" % (N * F * 4 * 22 / 1e9),
                   "cuda_graph": True, "early_stop": False,
                   "parity_tolerance": "max|err|/max|ref| < 1e-5 per tensor vs fp64 (tests/test_gpu_parity.py); "
                                       "positions 1e-6 A after one step"},
        "e2e": {"value": S_total * args.steps / (e2e_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": h2d_bytes / args.steps, "d2h_bytes_per_step": N * 12 / args.steps,
                "what": f"Denoiser.run() of {args.steps} steps on a pinned HOST batch per rank: batch H2D, plan, placement, "
                        "eager first step, CUDA-graph capture, replays, final positions D2H -- all timed; bytes are "
                        "per run / K",
                "full_forward": ({"value": S_total * args.steps / (e2e_full_ms / 1e3), "unit": UNIT,
                                  "what": "same with denoising_pos_params['full_forward']=True (every atom's scores "
                                          "each step): the figure comparable with `value`"} if e2e_full_ms else None)},
        "sampler_step": {"value": S_total * args.steps / (sampler_ms / 1e3), "unit": UNIT,
                         "ms_per_step": sampler_ms / args.steps, "launches_per_step": sampler_launches,
                         "what": "the step as adsorbdiff_b200.Denoiser replays it: last layer + heads on the adsorbate "
                                 "rows only, weight operand planes prepared once per run; positions identical to the "
                                 "full forward timed by `value`"},
        "gpu_launches": loop.launches_per_step * args.steps,
        "launches_per_step": loop.launches_per_step,
        "clocks": clocks,
        "roofline": roofline,
        "kernels": kernels,
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": cpu_kind,
                         "sample": f"{cpu_sys} systems x {cpu_steps} steps (+1 warm-up) of the same workload in {cpu_dt:.1f} s, "
                                   f"fp32, {cores} threads"},
    }
    if weak is not None:
        out["weak"] = weak
    if partition_check is not None:
        out["partition_check"] = partition_check
    if reference_gpu is not None:
        out["reference_gpu"] = reference_gpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- training step
TRAIN_OPTIM = dict(   # configs/denoising/painn_so3.yml:56-87
    optimizer="AdamW", optimizer_params=dict(weight_decay=0.001), lr_initial=1e-4, clip_grad_norm=100, ema_decay=0.999,
    denoising_pos_params=dict(num_steps=100, ads_std_low=0.1, ads_std_high=10, rot_std_low=0.01, rot_std_high=1.55,
                              free_std_low=0.01, free_std_high=0.1))
TRAIN_BATCH = 48      # systems per GPU (optim.batch_size of the same file)


def reference_train_rate(batch, steps, warmup, device):
    """systems/s of the reference's model code (PaiNN from baseline/_ref, its torch ops + torch autograd) doing the
    same optimisation step on `device`: forward, loss, backward, clip, AdamW, EMA.  The trainer class cannot be
    imported without the dataset/registry stack, so noising and the loss are this repo's torch restatement of them
    (adsorbdiff_b200.train) -- they are a negligible share of the step on either side."""
    import logging

    from adsorbdiff_b200 import synthetic as S, train as T
    from oracle import ref_import

    logging.disable(logging.INFO)
    ns = ref_import.load()
    model = ns.PaiNN(None, 0, 1, scale_file=ns.scale_file, so3_denoising=True)
    model.load_state_dict(S.random_state_dict(0), strict=True)
    model = model.to(device).train()
    ns.utils.radius_graph_pbc.__defaults__[-1][:] = [True, True, True]
    params = [q for q in model.parameters() if q.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.001)
    shadow = [q.detach().clone() for q in params]
    tables = T.IGSO3Tables(device)
    t0 = None
    for i in range(steps + warmup):
        if i == warmup:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        nb = T.tr_so3_schedule(batch.clone().to(device), TRAIN_OPTIM["denoising_pos_params"], tables)
        out = model(nb)
        loss = T.denoising_loss(out, nb, tables)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, max_norm=100)
        opt.step()
        torch._foreach_lerp_(shadow, [q.detach() for q in params], 1e-3)
        float(loss)
    torch.cuda.synchronize()
    return batch.num_graphs * steps / (time.perf_counter() - t0), torch.cuda.max_memory_allocated(device) / 2**30


def run_train(args):
    """`python bench.py --train`: the optimisation step of SURVEY.md section 8 row f-1 at the reference's batch size
    (48 systems per GPU, weak scaling: data parallel, one gradient all-reduce per step).  A step = noising
    (tr_so3_schedule) + forward + loss + backward + all-reduce + clip + AdamW + EMA, on a batch copied from pinned
    host memory inside the timed region, every step's loss read on the host (one step late; the reference logs
    `loss.item()` every step)."""
    import torch.distributed as dist

    from adsorbdiff_b200 import PaiNN, synthetic as S, train as T
    from adsorbdiff_b200 import _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.train_batch
    net = PaiNN(None, 0, 1, so3_denoising=True).to(dev)
    net.load_state_dict(S.random_state_dict(0), strict=True)
    step = T.TrainStep(net, TRAIN_OPTIM, T.IGSO3Tables(dev), generator=torch.Generator(device=dev).manual_seed(rank))
    # a few different pinned host batches, cycled (every step sees a batch structure it has not just seen)
    hosts = []
    for k in range(4):
        h = S.collate([S.make_system((rank * 4 + k) * B + i) for i in range(B)])
        for name, v in list(h.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(h, name, v.pin_memory())
        hosts.append(h)
    h2d = sum(v.numel() * v.element_size() for v in hosts[0].__dict__.values() if isinstance(v, torch.Tensor))
    losses = []

    def one(i):
        # every step's loss is read on the host (the reference logs loss.item() every step) -- one step late, through
        # TrainStep.read_loss, so that the read does not drain the stream the next step is already queued on
        # (the HOST batch goes in: TrainStep.to_device copies it to the GPU without touching it -- `SystemBatch.to`
        # itself moves a batch in place, like PyG's -- and keeps the plan's metadata on the host)
        step(hosts[i % len(hosts)])
        if step.step_count > 1:
            losses.append(step.read_loss(lag=1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        one(i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _cabi.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one(i)
    losses.append(step.read_loss(lag=0))   # the last step's loss: inside the timed region as well
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    launches = (_cabi.launch_count - launches0) // args.steps
    mem = torch.cuda.max_memory_allocated(dev) / 2**30
    in_sync = step.params_in_sync()      # every rank holds bit-identical parameters after the run
    step.check_gemm_status()
    ref = None
    if rank == 0 and not args.quick:
        try:
            ref_b = min(B, args.train_ref_batch)
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
            ref_rate, ref_mem = reference_train_rate(S.collate([S.make_system(i) for i in range(ref_b)]), 3, 2, str(dev))
            ref = {"value": ref_rate, "unit": "systems/s", "batch": ref_b, "peak_mem_gib": ref_mem,
                   "what": "reference PaiNN module (its torch ops, torch autograd) on this GPU, same step"}
        except Exception as e:
            ref = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if rank == 0:
        print(json.dumps({
            "metric": "train_systems_per_s", "value": world * B * args.steps / (ms / 1e3), "unit": "systems/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"painn_so3 training step, {B} systems/GPU (configs/denoising/painn_so3.yml), "
                                   "AdamW + clip + EMA, noising included", "l2": "a different batch every step"},
            "e2e": {"value": world * B * args.steps / (ms / 1e3), "unit": "systems/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "peak_mem_gib": mem, "loss_first_last": [losses[0], losses[-1]],
            "params_in_sync": in_sync,
            "reference_gpu": ref, "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--systems", type=int, default=1024, help="systems in the whole job (split over the ranks)")
    ap.add_argument("--quick", action="store_true", help="skip the secondary legs (weak, e2e full_forward, reference_gpu)")
    ap.add_argument("--train", action="store_true", help="time the training step (row f-1) instead of the sampler")
    ap.add_argument("--train-batch", type=int, default=TRAIN_BATCH, help="systems per GPU for --train")
    ap.add_argument("--train-ref-batch", type=int, default=TRAIN_BATCH, help="batch of the reference_gpu leg of --train")
    args = ap.parse_args()
    if os.environ.get("BENCH_WATCHDOG"):
        # debugging aid: dump every thread's Python stack and exit if the run is still going after that many seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    if args.train:
        run_train(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
