#!/usr/bin/env python
"""bench.py — SQP-RTI iterations/s of the batched pusher-slider NMPC hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement on the host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, independent shards, no collective

Workload (config.workload)
  N = 1 : BASELINE config 3 — 4096 independent NMPC instances, N = 40, santal, random initial poses (seed 2).
  N > 1 : BASELINE config 4 — 65 536 instances in total, Monte-Carlo over santal / balea / montana / pulirapid
          (object_id = i mod 4, seed 3), cut into contiguous shards (sharding.shard_range), every shard bucketed by object;
          STRONG scaling: value = 65 536 * steps / (max over ranks of the device time).
One step = one control period for every instance = NMPC_controller.solve: x0 wrap, reference window, v_bound clip, Euler
rollout, linearisation (ERK4 + forward sensitivities), QP (Riccati IPM: residuals 1e-11, complementarity 1e-18), full step,
u0 and status out — qspush_step, ONE CUDA graph of four kernels.

`value` : x0 and (u0, status) in device memory, CUDA events on the solver's stream around every step, L2 flushed in between.
`e2e`   : the same call with HOST buffers: x0 from pinned host memory (H2D), u0 and status back to pinned host memory (D2H),
          host wall clock around the call.  The reference trajectory lives on the device like controller.y_ref lives in the
          controller (handed over once), the warm start is the solver's own state.
          `e2e_all_fields_from_host` re-sends every acados-level field (x0, all stage references, init_u) each step.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC, UNIT = "sqp_rti_iterations_per_sec", "iterations/s"
HORIZON, DT = 40, 0.05
C3_BATCH, C3_OBJECT = 4096, "santal"
C4_TOTAL = 65536
C5_TOTAL, C5_HORIZON = 262144, 100
QP_TOL, QP_TOL_COMP = 1e-11, 1e-18
ALG_BYTES_PER_ITER = 8 * (26 * HORIZON + 16)          # 8448 B at N = 40
FLOP_PER_STAGE_DYN, FLOP_PER_STAGE_LIN, FLOP_PER_STAGE_QP = 2300.0, 40.0, 1100.0
ERK4_BYTES_PER_SAMPLE, ERK4_FLOP_PER_SAMPLE = 48 + 224, 2300.0                      # config 2: (x, u) in, (Phi, A, B) out


def _load_by_path(name, relpath):
    """Import one module file WITHOUT its package __init__ (the CPU arm must not dlopen libqspush.so)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


WL = _load_by_path("_qs_workloads", "uclv_qs_pushing_matlab_b200/workloads.py")
OBJ = _load_by_path("_qs_objects", "uclv_qs_pushing_matlab_b200/object_selection.py")
SHARD = _load_by_path("_qs_sharding", "uclv_qs_pushing_matlab_b200/sharding.py")
OBJECT_ORDER = WL.OBJECT_ORDER


def alg_flops_per_iter(k_ipm: float, horizon: int = HORIZON) -> float:
    return horizon * (FLOP_PER_STAGE_DYN + FLOP_PER_STAGE_LIN) + k_ipm * (horizon + 1) * FLOP_PER_STAGE_QP


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_workload(world, rank):
    """The rank's share of the benchmark workload: (workload dict, object names, instances in the whole job)."""
    if world == 1:
        wl = WL.make_rti_workload(C3_BATCH, HORIZON, dt=DT, seed=2)
        return wl, [C3_OBJECT], C3_BATCH
    wl = WL.make_rti_workload(C4_TOTAL, HORIZON, dt=DT, seed=3, n_objects=4)
    lo, hi = SHARD.shard_range(C4_TOTAL, world, rank)
    sub = {k: v[lo:hi] for k, v in wl.items()}
    perm, _ = SHARD.bucket_by_object(sub["object_id"], 1)          # contiguous per-object buckets inside the shard
    return {k: np.ascontiguousarray(v[perm]) for k, v in sub.items()}, list(OBJECT_ORDER), C4_TOTAL


def config_dict(world):
    common = {"horizon": HORIZON, "qp_tol": QP_TOL, "qp_tol_comp": QP_TOL_COMP, "mode": "sqp_rti",
              "l2": "256 MiB device memset between timed steps (outside each step's CUDA-event window)"}
    if world == 1:
        return dict(common, workload=f"config3: {C3_BATCH} independent SQP-RTI NMPC instances, N={HORIZON}, dt={DT}, {C3_OBJECT}, seed 2",
                    batch_per_gpu=C3_BATCH, global_batch=C3_BATCH, parallelism="1 GPU")
    return dict(common, workload=f"config4: {C4_TOTAL} SQP-RTI NMPC instances Monte-Carlo over santal/balea/montana/pulirapid "
                                 f"(object_id = i mod 4, seed 3), N={HORIZON}, dt={DT}, sharded over {world} GPUs",
                batch_per_gpu=C4_TOTAL // world, global_batch=C4_TOTAL,
                parallelism=f"{world} contiguous shards (shard_range), per-object buckets inside a shard, no collective on the solve path")


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_models(names):
    from oracle import oracle as orc
    tabs = WL.packaged_tables()
    return [orc.Model.create(tabs[n]["knots"], tabs[n]["ctrl_xy"], 3, OBJ.OBJECT_TABLE[n]["mu_sp"], tabs[n]["c_ellipse"], True) for n in names]


def cpu_oracle_rate(wl, names, nthreads, passes=1, sample=None, **opts):
    """Oracle restatement (NMPC_controller.solve pre-processing + one RTI iteration) on the host cores over the first `sample`
    instances of the workload (all when None); returns (instances per pass, [seconds per pass])."""
    from oracle import oracle as orc
    oms = oracle_models(names)
    n = len(wl["x0"]) if sample is None else min(sample, len(wl["x0"]))
    parts = []
    for o, om in enumerate(oms):
        idx = np.where(wl["object_id"][:n] == o)[0]
        if len(idx):
            parts.append((orc.Ocp(om, HORIZON, DT, **opts), {k: np.ascontiguousarray(v[idx]) for k, v in wl.items()}))
    times = []
    for _ in range(passes):
        t0 = time.perf_counter()
        for ocp, w in parts:
            nb = len(w["x0"])
            pr = ocp.prepare(w["x0"], np.zeros(nb, dtype=np.int32), np.zeros((nb, HORIZON + 1, 4)), w["u_init"], nthreads=nthreads)
            ocp.solve("rti", pr["x0"], w["yref"], w["yref_e"], pr["x"], pr["u"], nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    return n, times


R01_QP_RULE = dict(qp_tol=1e-6, qp_tol_comp=1e-6, qp_t_min=0.0, qp_gamma_f=0.0, qp_stall=5)   # the reference's own QP tolerance (NMPC_controller.m:276)


def cpu_baseline_object(wl, names, passes, sample, what):
    from oracle import oracle as orc
    orc.select_build("fast")
    cores = os.cpu_count() or 1
    cpu_oracle_rate(wl, names, cores, 1, sample=min(sample, 512))          # warm-up (thread pool, page faults)
    n, t = cpu_oracle_rate(wl, names, cores, passes, sample=sample)
    n1, t1 = cpu_oracle_rate(wl, names, 1, 1, sample=max(64, sample // 16))
    n6, t6 = cpu_oracle_rate(wl, names, cores, 1, sample=sample, **R01_QP_RULE)
    return n, t, {"value": n * len(t) / sum(t), "unit": UNIT, "cores": cores, "kind": "port",
                  "sample": f"{what}: {len(t)} pass(es) over the first {n} instances of the workload, restated C++ oracle (not acados) built "
                            f"-O3 -march=x86-64-v3, {cores} threads, same QP tolerances as the GPU arm",
                  "single_thread": {"value": n1 / sum(t1), "instances": n1},
                  "at_reference_qp_tol_1e-6": {"value": n6 / sum(t6), "instances": n6,
                                               "what": "all four IPM tolerances at 1e-6 (NMPC_controller.m:276), fixed fraction to the boundary"}}


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0                                             # rank 0 alone runs the CPU arm
    wl, names, total = shard_workload(args.gpus, 0)
    sample = min(len(wl["x0"]), args.cpu_sample)
    n, times, cb = cpu_baseline_object(wl, names, max(1, args.steps), sample, "each step")
    t = sum(times)
    value = n * len(times) / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": 1, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
            "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "acados v0.2.1 / MATLAB are not installable offline; the CPU arm is the C++ oracle restatement of the same algorithm "
                    "(the throughput of a CPU does not depend on the batch size, so a bounded sample stands for the workload)"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle-reason sampler for the timed region.

    NVML in a thread (every 5 ms, initialised before the region starts, so that even a 40 ms region is sampled under load);
    `nvidia-smi -lms` as the fallback when the NVML binding is missing (its first row arrives only after its own start-up).
    """
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, uuid=None):
        self.index, self.rows, self.proc = index, [], None
        self.nv, self.h, self.samples, self.run, self.thr = None, None, [], False, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if uuid is not None:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID("GPU-" + str(uuid))
                except Exception:
                    h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                h = nv.nvmlDeviceGetHandleByIndex(int(ids[index]) if index < len(ids) else index)
            self.nv, self.h = nv, h
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv, h = self.nv, self.h
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.run:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(reasons(h))))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self.run = True
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.nv is not None:
            self.run = False
            self.thr.join(timeout=1.0)
            sm = [s[0] for s in self.samples]
            bits = 0
            for s in self.samples:
                bits |= s[1]
            # NVML clocks-event-reason bits: sw_power_cap 0x4, hw_slowdown 0x8, sw_thermal 0x20, hw_thermal 0x40
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(n for b, n in names.items() if bits & b), "samples": len(sm), "source": "nvml, 5 ms period, inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import uclv_qs_pushing_matlab_b200 as q
    from uclv_qs_pushing_matlab_b200.workloads import packaged_model

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl, names, total = shard_workload(world, rank)
    B, N = len(wl["x0"]), HORIZON
    gms = [packaged_model(n) for n in names]
    solver = q.Solver(gms, N, DT, B, device=local_rank, qp_tol=QP_TOL, qp_tol_comp=QP_TOL_COMP, problems_per_warp=args.ppw, qp_kernel=args.qp_kernel)
    solver.order_with_torch = False                              # this file orders the streams itself (synchronize before every timed step)
    stream = torch.cuda.ExternalStream(solver.stream, device=dev)
    if len(names) > 1:
        solver.set_int("object_id", wl["object_id"])

    d_x0 = torch.from_numpy(wl["x0"]).to(dev)
    d_u0 = torch.empty(B, 2, dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    h_in = {k: torch.from_numpy(wl[k]).pin_memory() for k in ("x0", "yref", "yref_e", "u_init")}
    h_cold = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_u0 = torch.empty(B, 2, dtype=torch.float64).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # controller state: reference trajectory handed over once (NMPC_controller.set_reference_trajectory, :425-431), the
    # initial guess saved on the device so that every timed period solves the SAME problems
    traj = np.zeros((N, 6)); traj[:, 0] = 0.01 * (np.arange(N) * DT)
    off = np.zeros((B, 6)); off[:, :2] = wl["x0"][:, :2]
    solver.set_reference_trajectory(traj, off)
    solver.set("u", wl["u_init"]); solver.set_int("cold", h_cold); solver.snapshot_guess()

    def step_fields():                                           # acados_ocp-level calls, every field from the host
        solver.set("x0", h_in["x0"]); solver.set("yref", h_in["yref"]); solver.set("yref_e", h_in["yref_e"])
        solver.set("u", h_in["u_init"]); solver.set_int("cold", h_cold)
        solver.prepare(); solver.solve()
        solver.get("u", stage=0, out=h_u0); solver.get_int("status", out=h_status)     # host gets synchronise the stream

    def step_device():
        solver.step(d_x0, 1, d_u0, d_status, restore_guess=True)

    def step_host():
        solver.step(h_in["x0"], 1, h_u0, h_status, restore_guess=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the three variants are the same arithmetic: bit-identical u0
    step_fields(); u0_ref = h_u0.numpy().copy()
    solver.set_reference_window(1)
    if not (np.array_equal(solver.get("yref"), wl["yref"]) and np.array_equal(solver.get("yref_e"), wl["yref_e"])):
        raise SystemExit("bench: the device-side reference window differs from the workload's references")
    for _ in range(max(args.warmup, 3)):
        step_device()
    solver.sync()
    step_host()
    if not (np.array_equal(h_u0.numpy(), u0_ref) and np.array_equal(d_u0.cpu().numpy(), u0_ref)):
        raise SystemExit("bench: controller-level step and field-by-field step disagree")

    # ---- value: K steps, CUDA events on the solver's stream around each step, L2 flushed in between
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    qp_ms, lin_ms, prep_ms, launches0 = [], [], [], solver.launches
    barrier()
    sampler.start()
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        ev[i][0].record(stream)
        step_device()
        ev[i][1].record(stream)
        solver.sync()
        qp_ms.append(1e3 * solver.stat("time_qp_sol")); lin_ms.append(1e3 * solver.stat("time_lin")); prep_ms.append(1e3 * solver.stat("time_prep"))
    barrier()
    clocks = sampler.stop()
    launches = solver.launches - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_local = sum(step_ms) / 1e3

    # ---- e2e: host buffers through the C-ABI, wall clock
    def timed(fn):
        ms = []
        barrier()
        for i in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fn()
            ms.append(1e3 * (time.perf_counter() - t0))
        barrier()
        return ms
    e2e_ms = timed(step_host)
    t_e2e_local = sum(e2e_ms) / 1e3
    t_e2e_fields_local = sum(timed(step_fields)) / 1e3
    it = solver.get_int("qp_iter")
    st = solver.get_int("status")
    res = solver.get("res")
    k_ipm = float(it.mean())
    per_rank = None
    if world > 1:
        tt = torch.tensor([t_local, t_e2e_local, t_e2e_fields_local], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        per_rank = [[float(v) for v in a] for a in allt]
        t_max, t_e2e_max, t_e2e_fields_max = (max(a[j] for a in per_rank) for j in range(3))
        # final host gather of the small per-problem result (outside every timed region; no collective on the solve path)
        u0_all = SHARD.gather_to_rank0(h_u0.numpy(), total, world, rank)
        assert rank != 0 or u0_all.shape == (total, 2)
    else:
        t_max, t_e2e_max, t_e2e_fields_max = t_local, t_e2e_local, t_e2e_fields_local

    # ---- config 5 (full SQP to convergence, N = 100, mixed sticking / sliding start, four shapes): this rank's share
    c5 = c5f = None
    if world > 1 or args.config5:
        for feasible in (False, True):
            part = config5_share(q, torch, dev, local_rank, world, rank, feasible=feasible)
            parts = [part]
            if world > 1:
                parts = [None] * world
                dist.all_gather_object(parts, part)
            if feasible:
                c5f = merge_config5(parts, feasible=True)
            else:
                c5 = merge_config5(parts)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    total_iters = total * args.steps
    value = total_iters / t_max
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    qp_avg_ms = sum(qp_ms) / len(qp_ms)
    lin_avg_ms, prep_avg_ms = sum(lin_ms) / len(lin_ms), sum(prep_ms) / len(prep_ms)
    step_avg_ms = sum(step_ms) / len(step_ms)
    achieved_gbs = B * ALG_BYTES_PER_ITER / (qp_avg_ms * 1e-3) / 1e9
    traffic, pipe_pct, prof_src = None, None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_qp_traffic.json")))
        traffic, pipe_pct, prof_src = prof["dram_bytes_per_launch"], prof.get("fp64_pipe_active_pct_of_elapsed"), prof.get("source")
    except Exception:
        pass
    # FP64 roofline denominator measured here: DFMA issue-rate microbenchmark of the library (cuBLAS DGEMM beside it)
    fp64_peak_tf = q.measure_fp64_peak(local_rank)
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev); b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(dev)
    dgemm_tf = 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    fp64_ach_tf = B * alg_flops_per_iter(k_ipm) / (qp_avg_ms * 1e-3 + lin_avg_ms * 1e-3) / 1e12

    extras = {}
    if world == 1:
        extras = single_gpu_extras(q, torch, dev, local_rank, gms[0], wl, solver, fp64_peak_tf, hbm_peak, args)
        extras["cpu_baseline"] = cpu_baseline_object(wl, names, args.cpu_passes, args.cpu_sample, "bounded sample")[2]
    srt = sorted(step_ms)
    srt_e = sorted(e2e_ms)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(world),
        "e2e": {"value": total_iters / t_e2e_max, "unit": UNIT,
                "h2d_bytes_per_step": int(h_in["x0"].numel() * 8),
                "d2h_bytes_per_step": int(h_u0.numel() * 8 + h_status.numel() * 4), "ms_per_step": 1e3 * t_e2e_max / args.steps,
                "p50_ms": srt_e[len(srt_e) // 2], "p99_ms": srt_e[min(len(srt_e) - 1, int(0.99 * len(srt_e)))],
                "what": "controller-level call = NMPC_controller.solve(x0, index_time) for the batch through the C-ABI (qspush_step, one CUDA graph): "
                        "x0 from pinned host memory, reference window on the device (trajectory set once), initial guess restored from its device "
                        "snapshot, prepare + linearise + QP + step, u0 and status to pinned host memory; same problems and bit-identical u0 as `value`"},
        "e2e_all_fields_from_host": {"value": total_iters / t_e2e_fields_max, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(v.numel() * 8 for v in h_in.values()) + h_cold.numel() * 4),
                "d2h_bytes_per_step": int(h_u0.numel() * 8 + h_status.numel() * 4), "ms_per_step": 1e3 * t_e2e_fields_max / args.steps,
                "what": "acados_ocp-level calls: x0, every stage's cost_y_ref, cost_y_ref_e and init_u re-sent from pinned host memory each step (9 launches)"},
        "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": ("k_qp_warp<3,0,16> (Mehrotra IPM, two problems per warp, parallel-in-time Riccati scans, state in shared memory + TMEM, ordered work queue)" if args.qp_kernel else "k_qp (Riccati/Mehrotra IPM, one problem per thread)"), "achieved": achieved_gbs, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_source": prof_src, "peak_source": peak_src + " (of measured)",
                     "algorithmic_bytes_per_launch": B * ALG_BYTES_PER_ITER, "kernel_ms": qp_avg_ms,
                     "kernel_share_of_step": qp_avg_ms / step_avg_ms,
                     "fp64": {"achieved_tflops": fp64_ach_tf, "peak_tflops": fp64_peak_tf, "frac": fp64_ach_tf / fp64_peak_tf,
                              "peak_source": "qspush_measure_fp64_peak: hand-written DFMA issue-rate microbenchmark, measured in this run",
                              "cublas_dgemm_tflops": dgemm_tf, "flops_per_iteration": alg_flops_per_iter(k_ipm),
                              "ncu_pipe_fp64_active_pct": pipe_pct,
                              "what": "achieved = algorithmic flops of the serial Riccati IPM (SURVEY 8d) over (linearise + QP) time; ncu_pipe_fp64_active_pct = "
                                      "sm__pipe_fp64_cycles_active of the executed parallel-in-time algorithm"},
                     "note": "the path is FP64-pipe / dependent-chain bound, not HBM bound (SURVEY 8d): the HBM fraction is reported "
                             "because the schema asks for it, the fp64 object is the relevant roofline (see DESIGN.md)"},
        "k_ipm_mean": k_ipm, "k_ipm_max": int(it.max()), "status_ok_frac": float((st == 0).mean()),
        "qp_residuals_max": {"stat_eq_ineq": float(res[:, :3].max()), "comp": float(res[:, 3].max())},
        "phase_ms": {"prepare": prep_avg_ms, "linearise": lin_avg_ms, "qp": qp_avg_ms,
                     "step_minus_phases": step_avg_ms - (prep_avg_ms + lin_avg_ms + qp_avg_ms)},
        "latency_ms": {"p50": srt[len(srt) // 2], "p99": srt[min(len(srt) - 1, int(0.99 * len(srt)))], "max": srt[-1],
                       "what": "per-step device time of one batched control period (all instances of the batch finish together)"},
    }
    if per_rank is not None:
        line["per_rank_s"] = {"device": [a[0] for a in per_rank], "e2e": [a[1] for a in per_rank]}
    if c5 is not None:
        line["config5"] = c5
        line["config5_feasible_start"] = c5f
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def config5_share(q, torch, dev, local_rank, world, rank, feasible=False):
    """This rank's share of BASELINE config 5: 262 144 instances over the job, N = 100, full SQP (<= 30 iterations, merit
    backtracking) from a mixed sticking / sliding start, four shapes.  One timed solve after one warm-up solve.
    feasible = True: the feasible-start variant (tracking-size errors, initial guess inside the friction cone, the symmetric
    outline only — workloads.make_feasible_start_workload), a quarter of the job size, where full SQP converges."""
    from uclv_qs_pushing_matlab_b200.workloads import packaged_model
    n_job = C5_TOTAL if world > 1 else C5_TOTAL // 8              # a single GPU runs the share it would own on 8
    if feasible:
        n_job //= 4
        w5 = WL.make_feasible_start_workload(n_job, C5_HORIZON, dt=DT, seed=4)
        w5["object_id"][:] = OBJECT_ORDER.index("balea")
    else:
        w5 = WL.make_rti_workload(n_job, C5_HORIZON, dt=DT, seed=4, n_objects=4, mixed_modes=True)
    lo, hi = SHARD.shard_range(n_job, world, rank)
    sub = {k: v[lo:hi] for k, v in w5.items()}
    del w5
    perm, _ = SHARD.bucket_by_object(sub["object_id"], 1)
    sub = {k: np.ascontiguousarray(v[perm]) for k, v in sub.items()}
    B5 = len(sub["x0"])
    s5 = q.Solver([packaged_model(n) for n in OBJECT_ORDER], C5_HORIZON, DT, B5, device=local_rank, mode=1)
    s5.set_int("object_id", sub["object_id"])
    zeros = np.zeros(B5, dtype=np.int32)
    ms = None
    st = torch.cuda.ExternalStream(s5.stream, device=dev)
    for rep in range(2):
        s5.set("x0", sub["x0"]); s5.set("yref", sub["yref"]); s5.set("yref_e", sub["yref_e"]); s5.set("u", sub["u_init"]); s5.set_int("cold", zeros)
        s5.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); s5.prepare(); s5.solve(); e1.record(st); s5.sync()
        ms = e0.elapsed_time(e1)
    status, sqp_it, qp_it = s5.get_int("status"), s5.get_int("sqp_iter"), s5.get_int("qp_iter")
    per_shape = {}
    for o, n in enumerate(OBJECT_ORDER):
        m = sub["object_id"] == o
        per_shape[n] = {"instances": int(m.sum()), "status": {str(k): int(((status == k) & m).sum()) for k in (0, 1, 2, 3, 4)},
                        "sqp_iter_sum": int(sqp_it[m].sum()), "qp_iter_sum": int(qp_it[m].sum())}
    del s5
    return {"rank": rank, "instances": B5, "ms": ms, "sqp_iter_sum": int(sqp_it.sum()), "per_shape": per_shape, "n_job": n_job}


def merge_config5(parts, feasible=False):
    tot = sum(p["instances"] for p in parts)
    t = max(p["ms"] for p in parts) * 1e-3
    its = sum(p["sqp_iter_sum"] for p in parts)
    shapes = {}
    for n in OBJECT_ORDER:
        inst = sum(p["per_shape"][n]["instances"] for p in parts)
        stt = {k: sum(p["per_shape"][n]["status"][k] for p in parts) for k in ("0", "1", "2", "3", "4")}
        sq = sum(p["per_shape"][n]["sqp_iter_sum"] for p in parts)
        shapes[n] = {"instances": inst, "status": stt, "converged_frac": stt["0"] / max(inst, 1), "mean_sqp_iter": sq / max(inst, 1),
                     "mean_ipm_iter_per_qp": sum(p["per_shape"][n]["qp_iter_sum"] for p in parts) / max(1, sq)}
    conv = sum(s["status"]["0"] for s in shapes.values())
    shapes = {n: v for n, v in shapes.items() if v["instances"]}
    what = ("feasible start (tracking-size errors, guess inside the friction cone), symmetric outline (balea), seed 4" if feasible
            else "mixed sticking/sliding start, 4 shapes, seed 4")
    return {"workload": f"config5{'-feasible-start' if feasible else ''}: {tot} instances (job size {parts[0]['n_job']}), N={C5_HORIZON}, full SQP <= 30 iterations, "
                        f"merit backtracking, {what}",
            "instances": tot, "seconds": t, "sqp_iterations_per_s": its / t, "instances_per_s": tot / t, "converged_frac": conv / tot,
            "status_legend": "0 converged, 1 NaN, 2 iteration limit, 3 minimum step, 4 QP failure (acados v0.2.1 enum)",
            "per_shape": shapes, "per_rank_ms": [p["ms"] for p in parts]}


def single_gpu_extras(q, torch, dev, local_rank, gm, wl, solver, fp64_peak_tf, hbm_peak, args):
    """N = 1 only: config 2 (1M-sample ERK4 + sensitivity kernel), config 4 on one GPU (strong-scaling base), single-instance
    latency (the MEX drop-in case), device-resident closed loop."""
    from uclv_qs_pushing_matlab_b200.workloads import packaged_model
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # ---- config 2: 1M (x, u) samples through k_eval_erk4, device buffers, CUDA events
    n2 = 1 << 20
    x2, u2 = WL.make_samples_config2(gm.b, n2, seed=1, knots=gm.S)
    dx, du = torch.from_numpy(x2).to(dev), torch.from_numpy(u2).to(dev)
    Phi = torch.empty(n2, 4, dtype=torch.float64, device=dev); A = torch.empty(n2, 4, 4, dtype=torch.float64, device=dev); Bm = torch.empty(n2, 4, 2, dtype=torch.float64, device=dev)
    for _ in range(3):
        gm.eval_erk4_sens_device(dx, du, DT, Phi, A, Bm)
    ms2 = []
    for _ in range(10):
        flush.zero_(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gm.eval_erk4_sens_device(dx, du, DT, Phi, A, Bm); e1.record(); torch.cuda.synchronize(dev)
        ms2.append(e0.elapsed_time(e1))
    m2 = statistics.median(ms2)
    tf2, gb2 = n2 * ERK4_FLOP_PER_SAMPLE / (m2 * 1e-3) / 1e12, n2 * ERK4_BYTES_PER_SAMPLE / (m2 * 1e-3) / 1e9
    out["config2"] = {"workload": "config2: 1M (x, u) samples, ERK4 + forward sensitivities + B-spline contact geometry (k_eval_erk4), santal, seed 1",
                      "ms": m2, "gsamples_per_s": n2 / (m2 * 1e-3) / 1e9,
                      "fp64": {"achieved_tflops": tf2, "frac": tf2 / fp64_peak_tf, "flop_per_sample": ERK4_FLOP_PER_SAMPLE},
                      "hbm": {"achieved_gbs": gb2, "frac": gb2 / hbm_peak, "bytes_per_sample": ERK4_BYTES_PER_SAMPLE},
                      "timing": "median of 10 launches, CUDA events on the legacy default stream the stateless kernels run on, L2 flushed between launches"}
    del dx, du, Phi, A, Bm
    # ---- config 4 on ONE GPU: the base of the strong-scaling run
    w4 = WL.make_rti_workload(C4_TOTAL, HORIZON, dt=DT, seed=3, n_objects=4)
    perm, _ = SHARD.bucket_by_object(w4["object_id"], 1)
    w4 = {k: np.ascontiguousarray(v[perm]) for k, v in w4.items()}
    s4 = q.Solver([packaged_model(n) for n in OBJECT_ORDER], HORIZON, DT, C4_TOTAL, device=local_rank, qp_tol=QP_TOL, qp_tol_comp=QP_TOL_COMP)
    s4.set_int("object_id", w4["object_id"]); s4.order_with_torch = False
    traj = np.zeros((HORIZON, 6)); traj[:, 0] = 0.01 * (np.arange(HORIZON) * DT)
    off = np.zeros((C4_TOTAL, 6)); off[:, :2] = w4["x0"][:, :2]
    s4.set_reference_trajectory(traj, off)
    s4.set("u", w4["u_init"]); s4.set_int("cold", np.zeros(C4_TOTAL, dtype=np.int32)); s4.snapshot_guess()
    dx0 = torch.from_numpy(w4["x0"]).to(dev); du0 = torch.empty(C4_TOTAL, 2, dtype=torch.float64, device=dev); dst = torch.empty(C4_TOTAL, dtype=torch.int32, device=dev)
    st4 = torch.cuda.ExternalStream(s4.stream, device=dev)
    for _ in range(3):
        s4.step(dx0, 1, du0, dst, restore_guess=True)
    s4.sync()
    ms4 = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st4); s4.step(dx0, 1, du0, dst, restore_guess=True); e1.record(st4); s4.sync()
        ms4.append(e0.elapsed_time(e1))
    it4 = s4.get_int("qp_iter"); stat4 = s4.get_int("status")
    out["config4_one_gpu"] = {"workload": f"config4 on one GPU: {C4_TOTAL} instances, 4 shapes in contiguous buckets, N={HORIZON} (base of the strong-scaling run at N > 1)",
                              "ms_per_step": statistics.mean(ms4), "value": C4_TOTAL / (statistics.mean(ms4) * 1e-3), "unit": UNIT,
                              "k_ipm_mean": float(it4.mean()), "status_ok_frac": float((stat4 == 0).mean()),
                              "k_ipm_mean_per_shape": {n: float(it4[w4["object_id"] == o].mean()) for o, n in enumerate(OBJECT_ORDER)}}
    del s4, dx0, du0, dst
    # ---- per-solve latency of a single NMPC instance (the MEX drop-in case): host buffers through the C-ABI, wall clock
    lat_b1 = {}
    for tag, n1, mode in (("rti_N40", HORIZON, 0), ("rti_N10", 10, 0), ("sqp_N10", 10, 1)):
        w1 = WL.make_rti_workload(1, n1, dt=DT, seed=7)
        s1 = q.Solver([gm], n1, DT, 1, device=local_rank, qp_tol=QP_TOL, qp_tol_comp=QP_TOL_COMP, mode=mode)
        u1 = np.zeros((1, 2)); c1 = np.zeros(1, dtype=np.int32); st1 = np.zeros(1, dtype=np.int32)
        ts = []
        for i in range(args.latency_solves + 10):
            t0 = time.perf_counter()
            s1.set("x0", w1["x0"]); s1.set("yref", w1["yref"]); s1.set("yref_e", w1["yref_e"]); s1.set("u", w1["u_init"]); s1.set_int("cold", c1)
            s1.prepare(); s1.solve(); s1.get("u", stage=0, out=u1)
            if i >= 10:
                ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        lat_b1[tag] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[min(len(ts) - 1, int(0.99 * len(ts)))], "solves": len(ts),
                       "sqp_iter": int(s1.get_int("sqp_iter")[0]), "what": "acados_ocp-level calls, all fields from the host"}
        if mode == 0:                                            # controller-level: qspush_step (one graph launch per solve)
            tr1 = np.zeros((n1, 6)); tr1[:, 0] = 0.01 * (np.arange(n1) * DT)
            of1 = np.zeros((1, 6)); of1[:, :2] = w1["x0"][:, :2]
            s1.set_reference_trajectory(tr1, of1)
            s1.set("u", w1["u_init"]); s1.set_int("cold", c1); s1.snapshot_guess()
            ts = []
            for i in range(args.latency_solves + 10):
                t0 = time.perf_counter()
                s1.step(w1["x0"], 1, u1, st1, restore_guess=True)
                if i >= 10:
                    ts.append(1e3 * (time.perf_counter() - t0))
            ts.sort()
            lat_b1[tag + "_step"] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[min(len(ts) - 1, int(0.99 * len(ts)))], "solves": len(ts),
                                     "what": "qspush_step: x0 in, u0 / status out, one CUDA graph"}
        del s1
    out["latency_b1"] = lat_b1
    # ---- Monte-Carlo of whole pushes, device-resident (qspush_closed_loop): 20 control periods, nothing crosses PCIe inside the loop
    B = len(wl["x0"])
    Tn = 64
    traj = np.zeros((Tn, 6)); traj[:, 0] = 0.01 * DT * np.arange(Tn)
    off = np.zeros((B, 6)); off[:, :2] = wl["x0"][:, :2]
    trd, ofd = torch.from_numpy(traj).to(dev), torch.from_numpy(off).to(dev)
    for rep in range(2):
        xd = torch.from_numpy(wl["x0"].copy()).to(dev)
        solver.set_int("cold", torch.ones(B, dtype=torch.int32, device=dev)); solver.sync()
        t0 = time.perf_counter()
        rl = solver.closed_loop(trd, xd, 20, offset=ofd); solver.sync()
        tl = time.perf_counter() - t0
    out["closed_loop_device"] = {"periods": 20, "instances": B, "ms_per_period": 1e3 * tl / 20, "controller_solves_per_s": B * 20 / tl,
                                 "status_ok_frac": float((rl["status_log"] == 0).float().mean()),
                                 "what": "qspush_closed_loop: reference window, prepare, solve, plant step, shift per period on the device"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ppw", type=int, default=0, help="QP kernel packing (problems per warp), 0 = auto")
    ap.add_argument("--qp-kernel", type=int, default=1, help="1 = warp kernel (parallel-in-time, one or two problems per warp; what auto picks), 0 = one problem per thread")
    ap.add_argument("--latency-solves", type=int, default=1000, help="solves per B=1 latency measurement")
    ap.add_argument("--cpu-passes", type=int, default=2, help="passes of the oracle over the sample for cpu_baseline")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="instances per CPU pass (bounded sample of the workload)")
    ap.add_argument("--config5", action="store_true", help="also run this GPU's config-5 share at N = 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    # Exactly ONE line on stdout (the JSON line): libraries that chat on fd 1 (NCCL prints its version there under torchrun)
    # are sent to stderr for the whole run, the JSON line goes to the original stdout.
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _json_out
    sys.exit(main())
