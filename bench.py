#!/usr/bin/env python
"""bench.py — SQP-RTI iterations/s of the batched pusher-slider NMPC hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C-ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement on the host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, independent shards, no collective

Workload (config.workload): BASELINE config 3 — 4096 independent NMPC instances per GPU, N = 40, santal,
random initial poses (seed 2 + rank), one SQP-RTI iteration each = NMPC_controller.solve pre-processing
(x0 wrap, v_bound clip, Euler rollout) + linearisation (ERK4 + forward sensitivities) + QP (Riccati IPM to
KKT residuals 1e-11, complementarity 1e-18) + full step.  Weak scaling: every rank owns its own 4096 instances.

`value`  : inputs already resident in HBM, CUDA events on the solver's stream around every step.
`e2e`    : the same step through the C-ABI at the NMPC_controller.solve(x0, idx) boundary: x0 from pinned HOST memory
           (H2D) and (u0, status) back to pinned host memory (D2H) inside the timed region, host wall clock; the reference
           trajectory lives on the device like controller.y_ref lives in the controller (set once).
           `e2e_all_fields_from_host` re-sends every acados-level field (x0, all stage references, init_u) each step.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
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
BATCH_PER_GPU, HORIZON, DT, OBJECT = 4096, 40, 0.05, "santal"
QP_TOL = 1e-11            # stationarity / dynamics / inequality residuals; complementarity goes to qp_tol_comp = 1e-18
ALG_BYTES_PER_ITER = 8 * (26 * HORIZON + 16)          # 8448 B at N = 40
FLOP_PER_STAGE_DYN, FLOP_PER_STAGE_LIN, FLOP_PER_STAGE_QP = 2300.0, 40.0, 1100.0


def alg_flops_per_iter(k_ipm: float) -> float:
    return HORIZON * (FLOP_PER_STAGE_DYN + FLOP_PER_STAGE_LIN) + k_ipm * (HORIZON + 1) * FLOP_PER_STAGE_QP


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def config_dict(n_gpus):
    return {"workload": f"config3: {BATCH_PER_GPU} independent SQP-RTI NMPC instances per GPU, N={HORIZON}, dt={DT}, {OBJECT}, seed 2+rank",
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * n_gpus, "horizon": HORIZON, "qp_tol": QP_TOL,
            "mode": "sqp_rti", "parallelism": f"{n_gpus} independent shards, no collective",
            "l2": "256 MiB device memset between timed steps (outside each step's CUDA-event window); slabs 148 MB > 126 MB L2"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(passes: int, nthreads: int, seed: int = 2):
    """Oracle restatement (prepare + one RTI iteration) over `passes` passes of the config-3 batch on the host cores."""
    from oracle import oracle as orc
    from tests.workloads import oracle_model
    from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload
    om = oracle_model(OBJECT)
    wl = make_rti_workload(BATCH_PER_GPU, HORIZON, dt=DT, seed=seed)
    ocp = orc.Ocp(om, HORIZON, DT, qp_tol=QP_TOL)
    zeros_x = np.zeros((BATCH_PER_GPU, HORIZON + 1, 4))
    cold = np.zeros(BATCH_PER_GPU, dtype=np.int32)
    times = []
    for _ in range(passes):
        t0 = time.perf_counter()
        pr = ocp.prepare(wl["x0"], cold, zeros_x, wl["u_init"], nthreads=nthreads)
        ocp.solve("rti", pr["x0"], wl["yref"], wl["yref_e"], pr["x"], pr["u"], nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0                                             # rank 0 alone runs the CPU arm
    cores = os.cpu_count() or 1
    cpu_oracle_rate(1, cores)                                # warm-up (thread pool, page faults)
    for _ in range(max(0, args.warmup - 1)):
        cpu_oracle_rate(1, cores)
    times = cpu_oracle_rate(args.steps, cores)
    t = sum(times)
    value = BATCH_PER_GPU * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{args.steps} passes over the {BATCH_PER_GPU}-instance config-3 batch, restated oracle (not acados), {cores} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "note": "acados v0.2.1 / MATLAB are not installable offline; the CPU arm is the C++ oracle restatement of the same algorithm"}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock / throttle-reason sampler for the timed region.

    NVML in a thread (every 2 ms, initialised before the region starts, so that even a 40 ms region is sampled under load);
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
            time.sleep(0.002)

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
                    "reasons": sorted(n for b, n in names.items() if bits & b), "samples": len(sm), "source": "nvml, 2 ms period, inside the timed region"}
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
    from tests.workloads import gpu_model
    from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload

    rank, local_rank, world = env_rank()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N = BATCH_PER_GPU, HORIZON
    gm = gpu_model(OBJECT)
    wl = make_rti_workload(B, N, dt=DT, seed=2 + rank)
    solver = q.Solver([gm], N, DT, B, device=local_rank, qp_tol=QP_TOL, problems_per_warp=args.ppw, qp_kernel=args.qp_kernel)
    stream = torch.cuda.ExternalStream(solver.stream, device=dev)

    # device-resident inputs / outputs (value) and pinned host buffers (e2e)
    d_in = {k: torch.from_numpy(wl[k]).to(dev) for k in ("x0", "yref", "yref_e", "u_init")}
    d_cold = torch.zeros(B, dtype=torch.int32, device=dev)
    d_u0 = torch.empty(B, 2, dtype=torch.float64, device=dev)
    d_status = torch.empty(B, dtype=torch.int32, device=dev)
    h_in = {k: torch.from_numpy(wl[k]).pin_memory() for k in ("x0", "yref", "yref_e", "u_init")}
    h_cold = torch.zeros(B, dtype=torch.int32).pin_memory()
    h_u0 = torch.empty(B, 2, dtype=torch.float64).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        solver.set("x0", d_in["x0"]); solver.set("yref", d_in["yref"]); solver.set("yref_e", d_in["yref_e"])
        solver.set("u", d_in["u_init"]); solver.set_int("cold", d_cold)
        solver.prepare(); solver.solve()
        solver.get("u", stage=0, out=d_u0); solver.get_int("status", out=d_status)

    def step_host():
        solver.set("x0", h_in["x0"]); solver.set("yref", h_in["yref"]); solver.set("yref_e", h_in["yref_e"])
        solver.set("u", h_in["u_init"]); solver.set_int("cold", h_cold)
        solver.prepare(); solver.solve()
        solver.get("u", stage=0, out=h_u0); solver.get_int("status", out=h_status)     # host gets synchronise the stream

    # controller-level step = NMPC_controller.solve(x0, index_time) (NMPC_controller.m:329-423): the reference trajectory was handed
    # over once (set_reference_trajectory, :425-431) and the warm start is the controller's own state, so a control period moves
    # only x0 (in) and u0 / status (out) between host and device.  Same arithmetic as step_device / step_host: the window of
    # period 1 is bit-identical to the per-stage references of the workload (checked below), the initial guess is restored from
    # its device-resident copy.
    speed_t = (wl["yref"][0, :, 0] - wl["x0"][0, 0])
    traj = np.zeros((N, 6)); traj[:, 0] = 0.01 * (np.arange(N) * DT)
    off = np.zeros((B, 6)); off[:, 0] = wl["x0"][:, 0]; off[:, 1] = wl["x0"][:, 1]
    del speed_t
    solver.set_reference_trajectory(traj, off)

    def step_ctrl():
        solver.set("x0", h_in["x0"])
        solver.set_reference_window(1)
        solver.set("u", d_in["u_init"]); solver.set_int("cold", d_cold)
        solver.prepare(); solver.solve()
        solver.get("u", stage=0, out=h_u0); solver.get_int("status", out=h_status)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    solver.sync()
    step_host()
    u0_ref = h_u0.numpy().copy()
    solver.set_reference_window(1)
    if not (np.array_equal(solver.get("yref"), wl["yref"]) and np.array_equal(solver.get("yref_e"), wl["yref_e"])):
        raise SystemExit("bench: the device-side reference window differs from the workload's references")
    step_ctrl()
    if not np.array_equal(h_u0.numpy(), u0_ref):
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
    barrier()
    e2e_ms = []
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_host()
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    barrier()
    t_e2e_fields_local = sum(e2e_ms) / 1e3
    e2e_ms = []
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_ctrl()
        e2e_ms.append(1e3 * (time.perf_counter() - t0))
    barrier()
    t_e2e_local = sum(e2e_ms) / 1e3
    it = solver.get_int("qp_iter")
    st = solver.get_int("status")
    k_ipm = float(it.mean())
    if world > 1:
        tt = torch.tensor([t_local, t_e2e_local, t_e2e_fields_local], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_max, t_e2e_max, t_e2e_fields_max = float(tt[0]), float(tt[1]), float(tt[2])
        # final host gather of the small per-problem result (outside every timed region; no collective on the solve path)
        from uclv_qs_pushing_matlab_b200 import sharding
        u0_all = sharding.gather_to_rank0(h_u0.numpy(), B * world, world, rank)
        assert rank != 0 or u0_all.shape == (B * world, 2)
    else:
        t_max, t_e2e_max, t_e2e_fields_max = t_local, t_e2e_local, t_e2e_fields_local
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    total_iters = B * world * args.steps
    value = total_iters / t_max
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    qp_avg_ms = sum(qp_ms) / len(qp_ms)
    achieved_gbs = B * ALG_BYTES_PER_ITER / (qp_avg_ms * 1e-3) / 1e9
    traffic, pipe_pct = None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_qp_traffic.json")))["qp_kernel_%d" % args.qp_kernel]
        traffic = prof["dram_bytes_per_launch"]
        pipe_pct = prof.get("fp64_pipe_active_pct_of_elapsed")
    except Exception:
        pass
    # FP64 reference rate measured here with a cuBLAS DGEMM (MEASURED_PEAKS.json has no FP64 entry)
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev); b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(dev)
    fp64_peak_tf = 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    fp64_ach_tf = B * alg_flops_per_iter(k_ipm) / (qp_avg_ms * 1e-3 + sum(lin_ms) / len(lin_ms) * 1e-3) / 1e12
    # per-solve latency of a single NMPC instance (the MEX drop-in case): host buffers through the C-ABI, wall clock
    lat_b1 = {}
    for tag, n1, mode in ((("rti_N40", HORIZON, 0), ("rti_N10", 10, 0), ("sqp_N10", 10, 1)) if world == 1 else ()):   # N = 1 runs only
        w1 = make_rti_workload(1, n1, dt=DT, seed=7)
        s1 = q.Solver([gm], n1, DT, 1, device=local_rank, qp_tol=QP_TOL, mode=mode)
        u1 = np.zeros((1, 2)); c1 = np.zeros(1, dtype=np.int32)
        ts = []
        for i in range(args.latency_solves + 10):
            t0 = time.perf_counter()
            s1.set("x0", w1["x0"]); s1.set("yref", w1["yref"]); s1.set("yref_e", w1["yref_e"]); s1.set("u", w1["u_init"]); s1.set_int("cold", c1)
            s1.prepare(); s1.solve(); s1.get("u", stage=0, out=u1)
            if i >= 10:
                ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        lat_b1[tag] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[min(len(ts) - 1, int(0.99 * len(ts)))], "solves": len(ts),
                       "sqp_iter": int(s1.get_int("sqp_iter")[0])}
        # the same solve at the NMPC_controller.solve(x0, idx) boundary: reference window and initial guess stay on the device
        tr1 = np.zeros((n1, 6)); tr1[:, 0] = 0.01 * (np.arange(n1) * DT)
        of1 = np.zeros((1, 6)); of1[:, :2] = w1["x0"][:, :2]
        s1.set_reference_trajectory(tr1, of1)
        du1 = torch.from_numpy(w1["u_init"]).to(dev); dc1 = torch.zeros(1, dtype=torch.int32, device=dev)
        ts = []
        for i in range(args.latency_solves + 10):
            t0 = time.perf_counter()
            s1.set("x0", w1["x0"]); s1.set_reference_window(1); s1.set("u", du1); s1.set_int("cold", dc1)
            s1.prepare(); s1.solve(); s1.get("u", stage=0, out=u1)
            if i >= 10:
                ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        lat_b1[tag + "_ctrl"] = {"p50_ms": ts[len(ts) // 2], "p99_ms": ts[min(len(ts) - 1, int(0.99 * len(ts)))], "solves": len(ts),
                                 "sqp_iter": int(s1.get_int("sqp_iter")[0])}
        del s1
    # Monte-Carlo of whole pushes, device-resident (qspush_closed_loop): the same batch in closed loop for 20 control
    # periods, nothing crosses PCIe inside the loop (extra information, not the contract's e2e)
    loop_info = None
    if world == 1:
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
        loop_info = {"periods": 20, "instances": B, "ms_per_period": 1e3 * tl / 20, "controller_solves_per_s": B * 20 / tl,
                     "status_ok_frac": float((rl["status_log"] == 0).float().mean()),
                     "what": "qspush_closed_loop: reference window, prepare, solve, plant step, shift per period on the device"}
    cores = os.cpu_count() or 1
    cpu_base = None
    if world == 1:                                           # the CPU baseline is timed on rank 0 of the N = 1 run only
        cpu_t = cpu_oracle_rate(args.cpu_passes, cores)
        cpu_base = {"value": BATCH_PER_GPU * len(cpu_t) / sum(cpu_t), "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{len(cpu_t)} pass(es) over the {BATCH_PER_GPU}-instance config-3 batch, restated oracle (not acados), {cores} threads"}
    srt = sorted(step_ms)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_dict(world),
        "e2e": {"value": total_iters / t_e2e_max, "unit": UNIT,
                "h2d_bytes_per_step": int(h_in["x0"].numel() * 8),
                "d2h_bytes_per_step": int(h_u0.numel() * 8 + h_status.numel() * 4), "ms_per_step": 1e3 * t_e2e_max / args.steps,
                "what": "controller-level call = NMPC_controller.solve(x0, index_time) for the batch through the C-ABI: x0 from pinned host memory, "
                        "reference window on the device (trajectory set once, qspush_set_reference_trajectory / _window), initial guess restored "
                        "from its device copy, prepare + solve, u0 and status to pinned host memory; same problems and bit-identical u0 as `value`"},
        "e2e_all_fields_from_host": {"value": total_iters / t_e2e_fields_max, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(v.numel() * 8 for v in h_in.values()) + h_cold.numel() * 4),
                "d2h_bytes_per_step": int(h_u0.numel() * 8 + h_status.numel() * 4), "ms_per_step": 1e3 * t_e2e_fields_max / args.steps,
                "what": "acados_ocp-level calls: x0, every stage's cost_y_ref, cost_y_ref_e and init_u re-sent from pinned host memory each step"},
        "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": ("k_qp_warp<3,0,16> (Mehrotra IPM, two problems per warp, parallel-in-time Riccati scans, state in shared memory + TMEM, ordered work queue)" if args.qp_kernel else "k_qp (Riccati/Mehrotra IPM, one problem per thread)"), "achieved": achieved_gbs, "peak": hbm_peak,
                     "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src + " (of measured)",
                     "algorithmic_bytes_per_launch": B * ALG_BYTES_PER_ITER, "kernel_ms": qp_avg_ms,
                     "kernel_share_of_step": qp_avg_ms / (sum(step_ms) / len(step_ms)),
                     "fp64": {"achieved_tflops": fp64_ach_tf, "peak_tflops": fp64_peak_tf, "frac": fp64_ach_tf / fp64_peak_tf,
                              "peak_source": "cuBLAS DGEMM 4096^3 measured in this run", "flops_per_iteration": alg_flops_per_iter(k_ipm),
                              "ncu_pipe_fp64_active_pct": pipe_pct,
                              "what": "achieved = algorithmic flops of the serial Riccati IPM (SURVEY 8d); ncu_pipe_fp64_active_pct = "
                                      "sm__pipe_fp64_cycles_active of the executed parallel-in-time algorithm (profiles/r01_v9_qp_ncu_summary.md)"},
                     "note": "the path is FP64-pipe / dependent-chain bound, not HBM bound (SURVEY 8d): the HBM fraction is reported "
                             "because the schema asks for it, the fp64 object is the relevant roofline (see DESIGN.md)"},
        "cpu_baseline": cpu_base,
        "k_ipm_mean": k_ipm, "k_ipm_max": int(it.max()), "status_ok_frac": float((st == 0).mean()),
        "phase_ms": {"prepare": sum(prep_ms) / len(prep_ms), "linearise": sum(lin_ms) / len(lin_ms), "qp": qp_avg_ms},
        "latency_ms": {"p50": srt[len(srt) // 2], "p99": srt[min(len(srt) - 1, int(0.99 * len(srt)))], "max": srt[-1],
                       "what": "per-step device time of one batched solve (all instances of the batch finish together)"},
        "latency_b1": lat_b1,
        "closed_loop_device": loop_info,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ppw", type=int, default=0, help="QP kernel packing (problems per warp), 0 = auto")
    ap.add_argument("--qp-kernel", type=int, default=1, help="1 = warp kernel (parallel-in-time, one or two problems per warp; what auto picks), 0 = one problem per thread")
    ap.add_argument("--latency-solves", type=int, default=1000, help="solves per B=1 latency measurement")
    ap.add_argument("--cpu-passes", type=int, default=3, help="passes of the oracle over the batch for cpu_baseline")
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
