#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- IHMPC closed-loop scenario-steps/s on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 20 --warmup 3      # CPU arm (oracle port)

A "step" is ONE control step (QP build + solve + first-move plant update + warm start + weight
update) of EVERY scenario of the batch.  Headline workload (config.workload): the 2x2 distillation
controller of example_distillation.py on the plant-model-mismatch Monte-Carlo scenario set
(BASELINE.json configs[4], 1M scenarios over 8 GPUs = 131072 scenarios per GPU, weak scaling: at
N=8 the run IS configs[4] including the NCCL statistics exchange).  The same line carries, under
"configs", the other named workloads measured in the same run the same way: configs[2]
(distillation4096), configs[3] (mimo10, 65536 scenarios per GPU) and configs[1] (the log-barrier
controller: batch 4096 throughput and batch-1 p50 latency through sihmpc_log_loop).

Timed region = EXACTLY K control steps of the window [k0, k0+K) of the 200-step closed loop (k0 chosen
so that the window holds steps where bounds bind; k0 = 0 when K >= 200) + the on-device statistics
reduction (+ the NCCL exchange when N > 1), on CUDA events, max over ranks.  It is repeated
`--repeats` times from the same snapshot of the scenario state (L2 flushed before each); `value` is
the MEDIAN repetition, min/max are reported next to it.

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ihmpc_closed_loop_steps_per_sec"
UNIT = "scenario-steps/s"
LOOP_STEPS = 200                      # length of the closed loops of SURVEY.md section 8(d)


# ----------------------------------------------------------------------------- workloads
def build_workload(name, batch, first):
    from sihmpc_b200 import scenarios as sc
    from sihmpc_b200.ihmpccontroller import IHMPCController
    if name == "distillation_mismatch_mc":
        sys_ = sc.distillation_system()
        ctrl = sc.declare_distillation(IHMPCController, sys_)
        scen = sc.mismatch_scenarios(batch, first=first)
    elif name == "distillation4096":
        sys_ = sc.distillation_system()
        ctrl = sc.declare_distillation(IHMPCController, sys_)
        scen = sc.distillation_scenarios(batch, first=first)
    elif name == "mimo10":
        sys_ = sc.mimo10_system()
        ctrl = sc.declare_mimo10(IHMPCController, sys_)
        scen = sc.mimo10_scenarios(batch, sys_, first=first)
    else:
        raise SystemExit("unknown workload " + name)
    return sys_, ctrl, scen


DEFAULT_BATCH = {"distillation_mismatch_mc": 131072, "distillation4096": 4096, "mimo10": 65536}
# first step of the timed window when K < 200: mismatch Monte-Carlo -- scenarios sit on u >= 0 from k ~ 120 on;
# distillation4096 -- the set-point change at k = 11; mimo10 -- the second set-point change at k = 101, after which
# ~3 % of the scenarios sit on the |u| <= 5 box (the large active-set transients)
WINDOW_K0 = {"distillation_mismatch_mc": 115, "distillation4096": 5, "mimo10": 100}


def algorithmic_flops(plan):
    """FP64 flop per scenario-step (one Cholesky, no extra iterations)."""
    N, nx, nu, ny, nv = plan.N, plan.nx, plan.nu, plan.ny, plan.nv
    Ny = N * ny
    m_y = plan.mh
    return (2 * Ny * nx) + (2 * nv * (Ny + ny) + 2 * nv * nx) + (m_y * nv * (nv + 1) + nv) + \
           (nv ** 3 / 3.0 + 4 * nv ** 2) + (2 * nx * (nx + nu)) + (2 * Ny * (nx + nv) + 3 * Ny + 2 * nv)


def algorithmic_bytes(plan):
    """State round trip of one launch per scenario: in x,u,ysp,pesos,ViNant / out x,u,pesos,du,J,stats."""
    p = plan
    n_in = p.nx + p.nu + p.ny + p.m + p.n_vin
    n_out = p.nx + p.nu + p.m + p.n_vin + p.nu + 1 + (p.ny + 5)
    return 8 * (n_in + n_out)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML from a thread (about
    1 ms per sample, so a 50 ms region still gets tens of samples); nvidia-smi as the fall-back."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None, period_s=0.001):
        self.samples = []          # (t, sm_mhz, sm_max_mhz, power_w, [hw, hw_thermal, sw_thermal, sw_power_cap])
        self.proc = None
        self.stop_flag = False
        self.source = None
        try:
            import pynvml as N
            N.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = N.nvmlDeviceGetHandleByUUID(cand if isinstance(cand, bytes) else cand.encode())
                        break
                    except Exception:
                        h = None
            if h is None:
                h = N.nvmlDeviceGetHandleByIndex(index)
            self.N, self.h = N, h
            self.max_sm = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            self.period = period_s
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.source = None
        if self.source is None:
            try:
                self.proc = subprocess.Popen(
                    ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                     "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.source = "nvidia-smi"
                self.t = threading.Thread(target=self._read, daemon=True)
                self.t.start()
            except Exception:
                self.proc = None

    def _poll(self):
        N, h = self.N, self.h
        bits = [getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)]
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                sm = float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
                r = int(get_reasons(h))
                self.samples.append((time.time(), sm, self.max_sm, pw, [bool(r & b) for b in bits]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                self.samples.append((time.time(), float(f[0]), float(f[1]), float(f[2]),
                                     [x.lower().startswith("active") for x in f[3:7]]))
            except ValueError:
                continue

    def stop(self, t0, t1):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        time.sleep(0.01)
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = list(self.samples)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        use = inside if len(inside) >= 2 else [r for r in rows if r[3] > 250.0] or rows
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in use for i in range(4) if r[4][i]})
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": use[0][2],
                "power_w_max": max(r[3] for r in use), "samples": len(use),
                "samples_in_timed_region": len(inside), "source": self.source, "reasons": reasons}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    """One host core: closed loops of `idxs` scenarios with the NumPy oracle, step by step."""
    workload, idxs, n_steps, warm = args
    import numpy as _np
    from oracle.ihmpc_oracle import OracleIHMPC
    from sihmpc_b200 import scenarios as sc
    from tests.util import schedule_fn
    sys_, _, scen, declare = None, None, None, None
    if workload == "mimo10":
        sys_ = sc.mimo10_system(); declare = sc.declare_mimo10
        scen = sc.mimo10_scenarios(max(idxs) + 1, sys_)
    elif workload == "distillation4096":
        sys_ = sc.distillation_system(); declare = sc.declare_distillation
        scen = sc.distillation_scenarios(max(idxs) + 1)
    else:
        sys_ = sc.distillation_system(); declare = sc.declare_distillation
        scen = sc.mismatch_scenarios(max(idxs) + 1)
    states = []
    for i in idxs:
        c = declare(OracleIHMPC, sys_)
        plant = sc.plant_matrices_from_values(sys_, scen["plant_val"][i]) if "plant_val" in scen else (sys_.A, sys_.B)
        states.append(dict(c=c, x=scen["x0"][i].copy(), u=scen["u0"][i].copy(), pesos=[], w0=[],
                           plant=plant, f=schedule_fn(scen["ysp_seq"][i], scen["seg_start"][i]),
                           dist=(int(scen["dist_step"][i]), scen["dist"][i]) if "dist_step" in scen else None))
    lat = []

    def one_step(k):
        for s in states:
            t0 = time.perf_counter()
            c = s["c"]
            ysp = _np.asarray(s["f"](k), float)
            sol = c.mpc(x0=s["x"], ySP=ysp, w0=s["w0"], u0=s["u"], pesos=s["pesos"])
            du = c.du.reshape(-1)
            s["x"] = s["plant"][0] @ s["x"] + s["plant"][1] @ du
            if s["dist"] is not None and k == s["dist"][0]:
                s["x"][:c.ny] += s["dist"][1]
            s["u"] = s["u"] + du
            s["w0"] = c.warmStart(sol, ysp)
            s["pesos"], _ = c.satWeights2(s["x"], s["u"], s["w0"], ysp)
            lat.append(time.perf_counter() - t0)
    k = 0
    for _ in range(warm):
        one_step(k); k += 1
    lat.clear()
    t0 = time.time()
    for _ in range(n_steps):
        one_step(k); k += 1
    return t0, time.time(), lat


def cpu_arm(workload, per_core, n_steps, warm, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [(workload, list(range(c * per_core, (c + 1) * per_core)), n_steps, warm) for c in range(cores)]
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    t0 = min(r[0] for r in res); t1 = max(r[1] for r in res)
    lat = np.concatenate([np.asarray(r[2]) for r in res])
    units = cores * per_core * n_steps
    return dict(value=units / (t1 - t0), seconds=t1 - t0, units=units, cores=cores,
                p50_us=float(np.percentile(lat, 50) * 1e6), p99_us=float(np.percentile(lat, 99) * 1e6))


# ----------------------------------------------------------------------------- GPU measurement
class Ctx:
    """Rank layout, device and the shared L2-flush buffer."""

    def __init__(self, rank, world, local_rank, dev, torch, dist):
        self.rank, self.world, self.local_rank, self.dev, self.torch, self.dist = rank, world, local_rank, dev, torch, dist
        self.flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(device_ids=[self.local_rank])
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def sum_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]


def measure_workload(cx, name, batch, K, W, spl, repeats, fp64_peak, sampler_factory=None):
    """K control steps of `name` on every rank's shard, timed `repeats` times from the same snapshot.
    Returns (record, engine pieces) -- record = value / ms / roofline / counters of this workload."""
    torch = cx.torch
    from sihmpc_b200.dist import reduce_moments
    from sihmpc_b200.engine import BatchedEngine
    sys_, ctrl, scen = build_workload(name, batch, first=cx.rank * batch)
    plan = ctrl.compile()
    eng = BatchedEngine(plan, batch, device=cx.dev)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    if "dist_step" in scen:
        eng.set_disturbance(scen["dist_step"], scen["dist"])
    if "plant_val" in scen:
        eng.set_plant_values(scen["plant_val"])
    spl = spl or K
    k0 = WINDOW_K0[name] if K + WINDOW_K0[name] <= LOOP_STEPS else 0
    # ---- untimed: pre-roll to the first step of the window, snapshot, then W warm-up steps through the same
    # kernels and the same collective as the timed region (first-use NCCL channel set-up is not billed to a step)
    if k0 > 0:
        eng.run(k0)
    snap = eng.snapshot()
    eng.run(W, steps_per_launch=min(spl, W))
    reduce_moments(eng.reduce_stats())
    eng.restore(snap)
    eng.run(K, steps_per_launch=spl)                    # one full-shape launch as well (W may be < K)
    reduce_moments(eng.reduce_stats())
    torch.cuda.synchronize(cx.dev)
    sampler = sampler_factory() if sampler_factory else None
    ms_total, ms_kernel, agg = [], [], None
    t_wall0 = time.time()
    for _ in range(repeats):
        eng.restore(snap)
        cx.flush.fill_(1)                               # L2 flush (512 MiB write) before every repetition
        cx.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        eng.run(K, steps_per_launch=spl)                # ---- timed: EXACTLY K control steps ...
        ev[1].record()
        agg = reduce_moments(eng.reduce_stats())        # ... + statistics reduction (+ ONE NCCL collective)
        ev[2].record()
        cx.barrier()
        ms_total.append(ev[0].elapsed_time(ev[2]))
        ms_kernel.append(ev[0].elapsed_time(ev[1]))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    counters = eng.dp.counters()
    tot = np.asarray(cx.max_over_ranks(ms_total))       # per repetition: the slowest rank
    ker = np.asarray(cx.max_over_ranks(ms_kernel))
    med = int(np.argsort(tot)[len(tot) // 2])
    flagged = int(cx.sum_over_ranks([float((eng.stats[:, plan.ny + 3] != 0).sum().item())])[0])
    handed, solved = [int(v) for v in cx.sum_over_ranks([float(counters[0]), float(counters[1])])]
    n_launch = (K + spl - 1) // spl
    kname = eng.dp.kernel_name()
    has_bounds = bool(np.isfinite(plan.ulb).any() or np.isfinite(plan.uub).any() or np.isfinite(plan.dulb).any()
                      or np.isfinite(plan.duub).any())
    stream_min = int(os.environ.get("SIHMPC_STREAM_MIN_BATCH", "16384"))     # mma_units.cu: workers only for long first passes
    if kname == "mma_step_kernel" and has_bounds and not os.environ.get("SIHMPC_NO_STREAM_RESUME") and batch >= stream_min:
        kpl, kdesc = 4, "mma_step_kernel resume workers (concurrent) + first pass + sweep pass, ihmpc_step_kernel"
    elif kname in ("mma_step_kernel", "tile_step_kernel") and has_bounds:
        kpl, kdesc = 3, "%s first pass, %s resume pass, ihmpc_step_kernel" % (kname, kname)
    else:
        kpl, kdesc = 1, kname
    F_alg, B_alg = algorithmic_flops(plan), algorithmic_bytes(plan)
    per_launch_s = float(np.median(ker)) * 1e-3 / n_launch
    achieved = F_alg * batch * (K / n_launch) / per_launch_s / 1e12
    rec = {
        "workload": name, "value": cx.world * batch * K / (tot[med] * 1e-3), "unit": UNIT,
        "value_min": cx.world * batch * K / (tot.max() * 1e-3), "value_max": cx.world * batch * K / (tot.min() * 1e-3),
        "repeats": repeats, "ms_per_step": float(tot[med]) / K,
        "ms_timed_region": {"median": float(tot[med]), "min": float(tot.min()), "max": float(tot.max())},
        "scenarios_per_gpu": batch, "window": [k0, k0 + K], "steps_per_launch": spl,
        "plant": "%dx%d" % (plan.ny, plan.nu), "nx": plan.nx, "N": plan.N, "nv": plan.nv,
        "kernel": "sih::" + kname, "kernel_ms_per_launch": per_launch_s * 1e3, "launches": n_launch,
        "gpu_launches": n_launch * kpl + 2,
        "gpu_launches_detail": "%d sihmpc_step call(s) x %d kernels (%s) + stats_init_kernel + stats_reduce_kernel"
                               % (n_launch, kpl, kdesc),
        "handed_over_to_general_kernel": handed, "constrained_steps_resolved_in_fast_path": solved,
        "flagged_scenarios": flagged, "status_or": int(agg[2, plan.ny + 3].item()),
        "flop_per_scenario_step": F_alg, "achieved_tflops": achieved, "frac_of_fp64_peak": achieved / (fp64_peak / 1e12),
        "hbm_gbs": B_alg * batch / per_launch_s / 1e9, "bytes_per_scenario_launch": B_alg,
    }
    return rec, dict(plan=plan, scen=scen, eng=eng, clocks=clocks, snap=snap)


def measure_log_variant(cx, K, W, repeats, fp64_peak=None):
    """configs[1]: the log-barrier controller (example_distillation_log.py) through sihmpc_log_loop --
    throughput at batch 4096 and p50 latency of one step at batch 1 (host call to synchronised result)."""
    torch = cx.torch
    from sihmpc_b200 import scenarios as sc
    from sihmpc_b200.ihmpccontroller_log import IHMPCController as LogController
    sys_ = sc.distillation_system()
    c = sc.declare_distillation_log(LogController, sys_)
    batch = 4096
    scen = sc.distillation_scenarios(batch)
    seg = scen["seg_start"].copy(); seg[:, 1] = 51; seg[:, 2] = 401          # example_distillation_log.py:90-94
    eng = c.engine(batch)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], seg)
    k0 = 45                                              # window over the set-point change at k = 51
    eng.run(k0)
    snap = eng.snapshot()
    eng.run(max(W, 1)); eng.restore(snap); eng.run(K)
    torch.cuda.synchronize(cx.dev)
    ms = []
    for _ in range(repeats):
        eng.restore(snap)
        cx.flush.fill_(1)
        torch.cuda.synchronize(cx.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(K); e1.record()
        torch.cuda.synchronize(cx.dev)
        ms.append(e0.elapsed_time(e1))
    ms = np.asarray(ms)
    iters = float(eng.iters.to(torch.float64).mean().item())
    st_or = int(eng.status.max().item())
    # batch 1, one step per call, host to host
    e1_ = c.engine(1)
    e1_.set_state(scen["x0"][:1], scen["u0"][:1])
    e1_.set_schedule(scen["ysp_seq"][:1], seg[:1])
    ts = []
    for i in range(260):
        a = time.perf_counter()
        e1_.run(1)
        torch.cuda.synchronize(cx.dev)
        ts.append(time.perf_counter() - a)
    ts = np.asarray(ts[60:]) * 1e6
    # algorithmic work of one damped-Newton iteration on (dU, ss): gradients of the m sub-objectives
    # and of Vt through E, lower triangle of the (nv+1)^2 Hessian (mh weighted blocks + rank-m term), Cholesky, two
    # triangular solves, one line-search evaluation of phi
    p = c.compile()
    n = p.nv + 1
    f_it = (2 * p.R * p.nv * (p.m + 1) + (n * (n + 1) // 2) * (2 * p.mh + 3 * p.m) + n ** 3 / 3 + 4 * n * n
            + 2 * p.R * p.nv + 2 * (p.R + p.nv) * p.m)
    value = batch * K / (float(np.median(ms)) * 1e-3)
    roof = None
    if fp64_peak:
        roof = {"bound": "fp64", "flop_per_newton_iteration": float(f_it), "newton_iterations_per_step": iters,
                "achieved": value * f_it * iters / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                "frac": value * f_it * iters / fp64_peak,
                "note": "one warp per scenario, instruction-issue bound (profiles/r02_log_kernel_ncu_summary.md); "
                        "iterations = mean of the last step of the window"}
    return {"roofline": roof, "workload": "distillation_log (example_distillation_log.py controller, ihmpc_log_kernel via sihmpc_log_loop)",
            "value": value, "unit": UNIT,
            "value_min": batch * K / (ms.max() * 1e-3), "value_max": batch * K / (ms.min() * 1e-3),
            "scenarios": batch, "window": [k0, k0 + K], "repeats": repeats, "ms_per_step": float(np.median(ms)) / K,
            "mean_newton_iterations_last_step": iters, "status_or": st_or, "kernel": "sih::ihmpc_log_kernel",
            "latency_batch1": {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
                               "what": "one closed-loop step (mpc + plant + warmStart) at batch 1, device-resident state, "
                                       "call to synchronised result"}}


def measure_e2e(cx, plan, scen, eng, snap, batch, steps):
    """The same metric end to end through the C-ABI calls a host application makes, host<->device copies inside the
    timed region, every step:
      * e2e        -- sihmpc_resident_step: the plant lives with the caller; measured state + set-point go up from
                      pinned host memory, move + cost + status come down, the controller's own state (u, weights,
                      ViN_ant, warm start) stays on the device.  The measurements fed in are the states of the SAME
                      closed loop (recorded beforehand, untimed, with the batched engine), so every solve is the one of
                      the timed device run;
      * full_state -- sihmpc_step_host: the whole scenario state both ways every step (the reference's numpy-in /
                      numpy-out convention)."""
    torch = cx.torch
    from sihmpc_b200 import _capi
    from sihmpc_b200.engine import HostStepper, ResidentStepper
    # ---- record the measured states of the window, one step per launch (untimed)
    eng.restore(snap)
    k0 = eng.k
    xs = torch.empty(steps, batch, plan.nx, dtype=torch.float64, pin_memory=True)
    ys = torch.empty(steps, batch, plan.ny, dtype=torch.float64, pin_memory=True)
    seq, start = scen["ysp_seq"], scen["seg_start"]
    u_k0 = eng.u.cpu().numpy().copy(); pes_k0 = eng.pesos.cpu().numpy().copy()
    for i in range(steps):
        xs[i].copy_(eng.x)
        seg = (start <= k0 + i).sum(axis=1) - 1
        ys[i].copy_(torch.as_tensor(np.ascontiguousarray(seq[np.arange(batch), seg])))
        eng.run(1)
    # the resident controller starts from u and the weights of step k0 (the loop has been running for k0 steps)
    rs_pes = torch.as_tensor(pes_k0)
    rs = ResidentStepper(plan, batch, u_k0)
    _set_resident_weights(rs, rs_pes)
    for i in range(min(3, steps)):                         # warm-up (untimed), then start over from step k0
        rs.step_from(xs[i], ys[i])
    del rs
    rs = ResidentStepper(plan, batch, u_k0)
    _set_resident_weights(rs, rs_pes)
    cx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        rs.step_from(xs[i], ys[i])                          # synchronises the stream itself
    cx.barrier()
    dt = cx.max_over_ranks([time.perf_counter() - t0])[0]
    e2e = {"value": cx.world * batch * steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": rs.h2d_bytes(), "d2h_bytes_per_step": rs.d2h_bytes(), "steps": steps,
           "api": "sihmpc_resident_step (controller-resident: measured x and ysp up from pinned host memory, du + J + status "
                  "down, every step; u, weights, ViN_ant and the warm start stay on the device; chunked cudaMemcpyAsync "
                  "pipeline over three streams).  Measurements = the recorded states of the same closed loop"}
    # ---- the full-state host path
    hs = HostStepper(plan, batch, device=cx.dev)
    hs.x.copy_(torch.as_tensor(scen["x0"])); hs.u.copy_(torch.as_tensor(scen["u0"]))
    ysp_np = scen["ysp_seq"][:, 1, :]
    hs.ysp.copy_(torch.as_tensor(np.ascontiguousarray(ysp_np)))
    fl = _capi.F_ADVANCE | _capi.F_UPDATE_VIN | _capi.F_WEIGHTS_SAT2
    for _ in range(3):
        hs.step(flags=fl, want_w_opt=False)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        hs.step(flags=fl, want_w_opt=False)
    cx.barrier()
    dt = cx.max_over_ranks([time.perf_counter() - t0])[0]
    e2e["full_state"] = {"value": cx.world * batch * steps / dt, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes(),
                         "d2h_bytes_per_step": hs.d2h_bytes(advance=True, want_w_opt=False),
                         "api": "sihmpc_step_host (x,u,pesos,ViNant,ysp up; du,J,status + advanced x,u,pesos down; zero-copy)"}
    lat = None
    if cx.rank == 0:
        h1 = HostStepper(plan, 1, device=cx.dev)
        h1.x.copy_(torch.as_tensor(scen["x0"][:1])); h1.u.copy_(torch.as_tensor(scen["u0"][:1]))
        h1.ysp.copy_(torch.as_tensor(np.ascontiguousarray(ysp_np[:1])))
        ts = []
        for i in range(300):
            a = time.perf_counter(); h1.step(flags=fl); ts.append(time.perf_counter() - a)
        ts = np.asarray(ts[50:]) * 1e6
        r1 = ResidentStepper(plan, 1, scen["u0"][:1])
        x1 = torch.as_tensor(scen["x0"][:1]).pin_memory(); y1 = torch.as_tensor(np.ascontiguousarray(ysp_np[:1])).pin_memory()
        tr_ = []
        for i in range(300):
            a = time.perf_counter(); r1.step_from(x1, y1); tr_.append(time.perf_counter() - a)
        tr_ = np.asarray(tr_[50:]) * 1e6
        lat = {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
               "what": "one closed-loop step at batch 1 through sihmpc_step_host, host to host",
               "resident_p50_us": float(np.percentile(tr_, 50)), "resident_p99_us": float(np.percentile(tr_, 99))}
    return e2e, lat


def _set_resident_weights(rs, pes):
    """Weights of a loop that is already running: written into the resident state (the C ABI initialises every scenario
    with the same initial weights; a bench window that starts at step k0 carries per-scenario ones)."""
    import ctypes as C
    from sihmpc_b200 import _capi
    lib = rs.dp.lib
    off = C.c_size_t(0)
    _capi.check(lib.sihmpc_resident_weights_offset(rs.dp.handle, rs.batch, C.byref(off)), "sihmpc_resident_weights_offset")
    view = rs.state[off.value:off.value + pes.numel() * 8].view(rs.x.dtype).reshape(pes.shape)
    view.copy_(pes.to(rs.state.device))


# ----------------------------------------------------------------------------- main
def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints to fd 1 (NCCL's
    version banner, ...) has been redirected to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="distillation_mismatch_mc")
    ap.add_argument("--batch", type=int, default=0, help="scenarios per GPU (default: the named config's)")
    ap.add_argument("--steps-per-launch", type=int, default=0, help="control steps fused in one launch (0 = all)")
    ap.add_argument("--repeats", type=int, default=7, help="repetitions of the timed region (median reported)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other named workloads")
    ap.add_argument("--cpu-per-core", type=int, default=48,
                    help="scenarios per host core in the CPU arm (48 x 200 steps ~ 10 s per core)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(3, args.warmup)
    K = args.steps

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference's own CPU path cannot run (casadi/opom absent: SURVEY.md F3); the arm is the
        # oracle port on every host core, one control step over a bounded sample per bench step
        r = cpu_arm(args.workload, args.cpu_per_core, K, W)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": 1e3 * r["seconds"] / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload, "sample_scenarios": r["cores"] * args.cpu_per_core},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": "%d scenarios x %d steps, NumPy oracle, one process per core"
                                           % (r["cores"] * args.cpu_per_core, K),
                                 "p50_step_latency_us": r["p50_us"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return 0

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    from sihmpc_b200.dist import gather_stats
    from sihmpc_b200.engine import fp64_peak_flops
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cx = Ctx(rank, world, local_rank, dev, torch, dist)

    fp64_peak = fp64_peak_flops()                      # measured DFMA throughput of this GPU
    batch = args.batch or DEFAULT_BATCH[args.workload]

    def sampler_factory():
        if rank != 0:
            return None
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        return ClockSampler(local_rank, uuid)

    main_rec, m = measure_workload(cx, args.workload, batch, K, W, args.steps_per_launch, args.repeats, fp64_peak,
                                   sampler_factory)
    plan, scen, eng, clocks = m["plan"], m["scen"], m["eng"], m["clocks"]

    # ---- the statistics GATHER of configs[4] (every scenario's block to every rank), once, timed on its own
    gather_ms = None
    if world > 1:
        full = gather_stats(eng.stats, world * batch)          # warm
        torch.cuda.synchronize()                               # every rank idle before the timed one (else: the peer's backlog is billed to it)
        cx.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); full = gather_stats(eng.stats, world * batch); g1.record()
        cx.barrier()
        gather_ms = {"ms": cx.max_over_ranks([g0.elapsed_time(g1)])[0], "bytes_total": int(full.numel() * 8),
                     "what": "NCCL all-gather of the per-scenario statistics block [scenarios][ny+5] FP64 to every rank "
                             "(sihmpc_b200.dist.gather_stats), max over ranks"}
        del full

    e2e, lat = measure_e2e(cx, plan, scen, eng, m["snap"], batch, min(args.e2e_steps, K))
    del eng, m
    torch.cuda.empty_cache()

    # ---- the other named workloads, same method, same run
    configs = {}
    if not args.no_configs:
        for name in ("distillation4096", "mimo10"):
            if name == args.workload:
                continue
            rec, mm = measure_workload(cx, name, DEFAULT_BATCH[name], K, W, 0, max(3, args.repeats // 2), fp64_peak)
            configs[name] = rec
            del mm
            torch.cuda.empty_cache()
        if rank == 0:
            configs["distillation_log"] = measure_log_variant(cx, K, W, max(3, args.repeats // 2), fp64_peak)
        cx.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM traffic of the same kernel from the committed ncu --set full capture (profiles/ncu_traffic.json:
    # bytes per scenario per launch, measured at the listed batch), scaled to this launch; None if not captured
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tj.get(main_rec["kernel"].replace("sih::", "") + ":" + args.workload)
        if ent:
            traffic = float(ent["dram_bytes_per_scenario_launch"]) * batch
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": main_rec["achieved_tflops"], "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                "frac": main_rec["frac_of_fp64_peak"], "traffic": traffic,
                "peak_source": "measured in this run (sihmpc_fp64_peak: dependent-free DFMA loop, all SMs)",
                "flop_per_scenario_step": main_rec["flop_per_scenario_step"], "kernel": main_rec["kernel"],
                "kernel_ms_per_launch": main_rec["kernel_ms_per_launch"], "launches": main_rec["launches"],
                "hbm": {"achieved": main_rec["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s",
                        "bytes_per_scenario_launch": main_rec["bytes_per_scenario_launch"],
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(args.workload, args.cpu_per_core, 200, 0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": "%d scenarios x 200 closed-loop steps of the same workload, NumPy oracle, one process per core (%.1f s)"
                         % (r["cores"] * args.cpu_per_core, r["seconds"]),
               "p50_step_latency_us": r["p50_us"]}

    line = {"metric": METRIC, "value": main_rec["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "scenarios_per_gpu": batch, "plant": main_rec["plant"],
                       "nx": plan.nx, "N": plan.N, "nv": plan.nv, "steps_per_launch": main_rec["steps_per_launch"],
                       "window": "control steps [%d, %d) of the %d-step closed loop (pre-rolled untimed)"
                                 % (main_rec["window"][0], main_rec["window"][1], LOOP_STEPS),
                       "repeats": "%d repetitions from the same state snapshot; value = median" % args.repeats,
                       "l2": "flushed (512 MiB write) before every repetition; per-scenario state is read once "
                             "and written once per launch, everything else stays on chip",
                       "includes": "final on-device statistics reduction" + (" + one NCCL all-gather of the moments" if world > 1 else "")},
            "value_min": main_rec["value_min"], "value_max": main_rec["value_max"],
            "ms_timed_region": main_rec["ms_timed_region"],
            "clocks": clocks, "e2e": e2e, "gpu_launches": main_rec["gpu_launches"],
            "gpu_launches_detail": main_rec["gpu_launches_detail"],
            "handed_over_to_general_kernel": main_rec["handed_over_to_general_kernel"],
            "constrained_steps_resolved_in_fast_path": main_rec["constrained_steps_resolved_in_fast_path"],
            "roofline": roofline, "cpu_baseline": cpu, "latency_batch1": lat, "status_or": main_rec["status_or"],
            "flagged_scenarios": main_rec["flagged_scenarios"], "fp64_peak_tflops_measured": fp64_peak / 1e12,
            "gather": gather_ms, "configs": configs}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
