#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- IHMPC closed-loop scenario-steps/s on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 20 --warmup 3      # CPU arm (oracle port)

A "step" is ONE control step (QP build + solve + first-move plant update + warm start + weight
update) of EVERY scenario of the batch.  Workload (config.workload): the 2x2 distillation
controller of example_distillation.py on the plant-model-mismatch Monte-Carlo scenario set
(BASELINE.json configs[4], 1M scenarios over 8 GPUs = 131072 scenarios per GPU, weak scaling:
at N=8 the run IS configs[4] including the final NCCL statistics reduction).  Other workloads:
--workload distillation4096 (configs[2]) and --workload mimo10 (configs[3]).

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


# ----------------------------------------------------------------------------- workloads
def build_workload(name, batch, first, n_total_steps):
    from sihmpc_b200 import scenarios as sc
    from sihmpc_b200.ihmpccontroller import IHMPCController
    if name == "distillation_mismatch_mc":
        sys_ = sc.distillation_system()
        ctrl = sc.declare_distillation(IHMPCController, sys_)
        scen = sc.mismatch_scenarios(batch, first=first)
        declare = sc.declare_distillation
    elif name == "distillation4096":
        sys_ = sc.distillation_system()
        ctrl = sc.declare_distillation(IHMPCController, sys_)
        scen = sc.distillation_scenarios(batch, first=first)
        declare = sc.declare_distillation
    elif name == "mimo10":
        sys_ = sc.mimo10_system()
        ctrl = sc.declare_mimo10(IHMPCController, sys_)
        scen = sc.mimo10_scenarios(batch, sys_, first=first)
        declare = sc.declare_mimo10
    else:
        raise SystemExit("unknown workload " + name)
    return sys_, ctrl, scen, declare


DEFAULT_BATCH = {"distillation_mismatch_mc": 131072, "distillation4096": 4096, "mimo10": 65536}


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
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="distillation_mismatch_mc")
    ap.add_argument("--batch", type=int, default=0, help="scenarios per GPU (default: the named config's)")
    ap.add_argument("--steps-per-launch", type=int, default=0, help="control steps fused in one launch (0 = all)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from sihmpc_b200 import _capi
    from sihmpc_b200.engine import BatchedEngine, HostStepper, fp64_peak_flops
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    batch = args.batch or DEFAULT_BATCH[args.workload]
    sys_, ctrl, scen, _ = build_workload(args.workload, batch, first=rank * batch, n_total_steps=K + W)
    plan = ctrl.compile()
    eng = BatchedEngine(plan, batch, device=dev)
    eng.set_state(scen["x0"], scen["u0"])
    eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    if "dist_step" in scen:
        eng.set_disturbance(scen["dist_step"], scen["dist"])
    if "plant_val" in scen:
        eng.set_plant_values(scen["plant_val"])
    spl = args.steps_per_launch or K

    fp64_peak = fp64_peak_flops()                      # measured DFMA throughput of this GPU
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    # ---- warm-up: W control steps, untimed (same kernel, same launch shape as the timed region)
    eng.run(W, steps_per_launch=min(spl, W))
    eng.reduce_stats()
    torch.cuda.synchronize(dev)
    flush.fill_(1)                                     # L2 flush before the timed region
    sampler = None
    if rank == 0:
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        except Exception:
            uuid = None
        sampler = ClockSampler(local_rank, uuid)
    barrier()
    # ---- timed region: EXACTLY K control steps (+ the statistics reduction / NCCL exchange)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_wall0 = time.time()
    ev[0].record()
    eng.run(K, steps_per_launch=spl)
    ev[1].record()
    agg = eng.reduce_stats()
    if world > 1:
        dist.all_reduce(agg[:2], op=dist.ReduceOp.SUM)
        dist.all_reduce(agg[2], op=dist.ReduceOp.MAX)
    ev[2].record()
    barrier()
    t_wall1 = time.time()
    ms_total = ev[0].elapsed_time(ev[2])
    ms_kernel = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    n_launch = (K + spl - 1) // spl
    # kernels of OURS per sihmpc_step call: a fast path with bounds runs first pass + resume pass + general kernel
    # (the last two leave at once when nothing was listed); without bounds, or on the general kernel alone, one
    kname = eng.dp.kernel_name()
    has_bounds = bool(np.isfinite(plan.ulb).any() or np.isfinite(plan.uub).any() or np.isfinite(plan.dulb).any()
                      or np.isfinite(plan.duub).any())
    if kname in ("mma_step_kernel", "tile_step_kernel") and has_bounds:
        kernels_per_launch, kernels_desc = 3, "%s first pass, %s resume pass, ihmpc_step_kernel" % (kname, kname)
    else:
        kernels_per_launch, kernels_desc = 1, kname
    counters = eng.dp.counters()
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_kernel = float(t[0]), float(t[1])
    status_or = int(agg[2, plan.ny + 3].item())
    # scenarios whose loop raised a status bit (e.g. an unstable mismatched plant: 1 of the 1M of configs[4])
    flagged = (eng.stats[:, plan.ny + 3] != 0).sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(flagged, op=dist.ReduceOp.SUM)
    flagged = int(flagged.item())
    value = world * batch * K / (ms_total * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host state, H2D + D2H every step)
    e2e = None
    lat = None
    if True:
        hs = HostStepper(plan, batch, device=dev)
        hs.x.copy_(torch.as_tensor(scen["x0"])); hs.u.copy_(torch.as_tensor(scen["u0"]))
        ysp_np = scen["ysp_seq"][:, 1, :]
        hs.ysp.copy_(torch.as_tensor(np.ascontiguousarray(ysp_np)))
        fl = _capi.F_ADVANCE | _capi.F_UPDATE_VIN | _capi.F_WEIGHTS_SAT2
        for _ in range(3):
            hs.step(flags=fl, want_w_opt=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            hs.step(flags=fl, want_w_opt=False)        # synchronises the stream itself
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * batch * args.e2e_steps / float(tt[0]), "unit": UNIT,
               "h2d_bytes_per_step": hs.h2d_bytes(), "d2h_bytes_per_step": hs.d2h_bytes(advance=True, want_w_opt=False),
               "steps": args.e2e_steps, "api": "sihmpc_step_host (pinned host buffers; in: x,u,pesos,ViNant,ysp; out: du,J,status,ViNant + advanced x,u,pesos; "
                      "zero-copy: the step kernel reads and writes the pinned arrays over PCIe)"}
        if rank == 0:
            h1 = HostStepper(plan, 1, device=dev)
            h1.x.copy_(torch.as_tensor(scen["x0"][:1])); h1.u.copy_(torch.as_tensor(scen["u0"][:1]))
            h1.ysp.copy_(torch.as_tensor(np.ascontiguousarray(ysp_np[:1])))
            ts = []
            for i in range(300):
                a = time.perf_counter(); h1.step(flags=fl); ts.append(time.perf_counter() - a)
            ts = np.asarray(ts[50:]) * 1e6
            lat = {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
                   "what": "one closed-loop step at batch 1 through sihmpc_step_host, host to host"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    F_alg = algorithmic_flops(plan)
    B_alg = algorithmic_bytes(plan)
    per_launch_s = ms_kernel * 1e-3 / n_launch
    steps_per_launch_units = batch * (K / n_launch)
    achieved_tflops = F_alg * steps_per_launch_units / per_launch_s / 1e12
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
        ent = tj.get(eng.dp.kernel_name() + ":" + args.workload)
        if ent:
            traffic = float(ent["dram_bytes_per_scenario_launch"]) * batch
    except Exception:
        pass
    roofline = {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved_tflops / (fp64_peak / 1e12), "traffic": traffic,
                "peak_source": "measured in this run (sihmpc_fp64_peak: dependent-free DFMA loop, all SMs)",
                "flop_per_scenario_step": F_alg, "kernel": "sih::" + eng.dp.kernel_name(),
                "kernel_ms_per_launch": per_launch_s * 1e3, "launches": n_launch,
                "hbm": {"achieved": B_alg * batch / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "bytes_per_scenario_launch": B_alg,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(args.workload, args.cpu_per_core, 200, 0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": "%d scenarios x 200 closed-loop steps of the same workload, NumPy oracle, one process per core (%.1f s)"
                         % (r["cores"] * args.cpu_per_core, r["seconds"]),
               "p50_step_latency_us": r["p50_us"]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "scenarios_per_gpu": batch, "plant": "%dx%d" % (plan.ny, plan.nu),
                       "nx": plan.nx, "N": plan.N, "nv": plan.nv, "steps_per_launch": spl,
                       "l2": "flushed (512 MiB write) before the timed region; per-scenario state is read once "
                             "and written once per launch, everything else stays on chip",
                       "includes": "final on-device statistics reduction" + (" + NCCL all-reduce" if world > 1 else "")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch * kernels_per_launch + 1,
            "gpu_launches_detail": "%d sihmpc_step call(s) x %d kernels (%s) + 1 stats_reduce_kernel"
                                   % (n_launch, kernels_per_launch, kernels_desc),
            "handed_over_to_general_kernel": counters[0], "constrained_steps_resolved_in_fast_path": counters[1],
            "roofline": roofline,
            "cpu_baseline": cpu, "latency_batch1": lat, "status_or": status_or, "flagged_scenarios": flagged,
            "fp64_peak_tflops_measured": fp64_peak / 1e12}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
