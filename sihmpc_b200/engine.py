# -*- coding: utf-8 -*-
"""Batched closed-loop engine: thousands of independent controller/plant scenarios stepping in
lockstep on one GPU.  PyTorch is used for device memory, streams and (in dist.py)
torch.distributed only; all arithmetic runs in libsihmpc_b200.so through the C ABI.

The per-scenario step is the body of the reference's driver loop
(/root/reference/example_distillation.py:81-113):
    mpc (ihmpccontroller.py:567-604) -> dynF (:93-113) -> warmStart (:534-564) -> satWeights2 (:619-636)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _capi
from .plan import Plan


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class DevicePlan:
    """A compiled controller resident on one CUDA device (sihmpc_plan_create)."""

    def __init__(self, plan: Plan, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _capi.SihmpcError("sihmpc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _capi.load()
        self.plan = plan
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.handle = C.c_void_p()
        desc, keep = _capi.make_desc(plan)
        with torch.cuda.device(self.device):
            _capi.check(self.lib.sihmpc_plan_create(C.byref(desc), C.byref(self.handle)), "sihmpc_plan_create")
        del keep
        self._ws = None
        self._ws_bytes = 0

    def workspace(self, batch: int, host: bool = False):
        need = (self.lib.sihmpc_host_staging_bytes if host else self.lib.sihmpc_workspace_bytes)(self.handle, batch)
        if self._ws is None or self._ws_bytes < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_bytes = need
        return self._ws, self._ws_bytes

    def kernel_name(self) -> str:
        """Kernel that serves sihmpc_step for this plan (sihmpc_plan_kernel)."""
        return self.lib.sihmpc_plan_kernel(self.handle).decode()

    def counters(self, stream=None):
        """(scenarios handed over to the general kernel, scenario-steps with active bounds solved inside the
        fast path) of the last sihmpc_step launch (sihmpc_plan_counters; synchronises the stream)."""
        h, k = C.c_int64(0), C.c_int64(0)
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        rc = self.lib.sihmpc_plan_counters(self.handle, C.byref(h), C.byref(k), C.c_void_p(st))
        if rc:
            raise _capi.SihmpcError(self.lib.sihmpc_last_error().decode())
        return int(h.value), int(k.value)

    def close(self):
        if self.handle:
            self.lib.sihmpc_plan_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class BatchedEngine:
    """State of `batch` scenarios on the device + the step calls."""

    def __init__(self, plan: Plan, batch: int, device=None, weights: str = "sat2", alfa: float = 0.0):
        self.dp = DevicePlan(plan, device)
        self.plan = plan
        self.batch = int(batch)
        self.device = self.dp.device
        dev, f64 = self.device, torch.float64
        B, p = self.batch, plan
        self.nw = p.nv + 2 * p.ny
        self.nstat = p.ny + _capi.NSTATS_EXTRA
        self.x = torch.zeros(B, p.nx, dtype=f64, device=dev)
        self.u = torch.zeros(B, p.nu, dtype=f64, device=dev)
        self.pesos = torch.tensor(p.peso0, dtype=f64, device=dev).repeat(B, 1).contiguous()
        self.vinant = torch.full((B, max(p.n_vin, 1)), float("inf"), dtype=f64, device=dev)
        self.ysp = torch.zeros(B, p.ny, dtype=f64, device=dev)
        self.du = torch.zeros(B, p.nu, dtype=f64, device=dev)
        self.J = torch.zeros(B, dtype=f64, device=dev)
        self.w_opt = torch.zeros(B, self.nw, dtype=f64, device=dev)
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.n_active = torch.zeros(B, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(B, self.nstat, dtype=f64, device=dev)
        self.ysp_seq = None
        self.seg_start = None
        self.dist_step = None
        self.dist = None
        self.plant_val = None
        self.k = 0
        self.alfa = float(alfa)
        self.wflag = {"sat2": _capi.F_WEIGHTS_SAT2, "sat1": _capi.F_WEIGHTS_SAT1, "fixed": 0}[weights]

    # ---- scenario set-up --------------------------------------------------------------------
    def set_state(self, x, u, pesos=None):
        self.x.copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)).reshape(self.batch, -1))
        self.u.copy_(torch.as_tensor(np.asarray(u, dtype=np.float64)).reshape(self.batch, -1))
        if pesos is not None:
            self.pesos.copy_(torch.as_tensor(np.asarray(pesos, dtype=np.float64)).reshape(self.batch, -1))
        self.k = 0
        self.stats.zero_()
        self.vinant.fill_(float("inf"))

    def set_schedule(self, ysp_seq, seg_start):
        """ysp_seq [B][n_seg][ny]; seg_start [B][n_seg] (segment s applies for k >= seg_start)."""
        self.ysp_seq = torch.as_tensor(np.ascontiguousarray(ysp_seq, dtype=np.float64)).to(self.device)
        self.seg_start = torch.as_tensor(np.ascontiguousarray(seg_start, dtype=np.int32)).to(self.device)
        assert self.ysp_seq.shape[0] == self.batch and self.seg_start.shape == self.ysp_seq.shape[:2]

    def set_disturbance(self, dist_step, dist):
        self.dist_step = torch.as_tensor(np.ascontiguousarray(dist_step, dtype=np.int32)).to(self.device)
        self.dist = torch.as_tensor(np.ascontiguousarray(dist, dtype=np.float64)).to(self.device)

    def set_plant_values(self, plant_val):
        """Per-scenario values of the plant CSR ([B][nnz]); None = nominal model."""
        if plant_val is None:
            self.plant_val = None
        else:
            self.plant_val = torch.as_tensor(np.ascontiguousarray(plant_val, dtype=np.float64)).to(self.device)
            assert self.plant_val.shape == (self.batch, self.plan.pl_val.size)

    # ---- calls ------------------------------------------------------------------------------
    def _io(self, flags, traj=None, use_schedule=True):
        io = _capi.IO()
        io.x, io.u, io.pesos = self.x.data_ptr(), self.u.data_ptr(), self.pesos.data_ptr()
        io.vinant = self.vinant.data_ptr() if self.plan.n_vin > 0 else None
        if use_schedule and self.ysp_seq is not None:
            io.ysp_seq, io.seg_start = self.ysp_seq.data_ptr(), self.seg_start.data_ptr()
            io.n_seg = int(self.ysp_seq.shape[1])
        else:
            io.ysp = self.ysp.data_ptr()
        if self.dist_step is not None:
            io.dist_step, io.dist = self.dist_step.data_ptr(), self.dist.data_ptr()
        if self.plant_val is not None:
            io.plant_val = self.plant_val.data_ptr()
        io.du, io.J, io.w_opt = self.du.data_ptr(), self.J.data_ptr(), self.w_opt.data_ptr()
        io.status, io.n_active = self.status.data_ptr(), self.n_active.data_ptr()
        io.stats = self.stats.data_ptr()
        if traj is not None:
            io.traj_du = traj["du"].data_ptr() if "du" in traj else None
            io.traj_J = traj["J"].data_ptr() if "J" in traj else None
            io.traj_y = traj["y"].data_ptr() if "y" in traj else None
        io.flags = flags
        io.k0 = self.k
        io.alfa = self.alfa
        return io

    def solve(self, ysp=None):
        """mpc() only: no plant update (ihmpccontroller.py:567-604)."""
        if ysp is not None:
            self.ysp.copy_(torch.as_tensor(np.asarray(ysp, dtype=np.float64)).reshape(self.batch, -1))
        io = self._io(_capi.F_UPDATE_VIN, use_schedule=(ysp is None))
        ws, nb = self.dp.workspace(self.batch)
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _capi.check(self.dp.lib.sihmpc_step(self.dp.handle, C.byref(io), self.batch, 1,
                                                C.c_void_p(ws.data_ptr()), nb, C.c_void_p(st)), "sihmpc_step")

    def run(self, n_steps: int, steps_per_launch: Optional[int] = None, record: bool = False):
        """Closed loop for n_steps (fused step + advance).  Returns traces when record=True."""
        spl = n_steps if steps_per_launch is None else int(steps_per_launch)
        traj_all = None
        if record:
            traj_all = {
                "du": torch.zeros(n_steps, self.batch, self.plan.nu, dtype=torch.float64, device=self.device),
                "J": torch.zeros(n_steps, self.batch, dtype=torch.float64, device=self.device),
                "y": torch.zeros(n_steps, self.batch, self.plan.ny, dtype=torch.float64, device=self.device),
            }
        ws, nb = self.dp.workspace(self.batch)
        st = torch.cuda.current_stream(self.device).cuda_stream
        flags = _capi.F_ADVANCE | _capi.F_UPDATE_VIN | self.wflag
        done = 0
        with torch.cuda.device(self.device):
            while done < n_steps:
                n = min(spl, n_steps - done)
                traj = None
                if record:
                    traj = {k: v[done:done + n] for k, v in traj_all.items()}
                io = self._io(flags, traj)
                _capi.check(self.dp.lib.sihmpc_step(self.dp.handle, C.byref(io), self.batch, n,
                                                    C.c_void_p(ws.data_ptr()), nb, C.c_void_p(st)), "sihmpc_step")
                self.k += n
                done += n
        return traj_all

    # ---- repeatable windows (bench.py: the same K steps timed several times) ------------------
    def snapshot(self):
        """Copy of everything a step advances (x, u, weights, ViN_ant, statistics, step index)."""
        return dict(x=self.x.clone(), u=self.u.clone(), pesos=self.pesos.clone(), vinant=self.vinant.clone(),
                    stats=self.stats.clone(), k=self.k)

    def restore(self, snap):
        self.x.copy_(snap["x"]); self.u.copy_(snap["u"]); self.pesos.copy_(snap["pesos"])
        self.vinant.copy_(snap["vinant"]); self.stats.copy_(snap["stats"])
        self.k = int(snap["k"])

    def reduce_stats(self):
        """[3][nstat] = (sum, sum of squares, max) over the scenarios of this device."""
        out = torch.empty(3, self.nstat, dtype=torch.float64, device=self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _capi.check(self.dp.lib.sihmpc_stats_reduce(C.c_void_p(self.stats.data_ptr()), self.batch, self.nstat,
                                                        C.c_void_p(out.data_ptr()), C.c_void_p(st)),
                        "sihmpc_stats_reduce")
        return out


class LogBatchedEngine:
    """Batched closed loop of the LOG-BARRIER controller (sihmpc_log_loop): `batch` scenarios of the
    driver loop of /root/reference/example_distillation_log.py:85-135 (mpc, first move to the plant,
    warmStart with ss = 1) stepping in lockstep, `n_steps` per launch."""

    def __init__(self, plan: Plan, batch: int, device=None):
        self.dp = DevicePlan(plan, device)
        self.plan = plan
        self.batch = int(batch)
        self.device = self.dp.device
        dev, f64 = self.device, torch.float64
        B, p = self.batch, plan
        self.nw = p.nv + 2 * p.ny + 1
        self.nstat = p.ny + _capi.NSTATS_EXTRA
        self.x = torch.zeros(B, p.nx, dtype=f64, device=dev)
        self.u = torch.zeros(B, p.nu, dtype=f64, device=dev)
        self.ysp = torch.zeros(B, p.ny, dtype=f64, device=dev)
        self.du = torch.zeros(B, p.nu, dtype=f64, device=dev)
        self.J = torch.zeros(B, dtype=f64, device=dev)
        self.w_opt = torch.zeros(B, self.nw, dtype=f64, device=dev)
        self.w_warm = torch.zeros(B, self.nw, dtype=f64, device=dev)
        self.w_warm[:, -1] = 1.0
        self.ss_last = torch.zeros(B, dtype=f64, device=dev)      # ss of the last optimum: start of the next Newton solve (io.pesos)
        self.vinant = torch.full((B, max(p.n_vin, 1)), float("inf"), dtype=f64, device=dev)   # contracting ViN_ant
        self.status = torch.zeros(B, dtype=torch.int32, device=dev)
        self.iters = torch.zeros(B, dtype=torch.int32, device=dev)
        self.stats = torch.zeros(B, self.nstat, dtype=f64, device=dev)
        self.ysp_seq = self.seg_start = self.dist_step = self.dist = self.plant_val = None
        self.k = 0

    def set_state(self, x, u):
        self.x.copy_(torch.as_tensor(np.asarray(x, dtype=np.float64)).reshape(self.batch, -1))
        self.u.copy_(torch.as_tensor(np.asarray(u, dtype=np.float64)).reshape(self.batch, -1))
        self.w_warm.zero_(); self.w_warm[:, -1] = 1.0
        self.ss_last.zero_()
        self.vinant.fill_(float("inf"))
        self.k = 0
        self.stats.zero_()

    set_schedule = BatchedEngine.set_schedule
    set_disturbance = BatchedEngine.set_disturbance
    set_plant_values = BatchedEngine.set_plant_values

    def snapshot(self):
        return dict(x=self.x.clone(), u=self.u.clone(), w_warm=self.w_warm.clone(), stats=self.stats.clone(),
                    vinant=self.vinant.clone(), ss_last=self.ss_last.clone(), k=self.k)

    def restore(self, snap):
        self.x.copy_(snap["x"]); self.u.copy_(snap["u"]); self.w_warm.copy_(snap["w_warm"])
        self.stats.copy_(snap["stats"]); self.vinant.copy_(snap["vinant"]); self.ss_last.copy_(snap["ss_last"])
        self.k = int(snap["k"])

    def run(self, n_steps: int, steps_per_launch: Optional[int] = None, record: bool = False):
        spl = n_steps if steps_per_launch is None else int(steps_per_launch)
        traj_all = None
        if record:
            traj_all = {
                "du": torch.zeros(n_steps, self.batch, self.plan.nu, dtype=torch.float64, device=self.device),
                "J": torch.zeros(n_steps, self.batch, dtype=torch.float64, device=self.device),
                "y": torch.zeros(n_steps, self.batch, self.plan.ny, dtype=torch.float64, device=self.device),
            }
        st = torch.cuda.current_stream(self.device).cuda_stream
        done = 0
        with torch.cuda.device(self.device):
            while done < n_steps:
                n = min(spl, n_steps - done)
                io = _capi.IO()
                io.x, io.u = self.x.data_ptr(), self.u.data_ptr()
                if self.ysp_seq is not None:
                    io.ysp_seq, io.seg_start = self.ysp_seq.data_ptr(), self.seg_start.data_ptr()
                    io.n_seg = int(self.ysp_seq.shape[1])
                else:
                    io.ysp = self.ysp.data_ptr()
                if self.dist_step is not None:
                    io.dist_step, io.dist = self.dist_step.data_ptr(), self.dist.data_ptr()
                if self.plant_val is not None:
                    io.plant_val = self.plant_val.data_ptr()
                io.du, io.J, io.w_opt = self.du.data_ptr(), self.J.data_ptr(), self.w_opt.data_ptr()
                io.status, io.stats = self.status.data_ptr(), self.stats.data_ptr()
                io.pesos = self.ss_last.data_ptr()
                if record:
                    io.traj_du = traj_all["du"][done:done + n].data_ptr()
                    io.traj_J = traj_all["J"][done:done + n].data_ptr()
                    io.traj_y = traj_all["y"][done:done + n].data_ptr()
                io.flags = _capi.F_ADVANCE
                if self.plan.n_vin > 0:
                    io.vinant = self.vinant.data_ptr()
                    io.flags |= _capi.F_UPDATE_VIN
                io.k0 = self.k
                _capi.check(self.dp.lib.sihmpc_log_loop(self.dp.handle, C.byref(io), C.c_void_p(self.w_warm.data_ptr()),
                                                        self.batch, n, C.c_void_p(self.iters.data_ptr()), C.c_void_p(st)),
                            "sihmpc_log_loop")
                self.k += n
                done += n
        return traj_all


class HostStepper:
    """The host-buffer entry point (sihmpc_step_host): numpy/pinned in, numpy/pinned out."""

    def __init__(self, plan: Plan, batch: int, device=None):
        self.dp = DevicePlan(plan, device)
        self.plan = plan
        self.batch = int(batch)
        B, p = self.batch, plan
        pin = dict(dtype=torch.float64, pin_memory=True)
        self.x = torch.zeros(B, p.nx, **pin)
        self.u = torch.zeros(B, p.nu, **pin)
        self.pesos = torch.tensor(p.peso0, dtype=torch.float64).repeat(B, 1).contiguous().pin_memory()
        self.vinant = torch.full((B, max(p.n_vin, 1)), float("inf"), dtype=torch.float64).pin_memory()
        self.ysp = torch.zeros(B, p.ny, **pin)
        self.du = torch.zeros(B, p.nu, **pin)
        self.J = torch.zeros(B, **pin)
        self.w_opt = torch.zeros(B, p.nv + 2 * p.ny, **pin)
        self.status = torch.zeros(B, dtype=torch.int32, pin_memory=True)
        self.n_active = torch.zeros(B, dtype=torch.int32, pin_memory=True)
        self.lam = torch.zeros(B, 2 * (p.nv + getattr(p, "n_extra", 0)) + p.ncd, **pin)   # multipliers per inequality row
        self.ws, self.ws_bytes = self.dp.workspace(B, host=True)

    def h2d_bytes(self):
        p = self.plan
        return self.batch * 8 * (p.nx + p.nu + p.m + p.ny + (p.n_vin if p.n_vin else 0))

    def d2h_bytes(self, advance=False, want_w_opt=True):
        p = self.plan
        b = self.batch * (8 * (p.nu + 1 + ((p.nv + 2 * p.ny) if want_w_opt else 0)) + 8) + self.batch * 8 * p.n_vin
        if advance:
            b += self.batch * 8 * (p.nx + p.nu + p.m)
        return b

    def step(self, flags=_capi.F_UPDATE_VIN, alfa=0.0, want_w_opt=True, want_lam=False):
        io = _capi.IO()
        io.x, io.u, io.pesos = self.x.data_ptr(), self.u.data_ptr(), self.pesos.data_ptr()
        io.vinant = self.vinant.data_ptr() if self.plan.n_vin > 0 else None
        io.ysp = self.ysp.data_ptr()
        io.du, io.J = self.du.data_ptr(), self.J.data_ptr()
        io.w_opt = self.w_opt.data_ptr() if want_w_opt else None
        io.status, io.n_active = self.status.data_ptr(), self.n_active.data_ptr()
        io.lam = self.lam.data_ptr() if want_lam else None
        io.flags = flags
        io.alfa = alfa
        st = torch.cuda.current_stream(self.dp.device).cuda_stream
        with torch.cuda.device(self.dp.device):
            _capi.check(self.dp.lib.sihmpc_step_host(self.dp.handle, C.byref(io), self.batch,
                                                     C.c_void_p(self.ws.data_ptr()), self.ws_bytes,
                                                     C.c_void_p(st)), "sihmpc_step_host")


class ResidentStepper:
    """Controller-resident stepping (sihmpc_resident_step): the plant lives with the caller.  Every step the measured
    state and the set-point go up from pinned host memory, the move and the cost come down; u, the weights, ViN_ant
    and the warm start stay on the device.  Same arithmetic as the reference's driver loop
    (/root/reference/example_distillation.py:81-113) with `dynF` replaced by the caller's plant."""

    def __init__(self, plan: Plan, batch: int, u0, pesos0=None, device=None, weights: str = "sat2", alfa: float = 0.0):
        self.dp = DevicePlan(plan, device)
        self.plan = plan
        self.batch = int(batch)
        B, p = self.batch, plan
        pin = dict(dtype=torch.float64, pin_memory=True)
        self.x = torch.zeros(B, p.nx, **pin)
        self.ysp = torch.zeros(B, p.ny, **pin)
        self.du = torch.zeros(B, p.nu, **pin)
        self.J = torch.zeros(B, **pin)
        self.status = torch.zeros(B, dtype=torch.int32, pin_memory=True)
        self.alfa = float(alfa)
        self.flags = _capi.F_UPDATE_VIN | {"sat2": _capi.F_WEIGHTS_SAT2, "sat1": _capi.F_WEIGHTS_SAT1, "fixed": 0}[weights]
        lib = self.dp.lib
        self.nbytes = lib.sihmpc_resident_bytes(self.dp.handle, B)
        self.state = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dp.device)
        u0 = np.ascontiguousarray(np.broadcast_to(np.asarray(u0, dtype=np.float64).reshape(-1, p.nu), (B, p.nu)))
        pes = np.ascontiguousarray(np.asarray(p.peso0 if pesos0 is None else pesos0, dtype=np.float64).reshape(-1))
        st = torch.cuda.current_stream(self.dp.device).cuda_stream
        with torch.cuda.device(self.dp.device):
            _capi.check(lib.sihmpc_resident_init(self.dp.handle, C.c_void_p(self.state.data_ptr()), self.nbytes, B,
                                                 u0.ctypes.data_as(C.c_void_p), pes.ctypes.data_as(C.c_void_p),
                                                 C.c_void_p(st)), "sihmpc_resident_init")

    def h2d_bytes(self):
        return self.batch * 8 * (self.plan.nx + self.plan.ny)

    def d2h_bytes(self):
        return self.batch * (8 * (self.plan.nu + 1) + 4)

    def step(self, x=None, ysp=None):
        """x [B][nx], ysp [B][ny] (tensors / arrays; None: whatever is in self.x / self.ysp).  Returns (du, J)
        as views of the pinned result buffers."""
        if x is not None:
            self.x.copy_(torch.as_tensor(x).reshape(self.batch, -1))
        if ysp is not None:
            self.ysp.copy_(torch.as_tensor(ysp).reshape(self.batch, -1))
        return self.step_from(self.x, self.ysp)

    def step_from(self, x_pinned: torch.Tensor, ysp_pinned: torch.Tensor):
        """The same with caller-owned pinned buffers (no host copy): [B][nx] and [B][ny] FP64."""
        io = _capi.ResidentIO()
        io.x, io.ysp = x_pinned.data_ptr(), ysp_pinned.data_ptr()
        io.du, io.J, io.status = self.du.data_ptr(), self.J.data_ptr(), self.status.data_ptr()
        io.flags, io.alfa = self.flags, self.alfa
        st = torch.cuda.current_stream(self.dp.device).cuda_stream
        with torch.cuda.device(self.dp.device):
            _capi.check(self.dp.lib.sihmpc_resident_step(self.dp.handle, C.c_void_p(self.state.data_ptr()), self.nbytes,
                                                         C.byref(io), self.batch, C.c_void_p(st)), "sihmpc_resident_step")
        return self.du, self.J


def fp64_peak_flops() -> float:
    """Measured DFMA throughput of the current device (FLOP/s)."""
    lib = _capi.load()
    v = C.c_double(0.0)
    _capi.check(lib.sihmpc_fp64_peak(C.byref(v)), "sihmpc_fp64_peak")
    return float(v.value)
