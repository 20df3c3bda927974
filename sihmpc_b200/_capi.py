# -*- coding: utf-8 -*-
"""ctypes binding of libsihmpc_b200.so (include/sihmpc_b200.h).  Fails loudly when the
library is missing or cannot be loaded: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIHMPC_B200_LIB") or os.path.join(HERE, "libsihmpc_b200.so")   # override: A/B builds only

ABI_VERSION = 3
F_ADVANCE, F_UPDATE_VIN, F_WEIGHTS_SAT2, F_WEIGHTS_SAT1 = 1, 2, 4, 8
ST_MAXITER, ST_INFEASIBLE, ST_NOTPD, ST_ACTIVE_UNSUPPORTED = 1, 2, 4, 8
NSTATS_EXTRA = 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class PlanDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("nx", C.c_int32), ("nu", C.c_int32), ("ny", C.c_int32), ("nd", C.c_int32), ("N", C.c_int32),
        ("m", C.c_int32), ("mh", C.c_int32), ("R", C.c_int32), ("ncd", C.c_int32),
        ("n_vin", C.c_int32), ("pl_nnz", C.c_int32),
        ("Ts", C.c_double),
        ("E", _dp), ("Px", _dp), ("Py", _dp), ("e0", _dp), ("Wrow", _dp), ("w0row", _dp), ("Wdu", _dp),
        ("HJ", _dp), ("hj_index", _ip), ("Ht", _dp),
        ("Myx", _dp), ("Myu", _dp), ("Mi1x", _dp), ("Mi1u", _dp), ("tau", _dp), ("gamma", _dp),
        ("dulb", _dp), ("duub", _dp), ("ulb", _dp), ("uub", _dp),
        ("Gc", _dp), ("Gcx", _dp), ("Gcy", _dp), ("gc0", _dp), ("Lc", _dp), ("Uc", _dp),
        ("gc_vin", _ip), ("gc_q", _dp), ("vin_row", _ip), ("vin_q", _dp),
        ("pl_rowptr", _ip), ("pl_col", _ip), ("pl_val", _dp),
        ("C", _dp), ("D", _dp), ("Phi", _dp), ("Gam", _dp),
        ("n_extra", C.c_int32),
    ]


class IO(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("u", C.c_void_p), ("pesos", C.c_void_p), ("vinant", C.c_void_p),
        ("ysp", C.c_void_p),
        ("ysp_seq", C.c_void_p), ("seg_start", C.c_void_p), ("n_seg", C.c_int32),
        ("dist_step", C.c_void_p), ("dist", C.c_void_p),
        ("plant_val", C.c_void_p),
        ("du", C.c_void_p), ("J", C.c_void_p), ("w_opt", C.c_void_p),
        ("status", C.c_void_p), ("n_active", C.c_void_p),
        ("stats", C.c_void_p), ("traj_du", C.c_void_p), ("traj_J", C.c_void_p), ("traj_y", C.c_void_p),
        ("flags", C.c_int32), ("k0", C.c_int32), ("alfa", C.c_double),
        ("lam", C.c_void_p),
    ]


class ResidentIO(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ysp", C.c_void_p), ("du", C.c_void_p), ("J", C.c_void_p), ("status", C.c_void_p),
                ("flags", C.c_int32), ("alfa", C.c_double)]


EXPORTS = [
    "sihmpc_abi_version", "sihmpc_last_error", "sihmpc_device_count", "sihmpc_plan_create",
    "sihmpc_plan_destroy", "sihmpc_workspace_bytes", "sihmpc_step", "sihmpc_step_host",
    "sihmpc_host_staging_bytes", "sihmpc_predict", "sihmpc_subobjectives", "sihmpc_stats_reduce",
    "sihmpc_fp64_peak", "sihmpc_log_step", "sihmpc_log_loop", "sihmpc_plan_kernel", "sihmpc_plan_counters",
    "sihmpc_resident_bytes", "sihmpc_resident_init", "sihmpc_resident_step", "sihmpc_resident_weights_offset",
]

_lib = None


class SihmpcError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises SihmpcError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SihmpcError(
            "libsihmpc_b200.so is not built (%s). Run `python -m sihmpc_b200.build` "
            "(needs nvcc; cross-compiles sm_100a without a GPU). There is no CPU fallback." % LIB_PATH)
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise SihmpcError("cannot load %s: %s" % (LIB_PATH, e))
    lib.sihmpc_abi_version.restype = C.c_int
    lib.sihmpc_last_error.restype = C.c_char_p
    lib.sihmpc_device_count.restype = C.c_int
    lib.sihmpc_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]
    lib.sihmpc_plan_create.restype = C.c_int
    lib.sihmpc_plan_destroy.argtypes = [C.c_void_p]
    lib.sihmpc_plan_destroy.restype = C.c_int
    lib.sihmpc_workspace_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.sihmpc_workspace_bytes.restype = C.c_size_t
    lib.sihmpc_step.argtypes = [C.c_void_p, C.POINTER(IO), C.c_int64, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sihmpc_step.restype = C.c_int
    lib.sihmpc_step_host.argtypes = [C.c_void_p, C.POINTER(IO), C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.sihmpc_step_host.restype = C.c_int
    lib.sihmpc_host_staging_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.sihmpc_host_staging_bytes.restype = C.c_size_t
    lib.sihmpc_predict.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 4
    lib.sihmpc_predict.restype = C.c_int
    lib.sihmpc_subobjectives.argtypes = [C.c_void_p] + [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_void_p]
    lib.sihmpc_subobjectives.restype = C.c_int
    lib.sihmpc_stats_reduce.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sihmpc_stats_reduce.restype = C.c_int
    lib.sihmpc_log_step.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 5
    lib.sihmpc_log_step.restype = C.c_int
    lib.sihmpc_log_loop.argtypes = [C.c_void_p, C.POINTER(IO), C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.sihmpc_log_loop.restype = C.c_int
    lib.sihmpc_plan_kernel.argtypes = [C.c_void_p]
    lib.sihmpc_plan_kernel.restype = C.c_char_p
    lib.sihmpc_plan_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
    lib.sihmpc_plan_counters.restype = C.c_int
    lib.sihmpc_resident_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.sihmpc_resident_bytes.restype = C.c_size_t
    lib.sihmpc_resident_init.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sihmpc_resident_init.restype = C.c_int
    lib.sihmpc_resident_step.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(ResidentIO), C.c_int64, C.c_void_p]
    lib.sihmpc_resident_step.restype = C.c_int
    lib.sihmpc_resident_weights_offset.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_size_t)]
    lib.sihmpc_resident_weights_offset.restype = C.c_int
    lib.sihmpc_fp64_peak.argtypes = [C.POINTER(C.c_double)]
    lib.sihmpc_fp64_peak.restype = C.c_int
    if lib.sihmpc_abi_version() != ABI_VERSION:
        raise SihmpcError("libsihmpc_b200.so ABI %d != binding ABI %d" % (lib.sihmpc_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().sihmpc_last_error().decode("utf-8", "replace")
        raise SihmpcError("%s failed (code %d): %s" % (what, rc, msg))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def make_desc(plan):
    """sihmpc_b200.plan.Plan -> (PlanDesc, keepalive list of numpy arrays)."""
    keep = []

    def d(a):
        a = _f64(a); keep.append(a)
        return a.ctypes.data_as(_dp) if a.size else C.cast(None, _dp)

    def i(a):
        a = _i32(a); keep.append(a)
        return a.ctypes.data_as(_ip) if a.size else C.cast(None, _ip)

    desc = PlanDesc(
        abi_version=ABI_VERSION, nx=plan.nx, nu=plan.nu, ny=plan.ny, nd=plan.nd, N=plan.N, m=plan.m,
        mh=plan.mh, R=plan.R, ncd=plan.ncd, n_vin=plan.n_vin, pl_nnz=int(plan.pl_val.size), Ts=plan.Ts,
        E=d(plan.E), Px=d(plan.Px), Py=d(plan.Py), e0=d(plan.e0), Wrow=d(plan.Wrow), w0row=d(plan.w0row),
        Wdu=d(plan.Wdu), HJ=d(plan.HJ), hj_index=i(plan.hj_index), Ht=d(plan.Ht),
        Myx=d(plan.Myx), Myu=d(plan.Myu), Mi1x=d(plan.Mi1x), Mi1u=d(plan.Mi1u), tau=d(plan.tau),
        gamma=d(plan.gamma), dulb=d(plan.dulb), duub=d(plan.duub), ulb=d(plan.ulb), uub=d(plan.uub),
        Gc=d(plan.Gc), Gcx=d(plan.Gcx), Gcy=d(plan.Gcy), gc0=d(plan.gc0), Lc=d(plan.Lc), Uc=d(plan.Uc),
        gc_vin=i(plan.gc_vin), gc_q=d(plan.gc_q), vin_row=i(plan.vin_row), vin_q=d(plan.vin_q),
        pl_rowptr=i(plan.pl_rowptr), pl_col=i(plan.pl_col), pl_val=d(plan.pl_val),
        C=d(plan.C), D=d(plan.D), Phi=d(plan.Phi), Gam=d(plan.Gam), n_extra=int(getattr(plan, 'n_extra', 0)))
    return desc, keep
