# -*- coding: utf-8 -*-
"""Histories of a scenario subset and the result export of the reference's log example.

The batched engine keeps the state of every scenario on the device and only accumulates statistics; the
per-step histories the reference keeps on the host -- `j_hist`, `fObj.v_hist`, `fObj.peso_hist`
(/root/reference/sihmpc/ihmpccontroller.py:592-601) and the arrays the examples stack for their plots and
pickle at the end (/root/reference/example_distillation_log.py:144-151,263-269) -- are produced here for a
SELECTED SUBSET of the scenarios: scenario draws are addressed by their global index, so a sub-engine over
the chosen indices retraces exactly the trajectories they have inside the big batch; it is stepped one control
step per launch and the device state (x, u, weights), the kernel's traces (du, J, y) and the sub-objective
values (sihmpc_subobjectives on the evaluation plan = `fObj.F`) are read back after every step.
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np
import torch

from . import _capi


def trace_subset(ctrl, scen, indices, n_steps, weights="sat2", alfa=0.0):
    """Closed-loop histories of scenarios `indices` of the scenario set `scen` (dict with x0, u0, ysp_seq,
    seg_start and optionally dist_step/dist, plant_val) for n_steps control steps.

    Returns a dict of arrays with a leading scenario axis: y [S][ny][n+1] (initial output first, like yPlot),
    du [S][nu][n], u [S][nu][n], x [S][nx][n], v [S][len(V)][n] (all sub-objectives), vJ [S][len(VJ)][n],
    J [S][n], pesos [S][len(VJ)][n] (weights used BY step k) -- the layouts of yPlot, duPlot, uPlot, xPlot, vPlot,
    vJ, JPlot, pesosPlot of example_distillation_log.py:144-151."""
    from .engine import DevicePlan
    idx = np.asarray(indices, dtype=np.int64).reshape(-1)
    S = idx.size
    sub = {k: (np.ascontiguousarray(v[idx]) if isinstance(v, np.ndarray) and v.shape[:1] == scen["x0"].shape[:1] else v)
           for k, v in scen.items()}
    is_log = hasattr(ctrl, "Vss")
    eng = ctrl.engine(S) if is_log else ctrl.engine(S, weights=weights, alfa=alfa)
    eng.set_state(sub["x0"], sub["u0"])
    eng.set_schedule(sub["ysp_seq"], sub["seg_start"])
    if "dist_step" in sub:
        eng.set_disturbance(sub["dist_step"], sub["dist"])
    if "plant_val" in sub:
        eng.set_plant_values(sub["plant_val"])
    plan = ctrl.compile()
    ep = ctrl._evaluation_plan()
    edp = DevicePlan(ep, eng.device)
    nv, ny, nu, nx = plan.N * plan.nu, plan.ny, plan.nu, plan.nx
    nV, nVJ = len(ctrl.V), len(ctrl.VJ)
    vj_pos = [ctrl.V.index(V) for V in ctrl.VJ]
    out = dict(y=np.zeros((S, ny, n_steps + 1)), du=np.zeros((S, nu, n_steps)), u=np.zeros((S, nu, n_steps)),
               x=np.zeros((S, nx, n_steps)), v=np.zeros((S, nV, n_steps)), vJ=np.zeros((S, nVJ, n_steps)),
               J=np.zeros((S, n_steps)), pesos=np.zeros((S, nVJ, n_steps)))
    out["y"][:, :, 0] = sub["x0"] @ np.asarray(plan.C).T
    Vd = torch.zeros(S, nV, dtype=torch.float64, device=eng.device)
    st = torch.cuda.current_stream(eng.device).cuda_stream
    seq, start = sub["ysp_seq"], sub["seg_start"]
    for k in range(n_steps):
        x_k = eng.x.clone()
        if not is_log:
            out["pesos"][:, :, k] = eng.pesos.cpu().numpy()
        tr = eng.run(1, record=True)
        w = eng.w_opt[:, :nv + 2 * ny].contiguous()
        seg = (start <= k).sum(axis=1) - 1
        ysp = torch.as_tensor(np.ascontiguousarray(seq[np.arange(S), seg])).to(eng.device)
        with torch.cuda.device(eng.device):
            _capi.check(edp.lib.sihmpc_subobjectives(edp.handle, C.c_void_p(x_k.data_ptr()), C.c_void_p(w.data_ptr()),
                                                     C.c_void_p(ysp.data_ptr()), S, C.c_void_p(Vd.data_ptr()),
                                                     C.c_void_p(st)), "sihmpc_subobjectives")
        vals = Vd.cpu().numpy().copy()
        if is_log:
            vals[:, ctrl.V.index(ctrl.Vss)] = eng.w_opt[:, -1].cpu().numpy() ** 2       # Vss = ss^2 (log.py:76-80)
        out["v"][:, :, k] = vals
        out["vJ"][:, :, k] = vals[:, vj_pos]
        out["du"][:, :, k] = tr["du"][0].cpu().numpy()
        out["J"][:, k] = tr["J"][0].cpu().numpy()
        out["y"][:, :, k + 1] = tr["y"][0].cpu().numpy()
        out["u"][:, :, k] = eng.u.cpu().numpy()
        out["x"][:, :, k] = eng.x.cpu().numpy()
        if is_log:                                                                      # theoretical weights (log.py:566-574)
            gam = np.array([V.gamma for V in ctrl.VJ], dtype=float)
            out["pesos"][:, :, k] = 1.0 / (gam[None, :] - vals[:, vj_pos])
    return out


def result_tuple(hist, s=0):
    """(yPlot, duPlot, uPlot, xPlot, vPlot, vJ, JPlot, pesosPlot) of scenario s of a trace_subset() result, the tuple
    example_distillation_log.py:268 pickles."""
    return (hist["y"][s], hist["du"][s], hist["u"][s], hist["x"][s], hist["v"][s], hist["vJ"][s], hist["J"][s],
            hist["pesos"][s])


def export_results(ctrl, hist, s=0, path=None):
    """Pickle the result tuple under the reference's file name 'dist_{N}_{parm}.dat', parm = sqrt(gamma_i / N) per
    weighted sub-objective (example_distillation_log.py:263-269).  Returns the path written."""
    N = ctrl.N
    parm = [np.sqrt(V.gamma / N) for V in ctrl.VJ]
    path = path or 'dist_{}_{}.dat'.format(N, parm)
    with open(path, 'wb') as f:
        pickle.dump(result_tuple(hist, s), f)
    return path
