"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/ihmpc_oracle.py).

PROVENANCE: restatement-generated.  The reference itself cannot run in this image (casadi and
opom are absent), so these vectors pin the ORACLE, not the reference; they keep
the oracle, the plan compiler and the CUDA path from drifting apart silently.  In an
environment that has casadi + opom, `python tests/golden/make_golden.py --reference` runs the
unmodified reference classes through the same loop instead.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def example_loop(ctrl_cls, sys_mod, n_steps=200, t_end=200):
    """example_distillation.py:65-113 verbatim in structure."""
    from sihmpc_b200 import scenarios as sc
    sys_ = sc.distillation_system() if sys_mod is None else sys_mod
    c = sc.declare_distillation(ctrl_cls, sys_)
    u = [1.95, 1.71]
    x = np.append([96, 0.5], np.zeros(sys_.nx - 2))
    ysp = [96, 0.5]
    w0 = []; pesos = []
    out = dict(du=[], x=[], u=[], y=[], J=[], pesos=[], w_opt=[])
    for k in range(n_steps):
        if k > 10:
            ysp = [96, 1]
        if k > t_end / 2:
            ysp = [95.5, 1]
        sol = c.mpc(x0=x, ySP=ysp, w0=w0, u0=u, pesos=pesos)
        du = np.asarray(c.du, float).reshape(-1)
        res = c.dynF(x0=x, u0=u, du0=du)
        full = lambda v: np.asarray(v.full() if hasattr(v, "full") else v, float).reshape(-1)
        x, u, y = full(res['xkp1']), full(res['ukp1']), full(res['ykp1'])
        w0 = c.warmStart(sol, ysp)
        pesos, _ = c.satWeights2(x, u, w0, ysp)
        out['du'].append(du); out['x'].append(x); out['u'].append(u); out['y'].append(y)
        out['J'].append(float(sol['J'])); out['pesos'].append(np.array(pesos, float))
        out['w_opt'].append(full(sol['w_opt']))
    return {k: np.array(v) for k, v in out.items()}


def main():
    if "--reference" in sys.argv:
        from opom import OPOM, TransferFunction  # noqa: F401  (the real packages)
        from sihmpc import IHMPCController
        r = example_loop(IHMPCController, None)
        np.savez_compressed(os.path.join(HERE, "example_distillation_reference.npz"), **r)
        return
    from oracle.ihmpc_oracle import OracleIHMPC
    from sihmpc_b200 import scenarios as sc
    from tests.util import oracle_closed_loop
    r = example_loop(OracleIHMPC, None)
    np.savez_compressed(os.path.join(HERE, "example_distillation_oracle.npz"), **r)
    # config-3 scenarios 0..3, and constrained variant
    sys_ = sc.distillation_system()
    scen = sc.distillation_scenarios(4)
    tight = dict(ulb=[1.9, 1.6], uub=[2.0, 1.8], dulb=[-0.02, -0.02], duub=[0.02, 0.02])
    pack = {}
    for idx in range(4):
        a = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, 200)
        b = oracle_closed_loop(sc.declare_distillation, sys_, scen, idx, 120, bounds=tight)
        pack["du_%d" % idx] = a["du"]; pack["J_%d" % idx] = a["J"]; pack["y_%d" % idx] = a["y"]
        pack["tight_du_%d" % idx] = b["du"]; pack["tight_nact_%d" % idx] = b["n_active"]
    np.savez_compressed(os.path.join(HERE, "config3_oracle.npz"), **pack)
    print("golden written")


if __name__ == "__main__":
    main()
