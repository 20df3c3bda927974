import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B = 60000
c = sc.declare_distillation(IHMPCController)
scen = sc.mismatch_scenarios(B)
pv = scen["plant_val"]
print("plant_val finite", np.isfinite(pv).all(), "shape", pv.shape)
for T in (1, 2):
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"]); eng.set_plant_values(pv)
    tr = eng.run(T, record=True)
    J = tr["J"].cpu().numpy(); x = eng.x.cpu().numpy(); st = eng.stats.cpu().numpy()
    badJ = np.nonzero(~np.isfinite(J).all(axis=0))[0]; badx = np.nonzero(~np.isfinite(x).all(axis=1))[0]
    print("T", T, "bad J", len(badJ), badJ[:8], "bad x", len(badx), badx[:8], "status", np.unique(st[:, 5]), "iters", np.unique(st[:, 6]))
    if len(badx):
        b = badx[0]; print(" x", x[b], "\n pv", pv[b], "\n pv ok", pv[b-1])
