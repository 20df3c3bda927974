import os, sys, subprocess, numpy as np
sys.path.insert(0, '.')
import torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B = 131072
c = sc.declare_distillation(IHMPCController)
scen = sc.mismatch_scenarios(B)
eng = c.engine(B)
eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
for n in (10, 200):
    eng.run(n)
    st = eng.stats.cpu().numpy()
    bad = np.nonzero(st[:, 5] != 0)[0]
    print("after", eng.k, "steps: bad", len(bad), bad[:10], "nonfinite x", (~np.isfinite(eng.x.cpu().numpy())).any(axis=1).sum())
    if len(bad):
        b = bad[0]
        print("stats", st[b], "x", eng.x[b].cpu().numpy()[:8], "pesos", eng.pesos[b].cpu().numpy(), "K", scen["K"][b].ravel(), "tau", scen["tau"][b].ravel())
