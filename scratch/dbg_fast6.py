import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B, T = 45000, 200
opts = sys.argv[1:]
c = sc.declare_distillation(IHMPCController) if "free" not in opts else sc.declare_distillation(IHMPCController, None, 10, ulb=[-np.inf, -np.inf])
scen = sc.mismatch_scenarios(B)
eng = c.engine(B, weights="fixed" if "fixed" in opts else "sat2")
eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
if "nominal" not in opts:
    eng.set_plant_values(scen["plant_val"])
tr = eng.run(T, record=("norec" not in opts))
x = eng.x.cpu().numpy()
bad = np.nonzero(~np.isfinite(x).all(axis=1))[0]
st = eng.stats.cpu().numpy()
print(opts, "bad x", len(bad), bad[:6], "status", np.unique(st[:, 5]), "max|x|", np.nanmax(np.abs(x)))
