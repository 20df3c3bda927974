import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B = int(sys.argv[1]); T = int(sys.argv[2]); spl = int(sys.argv[3])
c = sc.declare_distillation(IHMPCController)
scen = sc.mismatch_scenarios(B)
eng = c.engine(B)
eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
tr = eng.run(T, steps_per_launch=spl, record=True)
J = tr["J"].cpu().numpy()
st = eng.stats.cpu().numpy()
bad = np.nonzero(st[:, 5] != 0)[0]
print("B", B, "T", T, "spl", spl, "bad", len(bad))
if len(bad):
    first = np.array([np.argmax(~np.isfinite(J[:, b])) if (~np.isfinite(J[:, b])).any() else -1 for b in bad])
    print("first NaN step histogram", np.unique(first, return_counts=True))
    print("bad idx min/max", bad.min(), bad.max(), "mod 32 unique", np.unique(bad % 32)[:40], "blocks", np.unique(bad // 32)[:50])
    b = bad[0]; f = first[0]
    print("J around", J[max(0, f - 3):f + 2, b], "du", tr["du"][max(0, f - 3):f + 2, b].cpu().numpy())
    hs = eng.dp.lib
