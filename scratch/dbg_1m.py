import sys, numpy as np
sys.path.insert(0, '.')
import torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B = 1048576
c = sc.declare_distillation(IHMPCController)
scen = sc.mismatch_scenarios(B)
eng = c.engine(B)
eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
eng.run(210)
st = eng.stats.cpu().numpy()
bad = np.nonzero(st[:, 5] != 0)[0]
print("bad", len(bad), bad[:10], "status values", np.unique(st[:, 5]))
print("handed-over (iters>0):", int((st[:, 6] > 0).sum()), "max viol", st[:, 3].max())
for b in bad[:3]:
    print(b, "K", scen["K"][b].ravel(), "tau", scen["tau"][b].ravel(), "x", eng.x[b].cpu().numpy()[:6], "stats", st[b])
np.save("gpurun_out/bad_idx.npy", bad[:20])
