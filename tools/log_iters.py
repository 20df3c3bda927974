import sys, os; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller_log import IHMPCController as L
sys_ = sc.distillation_system(); c = sc.declare_distillation_log(L, sys_)
B = 4096; scen = sc.distillation_scenarios(B)
seg = scen["seg_start"].copy(); seg[:, 1] = 51; seg[:, 2] = 401
eng = c.engine(B); eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], seg)
eng.run(45)
it0 = eng.stats[:, sys_.ny + 4].cpu().numpy().copy()
per = []
for k in range(20):
    eng.run(1); per.append(eng.iters.cpu().numpy().copy())
per = np.array(per)          # [20][B] iterations per step
tot = per.sum(0)
print("iterations per scenario over the window: mean %.1f median %.0f p99 %.0f max %d" % (tot.mean(), np.median(tot), np.percentile(tot, 99), tot.max()))
print("per step: mean", per.mean(1).round(2).tolist())
print("per step: max ", per.max(1).tolist())
print("scenarios with > 200 iterations:", int((tot > 200).sum()), " > 100:", int((tot > 100).sum()))
print("status or", int(eng.stats[:, sys_.ny + 3].max().item()))
