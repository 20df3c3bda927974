import sys, os; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B=131072; T=200
scen=sc.mismatch_scenarios(B)
c=sc.declare_distillation(IHMPCController)
for w in ("sat2","fixed","sat1"):
    for rep in range(2):
        eng=c.engine(B, weights=w); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(T); e1.record(); torch.cuda.synchronize()
    print(w, "ms", e0.elapsed_time(e1), eng.dp.counters())
