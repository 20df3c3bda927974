import sys, os; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
for mode in ("default","general"):
    os.environ.pop("SIHMPC_NO_FAST",None)
    if mode=="general": os.environ["SIHMPC_NO_FAST"]="1"
    c=sc.declare_distillation(IHMPCController); scen=sc.mismatch_scenarios(16, first=843880)
    eng=c.engine(16); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
    eng.run(210)
    st=eng.stats.cpu().numpy()
    print(mode, "status", st[:,5], "x[10]", eng.x[10,:3].cpu().numpy(), "J", st[10,4], "iters", st[10,6])
