import sys, os; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
for gi in (62575, 36890, 9856):
    ref=np.load("scratch/or_%d.npy"%gi, allow_pickle=True).item()
    for mode in ("inkernel","handover","general"):
        for k in ("SIHMPC_NO_INKERNEL_AS","SIHMPC_NO_FAST"): os.environ.pop(k,None)
        if mode=="handover": os.environ["SIHMPC_NO_INKERNEL_AS"]="1"
        if mode=="general": os.environ["SIHMPC_NO_FAST"]="1"
        c=sc.declare_distillation(IHMPCController)
        scen=sc.mismatch_scenarios(8, first=gi)
        eng=c.engine(8); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
        tr=eng.run(200, record=True)
        du=tr["du"].cpu().numpy()[:,0]
        err=np.abs(du-ref["du"]).max(1)/np.abs(ref["du"]).max()
        first=np.argmax(ref["na"]>0)
        print(gi, mode, "counters", eng.dp.counters(), "first active", first, "rel err at steps", {k: float("%.2e"%err[k]) for k in (first-1, first, first+1, first+3, first+6, first+10, first+20, 199)})
