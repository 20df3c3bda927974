import sys, os; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B=131072
c=sc.declare_distillation(IHMPCController); scen=sc.mismatch_scenarios(B)
eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
hist=[]
for k in range(420):
    eng.run(1)
    na=eng.n_active.cpu().numpy(); st=eng.status.cpu().numpy()
    hist.append((int((na>0).sum()), int((st!=0).sum())))
    if hist[-1][0] or hist[-1][1]:
        idx=np.nonzero((na>0)|(st!=0))[0]
        if k%10==0 or len(hist)<3 or hist[-2]==(0,0): print(k, hist[-1], idx[:6], "u", eng.u[idx[0]].cpu().numpy())
h=np.array(hist); print("steps with any active:", int((h[:,0]>0).sum()), "max count", h[:,0].max(), "first", int(np.argmax(h[:,0]>0)))
