import sys; sys.path.insert(0,'.')
import numpy as np, torch, time
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B=131072
c=sc.declare_distillation(IHMPCController); scen=sc.mismatch_scenarios(B)
eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
eng.run(200); eng.run(200); torch.cuda.synchronize()
eng.run(50); torch.cuda.synchronize()
print("iters", eng.stats[:,6].sum().item(), "handed", int((eng.stats[:,6]>0).sum()))
