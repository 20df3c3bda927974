import sys; sys.path.insert(0,'.')
import numpy as np, torch, time
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B=131072
c=sc.declare_distillation(IHMPCController); scen=sc.mismatch_scenarios(B)
eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
eng.run(10); torch.cuda.synchronize()
for rep in range(6):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    if rep==3: time.sleep(1.0)
    e0.record(); eng.run(200); e1.record(); torch.cuda.synchronize()
    print(rep, "ms", e0.elapsed_time(e1), "steps/s %.3e"%(B*200/e0.elapsed_time(e1)*1e3))
# same with plant_val=None (nominal)
eng.set_plant_values(None)
for rep in range(3):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(200); e1.record(); torch.cuda.synchronize()
    print("nominal", rep, "ms", e0.elapsed_time(e1))
