"""Launch time of the log kernel against the batch (20 steps per launch, window [45,65)): python tools/log_batch_sweep.py"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller_log import IHMPCController as L
sys_ = sc.distillation_system(); c = sc.declare_distillation_log(L, sys_)
for B in (32, 592, 1184, 2368, 4096, 4736, 9472, 18944):
    scen = sc.distillation_scenarios(B)
    seg = scen["seg_start"].copy(); seg[:, 1] = 51; seg[:, 2] = 401
    eng = c.engine(B); eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], seg)
    eng.run(45); snap = eng.snapshot(); eng.run(20)
    ms = []
    for _ in range(3):
        eng.restore(snap); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(20); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    print("batch %6d  warps/SM %5.1f  launch %.3f ms  per step %.1f us  %.3g scenario-steps/s" % (B, B / 148, t, t * 50, B * 20 / t * 1e3))
