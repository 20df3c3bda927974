"""p50 of one sihmpc_resident_step at small batches (host to host): python tools/resident_latency.py [iterations]"""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
from sihmpc_b200.engine import ResidentStepper
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sys_ = sc.distillation_system(); c = sc.declare_distillation(IHMPCController, sys_); plan = c.compile()
for B in (1, 64, 1024, 4096):
    scen = sc.distillation_scenarios(B)
    r1 = ResidentStepper(plan, B, scen["u0"])
    x1 = torch.as_tensor(scen["x0"]).pin_memory(); y1 = torch.as_tensor(np.ascontiguousarray(scen["ysp_seq"][:, 1, :])).pin_memory()
    ts = []
    for i in range(n_it):
        a = time.perf_counter(); r1.step_from(x1, y1); ts.append(time.perf_counter() - a)
    ts = np.asarray(ts[n_it // 6:]) * 1e6
    print("batch %5d  resident step p50 %.1f us  p99 %.1f us  du[0] %s" % (B, np.percentile(ts, 50), np.percentile(ts, 99), r1.du[0].numpy()))
