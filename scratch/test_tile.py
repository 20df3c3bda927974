"""Dev harness: tile path (SIHMPC_FAST=tile) vs the general kernel (SIHMPC_NO_FAST=1), same inputs."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
from sihmpc_b200.engine import BatchedEngine

def make(which, B, mode):
    os.environ.pop("SIHMPC_NO_FAST", None); os.environ.pop("SIHMPC_FAST", None)
    if mode == "general": os.environ["SIHMPC_NO_FAST"] = "1"
    elif mode == "tile": os.environ["SIHMPC_FAST"] = "tile"
    if which == "mimo10":
        sys_ = sc.mimo10_system(); c = sc.declare_mimo10(IHMPCController, sys_); scen = sc.mimo10_scenarios(B, sys_)
    elif which == "dist":
        sys_ = sc.distillation_system(); c = sc.declare_distillation(IHMPCController, sys_); scen = sc.distillation_scenarios(B)
    else:
        sys_ = sc.distillation_system(); c = sc.declare_distillation(IHMPCController, sys_); scen = sc.mismatch_scenarios(B)
    eng = BatchedEngine(c.compile(), B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
    if "dist_step" in scen: eng.set_disturbance(scen["dist_step"], scen["dist"])
    if "plant_val" in scen: eng.set_plant_values(scen["plant_val"])
    return eng

def compare(which, B, T):
    a = make(which, B, "general"); ta = a.run(T, record=True)
    b = make(which, B, "tile"); tb = b.run(T, record=True)
    torch.cuda.synchronize()
    for k in ("du", "J", "y"):
        x, y = ta[k].cpu().numpy(), tb[k].cpu().numpy()
        print(which, k, "max abs diff %.3e  (ref max %.3e)" % (np.abs(x - y).max(), np.abs(x).max()), "nan", np.isnan(y).sum())
    sa, sb = a.stats.cpu().numpy(), b.stats.cpu().numpy()
    print(which, "stats diff", np.abs(sa - sb).max(axis=0), "iters(tile)", sb[:, -1].sum(), "iters(gen)", sa[:, -1].sum(), "status", np.unique(sb[:, -2]))
    print(which, "x diff %.3e pesos diff %.3e" % ((a.x - b.x).abs().max().item(), ((a.pesos - b.pesos).abs() / a.pesos.abs()).max().item()))

def timeit(which, B, T, mode):
    e = make(which, B, mode)
    e.run(5); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); e.run(T); t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    print("%s B=%d T=%d %s: %.2f ms  %.3e scenario-steps/s" % (which, B, T, mode, ms, B * T / ms * 1e3), "handed", int((e.stats[:, -1] > 0).sum()))

if __name__ == "__main__":
    what = sys.argv[1:] or ["cmp", "time"]
    if "cmp" in what:
        compare("dist", 70, 30)
        compare("mimo10", 70, 30)
        compare("mismatch", 64, 40)
    if "time" in what:
        timeit("mimo10", 16384, 50, "general")
        timeit("mimo10", 16384, 50, "tile")
        timeit("mimo10", 65536, 50, "tile")
        timeit("dist", 131072, 50, "tile")
