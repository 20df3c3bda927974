import sys, os, time; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B=int(os.environ.get("B","65536")); T=int(os.environ.get("T","200")); SPL=int(os.environ.get("SPL","200"))
sys_=sc.mimo10_system(); scen=sc.mimo10_scenarios(B, sys_)
for mode in os.environ.get("MODES","inkernel,handover").split(","):
    os.environ.pop("SIHMPC_NO_INKERNEL_AS",None)
    if mode=="handover": os.environ["SIHMPC_NO_INKERNEL_AS"]="1"
    c=sc.declare_mimo10(IHMPCController, sys_)
    for rep in range(2):
        eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_disturbance(scen["dist_step"],scen["dist"])
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(T, steps_per_launch=SPL); e1.record(); torch.cuda.synchronize()
    st=eng.stats.cpu().numpy()
    print(mode, "ms", e0.elapsed_time(e1), "counters(handed,inkernel)", eng.dp.counters(), "status_or", int(st[:,-2].max()), "viol", st[:,-4].max(), "n with iters", int((st[:,-1]>0).sum()), flush=True)
    idx=np.nonzero(st[:,-1]>0)[0]
    np.save("gpurun_out/mimo_%s_idx.npy"%mode, idx); np.save("gpurun_out/mimo_%s_iters.npy"%mode, st[idx,-1])
    print("iters per handed scenario: median", np.median(st[idx,-1]), "max", st[idx,-1].max(), "n_active last", np.bincount(eng.n_active.cpu().numpy()[idx])[:20])
