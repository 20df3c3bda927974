import sys, os, time; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
B=int(os.environ.get("B","131072")); T=200
scen=sc.mismatch_scenarios(B)
for mode in ("inkernel","handover"):
    os.environ.pop("SIHMPC_NO_INKERNEL_AS",None)
    if mode=="handover": os.environ["SIHMPC_NO_INKERNEL_AS"]="1"
    c=sc.declare_distillation(IHMPCController)
    for rep in range(2):
        eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(T); e1.record(); torch.cuda.synchronize()
    st=eng.stats.cpu().numpy()
    print(mode, "ms", e0.elapsed_time(e1), "counters(handed,inkernel)", eng.dp.counters(), "status_or", int(st[:,5].max()), "viol", st[:,3].max(), "sumdu2", st[:,2].sum(), "ise", st[:,0].sum(), st[:,1].sum())
    if mode=="inkernel": ref=(eng.x.cpu().numpy(), st)
    else:
        print("max |x diff|", np.abs(eng.x.cpu().numpy()-ref[0]).max(), "stats diff", np.abs(st[:,:5]-ref[1][:,:5]).max(0))
        d=np.abs(eng.x.cpu().numpy()-ref[0]).max(1)
        bad=np.argsort(-d)[:8]
        print("worst idx", bad, d[bad])
        np.save("gpurun_out/as_bad_idx.npy", bad)
