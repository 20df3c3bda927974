import os, sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc, _capi
from sihmpc_b200.ihmpccontroller import IHMPCController
from sihmpc_b200.engine import HostStepper
c=sc.declare_distillation(IHMPCController); plan=c.compile()
fl=_capi.F_ADVANCE|_capi.F_UPDATE_VIN|_capi.F_WEIGHTS_SAT2
for B in (1, 64, 512):
    scen=sc.mismatch_scenarios(B)
    for mode in ("zerocopy","staged"):
        if mode=="staged": os.environ["SIHMPC_NO_ZEROCOPY"]="1"
        else: os.environ.pop("SIHMPC_NO_ZEROCOPY",None)
        h=HostStepper(plan,B)
        h.x.copy_(torch.as_tensor(scen["x0"])); h.u.copy_(torch.as_tensor(scen["u0"])); h.ysp.copy_(torch.as_tensor(np.ascontiguousarray(scen["ysp_seq"][:,1,:])))
        ts=[]
        for i in range(400):
            a=time.perf_counter(); h.step(flags=fl); ts.append(time.perf_counter()-a)
        ts=np.array(ts[100:])*1e6
        print(B, mode, "p50 %.1f us p99 %.1f us"%(np.percentile(ts,50),np.percentile(ts,99)), "x0", h.x[0,:2].numpy(), "du", h.du[0].numpy())
