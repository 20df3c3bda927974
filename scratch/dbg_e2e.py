import os, sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc, _capi
from sihmpc_b200.ihmpccontroller import IHMPCController
from sihmpc_b200.engine import HostStepper
c=sc.declare_distillation(IHMPCController); plan=c.compile()
fl=_capi.F_ADVANCE|_capi.F_UPDATE_VIN|_capi.F_WEIGHTS_SAT2
B=131072; scen=sc.mismatch_scenarios(B)
os.environ["SIHMPC_ZEROCOPY_MAX"]=sys.argv[1] if len(sys.argv)>1 else "512"
for ch in (4,):
    os.environ["SIHMPC_HOST_CHUNKS"]=str(ch)
    h=HostStepper(plan,B)
    h.x.copy_(torch.as_tensor(scen["x0"])); h.u.copy_(torch.as_tensor(scen["u0"])); h.ysp.copy_(torch.as_tensor(np.ascontiguousarray(scen["ysp_seq"][:,1,:])))
    for i in range(3): h.step(flags=fl, want_w_opt=False)
    t0=time.perf_counter()
    for i in range(20): h.step(flags=fl, want_w_opt=False)
    dt=(time.perf_counter()-t0)/20
    print(ch, "ms/step %.3f  %.3e scenario-steps/s  %.1f GB/s each way"%(dt*1e3, B/dt, 34.5e6/dt/1e9))
