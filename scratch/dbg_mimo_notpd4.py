import os, sys; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
sys_=sc.mimo10_system(); ny=sys_.ny
B=64; scen=sc.mimo10_scenarios(B, sys_)
def mk(mode):
    os.environ.pop("SIHMPC_NO_FAST",None); os.environ.pop("SIHMPC_FAST",None)
    if mode=="general": os.environ["SIHMPC_NO_FAST"]="1"
    c=sc.declare_mimo10(IHMPCController, sys_)
    eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_disturbance(scen["dist_step"],scen["dist"]); return eng
eg=mk("general"); et=mk("tile")
for k in range(112):
    itg0=eg.stats[:,ny+4].clone(); itt0=et.stats[:,ny+4].clone()
    eg.run(1); et.run(1)
    dx=(eg.x-et.x).abs().max(dim=1).values; dp=((eg.pesos-et.pesos).abs()/eg.pesos.abs()).max(dim=1).values; du=(eg.u-et.u).abs().max(dim=1).values
    ddu=(eg.du-et.du).abs().max(dim=1).values
    if k>=100 or dx.max()>1e-9:
        b=int(torch.argmax(dx))
        print(k,"dx %.2e du %.2e dpesos %.2e ddu %.2e worst scen %d; iters this step gen %d tile %d ; worst scen iters gen %d tile %d"%(dx.max(),du.max(),dp.max(),ddu.max(),b,int((eg.stats[:,ny+4]-itg0).sum()),int((et.stats[:,ny+4]-itt0).sum()),int(eg.stats[b,ny+4]-itg0[b]),int(et.stats[b,ny+4]-itt0[b])))
    if dx.max()>1e-6: break
