import os, sys; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
sys_=sc.mimo10_system()
B=64; scen=sc.mimo10_scenarios(B, sys_)
def run(mode, T, spl):
    os.environ.pop("SIHMPC_NO_FAST",None); os.environ.pop("SIHMPC_FAST",None)
    if mode=="general": os.environ["SIHMPC_NO_FAST"]="1"
    c=sc.declare_mimo10(IHMPCController, sys_)
    eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_disturbance(scen["dist_step"],scen["dist"])
    tr=eng.run(T, steps_per_launch=spl, record=True)
    return tr["du"].cpu().numpy(), tr["J"].cpu().numpy(), eng.stats.cpu().numpy(), eng.pesos.cpu().numpy()
T=210
for spl in (1, 210):
    dt,Jt,st,pt=run("tile",T,spl); dg,Jg,sg,pg=run("general",T,spl)
    ny=sys_.ny
    print("spl",spl,"tile status", np.unique(st[:,ny+3]), "iters", st[:,ny+4].sum(), "general status", np.unique(sg[:,ny+3]))
    d=np.abs(dt-dg).max(axis=(1,2)); print(" du diff by step max", d.max(), "first step >1e-9:", int(np.argmax(d>1e-9)) if (d>1e-9).any() else None)
    print(" pesos max", pt.max(), pg.max())
