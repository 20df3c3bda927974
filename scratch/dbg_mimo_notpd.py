import os, sys; sys.path.insert(0,'.')
import numpy as np, torch
os.environ["SIHMPC_NO_FAST"]="1"
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
from tests.util import oracle_closed_loop
sys_=sc.mimo10_system(); c=sc.declare_mimo10(IHMPCController, sys_)
B=64; scen=sc.mimo10_scenarios(B, sys_)
eng=c.engine(B); eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"]); eng.set_disturbance(scen["dist_step"],scen["dist"])
T=210
Js=[]; sts=[]; pes=[]
for k in range(T):
    eng.run(1)
    sts.append(eng.stats[:,sys_.ny+3].cpu().numpy().copy()); Js.append(eng.J.cpu().numpy().copy()); pes.append(eng.pesos.cpu().numpy().copy())
sts=np.array(sts); Js=np.array(Js); pes=np.array(pes)
bad=np.nonzero(sts[-1]!=0)[0]; print("flagged", len(bad), "of", B)
if len(bad):
    b=bad[0]; k0=int(np.argmax(sts[:,b]!=0)); print("scenario",b,"first flagged step",k0, "dist_step", scen["dist_step"][b])
    for k in range(max(0,k0-4),min(T,k0+3)): print(k, "J",Js[k,b], "pesos min/max", pes[k,b].min(), pes[k,b].max(), "u", None)
    ref=oracle_closed_loop(sc.declare_mimo10, sys_, scen, b, k0+2)
    print("oracle J", ref["J"][max(0,k0-4):k0+2])
    print("oracle keys", list(ref.keys()))
