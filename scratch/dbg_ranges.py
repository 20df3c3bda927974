import sys, os; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
c=sc.declare_distillation(IHMPCController, sc.distillation_system(), rslb=[-0.1,-0.05], rsub=[0.1,0.05])
p=c.compile(); print("nv",p.nv,"n_extra",p.n_extra,"ncd",p.ncd,"E",p.E.shape,"Gc",p.Gc.shape,"Ht",p.Ht.shape,"HJ",p.HJ.shape,"Wdu",p.Wdu.shape,"Myu",p.Myu.shape,"Mi1u",p.Mi1u.shape,"Gam",p.Gam.shape)
scen=sc.distillation_scenarios(8)
eng=c.engine(8); print(eng.dp.kernel_name())
eng.set_state(scen["x0"],scen["u0"]); eng.set_schedule(scen["ysp_seq"],scen["seg_start"])
try:
    eng.run(5); torch.cuda.synchronize(); print("run ok", eng.stats[:,5].cpu().numpy())
except Exception as e: print("run failed", e)
x=np.append([96,0.5],np.zeros(20)); 
for kw in (dict(ViN_ant=[np.inf,np.inf]), dict(ViN_ant=[np.inf,np.inf]), dict(), dict(ViN_ant=[np.inf,np.inf])):
    try:
        sol=c.mpc(x0=x, ySP=[96,1.0], w0=[], u0=[1.95,1.71], pesos=[0.1,0.4,4.4,4.4], **kw); print("mpc ok", c.ViNant, float(sol['J']), sol['w_opt'].full().ravel()[:4])
    except Exception as e: print("mpc failed", kw, e)
