import os, sys; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from sihmpc_b200.ihmpccontroller import IHMPCController
sys_=sc.mimo10_system()
B=64; scen=sc.mimo10_scenarios(B, sys_)
def mk(mode):
    os.environ.pop("SIHMPC_NO_FAST",None); os.environ.pop("SIHMPC_FAST",None)
    if mode=="general": os.environ["SIHMPC_NO_FAST"]="1"
    c=sc.declare_mimo10(IHMPCController, sys_)
    eng=c.engine(B); return c, eng
cg,eg=mk("general"); eg.set_state(scen["x0"],scen["u0"]); eg.set_schedule(scen["ysp_seq"],scen["seg_start"]); eg.set_disturbance(scen["dist_step"],scen["dist"])
eg.run(104)
ct,et=mk("tile")
plan=cg.compile()
for k in range(104,110):
    x=eg.x.cpu().numpy().copy(); u=eg.u.cpu().numpy().copy(); p=eg.pesos.cpu().numpy().copy()
    ysp=scen["ysp_seq"][:,2,:] if k>=101 else scen["ysp_seq"][:,1,:]
    et.set_state(x,u,p); et.solve(ysp)
    wt=et.w_opt.cpu().numpy().copy(); Jt=et.J.cpu().numpy().copy(); stt=et.status.cpu().numpy().copy()
    # general single solve on same inputs
    cg2,eg2=mk("general"); eg2.set_state(x,u,p); eg2.solve(ysp); wg=eg2.w_opt.cpu().numpy(); Jg=eg2.J.cpu().numpy(); na=eg2.n_active.cpu().numpy()
    d=np.abs(wt[:,:50]-wg[:,:50]).max(axis=1)
    b=int(np.argmax(d))
    print(k,"max dU diff %.3e (scen %d) |dU| %.3e  J rel diff %.3e  pesos max %.3e min %.3e n_active(gen) %d status tile %s"%(d.max(), b, np.abs(wg[b,:50]).max(), np.abs(Jt-Jg).max()/np.abs(Jg).max(), p.max(), p.min(), na.max(), np.unique(stt)))
    if k==106:
        # Hessian conditioning of worst scenario
        pb=p[b]; H=plan.Ht.copy()
        for a in range(plan.mh): H+=pb[plan.hj_index[a]]*plan.HJ[a]
        H+=2*np.diag(plan.Wdu@pb)
        ev=np.linalg.eigvalsh(H); print(" cond(H) %.3e  min ev %.3e max %.3e"%(ev[-1]/ev[0],ev[0],ev[-1]))
    eg.run(1)
