import sys; sys.path.insert(0,'.')
import numpy as np, torch
from sihmpc_b200 import scenarios as sc
from oracle.ihmpc_log_oracle import OracleIHMPCLog, declare_distillation_log
from sihmpc.ihmpccontroller_log import IHMPCController as LogController
sys_ = sc.distillation_system()
c = declare_distillation_log(LogController, sys_)
B, T = 24, 30
scen = sc.distillation_scenarios(B)
scen["seg_start"][:, 1] = 4; scen["seg_start"][:, 2] = 18
eng = c.engine(B); eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"])
for k in range(T):
    eng.run(1)
    it=eng.iters.cpu().numpy(); st=eng.status.cpu().numpy()
    if st.any() or it.max()>20: print(k, "iters", it[[0,13]], "max", it.max(), "status", st[[0,13]], "J", eng.J[[0,13]].cpu().numpy(), "ss", eng.w_opt[[0,13],-1].cpu().numpy())
print("dist_step", scen["dist_step"][[0,13]])
