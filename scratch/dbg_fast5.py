import os, sys, subprocess, numpy as np
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    mode = sys.argv[1]
    import torch
    from sihmpc_b200 import scenarios as sc
    from sihmpc_b200.ihmpccontroller import IHMPCController
    B, T = 512, int(os.environ.get("T", "200"))
    c = sc.declare_distillation(IHMPCController)
    scen = sc.mismatch_scenarios(B)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
    tr = eng.run(T, record=True)
    torch.cuda.synchronize()
    np.savez("/tmp/dbg_%s.npz" % mode, du=tr["du"].cpu().numpy(), J=tr["J"].cpu().numpy(), x=eng.x.cpu().numpy(), stats=eng.stats.cpu().numpy())
    sys.exit(0)
env = dict(os.environ); env["SIHMPC_QUAD_GRID"] = "4"
subprocess.check_call([sys.executable, __file__, "fast"], env=env)
env["SIHMPC_NO_FAST"] = "1"
subprocess.check_call([sys.executable, __file__, "gen"], env=env)
a = np.load("/tmp/dbg_fast.npz"); b = np.load("/tmp/dbg_gen.npz")
d = np.abs(a["du"] - b["du"]); d = np.where(np.isfinite(d), d, 1e30).max(axis=(0, 2))
bad = np.nonzero(d > 1e-9)[0]
print("bad", len(bad), bad[:40])
if len(bad):
    J = a["J"]; first = [int(np.argmax(np.abs(np.where(np.isfinite(a['du'][:, s]), a['du'][:, s], 1e30) - b['du'][:, s]).max(axis=1) > 1e-9)) for s in bad]
    print("first differing step", np.unique(first, return_counts=True))
