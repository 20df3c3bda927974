import os, sys, subprocess, numpy as np
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    mode = sys.argv[1]
    import torch
    from sihmpc_b200 import scenarios as sc
    from sihmpc_b200.ihmpccontroller import IHMPCController
    B, T = 20000, 60
    c = sc.declare_distillation(IHMPCController)
    scen = sc.mismatch_scenarios(B)
    eng = c.engine(B)
    eng.set_state(scen["x0"], scen["u0"]); eng.set_schedule(scen["ysp_seq"], scen["seg_start"]); eng.set_plant_values(scen["plant_val"])
    tr = eng.run(T, record=True)
    np.savez("gpurun_out/dbg_%s.npz" % mode, du=tr["du"].cpu().numpy(), J=tr["J"].cpu().numpy(), x=eng.x.cpu().numpy(), stats=eng.stats.cpu().numpy(), pesos=eng.pesos.cpu().numpy())
    sys.exit(0)
env = dict(os.environ)
subprocess.check_call([sys.executable, __file__, "fast"], env=env)
env["SIHMPC_NO_FAST"] = "1"
subprocess.check_call([sys.executable, __file__, "gen"], env=env)
a = np.load("gpurun_out/dbg_fast.npz"); b = np.load("gpurun_out/dbg_gen.npz")
for k in ("du", "J", "x", "stats", "pesos"):
    d = np.abs(a[k] - b[k]); d = np.where(np.isfinite(d), d, 1e30)
    print(k, "max diff", d.max(), "argmax", np.unravel_index(d.argmax(), d.shape), "nonfinite fast", (~np.isfinite(a[k])).sum(), "gen", (~np.isfinite(b[k])).sum())
bad = np.nonzero(np.abs(a["du"] - b["du"]).max(axis=(0, 2)) > 1e-9)[0]
print("bad scenarios", len(bad), bad[:20])
print("status col fast", np.unique(a["stats"][:, 5]), "gen", np.unique(b["stats"][:, 5]))
