import sys; sys.path.insert(0,'.')
import numpy as np, pytest
from tests import test_gpu_tile as T
from sihmpc_b200 import scenarios as sc
class MP:
    def setenv(self,k,v): import os; os.environ[k]=v
    def delenv(self,k,raising=False): import os; os.environ.pop(k,None)
mp=MP()
for (ny,nu,N,dmax) in T.SHAPES:
    sys_ = T._plant(ny, nu, 100 + 7 * ny + nu, dmax)
    declare = lambda cls, s, **b: sc.declare_mimo10(cls, s, N=N, **(b or dict(ulb=[-50.0] * nu, uub=[50.0] * nu)))
    scen = T._scenarios(21, sys_, 5)
    t = T._run(mp, "tile", declare, sys_, scen, 30, spl=13)
    g = T._run(mp, "general", declare, sys_, scen, 30, spl=13)
    print((ny,nu,N), "iters tile", t["stats"][:, ny+4].sum(), "general", g["stats"][:, ny+4].sum(), "status", np.unique(t["stats"][:, ny+3]), np.unique(g["stats"][:, ny+3]),
          "du diff %.2e" % np.abs(t["du"]-g["du"]).max(), "umax", np.abs(g["u"]).max(), "dumax", np.abs(g["du"]).max())
