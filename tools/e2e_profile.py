"""Where does the end-to-end time of a fresh context go?  set_pencil (symbolic analysis, uploads, factor slots),
first Ros1 step (cold arena / workspaces), second step."""
import sys
import time
import warnings

import numpy as np
import scipy.sparse.linalg as spla

sys.path.insert(0, ".")
import dre_b200
from dre_b200 import api

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 79841
E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
L0 = spla.splu(E.tocsc()).solve(C.T)
t0 = time.perf_counter()
be = api.backend()
t1 = time.perf_counter()
be.ensure_pencil(E, A)
be.ctx.sync()
t2 = time.perf_counter()
print(f"context create {1e3 * (t1 - t0):.1f} ms; ensure_pencil (symbolic + upload + slots) {1e3 * (t2 - t1):.1f} ms", flush=True)


class Obs:
    def __init__(self):
        self.t = time.perf_counter()

    def observe_gdre_step(self, t, X, K):
        now = time.perf_counter()
        print(f"  step to t={t}: {1e3 * (now - self.t):.1f} ms (rank X {X.rank()})", flush=True)
        self.t = now


sol = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, 0.01 * np.eye(6)), (4500.0, 4100.0)), api.Ros1(), dt=-100.0,
                observer=Obs())
be.ctx.sync()
print(f"total {1e3 * (time.perf_counter() - t0):.1f} ms")
