"""Per-step wall times of consecutive Ros1 steps (n=79841) in one process: run-to-run noise check."""
import sys, time, warnings
import numpy as np, scipy.sparse.linalg as spla
sys.path.insert(0, ".")
from threadpoolctl import threadpool_limits
import dre_b200
from dre_b200 import api
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 79841
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
L0 = spla.splu(E.tocsc()).solve(C.T)
be = api.backend()

class Obs:
    def __init__(self):
        self.t = time.perf_counter(); self.k = 0; self.it = []; self.tc = 0.0; self.ts = 0.0
    def observe_gale_start(self, prob, alg):
        self.ti = time.perf_counter(); self.iter_ms = []
    def observe_gale_step(self, i, X, res, rn):
        now = time.perf_counter(); self.iter_ms.append((1e3 * (now - self.ti), i)); self.ti = now
    def observe_gale_done(self, iters, X, res, rn):
        self.it.append(iters)
        top = sorted(self.iter_ms, reverse=True)[:4]
        print("     slowest ADI iterations (ms, index):", [(round(a, 1), b) for a, b in top], "median", round(sorted(a for a, _ in self.iter_ms)[len(self.iter_ms) // 2], 2), flush=True)
    def observe_gdre_step(self, t, X, K):
        be.ctx.sync()
        now = time.perf_counter()
        st = be.ctx.stats()
        print(f"step {self.k:2d} t={t:7.1f}: {1e3 * (now - self.t):8.1f} ms  rank {X.rank():4d} iters {self.it[-1:] } launches {st['kernel_launches']}", flush=True)
        be.ctx.stats_reset(False)
        self.t = now; self.k += 1

with threadpool_limits(limits=2):
    api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, 0.01 * np.eye(6)), (4500.0, 4500.0 - 100.0 * nsteps)), api.Ros1(), dt=-100.0, observer=Obs())
