"""Where does the wall clock of an ADI iteration go?  Runs ADI iterations of the second Ros1 step at
n=79841 (same state as tools/profile_step.py) with every C-ABI call timed through a proxy (wall time
inside the call, which includes any stream synchronisation the call performs) and prints the split
C calls / Python host code, plus the library's own per-class GPU event times from a second pass."""
import collections
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
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30


class TimedLib:
    def __init__(self, lib):
        self._lib = lib
        self.t = collections.defaultdict(float)
        self.c = collections.defaultdict(int)

    def __getattr__(self, name):
        fn = getattr(self._lib, name)

        def wrapped(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            self.t[name] += time.perf_counter() - t0
            self.c[name] += 1
            return r

        return wrapped


E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
L0 = spla.splu(E.tocsc()).solve(C.T)
sol = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, 0.01 * np.eye(6)), (4500.0, 4400.0)), api.Ros1(),
                dt=-100.0)
X = sol.X[-1]
be = api.backend()
Bd = api.DeviceMatrix.from_host(B)
Ctd = api.DeviceMatrix.from_host(C.T)
alpha, L, D, BtLD, EtL, Kt = api._feedback(Bd, X)
tau = 100.0
F = api.LowRankUpdate(api.PencilCombo(1.0, -1 / (2 * tau)), -1.0, Bd, Kt)
G = api.hcat([Ctd, EtL])
S = api._dcat([np.eye(6), BtLD.T @ BtLD + D / tau])
R = api.compress_(api.LDLt([1.0], [G], [S]))
cache = api.init(api.GALEProblem(api.PencilCombo(0, 1), F, R), api.ADI(), initial_guess=X)
be.ctx.sync()
tl = TimedLib(be.lib)
be.lib = tl
t0 = time.perf_counter()
for i in range(iters):
    api.step_(cache)
be.ctx.sync()
wall = time.perf_counter() - t0
be.lib = tl._lib
tot_c = sum(tl.t.values())
print(f"{iters} ADI iterations: wall {wall * 1e3:.1f} ms = {wall / iters * 1e3:.2f} ms/iter; inside C calls "
      f"{tot_c * 1e3:.1f} ms; Python host code {(wall - tot_c) * 1e3:.1f} ms")
for k, v in sorted(tl.t.items(), key=lambda kv: -kv[1]):
    print(f"  {k:22s} calls {tl.c[k]:6d}  total {v * 1e3:9.2f} ms  avg {v / tl.c[k] * 1e6:9.1f} us")
print("residual cols", cache.residual.Ls[0].ncols, "rank X", cache.X.rank())
# second pass with the library's event timing (serialises every launch group; the absolute wall is not comparable)
be.ctx.stats_reset(True)
t0 = time.perf_counter()
for i in range(iters):
    api.step_(cache)
be.ctx.sync()
wall2 = time.perf_counter() - t0
st = be.ctx.stats()
print(f"event-timed pass: wall {wall2 * 1e3:.1f} ms")
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()})
