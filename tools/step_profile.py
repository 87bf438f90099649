"""Whole-step profile: one warm Ros1 step, then one more step with every C-ABI call timed through a proxy
and the Python host code under cProfile (n=79841 by default)."""
import cProfile
import collections
import pstats
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
be.ctx.sync()
tl = TimedLib(be.lib)
be.lib = tl
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
sol2 = api.solve(api.GDREProblem(E, A, B, C, X, (4400.0, 4300.0)), api.Ros1(), dt=-100.0)
be.ctx.sync()
pr.disable()
wall = time.perf_counter() - t0
be.lib = tl._lib
tot_c = sum(tl.t.values())
print(f"one Ros1 step: wall {wall * 1e3:.1f} ms; inside C calls {tot_c * 1e3:.1f} ms; Python {(wall - tot_c) * 1e3:.1f} ms")
for k, v in sorted(tl.t.items(), key=lambda kv: -kv[1]):
    print(f"  {k:22s} calls {tl.c[k]:6d}  total {v * 1e3:9.2f} ms  avg {v / tl.c[k] * 1e6:9.1f} us")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
print(be.ctx.stats())
