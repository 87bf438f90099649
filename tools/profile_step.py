"""Profiling workload (for ncu): n=79841 Rail pencil, one full Ros1 step as warm-up, then 12 ADI
iterations of the second step (includes one column compression of X)."""
import sys, time, warnings
import numpy as np, scipy.sparse.linalg as spla
sys.path.insert(0, ".")
import dre_b200
from dre_b200 import api

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 79841
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
L0 = spla.splu(E.tocsc()).solve(C.T)
sol = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, 0.01 * np.eye(6)), (4500.0, 4400.0)), api.Ros1(), dt=-100.0)
X = sol.X[-1]
be = api.backend()
Bd = api.DeviceMatrix.from_host(B); Ctd = api.DeviceMatrix.from_host(C.T)
alpha, L, D, BtLD, EtL, Kt = api._feedback(Bd, X)
tau = 100.0
F = api.LowRankUpdate(api.PencilCombo(1.0, -1 / (2 * tau)), -1.0, Bd, Kt)
G = api.hcat([Ctd, EtL]); S = api._dcat([np.eye(6), BtLD.T @ BtLD + D / tau])
R = api.compress_(api.LDLt([1.0], [G], [S]))
cache = api.init(api.GALEProblem(api.PencilCombo(0, 1), F, R), api.ADI(), initial_guess=X)
be.ctx.sync(); be.ctx.stats_reset(False)
import ctypes
_rt = ctypes.CDLL("libcudart.so.12")
_rt.cudaProfilerStart()   # with `ncu --profile-from-start off` only the region below is profiled
t0 = time.perf_counter()
print("PROFILE_REGION_BEGIN", flush=True)
for i in range(iters):
    api.step_(cache)
be.ctx.sync()
_rt.cudaProfilerStop()
print("PROFILE_REGION_END %.3f s for %d iterations; residual cols %d, rank X %d" % (time.perf_counter() - t0, iters, cache.residual.Ls[0].ncols, cache.X.rank()), flush=True)
print(be.ctx.stats())
import os
if os.environ.get("DRE_TIMELINE"):   # device-side timeline of the streams (context.cu: TlScope)
    ln = ctypes.c_int64()
    be.lib.dre_debug_export(be.h, b"timeline", None, 0, ctypes.byref(ln))
