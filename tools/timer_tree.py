"""Host-side timer tree (the reference's @timeit_debug section names, api.enable_debug_timings) of consecutive Ros1
steps at n = 79841: where the driving thread spends a step.  Sections that end in a device synchronisation
(compress!, norm) carry the GPU time they wait for; launch-only sections carry launch overhead."""
import sys, time, warnings
import numpy as np, scipy.sparse.linalg as spla
sys.path.insert(0, ".")
from threadpoolctl import threadpool_limits
import dre_b200
from dre_b200 import api
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 79841
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
L0 = spla.splu(E.tocsc()).solve(C.T)
be = api.backend()
with threadpool_limits(limits=2):
    sol = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, 0.01 * np.eye(6)), (4500.0, 4300.0)), api.Ros1(), dt=-100.0)
    be.ctx.sync()
    api.enable_debug_timings(True)
    t0 = time.perf_counter()
    api.solve(api.GDREProblem(E, A, B, C, sol.X[-1], (4300.0, 4300.0 - 100.0 * nsteps)), api.Ros1(), dt=-100.0)
    be.ctx.sync()
    wall = time.perf_counter() - t0
print(f"{nsteps} steps: {1e3 * wall / nsteps:.1f} ms per step (host timers on)")
for k, v in sorted(api.TIMERS.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} {1e3 * v / nsteps:9.2f} ms/step  {api.COUNTS[k] / nsteps:8.1f} calls/step  {1e3 * v / api.COUNTS[k]:8.3f} ms/call")
