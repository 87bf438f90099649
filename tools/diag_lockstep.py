"""Diagnostic (GPU): lock-step comparison of the GPU ADI with the CPU oracle.
 1. free run of both; per-ADI-solve iteration counts and K errors
 2. at every Projection take_many of the GPU run: Ritz values of api.orth_restrict vs the oracle formula
    evaluated on the downloaded V blocks
 3. GPU run driven by the oracle's recorded shifts (Cyclic): residual histories and K
"""
import sys, warnings, json
import numpy as np, scipy.sparse.linalg as spla, scipy.linalg as sla
sys.path.insert(0, ".")
import dre_b200
from dre_b200 import api
from oracle import dre_oracle as O

warnings.simplefilter("ignore")
n = int(sys.argv[1]); nsteps = int(sys.argv[2])
E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
L0 = spla.splu(E.tocsc()).solve(C.T)
D0 = 0.01 * np.eye(6)
tspan = (4500.0, 4500.0 - 100 * nsteps)

class Rec:
    def __init__(s): s.runs = []
    def observe_gale_start(s, p, a): s.cur = dict(res=[], shifts=[])
    def observe_gale_metadata(s, d, mu): s.cur["shifts"].append(complex(mu))
    def observe_gale_step(s, i, X, res, rn): s.cur["res"].append((i, float(rn), X.rank(), res.rank()))
    def observe_gale_done(s, it, X, res, rn): s.cur["iters"] = it; s.runs.append(s.cur)

ro = Rec()
so = O.solve_gdre(O.GDREProblem(E, A, B, C, O.lowrank(L0, D0), tspan), O.Ros1(), dt=-100.0, observer=ro)

# --- instrument the GPU Projection iterator ---
orig_take_many = api.ProjectionShiftIterator.take_many
ritz_log = []
def take_many(self):
    lam = orig_take_many(self)
    Vh = [V.to_host() for V in self.Vs]
    N = np.concatenate(Vh, axis=1)
    Q = O.orth(N)
    Aop = self.prob.A
    a_, e_ = Aop.A.a, Aop.A.e
    As = (a_ * A + e_ * E)
    U = Aop.U.to_host(); Vt = Aop.Vt.to_host()
    At = Q.T @ (As @ Q) + (1.0 / Aop.alpha) * ((Q.T @ U) @ (Vt.T @ Q))
    Et = Q.T @ (E @ Q)
    ref = sla.eigvals(At, Et)
    if np.all(np.imag(ref) == 0): ref = np.real(ref)
    ref = O.safe_sort(O.stabilize_ritz_values(ref, "x"))
    s = sla.svdvals(N)
    entry = dict(k=N.shape[1], nsel_ref=Q.shape[1], n_gpu=len(lam), n_ref=len(ref), smax=float(s[0]), smin_sel=float(s[Q.shape[1]-1]) if Q.shape[1] else None)
    if len(lam) == len(ref):
        la, lb = np.array(lam, dtype=complex), np.array(ref, dtype=complex)
        entry["max_rel_diff"] = float(np.max(np.abs(la - lb) / np.abs(lb)))
        entry["first10_rel"] = [float(x) for x in (np.abs(la - lb) / np.abs(lb))[:10]]
    ritz_log.append(entry)
    return lam
api.ProjectionShiftIterator.take_many = take_many

rg = Rec()
sg = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, D0), tspan), api.Ros1(), dt=-100.0, observer=rg)
api.ProjectionShiftIterator.take_many = orig_take_many
print("== free run ==")
print("iters oracle", [r["iters"] for r in ro.runs], "gpu", [r["iters"] for r in rg.runs])
for i, (a, b) in enumerate(zip(so.K, sg.K)):
    print(f"K[{i}] relerr {np.linalg.norm(a-b)/np.linalg.norm(a):.3e}")
for r_o, r_g in zip(ro.runs, rg.runs):
    m = min(len(r_o["res"]), len(r_g["res"]))
    d = [abs(r_o["res"][k][1] - r_g["res"][k][1]) / r_o["res"][0][1] for k in range(m)]
    first = next((k for k in range(m) if d[k] > 1e-10), None)
    print(" res0", r_o["res"][0][1], r_g["res"][0][1], "ranks0", r_o["res"][0][2:], r_g["res"][0][2:], "maxdiff/res0 %.3e" % max(d), "first>1e-10 at", first,
          "final", r_o["res"][-1][1] / r_o["res"][0][1], r_g["res"][-1][1] / r_g["res"][0][1])
    ns = min(len(r_o["shifts"]), len(r_g["shifts"]))
    sd = [abs(r_o["shifts"][k] - r_g["shifts"][k]) / abs(r_o["shifts"][k]) for k in range(ns)]
    firsts = next((k for k in range(ns) if sd[k] > 1e-8), None)
    print("   shifts first rel diff >1e-8 at", firsts, "max", max(sd) if sd else None, "nshifts", len(r_o["shifts"]), len(r_g["shifts"]))
print("== ritz log ==")
for e in ritz_log: print(json.dumps(e))

# --- forced shifts: per ADI solve use the oracle's shift list ---
print("== forced-shift run ==")
class ForcedADI(api.ADI):
    pass
shift_lists = [r["shifts"] for r in ro.runs]
call = {"i": 0}
class ForcedShifts(api.Shifts.Strategy):
    def init(self, prob):
        lst = shift_lists[call["i"]]; call["i"] += 1
        return api._ListIterator([s.real if s.imag == 0 else s for s in lst] + [-1.0] * 5)
rf = Rec()
sf = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, D0), tspan), api.Ros1(api.ADI(shifts=ForcedShifts())), dt=-100.0, observer=rf)
print("iters oracle", [r["iters"] for r in ro.runs], "gpu forced", [r["iters"] for r in rf.runs])
for i, (a, b) in enumerate(zip(so.K, sf.K)):
    print(f"K[{i}] relerr {np.linalg.norm(a-b)/np.linalg.norm(a):.3e}")
for r_o, r_g in zip(ro.runs, rf.runs):
    m = min(len(r_o["res"]), len(r_g["res"]))
    d = [abs(r_o["res"][k][1] - r_g["res"][k][1]) / r_o["res"][0][1] for k in range(m)]
    dr = [abs(r_o["res"][k][1] - r_g["res"][k][1]) / r_o["res"][k][1] for k in range(m)]
    print(" maxdiff/res0 %.3e  max self-relative diff %.3e" % (max(d), max(dr)), "final", r_o["res"][-1][1] / r_o["res"][0][1], r_g["res"][-1][1] / r_g["res"][0][1])
