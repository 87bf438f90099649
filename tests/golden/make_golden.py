"""Generates tests/golden/rail371_ros1.npz and rail371_ros2.npz from the CPU oracle (oracle/dre_oracle.py).

The reference is a Julia package and cannot be run in this image (no julia binary), so these are ORACLE
outputs -- the oracle itself is pinned against the reference's known-answer tests in tests/test_oracle_pins.py.
Stored per fixture: the saved feedback matrices K(t), and for every ADI solve the shift sequence the oracle
consumed, the residual norm after every iteration and the iteration count.  Consumers replay the shifts
(lock-step), which removes the round-off sensitivity of the Projection shift generation.

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dre_b200  # noqa: E402
from oracle import dre_oracle as O  # noqa: E402


class Recorder:
    def __init__(self):
        self.runs, self.cur = [], None

    def observe_gale_start(self, prob, alg):
        self.cur = {"shifts": [], "res": [], "iters": None}

    def observe_gale_metadata(self, desc, mu):
        self.cur["shifts"].append(complex(mu))

    def observe_gale_step(self, i, X, res, rn):
        self.cur["res"].append(float(rn))

    def observe_gale_done(self, iters, X, res, rn):
        self.cur["iters"] = iters
        self.runs.append(self.cur)


def generate(name, n, nsteps, ros, dt):
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    L0 = spla.splu(E.tocsc()).solve(C.T)
    D0 = 0.01 * np.eye(C.shape[0])
    rec = Recorder()
    alg = O.Ros1() if ros == 1 else O.Ros2()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = O.solve_gdre(O.GDREProblem(E, A, B, C, O.lowrank(L0, D0), (4500.0, 4500.0 + nsteps * dt)), alg, dt=dt,
                           observer=rec)
    shifts = np.concatenate([np.array(r["shifts"], dtype=complex) for r in rec.runs])
    res = np.concatenate([np.array(r["res"]) for r in rec.runs])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name)
    np.savez_compressed(out, n=n, nsteps=nsteps, ros=ros, dt=dt, K=np.stack(sol.K), t=np.array(sol.t),
                        shifts=shifts, nshifts=np.array([len(r["shifts"]) for r in rec.runs]),
                        res=res, nres=np.array([len(r["res"]) for r in rec.runs]),
                        iters=np.array([r["iters"] for r in rec.runs]))
    print(out, "K", np.stack(sol.K).shape, "ADI solves", len(rec.runs), "iterations", [r["iters"] for r in rec.runs])


if __name__ == "__main__":
    generate("rail371_ros1.npz", 371, 3, 1, -100.0)
    generate("rail371_ros2.npz", 371, 2, 2, -50.0)
