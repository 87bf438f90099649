"""Golden-vector tests (tests/golden/*.npz, generated from the oracle by tests/golden/make_golden.py).

CPU: the oracle, replaying the stored shift sequences, reproduces the stored K(t), residual norms and iteration
counts (guards the oracle against drift).  GPU: the CUDA path, replaying the same shifts, matches the stored
vectors within the north-star tolerances: K(t) 1e-8 relative, ADI residual norms 1e-10 relative, identical
iteration counts.  (File name: runs last, after the kernel and parity tests.)"""
import os
import warnings

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import dre_b200
from oracle import dre_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ["rail371_ros1.npz", "rail371_ros2.npz"]


def _load(name):
    g = np.load(os.path.join(GOLDEN, name))
    so = np.concatenate([[0], np.cumsum(g["nshifts"])])
    ro = np.concatenate([[0], np.cumsum(g["nres"])])
    shifts = [list(g["shifts"][so[i]:so[i + 1]]) for i in range(len(g["nshifts"]))]
    res = [g["res"][ro[i]:ro[i + 1]] for i in range(len(g["nres"]))]
    return g, shifts, res


class _Rec:
    def __init__(self):
        self.iters, self.res = [], [[]]

    def observe_gale_step(self, i, X, res, rn):
        self.res[-1].append(float(rn))

    def observe_gale_done(self, iters, X, res, rn):
        self.iters.append(int(iters))
        self.res.append([])


def _problem(n):
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(n)
    L0 = spla.splu(E.tocsc()).solve(C.T)
    return E, A, B, C, L0, 0.01 * np.eye(C.shape[0])


def _check(g, shifts, res, K, rec, ktol, rtol):
    assert len(K) == g["K"].shape[0]
    for Kg, Kref in zip(K, g["K"]):
        assert np.linalg.norm(np.asarray(Kg) - Kref) <= ktol * np.linalg.norm(Kref)
    assert rec.iters == [int(i) for i in g["iters"]]
    for a, b in zip(res, rec.res):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and np.max(np.abs(a - b) / a) <= rtol


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_golden(name):
    g, shifts, res = _load(name)

    class Replay(O.Strategy):
        def __init__(self):
            self.i = 0

        def init(self, prob):
            lst = shifts[self.i]
            self.i += 1
            return O._ListIterator([z.real if z.imag == 0 else z for z in lst] + [-1.0] * 4)

    n, nsteps, ros, dt = int(g["n"]), int(g["nsteps"]), int(g["ros"]), float(g["dt"])
    E, A, B, C, L0, D0 = _problem(n)
    rec = _Rec()
    adi = O.ADI(shifts=Replay())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = O.solve_gdre(O.GDREProblem(E, A, B, C, O.lowrank(L0, D0), (4500.0, 4500.0 + nsteps * dt)),
                           O.Ros1(adi) if ros == 1 else O.Ros2(adi), dt=dt, observer=rec)
    _check(g, shifts, res, sol.K, rec, 1e-9, 1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_matches_golden(name):
    from dre_b200 import api

    g, shifts, res = _load(name)

    class Replay(api.Shifts.Strategy):
        def __init__(self):
            self.i = 0

        def init(self, prob):
            lst = shifts[self.i]
            self.i += 1
            return api._ListIterator([z.real if z.imag == 0 else z for z in lst] + [-1.0] * 4)

    n, nsteps, ros, dt = int(g["n"]), int(g["nsteps"]), int(g["ros"]), float(g["dt"])
    E, A, B, C, L0, D0 = _problem(n)
    rec = _Rec()
    adi = api.ADI(shifts=Replay())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sol = api.solve(api.GDREProblem(E, A, B, C, api.lowrank(L0, D0), (4500.0, 4500.0 + nsteps * dt)),
                        api.Ros1(adi) if ros == 1 else api.Ros2(adi), dt=dt, observer=rec)
    _check(g, shifts, res, sol.K, rec, 1e-8, 1e-10)
