"""GPU diagnostic: accuracy and phase timings of the in-tree symmetric eigensolver (DRE_EIG_DEBUG=1 prints the phases)."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

os.environ["DRE_EIG_DEBUG"] = "1"
import dre_b200
from dre_b200 import api

ctx = api.backend().ctx
for grid in ["", "8", "16", "32", "64"]:
    if grid:
        os.environ["DRE_EIG_GRID"] = grid
    for k in [250, 330, 500, 700]:
        rng = np.random.default_rng(k)
        Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        S = (Q * np.logspace(0, -16, k) * rng.choice([-1.0, 1.0], k)) @ Q.T
        S = 0.5 * (S + S.T)
        for rep in range(2):
            w, V = ctx.debug_eigh(S)
        print("grid", grid or "auto", "k", k, "orth %.1e res %.1e ev %.1e" % (
            np.linalg.norm(V.T @ V - np.eye(k)), np.linalg.norm(S @ V - V * w), np.max(np.abs(w - np.linalg.eigvalsh(S)))),
            flush=True)
