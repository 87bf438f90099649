import sys, os
sys.path.insert(0, ".")
import numpy as np
import dre_b200
from dre_b200 import api
os.environ["DRE_EIG_DEBUG"] = "1"
ctx = api.backend().ctx
for grid in ["", "1", "8", "17"]:
    if grid: os.environ["DRE_EIG_GRID"] = grid
    for k in [64, 100, 129, 130, 200, 300]:
        rng = np.random.default_rng(k)
        Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        S = (Q * np.logspace(0, -16, k) * rng.choice([-1.0, 1.0], k)) @ Q.T
        S = 0.5 * (S + S.T)
        try:
            w, V = ctx.debug_eigh(S)
            print("grid", grid or "auto", "k", k, "orth %.1e res %.1e ev %.1e" % (np.linalg.norm(V.T @ V - np.eye(k)), np.linalg.norm(S @ V - V * w), np.max(np.abs(w - np.linalg.eigvalsh(S)))), flush=True)
        except Exception as ex:
            print("grid", grid or "auto", "k", k, "FAILED", ex, flush=True)
