"""k_tridiag grid-size sweep (DRE_EIG_GRID) on indefinite cores of the sizes compress! meets; prints the phase times
of the in-tree eigensolver (DRE_EIG_DEBUG)."""
import os, subprocess, sys
if len(sys.argv) > 1:
    import numpy as np
    sys.path.insert(0, ".")
    import dre_b200
    from dre_b200 import api
    E, A, B, C, _ = dre_b200.pencils.rail_pencil(371)
    api.upload_pencil(E, A)
    ctx = api.backend().ctx
    for k in (420, 690):
        rng = np.random.default_rng(k)
        Q, _ = np.linalg.qr(rng.standard_normal((k, k)))
        S = (Q * (np.logspace(0, -14, k) * rng.choice([-1.0, 1.0], k))) @ Q.T
        S = 0.5 * (S + S.T)
        for _ in range(3):
            w, V = ctx.debug_eigh(S)
    sys.exit(0)
for g in (8, 16, 32, 48, 64, 96, 148):
    env = dict(os.environ, DRE_EIG_GRID=str(g), DRE_EIG_DEBUG="1")
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    lines = [l for l in out.stderr.splitlines() if "dre eig" in l]
    print("grid", g)
    for l in lines[2::3]:
        print("   ", l)
