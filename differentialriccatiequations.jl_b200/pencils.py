"""Synthetic Rail-shaped / 3D-heat FEM pencils (input data, no solver logic).

The reference's tests and benchmarks run on the MORWiki ``SteelProfile`` model
(/root/reference/test/rail.jl:15, /root/reference/benchmark/benchmarks.jl:40-42), which is
downloaded at run time and is not available offline.  These generators build pencils of the same
shape: ``E`` SPD P1 mass matrix, ``A`` symmetric negative definite (diffusion + Robin boundary),
``B`` n x 7 boundary loads, ``C`` 6 x n temperature differences, ~7 nnz/row, with exactly the
SteelProfile node counts n in {371, 1357, 5177, 20209, 79841}.

Everything is deterministic given (n, seed); seeds are part of the returned ``meta`` dict.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

RAIL_SIZES = (371, 1357, 5177, 20209, 79841)


def _grid_shape(n: int) -> tuple[int, int]:
    nx = int(math.ceil(math.sqrt(n)))
    ny = int(math.ceil(n / nx))
    return nx, ny


def rail_pencil(n: int, *, m: int = 7, q: int = 6, seed: int = 0,
                kappa: float = 2.0e-4, robin: float = 5.0e-3, heat_cap: float = 1.0):
    """2D P1-FEM heat pencil on a structured triangulated grid trimmed to exactly ``n`` nodes.

    Returns ``E, A`` (scipy CSC, float64, symmetric), ``B`` (n x m dense), ``C`` (q x n dense) and a
    ``meta`` dict.  Nodes are numbered row by row; the trailing nodes of the last grid row (and every
    triangle touching them) are dropped so that the node count is exactly ``n``.
    """
    nx, ny = _grid_shape(n)
    h = 1.0 / (nx - 1)
    idx = np.arange(nx * ny).reshape(ny, nx)
    # two triangles per cell, diagonal from (i,j) to (i+1,j+1): 7-point pattern
    a = idx[:-1, :-1].ravel()
    b = idx[:-1, 1:].ravel()
    c = idx[1:, 1:].ravel()
    d = idx[1:, :-1].ravel()
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 0)
    tris = tris[(tris < n).all(1)]
    used = np.zeros(n, bool)
    used[tris.ravel()] = True
    if not used.all():  # a dangling node (cannot happen for the SteelProfile sizes, guarded anyway)
        raise ValueError(f"grid trimming left isolated nodes for n={n}")
    xy = np.stack([(np.arange(nx * ny) % nx) * h, (np.arange(nx * ny) // nx) * h], 1)[:n]

    p0, p1, p2 = xy[tris[:, 0]], xy[tris[:, 1]], xy[tris[:, 2]]
    area = 0.5 * np.abs((p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1])
                        - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1]))
    # P1 gradients
    bb = np.stack([p1[:, 1] - p2[:, 1], p2[:, 1] - p0[:, 1], p0[:, 1] - p1[:, 1]], 1)
    cc = np.stack([p2[:, 0] - p1[:, 0], p0[:, 0] - p2[:, 0], p1[:, 0] - p0[:, 0]], 1)
    Kloc = (bb[:, :, None] * bb[:, None, :] + cc[:, :, None] * cc[:, None, :]) / (4.0 * area)[:, None, None]
    Mloc = (area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
    rows = np.repeat(tris, 3, axis=1).ravel()
    cols = np.tile(tris, (1, 3)).ravel()
    Kst = sp.coo_matrix((Kloc.ravel(), (rows, cols)), shape=(n, n)).tocsc()
    Mss = sp.coo_matrix((Mloc.ravel(), (rows, cols)), shape=(n, n)).tocsc()

    # boundary edges = edges that belong to exactly one triangle
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]], 0)
    e.sort(axis=1)
    key = e[:, 0].astype(np.int64) * n + e[:, 1]
    uk, cnt = np.unique(key, return_counts=True)
    bk = uk[cnt == 1]
    be = np.stack([bk // n, bk % n], 1)
    elen = np.linalg.norm(xy[be[:, 0]] - xy[be[:, 1]], axis=1)
    # order boundary edges by angle of their midpoint around the centroid -> m segments
    mid = 0.5 * (xy[be[:, 0]] + xy[be[:, 1]])
    cen = xy.mean(0)
    ang = np.arctan2(mid[:, 1] - cen[1], mid[:, 0] - cen[0])
    order = np.argsort(ang, kind="stable")
    seg = np.empty(len(be), np.int64)
    seg[order] = (np.arange(len(be)) * m) // len(be)

    br = np.concatenate([be[:, 0], be[:, 0], be[:, 1], be[:, 1]])
    bc = np.concatenate([be[:, 0], be[:, 1], be[:, 0], be[:, 1]])
    bv = np.concatenate([elen / 3.0, elen / 6.0, elen / 6.0, elen / 3.0])
    Mbd = sp.coo_matrix((bv, (br, bc)), shape=(n, n)).tocsc()

    E = (heat_cap * Mss).tocsc()
    A = (-(kappa * Kst) - robin * Mbd).tocsc()
    E = ((E + E.T) * 0.5).tocsc()
    A = ((A + A.T) * 0.5).tocsc()
    E.sort_indices()
    A.sort_indices()

    B = np.zeros((n, m))
    for j in range(m):
        sel = seg == j
        np.add.at(B[:, j], be[sel, 0], robin * elen[sel] / 2.0)
        np.add.at(B[:, j], be[sel, 1], robin * elen[sel] / 2.0)
    rng = np.random.default_rng(seed)
    C = np.zeros((q, n))
    picks = rng.choice(n, size=2 * q, replace=False)
    for i in range(q):
        C[i, picks[2 * i]] = 1.0
        C[i, picks[2 * i + 1]] = -1.0
    meta = dict(kind="rail2d", n=n, nx=nx, ny=ny, m=m, q=q, seed=seed, kappa=kappa, robin=robin,
                heat_cap=heat_cap, nnz_E=int(E.nnz), nnz_A=int(A.nnz))
    return E, A, B, C, meta


def heat3d_pencil(nside: int, *, m: int = 8, q: int = 8, seed: int = 0,
                  kappa: float = 1.0e-3, robin: float = 1.0e-2):
    """3D finite-difference heat pencil on an ``nside^3`` grid (7-point Laplacian, lumped mass
    perturbed to a consistent-mass-like 7-point SPD matrix), m face inputs, q point outputs."""
    N = nside
    n = N ** 3
    h = 1.0 / (N + 1)
    e1 = np.ones(N)
    T = sp.diags([-e1[:-1], 2 * e1, -e1[:-1]], [-1, 0, 1]) / h ** 2
    Adj = sp.diags([e1[:-1], e1[:-1]], [-1, 1])
    I = sp.identity(N)
    Kst = sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)
    # 7-point SPD "mass": 2/3 on the diagonal, 1/18 towards each of the six neighbours; its eigenvalues
    # 2/3 + (1/9) sum_d cos(theta_d) lie in [1/3, 1] (interior row sums 1), so E is SPD and well conditioned
    Ms = ((2.0 / 3.0) * sp.identity(n)
          + (1.0 / 18.0) * (sp.kron(sp.kron(Adj, I), I) + sp.kron(sp.kron(I, Adj), I) + sp.kron(sp.kron(I, I), Adj)))
    E = (Ms * h ** 3).tocsc()
    A = (-(kappa * h ** 3) * Kst).tolil()
    idx = np.arange(n).reshape(N, N, N)
    faces = [idx[0].ravel(), idx[-1].ravel(), idx[:, 0].ravel(), idx[:, -1].ravel(),
             idx[:, :, 0].ravel(), idx[:, :, -1].ravel()]
    A = A.tocsc()
    dvec = np.zeros(n)
    for f in faces:
        dvec[f] += robin * h ** 2
    A = (A - sp.diags(dvec)).tocsc()
    B = np.zeros((n, m))
    for j in range(m):
        if j < 6:
            B[faces[j], j] = robin * h ** 2
        else:  # two interior patch heaters
            c0 = N // 3 if j == 6 else 2 * N // 3
            patch = idx[c0 - 1:c0 + 2, c0 - 1:c0 + 2, c0 - 1:c0 + 2].ravel()
            B[patch, j] = h ** 3
    rng = np.random.default_rng(seed)
    C = np.zeros((q, n))
    C[np.arange(q), rng.choice(n, size=q, replace=False)] = 1.0
    E.sort_indices()
    A.sort_indices()
    meta = dict(kind="heat3d", n=n, nside=N, m=m, q=q, seed=seed, kappa=kappa, robin=robin,
                nnz_E=int(E.nnz), nnz_A=int(A.nnz))
    return E, A, B, C, meta


def random_spd_pencil(n: int, *, seed: int = 0, density: float | None = None):
    """Small random symmetric pencil in the style of /root/reference/test/tiny_random.jl:60-70:
    ``E = S + S' + n I`` (SPD), ``A = S' + S'^T - n I`` (negative definite)."""
    rng = np.random.default_rng(seed)
    density = density if density is not None else 1.0 / n
    S = sp.random(n, n, density=density, random_state=rng, format="csc")
    E = (S + S.T + n * sp.identity(n)).tocsc()
    S2 = sp.random(n, n, density=density, random_state=rng, format="csc")
    A = (S2 + S2.T - n * sp.identity(n)).tocsc()
    E.sort_indices()
    A.sort_indices()
    return E, A
