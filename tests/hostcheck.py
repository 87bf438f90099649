"""TEST INFRASTRUCTURE: NumPy emulation of the device algorithms, following the exported symbolic
arrays index by index (levels, extend-add maps, blocked left-looking LDL^T with the diagonal blocks
in a side array, multifrontal forward/backward sweeps with ping-pong update vectors).  It validates
the host-side symbolic analysis and the algorithm design on the CPU, where no GPU is available.
Never imported by the product."""
from __future__ import annotations

import numpy as np

NB = 32


class Sym:
    def __init__(self, sa):
        self._sa = sa
        for name in ("perm", "iperm", "sn_first", "sn_rowptr", "sn_rows", "sn_parent", "sn_level", "level_ptr",
                     "level_sn", "panel_off", "upd_off", "rhs_off", "child_ptr", "child_idx", "relmap", "asm_dest",
                     "asm_a", "asm_e", "csr_ptr", "csr_col", "csr_a", "csr_e"):
            setattr(self, name, sa.export(name))
        self.info = sa.info()
        self.n = self.info["n"]
        self.nsn = self.info["nsupernodes"]
        self.nlevels = self.info["nlevels"]

    def s(self, J):
        return int(self.sn_first[J + 1] - self.sn_first[J])

    def u(self, J):
        return int(self.sn_rowptr[J + 1] - self.sn_rowptr[J])


def factor(S: Sym, a, emu, dtype):
    """Returns (L storage, dblk dict) exactly as the kernels would produce them."""
    L = np.zeros(S.info["nnz_L"], dtype=dtype)
    np.add.at(L, S.asm_dest, a * S.asm_a + emu * S.asm_e)  # duplicates cannot occur; add.at == assignment
    dblk = {}
    U = [{}, {}]
    for l in range(S.nlevels):
        Ucur, Uprev = U[l & 1], U[(l + 1) & 1]
        Ucur.clear()
        sns = S.level_sn[S.level_ptr[l]:S.level_ptr[l + 1]]
        for J in sns:
            J = int(J)
            s, u = S.s(J), S.u(J)
            f = s + u
            P = L[S.panel_off[J]:S.panel_off[J] + f * s].reshape(s, f).T  # column-major f x s view
            UJ = np.zeros((u, u), dtype=dtype)
            # extend-add (lower triangles only)
            for ci in range(S.child_ptr[J], S.child_ptr[J + 1]):
                c = int(S.child_idx[ci])
                uc = S.u(c)
                rel = S.relmap[S.sn_rowptr[c]:S.sn_rowptr[c + 1]]
                Uc = Uprev[c]
                for j in range(uc):
                    pc = rel[j]
                    pr = rel[j:]
                    if pc < s:
                        P[pr, pc] += Uc[j:, j]
                    else:
                        UJ[pr - s, pc - s] += Uc[j:, j]
            # blocked left-looking LDL^T, diagonal blocks to the side array
            nblk = (s + NB - 1) // NB
            D = np.zeros((nblk, NB, NB), dtype=dtype)
            dvec = np.zeros(s, dtype=dtype)
            for step in range(nblk):
                k0 = step * NB
                nb = min(NB, s - k0)
                Lprev = P[:, :k0]
                Brows = P[k0:k0 + nb, :k0] * dvec[:k0]
                Cd = P[k0:k0 + nb, k0:k0 + nb] - Lprev[k0:k0 + nb] @ Brows.T
                Cs = P[k0 + nb:, k0:k0 + nb] - Lprev[k0 + nb:] @ Brows.T
                Ds = Cd.copy()
                for j in range(nb):
                    d = Ds[j, j]
                    assert d != 0
                    w = Ds[j + 1:, j].copy()
                    Ds[j + 1:, j] = w / d
                    for k in range(j + 1, nb):
                        Ds[k:, k] -= Ds[k:, j] * w[k - j - 1]
                Ld = np.tril(Ds, -1) + np.eye(nb)
                dd = np.diag(Ds).copy()
                # rows: Y = S Ld^-T, L = Y / d
                Y = np.linalg.solve(Ld, Cs.T).T if Cs.shape[0] else Cs
                P[k0 + nb:, k0:k0 + nb] = Y / dd
                blk = np.eye(NB, dtype=dtype)
                blk[:nb, :nb] = np.tril(Ds)
                D[step] = blk
                dvec[k0:k0 + nb] = dd
            dblk[J] = D
            # Schur complement (lower triangle)
            if u:
                L21 = P[s:, :]
                UJ -= np.tril((L21 * dvec) @ L21.T)
            Ucur[J] = UJ
    return L, dblk


def solve(S: Sym, L, dblk, B):
    """B: n x r in solver ordering (rows permuted).  Returns the solution in solver ordering."""
    W = B.astype(L.dtype).copy()
    T = [{}, {}]
    for l in range(S.nlevels):
        tcur, tprev = T[l & 1], T[(l + 1) & 1]
        tcur.clear()
        for J in S.level_sn[S.level_ptr[l]:S.level_ptr[l + 1]]:
            J = int(J)
            s, u = S.s(J), S.u(J)
            f = s + u
            first = int(S.sn_first[J])
            P = L[S.panel_off[J]:S.panel_off[J] + f * s].reshape(s, f).T
            tJ = np.zeros((u, W.shape[1]), dtype=L.dtype)
            for ci in range(S.child_ptr[J], S.child_ptr[J + 1]):
                c = int(S.child_idx[ci])
                rel = S.relmap[S.sn_rowptr[c]:S.sn_rowptr[c + 1]]
                tc = tprev[c]
                own = rel < s
                W[first + rel[own]] += tc[own]
                tJ[rel[~own] - s] += tc[~own]
            D = dblk[J]
            for jb in range(0, s, NB):
                nb = min(NB, s - jb)
                Ld = np.tril(D[jb // NB][:nb, :nb], -1) + np.eye(nb)
                xb = np.linalg.solve(Ld, W[first + jb:first + jb + nb])
                W[first + jb:first + jb + nb] = xb
                W[first + jb + nb:first + s] -= P[jb + nb:s, jb:jb + nb] @ xb
                tJ -= P[s:, jb:jb + nb] @ xb
            tcur[J] = tJ
    for l in range(S.nlevels - 1, -1, -1):
        for J in S.level_sn[S.level_ptr[l]:S.level_ptr[l + 1]]:
            J = int(J)
            s, u = S.s(J), S.u(J)
            f = s + u
            first = int(S.sn_first[J])
            P = L[S.panel_off[J]:S.panel_off[J] + f * s].reshape(s, f).T
            rows = S.sn_rows[S.sn_rowptr[J]:S.sn_rowptr[J + 1]]
            D = dblk[J]
            for jb in range(((s - 1) // NB) * NB, -1, -NB):
                nb = min(NB, s - jb)
                X_below = np.concatenate([W[first + jb + nb:first + s], W[rows]], axis=0)
                acc = P[jb + nb:, jb:jb + nb].T @ X_below
                blk = D[jb // NB][:nb, :nb]
                z = W[first + jb:first + jb + nb] / np.diag(blk)[:, None] - acc
                Ld = np.tril(blk, -1) + np.eye(nb)
                W[first + jb:first + jb + nb] = np.linalg.solve(Ld.T, z)
    return W


def permuted_csr(S: Sym, which):
    import scipy.sparse as sp

    vals = S.csr_a if which == "A" else S.csr_e
    return sp.csr_matrix((vals, S.csr_col, S.csr_ptr), shape=(S.n, S.n))
