"""TEST INFRASTRUCTURE: NumPy emulation of the device algorithms, following the exported symbolic
arrays index by index (height levels, extend-add maps, dense supernodes with the explicit inverse of
the unit-lower diagonal block, multifrontal forward/backward sweeps as plain matrix products).  It validates
the host-side symbolic analysis and the algorithm design on the CPU, where no GPU is available.
Never imported by the product."""
from __future__ import annotations

import numpy as np

NB = 32


class Sym:
    def __init__(self, sa):
        self._sa = sa
        for name in ("perm", "iperm", "sn_first", "sn_rowptr", "sn_rows", "sn_parent", "sn_level", "level_ptr",
                     "level_sn", "panel_off", "linv_off", "upd_off", "rhs_off", "child_ptr", "child_idx", "relmap", "asm_dest",
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


def factor(S: Sym, a, emu, dtype, m21=False):
    """Returns (L storage, Linv dict, dvec) exactly as the kernels produce them: per level (leaves first)
    extend-add, LDL^T of the dense s x s diagonal block, explicit inverse of its unit-lower factor,
    L21 = A21 Linv' D^-1, U -= L21 D L21'.  m21: the panels end up holding M21 = L21 Linv (row-split sweeps)."""
    L = np.zeros(S.info["nnz_L"], dtype=dtype)
    np.add.at(L, S.asm_dest, a * S.asm_a + emu * S.asm_e)  # duplicates cannot occur; add.at == assignment
    Linv, U = {}, {}
    dvec = np.zeros(S.n, dtype=dtype)
    for l in range(S.nlevels):
        for J in S.level_sn[S.level_ptr[l]:S.level_ptr[l + 1]]:
            J = int(J)
            s, u = S.s(J), S.u(J)
            f = s + u
            first = int(S.sn_first[J])
            P = L[S.panel_off[J]:S.panel_off[J] + f * s].reshape(s, f).T  # column-major f x s view
            UJ = np.zeros((u, u), dtype=dtype)
            for ci in range(S.child_ptr[J], S.child_ptr[J + 1]):  # extend-add (lower triangles only)
                c = int(S.child_idx[ci])
                assert S.sn_level[c] < l
                uc = S.u(c)
                rel = S.relmap[S.sn_rowptr[c]:S.sn_rowptr[c + 1]]
                Uc = U[c]
                for j in range(uc):
                    pc = rel[j]
                    pr = rel[j:]
                    if pc < s:
                        P[pr, pc] += Uc[j:, j]
                    else:
                        UJ[pr - s, pc - s] += Uc[j:, j]
            # right-looking LDL^T of the diagonal block, 32 columns at a time (k_diag)
            Dg = np.tril(P[:s, :s]).copy()
            d = np.zeros(s, dtype=dtype)
            for jb in range(0, s, NB):
                nb = min(NB, s - jb)
                blk = Dg[jb:jb + nb, jb:jb + nb]
                for j in range(nb):
                    dj = blk[j, j]
                    assert dj != 0
                    w = blk[j + 1:, j].copy()
                    blk[j + 1:, j] = w / dj
                    for k in range(j + 1, nb):
                        blk[k:, k] -= blk[k:, j] * w[k - j - 1]
                dd = np.diag(blk).copy()
                d[jb:jb + nb] = dd
                Lbb = np.tril(blk, -1) + np.eye(nb)
                Li = np.linalg.inv(Lbb)
                below = Dg[jb + nb:, jb:jb + nb]
                below[:] = (below @ Li.T) / dd
                Dg[jb + nb:, jb + nb:] -= np.tril((below * dd) @ below.T)
            L11 = np.tril(Dg, -1) + np.eye(s)
            Linv[J] = np.linalg.inv(L11)
            dvec[first:first + s] = d
            P[:s, :s] = np.tril(Dg)
            if u:
                P[s:, :] = (P[s:, :] @ Linv[J].T) / d
                UJ -= np.tril((P[s:, :] * d) @ P[s:, :].T)
                if m21:
                    P[s:, :] = P[s:, :] @ Linv[J]
            U[J] = UJ
    return L, Linv, dvec


def solve(S: Sym, L, Linv, dvec, B):
    """B: n x r in solver ordering (rows permuted).  Returns the solution in solver ordering."""
    W = B.astype(L.dtype).copy()
    T = {}
    for l in range(S.nlevels):
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
                tc = T[c]
                own = rel < s
                W[first + rel[own]] += tc[own]
                tJ[rel[~own] - s] += tc[~own]
            y = Linv[J] @ W[first:first + s]
            W[first:first + s] = y
            T[J] = tJ - P[s:, :] @ y
    for l in range(S.nlevels - 1, -1, -1):
        for J in S.level_sn[S.level_ptr[l]:S.level_ptr[l + 1]]:
            J = int(J)
            s, u = S.s(J), S.u(J)
            f = s + u
            first = int(S.sn_first[J])
            P = L[S.panel_off[J]:S.panel_off[J] + f * s].reshape(s, f).T
            rows = S.sn_rows[S.sn_rowptr[J]:S.sn_rowptr[J + 1]]
            z = W[first:first + s] / dvec[first:first + s, None] - P[s:, :].T @ W[rows]
            W[first:first + s] = Linv[J].T @ z
    return W


def permuted_csr(S: Sym, which):
    import scipy.sparse as sp

    vals = S.csr_a if which == "A" else S.csr_e
    return sp.csr_matrix((vals, S.csr_col, S.csr_ptr), shape=(S.n, S.n))
