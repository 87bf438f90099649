// Host-callable launchers of the hand-written sm_100a kernels (dense_kernels.cu, sparse_kernels.cu).
// Every launcher queues work on `st` and returns; it bumps *launches by the kernels it started.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace dre {

// ---------------- dense tall-skinny toolbox (real FP64, DMMA) ----------------
// partial[s][i*b + j] = sum over the s-th row range of  w[row] * X[row][i] * Y[row][j]
// followed by a deterministic reduction over s into out1 (overwrite, may be null) and out2
// (accumulate, may be null); both row-major with leading dimensions ld1 / ld2.
struct GramPlan {
    int nsplit;
    int64_t rows_per_split;
    size_t partial_elems;  // required workspace (doubles)
};
GramPlan gram_plan(int64_t n, int a, int b, int sm_count);
void launch_gram(const double* X, int64_t ldx, int a, const double* Y, int64_t ldy, int b, int64_t n,
                 const double* roww, double* partial, const GramPlan& plan, double* out1, int64_t ld1,
                 double* out2, int64_t ld2, cudaStream_t st, int64_t* launches);

// Y[n x b] = beta*Y + alpha * X[n x a] * W,  W row-major a x b (ldw) or, if w_trans, stored b x a.
void launch_tall_gemm(double alpha, const double* X, int64_t ldx, int a, const double* W, int64_t ldw,
                      int w_trans, double beta, double* Y, int64_t ldy, int b, int64_t n, cudaStream_t st,
                      int64_t* launches);

// dst[row][c] = src[row][c] * (colscale ? colscale[c] : 1)
void launch_copy_scale(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                       const double* colscale, cudaStream_t st, int64_t* launches);
// Y = alpha*X + beta*Y
void launch_axpby(double alpha, const double* X, int64_t ldx, double beta, double* Y, int64_t ldy, int64_t n,
                  int cols, cudaStream_t st, int64_t* launches);
// gather/scatter rows with permutation + transpose between host-style column-major staging and row-major panels
// dst_rowmajor[rowmap ? rowmap[i] : i][c] = src_colmajor[i + c*lds]
void launch_colmajor_to_panel(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                              const int32_t* iperm, cudaStream_t st, int64_t* launches);
void launch_panel_to_colmajor(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                              const int32_t* iperm, cudaStream_t st, int64_t* launches);

// Pivoted Cholesky selection on a pb x pb Gram matrix (pb <= 64), single CTA.
//   thr = max(drop2, rel2 * max diag);  selects pivots while the Schur diagonal >= thr.
//   Wsel (pb x 64 row-major, ld 64): Q = P * Wsel orthonormalises the selected columns.
//   info[0] = nsel, dinfo[0] = initial max diagonal, dinfo[1] = largest remaining diagonal.
void launch_pivchol(const double* G, int64_t ldg, int pb, double drop2, double rel2, double* Wsel, int32_t* info,
                    double* dinfo, cudaStream_t st, int64_t* launches);

// out[0] = sum_ij G[i][j]^2 t[i] t[j]   (r x r, diagonal core)
void launch_norm_diag(const double* G, int64_t ldg, int r, const double* t, double* out, cudaStream_t st,
                      int64_t* launches);
// Wt[j][k] = V[ids[j]*ldv + k]  (gather selected eigenvectors, rows of the row-major view)
void launch_gather_rows(double* Wt, int64_t ldw, const double* V, int64_t ldv, const int32_t* ids, int nsel,
                        int len, cudaStream_t st, int64_t* launches);

// ---------------- sparse: CSR SpMM ----------------
// Y[row][c] = beta*Y[row][c] + alpha * sum_j val[row,j] * X[col_j][c]      (row-major panels)
void launch_spmm(const int32_t* ptr, const int32_t* col, const double* val, int64_t n, double alpha,
                 const double* X, int64_t ldx, double beta, double* Y, int64_t ldy, int cols, cudaStream_t st,
                 int64_t* launches);

// ---------------- sparse: supernodal LDL^T + block solves ----------------
struct DevSymbolic {
    int64_t n;
    int32_t nsn, nlevels, nsubtrees;
    const int32_t* sn_first;
    const int64_t* sn_rowptr;
    const int32_t* sn_rows;
    const int32_t* relmap;
    const int32_t* child_ptr;
    const int32_t* child_idx;
    const int64_t* panel_off;
    const int64_t* upd_off;    // absolute offsets of the update matrices (bottom region, then top ping-pong)
    const int64_t* rhs_off;    // row offsets of the update vectors (cumulative over all supernodes)
    const int64_t* dblk_off;   // offset of the supernode's 32x32 block: pivots + inverse unit-lower factor
    const int32_t* st_ptr;     // bottom subtrees: supernode lists (ascending)
    const int32_t* st_sn;
    // assembly
    int64_t nasm;
    const int64_t* asm_dest;
    const double* asm_a;
    const double* asm_e;
};

template <class T>
void launch_assemble(const DevSymbolic& S, T* L, double a, T emu, cudaStream_t st, int64_t* launches);

// top levels: extend-add of the children's update matrices into the parents of one level
template <class T>
void launch_extend_add(const DevSymbolic& S, const int32_t* parents, int nparents, int gy, T* L, T* U,
                       cudaStream_t st, int64_t* launches);
// top levels: panel factorization of every front of a level; items: (J, slab) pairs
template <class T>
void launch_front(const DevSymbolic& S, const int2* items, int nitems, T* L, T* dblk, int32_t* errflag,
                  cudaStream_t st, int64_t* launches);
// top levels: Schur complement  U_J -= L21 D L21'  (lower triangle, 64x64 tiles); items: (J, ti, tj)
template <class T>
void launch_schur(const DevSymbolic& S, const int4* items, int nitems, const T* L, const T* dblk, T* U,
                  cudaStream_t st, int64_t* launches);
// bottom subtrees: one CTA factors a whole subtree
template <class T>
void launch_factor_subtrees(const DevSymbolic& S, T* L, T* dblk, T* U, int32_t* errflag, cudaStream_t st,
                            int64_t* launches);

// forward / backward sweeps on the row-major RHS block W (n x ldw), nrhs columns
template <class T>
void launch_fwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, const T* L, const T* dblk, T* W,
                      int64_t ldw, int nrhs, T* tbuf, cudaStream_t st, int64_t* launches);
template <class T>
void launch_bwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, const T* L, const T* dblk, T* W,
                      int64_t ldw, int nrhs, cudaStream_t st, int64_t* launches);
template <class T>
void launch_fwd_subtrees(const DevSymbolic& S, const T* L, const T* dblk, T* W, int64_t ldw, int nrhs, T* tbuf,
                         cudaStream_t st, int64_t* launches);
template <class T>
void launch_bwd_subtrees(const DevSymbolic& S, const T* L, const T* dblk, T* W, int64_t ldw, int nrhs,
                         cudaStream_t st, int64_t* launches);

// W[:, 0:r] = R, W[:, r:r+m] = Vt   (real -> T)
template <class T>
void launch_load_rhs(T* W, int64_t ldw, const double* R, int64_t ldr, int r, const double* Vt, int64_t ldv, int m,
                     int64_t n, cudaStream_t st, int64_t* launches);

// SMW core: BtW (m x ldb, T) = B' [Z, Y];  S = alpha I + BtW[:, r:r+m];  Sol = S^-1 BtW[:, 0:r]  (m x r, ld r)
template <class T>
void launch_smw_core(const T* BtW, int64_t ldb, int m, int r, double alpha, T* Sol, int32_t* errflag,
                     cudaStream_t st, int64_t* launches);

// V = Z - Y*Sol, written as
//   mode 0 (real):             V1 = V
//   mode 1 (complex raw):      V1 = Re V, V2 = Im V
//   mode 2 (complex ADI pair): V1 = sqrt2 (Re V + d Im V), V2 = sqrt(2 d^2 + 2) Im V
template <class T>
void launch_smw_apply(const T* W, int64_t ldw, int r, int m, const T* Sol, int mode, double d, double* V1,
                      int64_t ld1, double* V2, int64_t ld2, int64_t n, cudaStream_t st, int64_t* launches);

}  // namespace dre
