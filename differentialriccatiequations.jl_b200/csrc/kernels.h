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
// waves > 0 overrides DRE_GRAM_WAVES: that many waves of shorter-lived CTAs (more row splits)
GramPlan gram_plan(int64_t n, int a, int b, int sm_count, int waves_override = 0);
void launch_gram(const double* X, int64_t ldx, int a, const double* Y, int64_t ldy, int b, int64_t n,
                 const double* roww, double* partial, const GramPlan& plan, double* out1, int64_t ld1,
                 double* out2, int64_t ld2, cudaStream_t st, int64_t* launches);

// Y[n x b] = beta*Y + alpha * X[n x a] * W,  W row-major a x b (ldw) or, if w_trans, stored b x a.
void launch_tall_gemm(double alpha, const double* X, int64_t ldx, int a, const double* W, int64_t ldw,
                      int w_trans, double beta, double* Y, int64_t ldy, int b, int64_t n, cudaStream_t st,
                      int64_t* launches);

// dst[row][c] = src[row][c] * (colscale ? colscale[c] : 1)
void launch_copy_scale(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                       const double* colscale, cudaStream_t st, int64_t* launches, const double* rowscale = nullptr);
// out[c] = sum_rows P[row][c]^2, cols <= 256; partial needs nblk*cols doubles
void launch_colnorm2(const double* P, int64_t ldp, int64_t n, int cols, double* partial, int nblk, double* out,
                     cudaStream_t st, int64_t* launches);
// dst[row][c] = w1[c]*src[row][c] + w2[c]*src[row][idx2[c]]
void launch_combine_cols(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                         const int32_t* idx2, const double* w1, const double* w2, cudaStream_t st, int64_t* launches);
// Y = alpha*X + beta*Y
// Arnoldi step of the Heuristic shifts (heuristic.jl:111-125): twice-repeated MGS of w against V[:, 0:nbasis] and
// vnext = w / ||w||, all on device; coef receives the 2*nbasis projection coefficients in order and then ||w||.
// partials: scratch of 2 * 296 doubles.
void launch_arnoldi_mgs(const double* V, int64_t ldv, int nbasis, double* w, int64_t ldw, double* vnext, int64_t ldvn,
                        int64_t n, double* partials, double* coef, cudaStream_t st, int64_t* launches);
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
//   Wsel (64 x 64 row-major, ld 64; all 64 rows are written, rows >= pb and columns >= nsel with zeros): Q = P * Wsel
//   orthonormalises the selected columns.
//   info[0] = nsel, dinfo[0] = initial max diagonal, dinfo[1] = largest remaining diagonal.
void launch_pivchol(const double* G, int64_t ldg, int pb, double drop2, double rel2, double* Wsel, int32_t* info,
                    double* dinfo, cudaStream_t st, int64_t* launches);

// out[0] = sum_ij G[i][j]^2 t[i] t[j]   (r x r, diagonal core)
void launch_norm_diag(const double* G, int64_t ldg, int r, const double* t, double* out, cudaStream_t st,
                      int64_t* launches);
// Wt[j][k] = V[ids[j]*ldv + k]  (gather selected eigenvectors, rows of the row-major view)
void launch_gather_rows(double* Wt, int64_t ldw, const double* V, int64_t ldv, const int32_t* ids, int nsel,
                        int len, cudaStream_t st, int64_t* launches);

// ---------------- small symmetric eigenproblem (eigensolver.cu) ----------------
// S (k x k symmetric, device, ld k) is overwritten: row j of the row-major view = eigenvector of the j-th smallest
// eigenvalue; h_evals (host) and d_evals (device) receive the ascending eigenvalues.  work: eig_sym_work_doubles(k)
// doubles of device scratch; grow_rots(ctx, bytes) returns a device buffer of at least `bytes` (rotations of the
// host QL iteration).  Synchronises st.  Returns 0 ok, 1 CUDA error, 2 no convergence.
size_t eig_sym_work_doubles(int k);
int eig_sym(double* S, int k, double* d_evals, double* h_evals, double* work, void* (*grow_rots)(void*, size_t),
            void* grow_ctx, int sm_count, cudaStream_t st, int64_t* launches);

// ---------------- sparse: CSR SpMM ----------------
extern int spmm_variant;   // 1 = k_spmm, 2 = k_spmm2 (indices broadcast by shuffles, two nonzeros in flight)
// Y[row][c] = beta*Y[row][c] + alpha * sum_j val[row,j] * X[col_j][c]      (row-major panels)
void launch_spmm(const int32_t* ptr, const int32_t* col, const double* val, int64_t n, double alpha,
                 const double* X, int64_t ldx, double beta, double* Y, int64_t ldy, int cols, cudaStream_t st,
                 int64_t* launches);

// ---------------- sparse: supernodal LDL^T + block solves ----------------
// Supernodes are dense and wide (a dissection leaf is one supernode of up to ~100 columns, separators up to
// SN_MAX columns).  Besides the panel L (f_J x s_J, column-major) the factorization stores the explicit
// inverse of the unit-lower diagonal block (s_J x s_J, column-major, ones on / zeros above the diagonal) and
// the pivots, so that every step of factorization and solves is a plain matrix product on the FP64 tensor
// cores -- no dependent chain inside a supernode.
constexpr int SN_MAX = 256;  // widest supernode the kernels accept (register accumulators of the sweeps)

struct DevSymbolic {
    int64_t n;
    int32_t nsn, nlevels;
    const int32_t* sn_first;
    const int64_t* sn_rowptr;
    const int32_t* sn_rows;
    const int32_t* relmap;
    const int32_t* child_ptr;
    const int32_t* child_idx;
    const int64_t* panel_off;
    const int64_t* linv_off;
    const int64_t* upd_off;
    const int64_t* rhs_off;    // row offsets of the update vectors (cumulative over all supernodes)
    // assembly
    int64_t nasm;
    const int64_t* asm_dest;
    const double* asm_a;
    const double* asm_e;
};

template <class T>
void launch_assemble(const DevSymbolic& S, T* L, double a, T emu, cudaStream_t st, int64_t* launches,
                     const double* prm = nullptr);
void launch_set_prm(double* prm, double a, double re, double im, cudaStream_t st, int64_t* launches);
// parents of one level gather their children's update matrices (grid.y = gy column classes)
template <class T>
void launch_extend_add(const DevSymbolic& S, const int32_t* parents, int nparents, int gy, T* L, T* U,
                       cudaStream_t st, int64_t* launches);
extern int diag_variant;      // launch_diag: 2 = k_diag2 (default), 1 = k_diag
extern int diag_narrow_min;   // launch_diag: levels with >= this many supernodes use 64-thread CTAs
// LDL^T of the s x s diagonal blocks of one level + inverse of the unit-lower factor (one CTA per supernode)
template <class T>
void launch_diag(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, T* L, T* Linv, T* dvec, int32_t* errflag,
                 cudaStream_t st, int64_t* launches);
// L21 = A21 Linv' D^-1; items: (J, 64-row slab)
template <class T>
void launch_l21(const DevSymbolic& S, const int2* items, int nitems, T* L, const T* Linv, const T* dvec,
                cudaStream_t st, int64_t* launches);
// U_J -= L21 D L21'  (lower triangle, 64x64 tiles); items: (J, ti, tj)
template <class T>
void launch_schur(const DevSymbolic& S, const int4* items, int nitems, const T* L, const T* dvec, T* U,
                  cudaStream_t st, int64_t* launches);

// where the forward sweep finds the right-hand side: columns [0, r) in R, [r, r+m) in Vt (real row-major panels)
struct RhsSource {
    const double* R;
    int64_t ldr;
    int r;
    const double* Vt;
    int64_t ldv;
};

// forward / backward sweeps of one level on the row-major RHS block W (n x ldw), nrhs columns.
// smax = widest supernode of the level (sizes the dynamic shared memory).
template <class T>
void launch_fwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, const T* L, const T* Linv,
                      T* W, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, cudaStream_t st, int64_t* launches);
template <class T>
void launch_bwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, const T* L, const T* Linv,
                      const T* dvec, T* W, int64_t ldw, int nrhs, cudaStream_t st, int64_t* launches);

// ---- row-split sweeps (sweep v2): see the comment block in sparse_kernels.cu ----
// M21 = L21 Linv in place; items: (J, 64-row slab) as for launch_l21, after launch_schur of the level
template <class T>
void launch_m21(const DevSymbolic& S, const int2* items, int nitems, T* L, const T* Linv, cudaStream_t st,
                int64_t* launches);
// items: (J, row block of 8*q strips) from schedule.h; q in {1, 3} forward, {1, 2} backward
template <class T>
void launch_fwd2_level(const DevSymbolic& S, const int2* items, int nitems, int q, int smax, const T* L, const T* Linv,
                       const T* dvec, T* Y, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, int has_children,
                       cudaStream_t st, int64_t* launches);
template <class T>
void launch_bwd2_level(const DevSymbolic& S, const int2* items, int nitems, int q, int smax, const T* L, const T* Linv,
                       const T* Y, T* W, int64_t ldw, int nrhs, cudaStream_t st, int64_t* launches);

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
