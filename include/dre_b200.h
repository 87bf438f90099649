/* dre_b200.h -- C ABI of libdre_b200.so: the B200-native LRSIF-ADI hot path of
 * DifferentialRiccatiEquations.jl (reference v0.5.5 under /root/reference, cited as file:line).
 *
 * Plain C: opaque handles, pointers and sizes only.  The Julia glue reaches CUDA
 * exclusively through `ccall` on these symbols; the Python host mirror (dre_b200.capi) binds the
 * same symbols with ctypes.
 *
 * Conventions
 *   - Every function returns int32 status: 0 ok; <0 argument/shape/state error; >0 CUDA / library /
 *     numerical failure (zero pivot, non-finite).  The message is available from dre_last_error().
 *     No C++ exception crosses the ABI.
 *   - Host arrays are column-major (Julia/Fortran order) and are only borrowed for the duration of
 *     the call.  Sparse matrices are CSC with 64-bit indices and index_base 0 or 1
 *     (SparseMatrixCSC{Float64,Int64} zero-copy).
 *   - Device "panels" are n x cols matrices (n = state dimension) owned by the context, addressed by
 *     an integer id; a dre_view is a column range of a panel (hcat without copies,
 *     cf. src/util/_hcat.jl:5-18).  Internally panels are stored row-major in the solver's
 *     fill-reducing row ordering; upload/download apply the permutation, so the ordering is invisible.
 *   - A context is not thread-safe; all work is queued on one CUDA stream and functions that do not
 *     return host data are asynchronous.
 */
#ifndef DRE_B200_H
#define DRE_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define DRE_API __attribute__((visibility("default")))
#else
#define DRE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DRE_OK 0
#define DRE_ERR_ARG (-1)
#define DRE_ERR_STATE (-2)
#define DRE_ERR_CUDA 1
#define DRE_ERR_NUMERIC 2
#define DRE_ERR_LIB 3

typedef struct dre_context dre_context;   /* device context (one GPU, one stream) */
typedef struct dre_symbolic dre_symbolic; /* host-only symbolic analysis */

typedef struct {
    int32_t id;    /* panel id from dre_mat_create */
    int32_t col0;  /* first column */
    int32_t ncols; /* number of columns (0 = empty view) */
} dre_view;

typedef struct {
    int64_t n;
    int64_t nnz_pattern; /* nnz(pattern(A) u pattern(E)), both triangles */
    int64_t nnz_L;       /* stored supernodal panel entries */
    int64_t sum_update_rows;
    double factor_flops; /* real flops of one LDL^T */
    int32_t nsupernodes, nlevels, max_front, max_supernode;
} dre_symbolic_info;

/* ---- errors ---- */
DRE_API const char* dre_last_error(const dre_context* ctx); /* ctx may be NULL: last error of ctx-less calls */
DRE_API const char* dre_version(void);

/* ---- host-only symbolic analysis (no GPU needed) ----
 * Replaces the analysis half of `factorize` (src/blocklinear/types.jl:41-42, backslash.jl:13):
 * nested-dissection ordering, supernode partition, elimination tree, level schedule, scatter maps
 * for a*A + (e+mu)*E.  E and A must be symmetric (values checked). */
DRE_API int32_t dre_symbolic_create(int64_t n, const int64_t* E_colptr, const int64_t* E_rowval, const double* E_nzval,
                            const int64_t* A_colptr, const int64_t* A_rowval, const double* A_nzval,
                            int32_t index_base, int32_t leaf_size, dre_symbolic** out);
DRE_API void dre_symbolic_destroy(dre_symbolic* s);
DRE_API int32_t dre_symbolic_get_info(const dre_symbolic* s, dre_symbolic_info* info);
/* diagnostic export of internal arrays (tests): `what` is one of the names below; copies at most
 * `cap` elements of 8 bytes (int64 or double) into buf and returns the full length in *len. */
DRE_API int32_t dre_symbolic_export(const dre_symbolic* s, const char* what, void* buf, int64_t cap, int64_t* len);

/* ---- context ---- */
DRE_API int32_t dre_create(int32_t device, dre_context** out);
DRE_API int32_t dre_destroy(dre_context* ctx);
DRE_API int32_t dre_sync(dre_context* ctx);
/* GALEProblem / GDREProblem matrices E, A (src/lyapunov/types.jl:10-16): runs the symbolic analysis
 * once and uploads the permuted pencil. */
DRE_API int32_t dre_set_pencil(dre_context* ctx, int64_t n, const int64_t* E_colptr, const int64_t* E_rowval,
                       const double* E_nzval, const int64_t* A_colptr, const int64_t* A_rowval,
                       const double* A_nzval, int32_t index_base);
DRE_API int32_t dre_get_symbolic_info(const dre_context* ctx, dre_symbolic_info* info);

/* ---- device panels (the outer factors L of LDLt, src/LDLt.jl:29-33; B, C', K') ---- */
DRE_API int32_t dre_mat_create(dre_context* ctx, int32_t cols, int32_t* id);
DRE_API int32_t dre_mat_free(dre_context* ctx, int32_t id);
DRE_API int32_t dre_mat_upload(dre_context* ctx, dre_view dst, const double* host, int64_t ld);
DRE_API int32_t dre_mat_download(dre_context* ctx, dre_view src, double* host, int64_t ld);
DRE_API int32_t dre_mat_copy(dre_context* ctx, dre_view dst, dre_view src);
/* Raw device address of a view for the host-side collective plumbing (NCCL through torch.distributed, or
 * CUDA.jl-free MPI/NCCL bindings on the Julia side): element (row, col) of the view is ptr[row * ld + col]
 * with rows in the solver's internal ordering (identical on every rank for the same pencil).  The caller must
 * dre_sync() before touching the memory from another stream and must not keep the pointer past dre_mat_free. */
DRE_API int32_t dre_mat_devptr(dre_context* ctx, dre_view v, void** ptr, int64_t* ld);
/* A second context on the same device used as a "compression lane": dre_set_dense_only gives it the row count of
 * the panels (no pencil, no sparse solver: only the dense low-rank algebra -- dre_ldlt_compress, dre_ldlt_norm,
 * dre_gemm_*, dre_mat_copy -- may be called on it), dre_mat_wrap registers memory owned by ANOTHER context (address
 * and leading dimension from dre_mat_devptr there) as a panel of this one without copying.  A wrapped panel is
 * never released by this context (dre_mat_free only forgets it); the caller keeps the owner alive and orders the
 * two contexts' streams (dre_sync on the producer before the consumer starts).  With both, the host side can run
 * compress!(X) of src/lyapunov/adi.jl:143-147 on its own stream and thread while the ADI iteration that produced
 * the terms carries on -- the following steps only append to X. */
DRE_API int32_t dre_set_dense_only(dre_context* ctx, int64_t n);
DRE_API int32_t dre_mat_wrap(dre_context* ctx, void* device_ptr, int64_t ld, int32_t cols, int32_t* id);
/* The context's main CUDA stream (a cudaStream_t).  Collectives and copies queued on it by the host side are
 * ordered with the library's own work, so the multi-GPU exchange needs no host synchronisation. */
DRE_API int32_t dre_get_stream(dre_context* ctx, void** stream);
/* Y = alpha*X + beta*Y (X may be an empty view when alpha == 0) */
DRE_API int32_t dre_mat_axpby(dre_context* ctx, double alpha, dre_view X, double beta, dre_view Y);

/* ---- Heuristic shifts on device ----
 * One Arnoldi step of compute_ritz_values (src/shifts/heuristic.jl:111-125) on device-resident vectors: the
 * twice-repeated modified Gram-Schmidt of w (n x 1, on entry the operator applied to the last basis vector; on exit
 * the unnormalised remainder) against the basis V (n x (j+1)), then vnext = (1/||w||) * w.  h receives the j+2
 * entries H[0..j, j] (sum of both sweeps, accumulated as the reference does) and H[j+1, j] = ||w||.  Synchronises
 * (j+2 doubles D2H).  DRE_ERR_NUMERIC when the remainder vanishes. */
DRE_API int32_t dre_arnoldi_orth(dre_context* ctx, dre_view V, dre_view w, dre_view vnext, double* h);

/* ---- products ----
 * op: 'E' or 'A' (the pencil is symmetric, so E' L / A' L of src/lyapunov/residual.jl:18,
 * lowrank_ros1.jl:42 use the same kernels).  Y = alpha*op*X + beta*Y.   */
DRE_API int32_t dre_spmm(dre_context* ctx, int32_t op, double alpha, dre_view X, double beta, dre_view Y);
/* out (X.ncols x Y.ncols, host column-major) = X' * Y   (B'L, Q'EQ, ...; synchronises) */
DRE_API int32_t dre_gemm_tn(dre_context* ctx, dre_view X, dre_view Y, double* out, int64_t ld);
/* Y = alpha * X * W + beta * Y with a small host matrix W (X.ncols x Y.ncols, column-major) */
DRE_API int32_t dre_gemm_nn(dre_context* ctx, double alpha, dre_view X, const double* W, int64_t ldw, double beta,
                    dre_view Y);

/* ---- the shifted closed-loop solve ----
 * Operator F = a*A + e*E + inv(alpha)*U*Vt'  (lr_update, src/LowRankUpdate.jl:38-39;
 * Ros1: a=1, e=-1/(2 tau), alpha=-1, U=B, Vt=K' -- src/riccati/lowrank_ros1.jl:39).
 * U, Vt are n x m panels (m may be 0: plain sparse operator, views with ncols = 0). */
DRE_API int32_t dre_set_operator(dre_context* ctx, double a, double e, double alpha, dre_view U, dre_view Vt);
/* Solve (F' + mu E') V = R   (src/lyapunov/adi.jl:156-159 real, :195-198 complex):
 * numeric supernodal LDL^T of a*A+(e+mu)*E, block forward/backward sweeps for [R, Vt] and the fused
 * Sherman-Morrison-Woodbury correction (src/blocklinear/sherman-morrison-woodbury.jl:10-45).
 * mu_im == 0: V1 = V (V2 ignored).  mu_im != 0: V1 = Re V, V2 = Im V.
 * SCOPE: E and A must be structurally and numerically symmetric (dre_set_pencil / dre_symbolic_create reject anything
 * else).  The LDL^T does NOT pivot (the reference's UMFPACK does): it is meant for the definite FEM pencils of this
 * path, whose shifted matrices a*A + (e+mu)*E are (complex-)symmetric with a definite real part; a zero or non-finite
 * pivot returns DRE_ERR_NUMERIC, a merely tiny pivot of an indefinite pencil does not -- verify one solve residual
 * (tests/c_abi_smoke.c does) before trusting such a pencil.  The low-rank update may have at most 32 columns
 * (m <= 32; the configurations of this path have 7-8), else DRE_ERR_ARG. */
DRE_API int32_t dre_shift_solve(dre_context* ctx, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2);
/* Hint: the shift the NEXT dre_adi_step / dre_shift_solve will use (the ADI shift buffer is known ahead,
 * src/shifts/helpers.jl:106-113).  Queues the numeric factorization for it on a side stream into the spare
 * factor slot, overlapping the current step's sweeps / Gram / compression work.  Purely a performance hint:
 * results are identical with or without it. */
DRE_API int32_t dre_prefactor(dre_context* ctx, double mu_re, double mu_im);
/* One ADI step (src/lyapunov/adi.jl:149-179 real, :181-225 complex pair):
 *   real:     V1 = (F'+mu E')^-1 R;                          R += -2 mu E' V1
 *   complex:  V  = (F'+mu E')^-1 R, d = Re mu / Im mu,
 *             V1 = sqrt2 (Re V + d Im V), V2 = sqrt(2 d^2+2) Im V;   R += -2 sqrt2 Re(mu) E' V1
 * R is updated in place; V1/V2 are freshly written panels. */
DRE_API int32_t dre_adi_step(dre_context* ctx, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2);
/* The solve half of dre_adi_step only (V1 / V2 as above, R untouched).  Used when the right-hand-side column
 * blocks are sharded over several GPUs: every rank solves its own column block, the blocks are exchanged
 * (NCCL all-gather driven by the host side) and the residual update R += c E' V1 runs on the full panel
 * through dre_spmm. */
DRE_API int32_t dre_adi_solve(dre_context* ctx, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2);

/* ---- low-rank algebra (src/LDLt.jl) ---- */
/* |alpha| * || L D L' ||_F  (norm(::LDLt), src/LDLt.jl:77-89).  D: k x k host column-major. */
DRE_API int32_t dre_ldlt_norm(dre_context* ctx, dre_view L, const double* D, int64_t ldd, double alpha, double* out);
/* The same norm in two halves, for a DIAGONAL core d[0..k): _begin queues the Gram product and the reduction on a
 * side stream of the library behind everything issued so far and returns at once; _end waits for it.  Between the
 * two the caller may queue work that only READS L (the host side uses the gap to start the shifted solve of the
 * next ADI iteration, src/lyapunov/adi.jl:97-128: the residual norm is only needed for the stopping test).  L must
 * not be written or freed before _end returns. */
DRE_API int32_t dre_ldlt_norm_begin(dre_context* ctx, dre_view L, const double* d, double alpha);
DRE_API int32_t dre_ldlt_norm_end(dre_context* ctx, double* out);
/* compress!(::LDLt) (src/LDLt.jl:204-225) of sum_i alphas[i] * Ls[i] * Ds[i] * Ls[i]':
 * orthonormal basis of hcat(Ls) (rank-revealing block Gram-Schmidt = orthf, :237-245), eigen-
 * decomposition of the projected core, truncation |lambda| >= tol_factor * max|lambda| * eps
 * (tol_factor = 100 in the reference).  Writes the new outer factor into `out` (capacity
 * out.ncols) and the new diagonal core into lam (capacity out.ncols); *newrank = kept columns. */
DRE_API int32_t dre_ldlt_compress(dre_context* ctx, int32_t nterms, const dre_view* Ls, const double* const* Ds,
                          const int64_t* ldds, const double* alphas, double tol_factor, dre_view out,
                          double* lam, int32_t* newrank);
/* The same compress! in three phases, for callers that receive the terms of X one by one (the compression lane of the
 * multi-GPU pipeline mode appends one ADI increment per step, src/lyapunov/adi.jl:170-176, and compresses every
 * `compression_interval` steps, :143-147): _begin reserves room for at most max_cols columns in total and resets the
 * job; every _add orthogonalises its terms against the basis built so far (so the work of an increment overlaps the
 * computation of the next one); _finish runs the core / eigen-decomposition / L <- Q V tail and ends the job.
 * dre_ldlt_compress(terms) == _begin(total columns) + _add(terms) + _finish.  One job per context at a time (a new
 * _begin or a dre_ldlt_compress call drops an open job; dre_rrqr, which shares the job's workspaces, refuses to run
 * while one is open); the panels passed to _add must stay alive and unchanged only for the duration of that call. */
DRE_API int32_t dre_compress_begin(dre_context* ctx, int32_t max_cols, double tol_factor);
/* Optional, between _begin and the first _add: a lower bound for the largest scaled column norm
 * max_j |alpha d_j|^(1/2) ||l_j|| the job will meet (for an orthonormal factor: sqrt(max |lambda|)).  Directions are
 * dropped relative to the largest column seen so far; a caller that adds the small terms first and the large one
 * last (two compression lanes) passes the scale of the large one here. */
DRE_API int32_t dre_compress_scale_hint(dre_context* ctx, double scale);
DRE_API int32_t dre_compress_add(dre_context* ctx, int32_t nterms, const dre_view* Ls, const double* const* Ds,
                         const int64_t* ldds, const double* alphas);
DRE_API int32_t dre_compress_finish(dre_context* ctx, dre_view out, double* lam, int32_t* newrank);
/* Hint for the NEXT dre_ldlt_compress / dre_compress_add only: the columns of `v` are orthonormal (they are the outer factor a
 * previous compress! produced, src/LDLt.jl:219-222).  If `v` is the first term of that call and its core is
 * diagonal, its columns are adopted as the first basis vectors without re-orthogonalisation.  Results agree
 * with the unhinted call to round-off. */
DRE_API int32_t dre_hint_orthonormal(dre_context* ctx, dre_view v);
/* Rank-revealing QR of N = hcat(views): N ~ Q * Rt' with orthonormal Q (n x rho) written to `Q`
 * (capacity Q.ncols) and Rt (ncols(N) x rho, host column-major, ld = ldr).  Directions whose
 * residual norm falls below max(drop_rel * max column norm, drop_abs) are discarded.  Building
 * block of orth (src/Stuff.jl:13-18) for the Projection shifts (src/shifts/projection.jl:54-61). */
DRE_API int32_t dre_rrqr(dre_context* ctx, int32_t nviews, const dre_view* views, double drop_rel, double drop_abs,
                 dre_view Q, double* Rt, int64_t ldr, int32_t* rho);

/* ---- diagnostics (tests) ----
 * Raw copy of the numeric factorization currently held: what = "L" (supernodal panels), "Linv" (inverse
 * unit-lower diagonal blocks), "dvec" (pivots), "U" (update matrices); elements are double (real shift) or
 * interleaved complex (complex shift).  Copies at most cap_bytes and returns the full size in *len_bytes. */
DRE_API int32_t dre_debug_export(dre_context* ctx, const char* what, void* buf, int64_t cap_bytes, int64_t* len_bytes);

/* The library's own symmetric eigensolver (Householder tridiagonalisation in one cooperative launch, implicit QL
 * on the host, rotations applied on the device; replaces LAPACK syevr of src/LDLt.jl:214) on a k x k symmetric
 * host matrix A (column-major): evals ascending, column j of evecs (column-major, k x k) = eigenvector j. */
DRE_API int32_t dre_debug_eigh(dre_context* ctx, int32_t k, const double* A, double* evals, double* evecs);

/* ---- timing / counters for bench.py ---- */
typedef struct {
    int64_t kernel_launches;    /* launches of this library's own kernels since the last reset */
    int64_t factorizations;     /* numeric factorizations */
    int64_t solves;             /* block solves (forward + backward sweep pairs) */
    int64_t spmms, grams, tallgemms;
    int64_t prefactors, prefactor_hits; /* side-stream factorizations queued / adopted by a solve */
    double ms_factor, ms_solve, ms_spmm, ms_gram, ms_tallgemm; /* CUDA-event times when timing is enabled */
    /* algorithmic work (SURVEY.md section 8d formulas), accumulated per call */
    double flops_factor, flops_gram, flops_tallgemm, flops_solve;
    double bytes_solve, bytes_spmm, bytes_gram, bytes_tallgemm;
} dre_stats;
/* CUDA-event stopwatch on the context's own stream (torch events cannot see this stream):
 * start records an event; stop records a second one, synchronises and returns the elapsed ms. */
DRE_API int32_t dre_timer_start(dre_context* ctx);
DRE_API int32_t dre_timer_stop(dre_context* ctx, double* ms);
DRE_API int32_t dre_stats_reset(dre_context* ctx, int32_t enable_event_timing);
DRE_API int32_t dre_stats_get(dre_context* ctx, dre_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* DRE_B200_H */
