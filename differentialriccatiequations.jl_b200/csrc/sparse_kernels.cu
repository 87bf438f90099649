// Per-shift numeric supernodal LDL^T (no pivoting; real symmetric definite or complex symmetric)
// and multifrontal block solves with many right-hand sides.
//
// Replaces, for the hot path, what the reference gets from SuiteSparse through
//   factorize(A' + mu E')   src/blocklinear/backslash.jl:13, src/blocklinear/types.jl:41-42
//   F \ R                    src/blocklinear/backslash.jl:19
//   SMW correction           src/blocklinear/sherman-morrison-woodbury.jl:10-45
// T = double for real shifts, T = cplx (complex SYMMETRIC, no conjugation) for complex shifts --
// the complex pair path the reference lists as broken on GPU (README.md:174-176).
//
// Structure (from csrc/symbolic.cpp): every supernode has at most 32 columns (wider dissection blocks
// are chains).  The tree is cut into
//   * bottom subtrees (<= a few hundred columns each): ONE CTA walks a whole subtree in elimination
//     order, so the thousands of tiny fronts of the lower levels cost one launch per phase;
//   * top supernodes, processed level by level (one launch per level and phase).
// Storage: panel_J = f_J x s_J column-major (ld f_J), L21 below the supernode's own rows; the pivots and
// the INVERSE of the unit-lower diagonal block live in a side array (32x32 per supernode) so that the
// block solves are plain multiplications.  Update matrices: one region per bottom supernode plus two
// ping-pong regions for the top levels.  Update vectors of the solves: one region per supernode,
// column-major (u_J contiguous per right-hand side).
#include <algorithm>

#include "kernels.h"

namespace dre {

constexpr int NB = 32;     // maximum supernode width
constexpr int SLAB = 96;   // L21 rows handled per slab

__device__ __forceinline__ int sn_s(const DevSymbolic& S, int J) { return S.sn_first[J + 1] - S.sn_first[J]; }
__device__ __forceinline__ int sn_u(const DevSymbolic& S, int J) { return (int)(S.sn_rowptr[J + 1] - S.sn_rowptr[J]); }

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx shfl(cplx v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// ------------------------------------------------------------------------------------------
// assembly: L[dest] = a*A + (e+mu)*E on the lower-triangular union pattern
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void k_assemble(int64_t nasm, const int64_t* __restrict__ dest, const double* __restrict__ va,
                           const double* __restrict__ ve, T* __restrict__ L, double a, T emu) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nasm; i += (int64_t)gridDim.x * blockDim.x) {
        T v;
        from_real(a * va[i], v);
        L[dest[i]] = add(v, mul(ve[i], emu));
    }
}

template <class T>
void launch_assemble(const DevSymbolic& S, T* L, double a, T emu, cudaStream_t st, int64_t* launches) {
    int blocks = (int)std::min<int64_t>((S.nasm + 255) / 256, 148 * 8);
    if (blocks < 1) blocks = 1;
    k_assemble<T><<<blocks, 256, 0, st>>>(S.nasm, S.asm_dest, S.asm_a, S.asm_e, L, a, emu);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// device bodies shared by the per-level (top) kernels and the subtree (bottom) kernels
// ------------------------------------------------------------------------------------------

// extend-add: supernode J gathers the update matrices of its children through the relative index maps.
// Deterministic: parent column pc is owned by column class (pc % gy == by); children in fixed order.
template <class T>
__device__ __forceinline__ void ea_body(const DevSymbolic& S, int J, int by, int gy, T* L, T* U) {
    const int sJ = sn_s(S, J), uJ = sn_u(S, J), fJ = sJ + uJ;
    T* P = L + S.panel_off[J];
    T* UJ = U + S.upd_off[J];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int ci = S.child_ptr[J]; ci < S.child_ptr[J + 1]; ++ci) {
        const int c = S.child_idx[ci];
        const int uc = sn_u(S, c);
        const int32_t* rel = S.relmap + S.sn_rowptr[c];
        const T* Uc = U + S.upd_off[c];
        for (int j = warp; j < uc; j += nwarps) {
            const int pc = rel[j];
            if (pc % gy != by) continue;
            for (int i = j + lane; i < uc; i += 32) {
                const int pr = rel[i];
                const T val = Uc[(int64_t)i + (int64_t)j * uc];
                T* t = (pc < sJ) ? (P + ((int64_t)pr + (int64_t)pc * fJ))
                                 : (UJ + ((int64_t)(pr - sJ) + (int64_t)(pc - sJ) * uJ));
                *t = add(*t, val);
            }
        }
        __syncthreads();
    }
}

// LDL^T of the (identity-padded) 32x32 diagonal block by ONE warp: lane i holds row i in registers.
template <class T>
__device__ __forceinline__ void warp_ldlt32(T (&a)[NB], int lane, int32_t* errflag) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const T d = shfl(a[j], j);
        if (lane == j && is_bad(d)) atomicExch(errflag, 1);
        const T w = a[j];                 // unscaled column entry of this row
        const T l = mul(w, recip(d));
#pragma unroll
        for (int k = j + 1; k < NB; ++k) {
            const T wk = shfl(w, k);      // w of row k = A[k][j]
            if (lane >= k) a[k] = sub(a[k], mul(l, wk));
        }
        if (lane > j) a[j] = l;
    }
}

// phase A: factor the diagonal block of front J (all threads call; warp 0 works), optionally store it.
// Ds: strictly lower = L_d, diagonal = pivots.  Li: strictly lower part of the inverse of the unit factor.
template <class T>
__device__ __forceinline__ void front_diag(const DevSymbolic& S, int J, T* L, T* dblk, int32_t* errflag,
                                           T (*Ds)[NB + 1], T (*Li)[NB + 1], bool store) {
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    const T* P = L + S.panel_off[J];
    const int tid = threadIdx.x;
    if (tid < 32) {
        const int lane = tid;
        T a[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) {
            T v = (lane == c) ? one<T>() : zero<T>();
            if (lane < s && c < s && c <= lane) v = P[(int64_t)lane + (int64_t)c * f];
            a[c] = v;
        }
        warp_ldlt32<T>(a, lane, errflag);
#pragma unroll
        for (int c = 0; c < NB; ++c) Ds[lane][c] = a[c];
        __syncwarp();
        // inverse of the unit lower factor: lane c owns column c
        const int c = lane;
        for (int i = c + 1; i < NB; ++i) {
            T v = Ds[i][c];
            for (int k = c + 1; k < i; ++k) fma_acc(v, Ds[i][k], Li[k][c]);
            Li[i][c] = sub(zero<T>(), v);
        }
        __syncwarp();
        if (store) {
            T* Db = dblk + S.dblk_off[J];
            for (int i = 0; i < NB; ++i) {
                // element (i, c): strictly lower -> Linv, diagonal -> pivot, upper -> 0
                T v = zero<T>();
                if (i > c) v = Li[i][c];
                else if (i == c) v = Ds[i][i];
                Db[i + c * 32] = v;
            }
        }
    }
    __syncthreads();
}

// phase B: L21 slab = S * Linv^T * D^-1 for rows [row0, row0 + SLAB) of the front (row0 >= s)
template <class T>
__device__ __forceinline__ void front_slab(const DevSymbolic& S, int J, int row0, T* L, T (*Ds)[NB + 1],
                                           T (*Li)[NB + 1], T (*Ss)[NB + 1]) {
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    T* P = L + S.panel_off[J];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < SLAB * NB; idx += 256) {
        const int m = idx % SLAB, c = idx / SLAB;
        Ss[m][c] = (row0 + m < f && c < s) ? P[(int64_t)(row0 + m) + (int64_t)c * f] : zero<T>();
    }
    __syncthreads();
    // out[m][c] = (S[m][c] + sum_{t<c} S[m][t] * Linv[c][t]) / d_c ; each thread computes 12 outputs
    T outv[SLAB * NB / 256];
#pragma unroll
    for (int q = 0; q < SLAB * NB / 256; ++q) {
        const int idx = tid + 256 * q;
        const int m = idx % SLAB, c = idx / SLAB;
        T v = Ss[m][c];
        for (int t = 0; t < c; ++t) fma_acc(v, Ss[m][t], Li[c][t]);
        outv[q] = mul(v, recip(Ds[c][c]));
    }
#pragma unroll
    for (int q = 0; q < SLAB * NB / 256; ++q) {
        const int idx = tid + 256 * q;
        const int m = idx % SLAB, c = idx / SLAB;
        if (row0 + m < f && c < s) P[(int64_t)(row0 + m) + (int64_t)c * f] = outv[q];
    }
    __syncthreads();
}

// Schur tile: U_J[i0:i0+64, j0:j0+64] -= L21 D L21^T  (lower triangle), K = s_J
template <class T>
__device__ __forceinline__ void schur_tile(const DevSymbolic& S, int J, int i0, int j0, const T* L, const T* dblk,
                                           T* U, T (*As)[64], T (*Bs)[64]) {
    constexpr int KC = 16;
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    const T* P = L + S.panel_off[J];
    const T* Dk = dblk + S.dblk_off[J];
    T* UJ = U + S.upd_off[J];
    const int tid = threadIdx.x, tr = tid & 15, tc = tid >> 4;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = zero<T>();
    for (int kk = 0; kk < s; kk += KC) {
        const int m = tid & 63;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int k = (tid >> 6) + 4 * it;
            const int kg = kk + k;
            T av = zero<T>(), bv = zero<T>();
            if (kg < s) {
                if (i0 + m < u) av = P[(int64_t)(s + i0 + m) + (int64_t)kg * f];
                if (j0 + m < u) bv = mul(P[(int64_t)(s + j0 + m) + (int64_t)kg * f], Dk[kg * 33]);
            }
            As[k][m] = av;
            Bs[k][m] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            T av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[k][tr + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tc + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) fma_acc(acc[i][j], av[i], bv[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + tr + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tc + 16 * j;
            if (gi < u && gj < u && gi >= gj) {
                T* t = UJ + ((int64_t)gi + (int64_t)gj * u);
                *t = sub(*t, acc[i][j]);
            }
        }
    }
}

// ---- top-level factorization kernels ----
template <class T>
__global__ void __launch_bounds__(256) k_extend_add(DevSymbolic S, const int32_t* __restrict__ parents, T* L, T* U) {
    ea_body<T>(S, parents[blockIdx.x], blockIdx.y, gridDim.y, L, U);
}

template <class T>
__global__ void __launch_bounds__(256) k_front(DevSymbolic S, const int2* __restrict__ items, T* L, T* dblk,
                                               int32_t* errflag) {
    extern __shared__ __align__(16) unsigned char dre_smem_raw[];
    T* smp = reinterpret_cast<T*>(dre_smem_raw);
    T (*Ds)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(smp);           smp += NB * (NB + 1);
    T (*Li)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(smp);           smp += NB * (NB + 1);
    T (*Ss)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(smp);
    const int2 item = items[blockIdx.x];
    const int J = item.x, slab = item.y;
    front_diag<T>(S, J, L, dblk, errflag, Ds, Li, slab == 0);
    const int s = sn_s(S, J), u = sn_u(S, J);
    if (slab * SLAB < u) front_slab<T>(S, J, s + slab * SLAB, L, Ds, Li, Ss);
}

template <class T>
__global__ void __launch_bounds__(256) k_schur(DevSymbolic S, const int4* __restrict__ items, const T* __restrict__ L,
                                               const T* __restrict__ dblk, T* U) {
    __shared__ T As[16][64];
    __shared__ T Bs[16][64];
    const int4 item = items[blockIdx.x];
    schur_tile<T>(S, item.x, item.y * 64, item.z * 64, L, dblk, U, As, Bs);
}

// ---- bottom subtrees: one CTA factors a whole subtree (zero U, extend-add, panel, Schur per front) ----
template <class T>
__global__ void __launch_bounds__(256) k_factor_subtree(DevSymbolic S, T* L, T* dblk, T* U, int32_t* errflag) {
    extern __shared__ __align__(16) unsigned char dre_smem_raw[];
    T* smp = reinterpret_cast<T*>(dre_smem_raw);
    T (*Ds)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(smp);           smp += NB * (NB + 1);
    T (*Li)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(smp);           smp += NB * (NB + 1);
    T (*Ss)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(smp);           smp += SLAB * (NB + 1);
    T (*As)[64] = reinterpret_cast<T (*)[64]>(smp);                   smp += 16 * 64;
    T (*Bs)[64] = reinterpret_cast<T (*)[64]>(smp);
    const int t = blockIdx.x;
    const int tid = threadIdx.x;
    for (int p = S.st_ptr[t]; p < S.st_ptr[t + 1]; ++p) {
        const int J = S.st_sn[p];
        const int s = sn_s(S, J), u = sn_u(S, J);
        T* UJ = U + S.upd_off[J];
        for (int64_t idx = tid; idx < (int64_t)u * u; idx += 256) UJ[idx] = zero<T>();
        __syncthreads();
        ea_body<T>(S, J, 0, 1, L, U);
        __syncthreads();
        front_diag<T>(S, J, L, dblk, errflag, Ds, Li, true);
        for (int r0 = 0; r0 < u; r0 += SLAB) front_slab<T>(S, J, s + r0, L, Ds, Li, Ss);
        const int nt = (u + 63) / 64;
        for (int ti = 0; ti < nt; ++ti)
            for (int tj = 0; tj <= ti; ++tj) schur_tile<T>(S, J, ti * 64, tj * 64, L, dblk, U, As, Bs);
        __syncthreads();
    }
}

template <class T>
void launch_extend_add(const DevSymbolic& S, const int32_t* parents, int nparents, int gy, T* L, T* U,
                       cudaStream_t st, int64_t* launches) {
    if (nparents <= 0) return;
    dim3 grid(nparents, gy);
    k_extend_add<T><<<grid, 256, 0, st>>>(S, parents, L, U);
    if (launches) *launches += 1;
}

template <class T>
void launch_front(const DevSymbolic& S, const int2* items, int nitems, T* L, T* dblk, int32_t* errflag,
                  cudaStream_t st, int64_t* launches) {
    if (nitems <= 0) return;
    const int smem = (int)sizeof(T) * (2 * NB * (NB + 1) + SLAB * (NB + 1));
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_front<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    k_front<T><<<nitems, 256, smem, st>>>(S, items, L, dblk, errflag);
    if (launches) *launches += 1;
}

template <class T>
void launch_schur(const DevSymbolic& S, const int4* items, int nitems, const T* L, const T* dblk, T* U,
                  cudaStream_t st, int64_t* launches) {
    if (nitems <= 0) return;
    k_schur<T><<<nitems, 256, 0, st>>>(S, items, L, dblk, U);
    if (launches) *launches += 1;
}

template <class T>
void launch_factor_subtrees(const DevSymbolic& S, T* L, T* dblk, T* U, int32_t* errflag, cudaStream_t st,
                            int64_t* launches) {
    if (S.nsubtrees <= 0) return;
    const int smem = (int)sizeof(T) * (2 * NB * (NB + 1) + SLAB * (NB + 1) + 2 * 16 * 64);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_factor_subtree<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    k_factor_subtree<T><<<S.nsubtrees, 256, smem, st>>>(S, L, dblk, U, errflag);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// solves.  W is row-major (n x ldw); the update vector of supernode J is column-major
// (u_J contiguous per RHS column) at  t + rhs_off[J]*ldw.
// forward:  y_J = L11^-1 (b_J + children),  t_J = children - L21 y_J
// backward: x_J = L11^-T (D^-1 y_J - L21^T x_struct)
// ------------------------------------------------------------------------------------------
template <class T, int CW>
struct SweepSmem {
    T (*xb)[CW + 1];   // [32][CW+1]
    T (*Li)[NB + 1];   // [32][33]
    T (*Ls)[64 + 1];   // [32][65]  (forward: L21 tile; backward uses it as [32][33])
    T (*Xs)[CW + 1];   // [32][CW+1]
    __device__ SweepSmem(unsigned char* raw) {
        T* smp = reinterpret_cast<T*>(raw);
        xb = reinterpret_cast<T (*)[CW + 1]>(smp);  smp += NB * (CW + 1);
        Li = reinterpret_cast<T (*)[NB + 1]>(smp);  smp += NB * (NB + 1);
        Ls = reinterpret_cast<T (*)[64 + 1]>(smp);  smp += NB * 65;
        Xs = reinterpret_cast<T (*)[CW + 1]>(smp);
    }
    static constexpr int bytes() { return (int)sizeof(T) * (2 * NB * (CW + 1) + NB * (NB + 1) + NB * 65); }
};

template <class T, int CW>
__device__ __forceinline__ void fwd_body(const DevSymbolic& S, int J, int c0, int ncw, const T* __restrict__ L,
                                         const T* __restrict__ dblk, T* W, int64_t ldw, T* tbuf,
                                         SweepSmem<T, CW>& sm) {
    const int first = S.sn_first[J];
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    const T* P = L + S.panel_off[J];
    const T* Dk = dblk + S.dblk_off[J];
    T* tJ = tbuf + S.rhs_off[J] * ldw;
    const int tid = threadIdx.x;

    for (int idx = tid; idx < u * ncw; idx += 256) {
        const int cc = idx / u, r = idx - cc * u;
        tJ[(int64_t)(c0 + cc) * u + r] = zero<T>();
    }
    __syncthreads();
    for (int ci = S.child_ptr[J]; ci < S.child_ptr[J + 1]; ++ci) {
        const int c = S.child_idx[ci];
        const int uc = sn_u(S, c);
        const int32_t* rel = S.relmap + S.sn_rowptr[c];
        const T* tch = tbuf + S.rhs_off[c] * ldw;
        for (int idx = tid; idx < uc * ncw; idx += 256) {
            const int cc = idx / uc, i = idx - cc * uc;
            const int pr = rel[i];
            const T val = tch[(int64_t)(c0 + cc) * uc + i];
            T* t = (pr < s) ? (W + (int64_t)(first + pr) * ldw + c0 + cc) : (tJ + (int64_t)(c0 + cc) * u + (pr - s));
            *t = add(*t, val);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < NB * CW; idx += 256) {
        const int i = idx / CW, cc = idx - i * CW;
        sm.xb[i][cc] = (i < s && cc < ncw) ? W[(int64_t)(first + i) * ldw + c0 + cc] : zero<T>();
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx & 31, k = idx >> 5;
        sm.Li[i][k] = (k < i) ? Dk[i + k * 32] : zero<T>();
    }
    __syncthreads();
    {   // y = Linv x  (unit lower): thread (row i, column group)
        constexpr int CPT = CW / 8;
        const int i = tid & 31, cg = tid >> 5;
        T acc[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[c] = sm.xb[i][cg * CPT + c];
        for (int k = 0; k < i; ++k) {
            const T l = sm.Li[i][k];
#pragma unroll
            for (int c = 0; c < CPT; ++c) fma_acc(acc[c], l, sm.xb[k][cg * CPT + c]);
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CPT; ++c) sm.xb[i][cg * CPT + c] = acc[c];
    }
    __syncthreads();
    for (int idx = tid; idx < NB * CW; idx += 256) {
        const int i = idx / CW, cc = idx - i * CW;
        if (i < s && cc < ncw) W[(int64_t)(first + i) * ldw + c0 + cc] = sm.xb[i][cc];
    }
    {   // t_J -= L21 y, 64-row tiles of L21 staged in shared memory
        constexpr int CPT = CW / 4;
        const int rl = tid & 63, cg = tid >> 6;
        for (int r0 = 0; r0 < u; r0 += 64) {
            __syncthreads();
            for (int idx = tid; idx < NB * 64; idx += 256) {
                const int rr = idx & 63, k = idx >> 6;
                sm.Ls[k][rr] = (r0 + rr < u && k < s) ? P[(int64_t)(s + r0 + rr) + (int64_t)k * f] : zero<T>();
            }
            __syncthreads();
            T acc[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[c] = zero<T>();
            for (int k = 0; k < s; ++k) {
                const T l = sm.Ls[k][rl];
#pragma unroll
                for (int c = 0; c < CPT; ++c) fma_acc(acc[c], l, sm.xb[k][cg * CPT + c]);
            }
            if (r0 + rl < u) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int cc = cg * CPT + c;
                    if (cc < ncw) {
                        T* t = tJ + (int64_t)(c0 + cc) * u + r0 + rl;
                        *t = sub(*t, acc[c]);
                    }
                }
            }
        }
    }
    __syncthreads();
}

template <class T, int CW>
__device__ __forceinline__ void bwd_body(const DevSymbolic& S, int J, int c0, int ncw, const T* __restrict__ L,
                                         const T* __restrict__ dblk, T* W, int64_t ldw, SweepSmem<T, CW>& sm) {
    constexpr int CPT = CW / 8;
    T (*Lt)[NB + 1] = reinterpret_cast<T (*)[NB + 1]>(sm.Ls);  // [32 rows][33] tile of L21
    const int first = S.sn_first[J];
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    const T* P = L + S.panel_off[J];
    const T* Dk = dblk + S.dblk_off[J];
    const int32_t* rows = S.sn_rows + S.sn_rowptr[J];
    const int tid = threadIdx.x;
    const int kq = tid & 31, cg = tid >> 5;

    T acc[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[c] = zero<T>();
    for (int r0 = 0; r0 < u; r0 += NB) {
        for (int idx = tid; idx < NB * NB; idx += 256) {
            const int r = idx & 31, k = idx >> 5;
            Lt[r][k] = (r0 + r < u && k < s) ? P[(int64_t)(s + r0 + r) + (int64_t)k * f] : zero<T>();
        }
        for (int idx = tid; idx < NB * CW; idx += 256) {
            const int r = idx / CW, cc = idx - r * CW;
            T v = zero<T>();
            if (r0 + r < u && cc < ncw) v = W[(int64_t)rows[r0 + r] * ldw + c0 + cc];
            sm.Xs[r][cc] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < NB; ++r) {
            const T l = Lt[r][kq];
#pragma unroll
            for (int c = 0; c < CPT; ++c) fma_acc(acc[c], l, sm.Xs[r][cg * CPT + c]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < NB * CW; idx += 256) {
        const int i = idx / CW, cc = idx - i * CW;
        sm.xb[i][cc] = (i < s && cc < ncw) ? W[(int64_t)(first + i) * ldw + c0 + cc] : zero<T>();
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx & 31, k = idx >> 5;
        sm.Li[i][k] = Dk[i + k * 32];  // strictly lower: Linv, diagonal: pivots
    }
    __syncthreads();
    {
        const T rd = (kq < s) ? recip(sm.Li[kq][kq]) : zero<T>();
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int cc = cg * CPT + c;
            sm.xb[kq][cc] = sub(mul(sm.xb[kq][cc], rd), acc[c]);
        }
    }
    __syncthreads();
    {   // x = Linv^T z : x_i = z_i + sum_{k>i} Linv[k][i] z_k
        T xv[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) xv[c] = sm.xb[kq][cg * CPT + c];
        for (int k = kq + 1; k < s; ++k) {
            const T l = sm.Li[k][kq];
#pragma unroll
            for (int c = 0; c < CPT; ++c) fma_acc(xv[c], l, sm.xb[k][cg * CPT + c]);
        }
        if (kq < s) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int cc = cg * CPT + c;
                if (cc < ncw) W[(int64_t)(first + kq) * ldw + c0 + cc] = xv[c];
            }
        }
    }
    __syncthreads();
}

template <class T, int CW>
__global__ void __launch_bounds__(256) k_fwd(DevSymbolic S, const int32_t* __restrict__ sns, const T* __restrict__ L,
                                             const T* __restrict__ dblk, T* W, int64_t ldw, int nrhs, T* tbuf) {
    extern __shared__ __align__(16) unsigned char dre_smem_raw[];
    SweepSmem<T, CW> sm(dre_smem_raw);
    const int c0 = blockIdx.y * CW;
    fwd_body<T, CW>(S, sns[blockIdx.x], c0, min(CW, nrhs - c0), L, dblk, W, ldw, tbuf, sm);
}

template <class T, int CW>
__global__ void __launch_bounds__(256) k_bwd(DevSymbolic S, const int32_t* __restrict__ sns, const T* __restrict__ L,
                                             const T* __restrict__ dblk, T* W, int64_t ldw, int nrhs) {
    extern __shared__ __align__(16) unsigned char dre_smem_raw[];
    SweepSmem<T, CW> sm(dre_smem_raw);
    const int c0 = blockIdx.y * CW;
    bwd_body<T, CW>(S, sns[blockIdx.x], c0, min(CW, nrhs - c0), L, dblk, W, ldw, sm);
}

// bottom subtrees: CTA (subtree, column chunk) walks the subtree (ascending for forward, descending for
// backward); right-hand-side columns are independent, so no inter-CTA synchronisation is needed.
template <class T, int CW>
__global__ void __launch_bounds__(256) k_fwd_subtree(DevSymbolic S, const T* __restrict__ L,
                                                     const T* __restrict__ dblk, T* W, int64_t ldw, int nrhs,
                                                     T* tbuf) {
    extern __shared__ __align__(16) unsigned char dre_smem_raw[];
    SweepSmem<T, CW> sm(dre_smem_raw);
    const int c0 = blockIdx.y * CW;
    const int ncw = min(CW, nrhs - c0);
    const int t = blockIdx.x;
    for (int p = S.st_ptr[t]; p < S.st_ptr[t + 1]; ++p)
        fwd_body<T, CW>(S, S.st_sn[p], c0, ncw, L, dblk, W, ldw, tbuf, sm);
}

template <class T, int CW>
__global__ void __launch_bounds__(256) k_bwd_subtree(DevSymbolic S, const T* __restrict__ L,
                                                     const T* __restrict__ dblk, T* W, int64_t ldw, int nrhs) {
    extern __shared__ __align__(16) unsigned char dre_smem_raw[];
    SweepSmem<T, CW> sm(dre_smem_raw);
    const int c0 = blockIdx.y * CW;
    const int ncw = min(CW, nrhs - c0);
    const int t = blockIdx.x;
    for (int p = S.st_ptr[t + 1] - 1; p >= S.st_ptr[t]; --p)
        bwd_body<T, CW>(S, S.st_sn[p], c0, ncw, L, dblk, W, ldw, sm);
}

template <class T, int CW>
static void set_sweep_attrs() {
    static bool done = false;
    if (done) return;
    const int smem = SweepSmem<T, CW>::bytes();
    cudaFuncSetAttribute(k_fwd<T, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bwd<T, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_fwd_subtree<T, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_bwd_subtree<T, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    done = true;
}

template <class T>
void launch_fwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, const T* L, const T* dblk, T* W,
                      int64_t ldw, int nrhs, T* tbuf, cudaStream_t st, int64_t* launches) {
    if (nsns <= 0 || nrhs <= 0) return;
    set_sweep_attrs<T, 32>();
    set_sweep_attrs<T, 8>();
    if ((int64_t)nsns * ((nrhs + 31) / 32) >= 2 * 148) {
        dim3 grid(nsns, (nrhs + 31) / 32);
        k_fwd<T, 32><<<grid, 256, SweepSmem<T, 32>::bytes(), st>>>(S, sns, L, dblk, W, ldw, nrhs, tbuf);
    } else {
        dim3 grid(nsns, (nrhs + 7) / 8);
        k_fwd<T, 8><<<grid, 256, SweepSmem<T, 8>::bytes(), st>>>(S, sns, L, dblk, W, ldw, nrhs, tbuf);
    }
    if (launches) *launches += 1;
}

template <class T>
void launch_bwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, const T* L, const T* dblk, T* W,
                      int64_t ldw, int nrhs, cudaStream_t st, int64_t* launches) {
    if (nsns <= 0 || nrhs <= 0) return;
    set_sweep_attrs<T, 32>();
    set_sweep_attrs<T, 8>();
    if ((int64_t)nsns * ((nrhs + 31) / 32) >= 2 * 148) {
        dim3 grid(nsns, (nrhs + 31) / 32);
        k_bwd<T, 32><<<grid, 256, SweepSmem<T, 32>::bytes(), st>>>(S, sns, L, dblk, W, ldw, nrhs);
    } else {
        dim3 grid(nsns, (nrhs + 7) / 8);
        k_bwd<T, 8><<<grid, 256, SweepSmem<T, 8>::bytes(), st>>>(S, sns, L, dblk, W, ldw, nrhs);
    }
    if (launches) *launches += 1;
}

template <class T>
void launch_fwd_subtrees(const DevSymbolic& S, const T* L, const T* dblk, T* W, int64_t ldw, int nrhs, T* tbuf,
                         cudaStream_t st, int64_t* launches) {
    if (S.nsubtrees <= 0 || nrhs <= 0) return;
    set_sweep_attrs<T, 32>();
    set_sweep_attrs<T, 8>();
    if ((int64_t)S.nsubtrees * ((nrhs + 31) / 32) >= 2 * 148) {
        dim3 grid(S.nsubtrees, (nrhs + 31) / 32);
        k_fwd_subtree<T, 32><<<grid, 256, SweepSmem<T, 32>::bytes(), st>>>(S, L, dblk, W, ldw, nrhs, tbuf);
    } else {
        dim3 grid(S.nsubtrees, (nrhs + 7) / 8);
        k_fwd_subtree<T, 8><<<grid, 256, SweepSmem<T, 8>::bytes(), st>>>(S, L, dblk, W, ldw, nrhs, tbuf);
    }
    if (launches) *launches += 1;
}

template <class T>
void launch_bwd_subtrees(const DevSymbolic& S, const T* L, const T* dblk, T* W, int64_t ldw, int nrhs,
                         cudaStream_t st, int64_t* launches) {
    if (S.nsubtrees <= 0 || nrhs <= 0) return;
    set_sweep_attrs<T, 32>();
    set_sweep_attrs<T, 8>();
    if ((int64_t)S.nsubtrees * ((nrhs + 31) / 32) >= 2 * 148) {
        dim3 grid(S.nsubtrees, (nrhs + 31) / 32);
        k_bwd_subtree<T, 32><<<grid, 256, SweepSmem<T, 32>::bytes(), st>>>(S, L, dblk, W, ldw, nrhs);
    } else {
        dim3 grid(S.nsubtrees, (nrhs + 7) / 8);
        k_bwd_subtree<T, 8><<<grid, 256, SweepSmem<T, 8>::bytes(), st>>>(S, L, dblk, W, ldw, nrhs);
    }
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// RHS staging, SMW core and epilogue
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void k_load_rhs(T* __restrict__ W, int64_t ldw, const double* __restrict__ R, int64_t ldr, int r,
                           const double* __restrict__ Vt, int64_t ldv, int m, int64_t n) {
    const int tot = r + m;
    const int64_t total = n * tot;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / tot;
        const int c = (int)(idx % tot);
        const double v = (c < r) ? R[row * ldr + c] : Vt[row * ldv + (c - r)];
        T t;
        from_real(v, t);
        W[row * ldw + c] = t;
    }
}

template <class T>
void launch_load_rhs(T* W, int64_t ldw, const double* R, int64_t ldr, int r, const double* Vt, int64_t ldv, int m,
                     int64_t n, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || r + m <= 0) return;
    int blocks = (int)std::min<int64_t>((n * (r + m) + 255) / 256, 148 * 16);
    k_load_rhs<T><<<blocks, 256, 0, st>>>(W, ldw, R, ldr, r, Vt, ldv, m, n);
    if (launches) *launches += 1;
}

__device__ __forceinline__ double abs2(double a) { return a * a; }
__device__ __forceinline__ double abs2(cplx a) { return a.x * a.x + a.y * a.y; }

// S = alpha I + BtW[:, r:r+m];  Sol = S^-1 BtW[:, 0:r]   (m <= 32; LU with partial pivoting)
template <class T>
__global__ void __launch_bounds__(256) k_smw_core(const T* __restrict__ BtW, int64_t ldb, int m, int r, double alpha,
                                                  T* __restrict__ Sol, int32_t* errflag) {
    __shared__ T Sm[32][33];
    __shared__ int piv[32];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < m * m; idx += 256) {
        const int i = idx / m, j = idx % m;
        T v = BtW[(int64_t)i * ldb + r + j];
        if (i == j) {
            T a;
            from_real(alpha, a);
            v = add(v, a);
        }
        Sm[i][j] = v;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < m; ++k) {
            int p = k;
            double best = abs2(Sm[k][k]);
            for (int i = k + 1; i < m; ++i) {
                const double a2 = abs2(Sm[i][k]);
                if (a2 > best) { best = a2; p = i; }
            }
            piv[k] = p;
            if (p != k)
                for (int j = 0; j < m; ++j) { T t = Sm[k][j]; Sm[k][j] = Sm[p][j]; Sm[p][j] = t; }
            if (is_bad(Sm[k][k])) atomicExch(errflag, 2);
            const T rd = recip(Sm[k][k]);
            for (int i = k + 1; i < m; ++i) {
                const T l = mul(Sm[i][k], rd);
                Sm[i][k] = l;
                for (int j = k + 1; j < m; ++j) Sm[i][j] = sub(Sm[i][j], mul(l, Sm[k][j]));
            }
        }
    }
    __syncthreads();
    for (int c = tid; c < r; c += 256) {
        T x[32];
        for (int i = 0; i < m; ++i) x[i] = BtW[(int64_t)i * ldb + c];
        for (int k = 0; k < m; ++k) {
            const int p = piv[k];
            if (p != k) { T t = x[k]; x[k] = x[p]; x[p] = t; }
        }
        for (int i = 1; i < m; ++i) {
            T v = x[i];
            for (int k = 0; k < i; ++k) v = sub(v, mul(Sm[i][k], x[k]));
            x[i] = v;
        }
        for (int i = m - 1; i >= 0; --i) {
            T v = x[i];
            for (int k = i + 1; k < m; ++k) v = sub(v, mul(Sm[i][k], x[k]));
            x[i] = mul(v, recip(Sm[i][i]));
        }
        for (int i = 0; i < m; ++i) Sol[(int64_t)i * r + c] = x[i];
    }
}

template <class T>
void launch_smw_core(const T* BtW, int64_t ldb, int m, int r, double alpha, T* Sol, int32_t* errflag,
                     cudaStream_t st, int64_t* launches) {
    if (m <= 0 || r <= 0) return;
    k_smw_core<T><<<1, 256, 0, st>>>(BtW, ldb, m, r, alpha, Sol, errflag);
    if (launches) *launches += 1;
}

__device__ __forceinline__ void emit(double v, int mode, double d, double* V1, double* V2, int64_t o1, int64_t o2) {
    V1[o1] = v;
}
__device__ __forceinline__ void emit(cplx v, int mode, double d, double* V1, double* V2, int64_t o1, int64_t o2) {
    if (mode == 2) {
        V1[o1] = 1.4142135623730951 * v.x + (1.4142135623730951 * d) * v.y;
        V2[o2] = sqrt(2.0 * d * d + 2.0) * v.y;
    } else {
        V1[o1] = v.x;
        V2[o2] = v.y;
    }
}

template <class T>
__global__ void k_smw_apply(const T* __restrict__ W, int64_t ldw, int r, int m, const T* __restrict__ Sol, int mode,
                            double d, double* __restrict__ V1, int64_t ld1, double* __restrict__ V2, int64_t ld2,
                            int64_t n) {
    const int64_t total = n * r;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / r;
        const int c = (int)(idx % r);
        const T* w = W + row * ldw;
        T v = w[c];
        for (int j = 0; j < m; ++j) v = sub(v, mul(w[r + j], Sol[(int64_t)j * r + c]));
        emit(v, mode, d, V1, V2, row * ld1 + c, row * ld2 + c);
    }
}

template <class T>
void launch_smw_apply(const T* W, int64_t ldw, int r, int m, const T* Sol, int mode, double d, double* V1,
                      int64_t ld1, double* V2, int64_t ld2, int64_t n, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || r <= 0) return;
    int blocks = (int)std::min<int64_t>((n * r + 255) / 256, 148 * 16);
    k_smw_apply<T><<<blocks, 256, 0, st>>>(W, ldw, r, m, Sol, mode, d, V1, ld1, V2, ld2, n);
    if (launches) *launches += 1;
}

// ---- explicit instantiations ----
#define DRE_INST(T)                                                                                                  \
    template void launch_assemble<T>(const DevSymbolic&, T*, double, T, cudaStream_t, int64_t*);                     \
    template void launch_extend_add<T>(const DevSymbolic&, const int32_t*, int, int, T*, T*, cudaStream_t,           \
                                       int64_t*);                                                                    \
    template void launch_front<T>(const DevSymbolic&, const int2*, int, T*, T*, int32_t*, cudaStream_t, int64_t*);   \
    template void launch_schur<T>(const DevSymbolic&, const int4*, int, const T*, const T*, T*, cudaStream_t,        \
                                  int64_t*);                                                                         \
    template void launch_factor_subtrees<T>(const DevSymbolic&, T*, T*, T*, int32_t*, cudaStream_t, int64_t*);       \
    template void launch_fwd_level<T>(const DevSymbolic&, const int32_t*, int, const T*, const T*, T*, int64_t, int, \
                                      T*, cudaStream_t, int64_t*);                                                   \
    template void launch_bwd_level<T>(const DevSymbolic&, const int32_t*, int, const T*, const T*, T*, int64_t, int, \
                                      cudaStream_t, int64_t*);                                                       \
    template void launch_fwd_subtrees<T>(const DevSymbolic&, const T*, const T*, T*, int64_t, int, T*, cudaStream_t, \
                                         int64_t*);                                                                  \
    template void launch_bwd_subtrees<T>(const DevSymbolic&, const T*, const T*, T*, int64_t, int, cudaStream_t,     \
                                         int64_t*);                                                                  \
    template void launch_load_rhs<T>(T*, int64_t, const double*, int64_t, int, const double*, int64_t, int, int64_t, \
                                     cudaStream_t, int64_t*);                                                        \
    template void launch_smw_core<T>(const T*, int64_t, int, int, double, T*, int32_t*, cudaStream_t, int64_t*);     \
    template void launch_smw_apply<T>(const T*, int64_t, int, int, const T*, int, double, double*, int64_t, double*, \
                                      int64_t, int64_t, cudaStream_t, int64_t*);
DRE_INST(double)
DRE_INST(cplx)

}  // namespace dre
