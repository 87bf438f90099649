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
// Design (from csrc/symbolic.cpp): supernodes are the nested-dissection blocks -- a leaf is ONE dense
// supernode of up to ~100 columns, separators are split into chains of at most SN_MAX = 256 columns -- so
// the elimination tree has ~a dozen levels and every level is a handful of fat launches:
//   factor  : extend-add -> k_diag (LDL^T of the s x s diagonal block + explicit inverse of its unit-lower
//             factor, one CTA per supernode) -> k_l21 (L21 = A21 Linv' D^-1, plain GEMM over row slabs)
//             -> k_schur (U -= L21 D L21', 64x64 tiles)
//   forward : y = Linv (b + children);  t = children - L21 y        (one launch per level)
//   backward: x = Linv' (D^-1 y - L21' x_struct)                    (one launch per level)
// Because the inverse of the diagonal block is explicit, no step contains a dependent chain: everything is
// a matrix product and runs on the FP64 tensor cores (DMMA m8n8k4) with operand fragments loaded straight
// from global memory / L2 (each factor entry is used by exactly one warp per right-hand-side chunk, so
// staging it in shared memory would only add a barrier); right-hand sides live in shared memory.
// Complex symmetric arithmetic uses the same 8x8x4 real tiles: an 8-column tile holds 4 complex columns
// (re, im interleaved) and  C += Ar*[Xr Xi] + Ai*[-Xi Xr]  is two DMMAs.
// Storage: panel_J = f_J x s_J column-major (ld f_J); Linv_J = s_J x s_J column-major (ones on, zeros above
// the diagonal inside the 32x32 diagonal blocks); pivots in dvec[n]; update matrices u_J x u_J (lower) and
// update vectors (u_J per right-hand side, column-major), one region per supernode.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace dre {

constexpr int NB = 32;     // block size of the in-supernode LDL^T

// Levels with at least this many supernodes run k_diag with 64-thread CTAs.  Opt-in until measured on the GPU:
// DRE_DIAG_NARROW_MIN=296 (two CTAs per SM) is the intended setting; the emulator tests lower it to reach the variant.
static int diag_narrow_min_default() {
    const char* ev = getenv("DRE_DIAG_NARROW_MIN");
    return ev ? std::max(1, atoi(ev)) : (1 << 30);
}
int diag_narrow_min = diag_narrow_min_default();
// 2: k_diag2 (block column in shared memory, inverse built inside the elimination loop); 1: k_diag
int diag_variant = (getenv("DRE_DIAG_V") && atoi(getenv("DRE_DIAG_V")) == 1) ? 1 : 2;
// thread-block cluster of 4 CTAs per supernode on the levels with few fat supernodes (DRE_DIAG_CLUSTER=0: off)
int diag_cluster = (getenv("DRE_DIAG_CLUSTER") && atoi(getenv("DRE_DIAG_CLUSTER")) == 0) ? 1 : 4;

__device__ __forceinline__ int sn_s(const DevSymbolic& S, int J) { return S.sn_first[J + 1] - S.sn_first[J]; }
__device__ __forceinline__ int sn_u(const DevSymbolic& S, int J) { return (int)(S.sn_rowptr[J + 1] - S.sn_rowptr[J]); }

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ cplx shfl(cplx v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-core tiles, generic over T.  m8n8k4 fragments: lane l holds
//   A[l>>2][l&3],  B[l&3][l>>2],  C[l>>2][2*(l&3) + {0,1}].
// ------------------------------------------------------------------------------------------
template <class T> struct MM;
template <> struct MM<double> {
    static constexpr int CPN = 8;  // T-columns covered by one 8x8 tile
    __device__ __forceinline__ static int bcol(int lane) { return lane >> 2; }
    __device__ __forceinline__ static void mma(double (&c)[2], double a, double b, int) { dmma884(c[0], c[1], a, b); }
    template <class F>
    __device__ __forceinline__ static void each(const double (&c)[2], int lane, F f) {
        f(2 * (lane & 3), c[0]);
        f(2 * (lane & 3) + 1, c[1]);
    }
};
template <> struct MM<cplx> {
    static constexpr int CPN = 4;
    __device__ __forceinline__ static int bcol(int lane) { return lane >> 3; }
    // b = X[k][complex column lane>>3]; the lane's real column is the re (even) or im (odd) part of it
    __device__ __forceinline__ static void mma(double (&c)[2], cplx a, cplx b, int lane) {
        const bool im = (lane >> 2) & 1;
        dmma884(c[0], c[1], a.x, im ? b.y : b.x);
        dmma884(c[0], c[1], a.y, im ? b.x : -b.y);
    }
    template <class F>
    __device__ __forceinline__ static void each(const double (&c)[2], int lane, F f) {
        f(lane & 3, mk(c[0], c[1]));
    }
};

// shared-memory leading dimension (in T) of a right-hand-side tile with CW columns: fragment loads hit the
// minimum of two wavefronts (real: ld = 8 mod 16 doubles; complex: ld = 4 mod 8 elements)
template <class T, int CW>
struct RhsLd {
    static constexpr int value = sizeof(T) == 8 ? (((CW + 15) & ~15) + 8) : (((CW + 7) & ~7) + 4);
};

// acc[nt] (8 rows x NT tiles) += sum_{k in [kbeg, kend)} A(row, k) * B(k, tile nt)
// fa(k): this lane's A element (row = lane>>2 fixed by the caller), fb(k, nt): this lane's B element.
// Both must return zero outside their valid range (kend need not be a multiple of 4).
template <class T, int NT, class FA, class FB>
__device__ __forceinline__ void strip_mma(double (&acc)[NT][2], int kbeg, int kend, int lane, FA fa, FB fb) {
    const int kk = lane & 3;
    int k0 = kbeg;
    if (k0 + 16 <= kend) {
        // software pipeline: the next four A fragments (global / L2 loads) are in flight while the current
        // four feed the tensor cores
        T a[4], an[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = fa(k0 + 4 * q + kk);
        for (; k0 + 32 <= kend; k0 += 16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) an[q] = fa(k0 + 16 + 4 * q + kk);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) MM<T>::mma(acc[nt], a[q], fb(k0 + 4 * q + kk, nt), lane);
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = an[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) MM<T>::mma(acc[nt], a[q], fb(k0 + 4 * q + kk, nt), lane);
        k0 += 16;
    }
    for (; k0 < kend; k0 += 4) {
        const T a = fa(k0 + kk);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) MM<T>::mma(acc[nt], a, fb(k0 + kk, nt), lane);
    }
}

// 32 x 32 (in T) block product per warp: acc[mt][nt] += sum_{k<K} A(i, k) B(k, j)
// fa(i, k), fb(k, j) with i, j in [0, 32) return zero outside their valid range.
template <class T>
struct Blk32 {
    static constexpr int NTW = 32 / MM<T>::CPN;
    double acc[4][NTW][2];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    }
    template <class FA, class FB>
    __device__ __forceinline__ void gemm(int K, int lane, FA fa, FB fb) {
        const int kk = lane & 3, r = lane >> 2, bc = MM<T>::bcol(lane);
        for (int k0 = 0; k0 < K; k0 += 4) {
            T a[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = fa(mt * 8 + r, k0 + kk);
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) {
                const T b = fb(k0 + kk, nt * MM<T>::CPN + bc);
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) MM<T>::mma(acc[mt][nt], a[mt], b, lane);
            }
        }
    }
    // The same product for operands that come from global memory / L2: the fragments of the next two k-steps are in
    // flight while the current one feeds the tensor cores (gemm() above issues the eight loads of a k-step and then
    // waits a full L2 latency for them: ~3 us per 32x32x32 product, against ~1.3 us with shared-memory operands).
    template <class FA, class FB>
    __device__ __forceinline__ void gemm_stream(int K, int lane, FA fa, FB fb) {
        const int kk = lane & 3, r = lane >> 2, bc = MM<T>::bcol(lane);
        T a[3][4], b[3][NTW];
        auto load = [&](int slot, int k0) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[slot][mt] = fa(mt * 8 + r, k0 + kk);
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt) b[slot][nt] = fb(k0 + kk, nt * MM<T>::CPN + bc);
        };
        auto mma_slot = [&](int slot) {
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) MM<T>::mma(acc[mt][nt], a[slot][mt], b[slot][nt], lane);
        };
        if (K <= 0) return;
        load(0, 0);
        if (K > 4) load(1, 4);
        int k0 = 0;
        for (; k0 + 12 <= K; k0 += 12) {   // slots rotate 0,1,2 (fully unrolled so that they stay in registers)
            if (k0 + 8 < K) load(2, k0 + 8);
            mma_slot(0);
            if (k0 + 12 < K) load(0, k0 + 12);
            mma_slot(1);
            if (k0 + 16 < K) load(1, k0 + 16);
            mma_slot(2);
        }
        if (k0 < K) {
            if (k0 + 8 < K) load(2, k0 + 8);
            mma_slot(0);
            if (k0 + 4 < K) mma_slot(1);
            if (k0 + 8 < K) mma_slot(2);
        }
    }
    // f(i, j, value) for every element of the block held by this lane
    template <class F>
    __device__ __forceinline__ void each(int lane, F f) const {
        const int r = lane >> 2;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NTW; ++nt)
                MM<T>::each(acc[mt][nt], lane, [&](int c, T v) { f(mt * 8 + r, nt * MM<T>::CPN + c, v); });
    }
};

// ------------------------------------------------------------------------------------------
// assembly: L[dest] = a*A + (e+mu)*E on the lower-triangular union pattern
// ------------------------------------------------------------------------------------------
// prm != nullptr: the two scalars come from a device parameter block {a, Re(e+mu), Im(e+mu)} written on the same stream
// just before (k_set_prm), so that the launch sequence of a numeric factorization is shift-independent and can be
// replayed as a CUDA graph.
__device__ __forceinline__ void emu_from(const double* prm, double& out) { out = prm[1]; }
__device__ __forceinline__ void emu_from(const double* prm, cplx& out) { out = mk(prm[1], prm[2]); }

template <class T>
__global__ void k_assemble(int64_t nasm, const int64_t* __restrict__ dest, const double* __restrict__ va,
                           const double* __restrict__ ve, T* __restrict__ L, double a, T emu,
                           const double* __restrict__ prm) {
    if (prm) {
        a = prm[0];
        emu_from(prm, emu);
    }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nasm; i += (int64_t)gridDim.x * blockDim.x) {
        T v;
        from_real(a * va[i], v);
        L[dest[i]] = add(v, mul(ve[i], emu));
    }
}

__global__ void k_set_prm(double* prm, double a, double re, double im) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        prm[0] = a;
        prm[1] = re;
        prm[2] = im;
    }
}

void launch_set_prm(double* prm, double a, double re, double im, cudaStream_t st, int64_t* launches) {
    DRE_LAUNCH((k_set_prm), 1, 32, 0, st, prm, a, re, im);
    if (launches) *launches += 1;
}

template <class T>
void launch_assemble(const DevSymbolic& S, T* L, double a, T emu, cudaStream_t st, int64_t* launches, const double* prm) {
    int blocks = (int)std::min<int64_t>((S.nasm + 255) / 256, 148 * 8);
    if (blocks < 1) blocks = 1;
    DRE_LAUNCH((k_assemble<T>), blocks, 256, 0, st, S.nasm, S.asm_dest, S.asm_a, S.asm_e, L, a, emu, prm);
    if (launches) *launches += 1;
}

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

// ------------------------------------------------------------------------------------------
// factorization kernels
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) k_extend_add(DevSymbolic S, const int32_t* __restrict__ parents, T* L, T* U) {
    ea_body<T>(S, parents[blockIdx.x], blockIdx.y, gridDim.y, L, U);
}

// LDL^T of the s x s diagonal block of supernode J (right-looking over 32-column blocks, in place in the
// panel) and the explicit inverse of its unit-lower factor.  One CTA per supernode; every 32x32x32 block
// product is done by one warp on the tensor cores.
// NW = warps per CTA: 8 for the few fat supernodes near the root; 2 on the populous leaf levels, where the CTA count
// exceeds what the register file holds at 256 threads x 128 registers (2 CTAs per SM -> 3.5 waves of mostly idle
// threads at 1024 leaves, and no registers left for the kernels of the other streams) -- 64-thread CTAs run 8 per SM.
template <class T, int NW>
__global__ void __launch_bounds__(32 * NW) k_diag(DevSymbolic S, const int32_t* __restrict__ sns, T* L, T* Linv, T* dvec,
                                                  int32_t* errflag) {
    __shared__ T Ds[NB][NB + 1];   // current diagonal block: strictly lower = L_bb, diagonal = pivots
    __shared__ T Li[NB][NB + 1];   // inverse of the unit-lower L_bb (ones on, zeros above the diagonal)
    const int J = sns[blockIdx.x];
    const int s = sn_s(S, J), f = s + sn_u(S, J);
    T* P = L + S.panel_off[J];
    T* LI = Linv + S.linv_off[J];
    T* dv = dvec + S.sn_first[J];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbk = (s + NB - 1) / NB;
    Blk32<T> blk;

    for (int b = 0; b < nbk; ++b) {
        const int jb = b * NB, nb = min(NB, s - jb);
        if (warp == 0) {
            T a[NB];
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                T v = (lane == c) ? one<T>() : zero<T>();
                if (lane < nb && c < nb && c <= lane) v = P[(int64_t)(jb + lane) + (int64_t)(jb + c) * f];
                a[c] = v;
            }
            warp_ldlt32<T>(a, lane, errflag);
#pragma unroll
            for (int c = 0; c < NB; ++c) Ds[lane][c] = a[c];
            __syncwarp();
            // inverse of the unit-lower factor: lane j owns column j, x_i = delta_ij - sum_{k<i} L[i][k] x_k
            T x[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                T v0 = zero<T>(), v1 = zero<T>();
#pragma unroll
                for (int k = 0; k + 1 < i; k += 2) {
                    fma_acc(v0, Ds[i][k], x[k]);
                    fma_acc(v1, Ds[i][k + 1], x[k + 1]);
                }
                if (i & 1) fma_acc(v0, Ds[i][i - 1], x[i - 1]);
                x[i] = sub((i == lane) ? one<T>() : zero<T>(), add(v0, v1));
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) Li[i][lane] = x[i];
            __syncwarp();
            if (lane < nb) dv[jb + lane] = Ds[lane][lane];
            // block (b, b) of the inverse, coalesced along rows
            for (int c = 0; c < nb; ++c)
                if (lane < nb) {
                    LI[(int64_t)(jb + lane) + (int64_t)(jb + c) * s] = Li[lane][c];
                    if (c <= lane) P[(int64_t)(jb + lane) + (int64_t)(jb + c) * f] = Ds[lane][c];  // L_bb and pivots
                }
        }
        __syncthreads();
        // block rows below (inside the diagonal block): L_Ib = A_Ib Linv_bb' D_b^-1
        for (int I = b + 1 + warp; I < nbk; I += NW) {
            blk.clear();
            blk.gemm(NB, lane,
                     [&](int i, int k) {
                         const int row = I * NB + i;
                         return (row < s && k < nb) ? P[(int64_t)row + (int64_t)(jb + k) * f] : zero<T>();
                     },
                     [&](int k, int j) { return Li[j][k]; });
            blk.each(lane, [&](int i, int j, T v) {
                const int row = I * NB + i;
                if (row < s && j < nb) P[(int64_t)row + (int64_t)(jb + j) * f] = mul(v, recip(Ds[j][j]));
            });
        }
        __syncthreads();
        // trailing update of the diagonal block: C_IK -= L_Ib D_b L_Kb'   for b < K <= I
        const int m = nbk - b - 1;
        for (int p = warp; p < m * (m + 1) / 2; p += NW) {
            int Ii = (int)((sqrtf(8.0f * p + 1.0f) - 1.0f) * 0.5f);
            while ((Ii + 1) * (Ii + 2) / 2 <= p) ++Ii;
            while (Ii * (Ii + 1) / 2 > p) --Ii;
            const int Ki = p - Ii * (Ii + 1) / 2;
            const int I = b + 1 + Ii, K = b + 1 + Ki;
            blk.clear();
            blk.gemm(NB, lane,
                     [&](int i, int k) {
                         const int row = I * NB + i;
                         return (row < s && k < nb) ? P[(int64_t)row + (int64_t)(jb + k) * f] : zero<T>();
                     },
                     [&](int k, int j) {
                         const int col = K * NB + j;
                         return (col < s && k < nb) ? mul(P[(int64_t)col + (int64_t)(jb + k) * f], Ds[k][k]) : zero<T>();
                     });
            blk.each(lane, [&](int i, int j, T v) {
                const int row = I * NB + i, col = K * NB + j;
                if (row < s && col < s) {
                    T* t = P + ((int64_t)row + (int64_t)col * f);
                    *t = sub(*t, v);
                }
            });
        }
        __syncthreads();
    }
    // off-diagonal blocks of the inverse, one block column per warp:
    //   Linv[I][Jc] = -Linv[I][I] * sum_{K=Jc}^{I-1} L[I][K] Linv[K][Jc]
    for (int Jc = warp; Jc < nbk; Jc += NW) {
        for (int I = Jc + 1; I < nbk; ++I) {
            blk.clear();
            for (int K = Jc; K < I; ++K)
                blk.gemm(NB, lane,
                         [&](int i, int k) {
                             const int row = I * NB + i, col = K * NB + k;
                             return (row < s && col < s) ? P[(int64_t)row + (int64_t)col * f] : zero<T>();
                         },
                         [&](int k, int j) {
                             const int row = K * NB + k, col = Jc * NB + j;
                             return (row < s && col < s) ? LI[(int64_t)row + (int64_t)col * s] : zero<T>();
                         });
            blk.each(lane, [&](int i, int j, T v) {
                const int row = I * NB + i, col = Jc * NB + j;
                if (row < s && col < s) LI[(int64_t)row + (int64_t)col * s] = v;
            });
            __syncwarp();
            blk.clear();
            blk.gemm(NB, lane,
                     [&](int i, int k) {
                         const int row = I * NB + i, col = I * NB + k;
                         return (row < s && col < s) ? LI[(int64_t)row + (int64_t)col * s] : zero<T>();
                     },
                     [&](int k, int j) {
                         const int row = I * NB + k, col = Jc * NB + j;
                         return (row < s && col < s) ? LI[(int64_t)row + (int64_t)col * s] : zero<T>();
                     });
            __syncwarp();
            blk.each(lane, [&](int i, int j, T v) {
                const int row = I * NB + i, col = Jc * NB + j;
                if (row < s && col < s) LI[(int64_t)row + (int64_t)col * s] = sub(zero<T>(), v);
            });
            __syncwarp();
        }
    }
}

// LDL^T of a 32x32 block held in shared memory (row i at A + i*LDA, LDA odd: lane-strided accesses are conflict
// free), by one warp, lane i = row i.  No register arrays: the register version above (warp_ldlt32) is compiled
// into a rolled loop over a LOCAL-memory copy of the row (ptxas keeps the dynamically indexed array in local memory:
// LDL -> DFMA -> STL per update), which made the 32x32 step ~50 us.  The reciprocal of the NEXT pivot is started
// right after the one update it depends on, so its latency hides behind the remaining updates of the column.
// On exit: strictly lower part = unit-lower factor, diagonal = pivots, rdiag[i] = 1 / pivot i.
template <class T, int LDA>
__device__ __forceinline__ void warp_ldlt32_smem(T* A, T* rdiag, int lane, int32_t* errflag) {
    T d = A[0];
    if (lane == 0 && is_bad(d)) atomicExch(errflag, 1);
    T rdv = recip(d);
    for (int j = 0; j < NB; ++j) {
        const T w = A[lane * LDA + j];          // unscaled column entry of this row (rows > j use it)
        const T l = mul(w, rdv);
        if (lane == j) rdiag[j] = rdv;
        T rdn = rdv;
        if (j + 1 < NB) {
            // the update that the next pivot waits for
            const T wk = A[(j + 1) * LDA + j];
            if (lane >= j + 1) A[lane * LDA + j + 1] = sub(A[lane * LDA + j + 1], mul(l, wk));
            __syncwarp();
            const T dn = A[(j + 1) * LDA + j + 1];
            if (lane == j + 1 && is_bad(dn)) atomicExch(errflag, 1);
            rdn = recip(dn);
            // remaining columns in groups of eight: all loads of a group are issued before its first store (the
            // compiler keeps shared-memory loads behind earlier stores; element by element the loop ran one
            // LDS -> DFMA -> STS latency chain per entry)
            for (int k0 = j + 2; k0 < NB; k0 += 8) {
                T wk2[8], av[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = k0 + q;
                    wk2[q] = (k < NB) ? A[k * LDA + j] : zero<T>();
                    av[q] = (k < NB && lane >= k) ? A[lane * LDA + k] : zero<T>();
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = k0 + q;
                    if (k < NB && lane >= k) A[lane * LDA + k] = sub(av[q], mul(l, wk2[q]));
                }
            }
        }
        __syncwarp();
        if (lane > j) A[lane * LDA + j] = l;
        rdv = rdn;
    }
    __syncwarp();
}

// X = inverse of the unit-lower factor held in A (as left by warp_ldlt32_smem); X row i at X + i*LDX.
// Lane c owns column c: X[i][c] = delta_ic - sum_{k<i} L[i][k] X[k][c]; every lane only reads back what it wrote.
template <class T, int LDA, int LDX>
__device__ __forceinline__ void warp_trinv32_smem(const T* A, T* X, int lane) {
    for (int i = 0; i < NB; ++i) {
        T v0 = zero<T>(), v1 = zero<T>(), v2 = zero<T>(), v3 = zero<T>();
        int k = 0;
        for (; k + 4 <= i; k += 4) {
            fma_acc(v0, A[i * LDA + k], X[k * LDX + lane]);
            fma_acc(v1, A[i * LDA + k + 1], X[(k + 1) * LDX + lane]);
            fma_acc(v2, A[i * LDA + k + 2], X[(k + 2) * LDX + lane]);
            fma_acc(v3, A[i * LDA + k + 3], X[(k + 3) * LDX + lane]);
        }
        for (; k < i; ++k) fma_acc(v0, A[i * LDA + k], X[k * LDX + lane]);
        X[i * LDX + lane] = sub((i == lane) ? one<T>() : zero<T>(), add(add(v0, v1), add(v2, v3)));
    }
    __syncwarp();
}

// k_diag, second version (default; DRE_DIAG_V=1 selects the kernel above).  Same results in the same arrays.
// What the first version spends its time on (ncu launch list, n = 79 841: 2.0 of the 3.2 ms of a factorization; a
// 200-column supernode near the root takes 413 us on its own): every 32x32x32 block product fetches its operands
// from global memory / L2 behind dependent loads, the inverse is finished by one serial chain of block products per
// block column after the factorization, and only one warp in eight works during the 32x32 steps.  Here
//  * the current block column (rows jb..s-1, 32 columns) lives in shared memory: the 32x32 LDL^T, its inverse, the
//    solve of the rows below and BOTH operands of the trailing update are served from there;
//  * the inverse is built row block by row block INSIDE the elimination loop,
//        Linv[b][Jc] = -Linv[b][b] * sum_{K=Jc}^{b-1} L[b][K] Linv[K][Jc]      (all Jc < b are independent),
//    as further work items of step b next to the trailing-update blocks, so it adds no chain of its own.
// LDL^T of a 32x32 block AND the inverse of its unit-lower factor by the whole CTA (NW warps): lane = row i, warp g owns
// the columns k = g (mod NW).  One barrier per column; per column every thread updates at most 32/NW entries of its row
// of A and of X (Gauss-Jordan: X <- (I - l_j e_j') X), all independent.  A single warp doing the same (warp_ldlt32_smem
// + warp_trinv32_smem) issues ~9000 dependent instructions per block (~30 us measured, 45 % of k_diag2 on the fat
// supernodes); here the chain per column is  pivot -> reciprocal -> l -> update -> barrier.
// A: [NB][LDA] (strictly lower <- unit-lower factor, diagonal <- pivots);  X: [NB][LDX];  rdiag[NB] <- 1 / pivot.
template <class T, int NW, int LDA, int LDX>
__device__ __forceinline__ void cta_ldlt32_inv(T* A, T* X, T* rdiag, int lane, int warp, int32_t* errflag) {
    constexpr int CPW = NB / NW;   // columns per warp
    for (int q = 0; q < CPW; ++q) {
        const int c = warp + q * NW;
        X[lane * LDX + c] = (lane == c) ? one<T>() : zero<T>();
    }
    __syncthreads();
    for (int j = 0; j < NB; ++j) {
        const T d = A[j * LDA + j];
        if (warp == 0 && lane == j && is_bad(d)) atomicExch(errflag, 1);
        const T rdv = recip(d);
        const T l = mul(A[lane * LDA + j], rdv);   // rows > j use it
        T wk[CPW], av[CPW], xj[CPW], xv[CPW];
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            const int k = warp + q * NW;
            const bool ua = k > j && lane >= k, ux = k <= j && lane > j;
            wk[q] = ua ? A[k * LDA + j] : zero<T>();
            av[q] = ua ? A[lane * LDA + k] : zero<T>();
            xj[q] = ux ? X[j * LDX + k] : zero<T>();
            xv[q] = ux ? X[lane * LDX + k] : zero<T>();
        }
        // (column j keeps its UNSCALED entries until the end of the loop, so nothing a slower thread still reads in
        // this column is overwritten here: one barrier per column)
#pragma unroll
        for (int q = 0; q < CPW; ++q) {
            const int k = warp + q * NW;
            if (k > j && lane >= k) A[lane * LDA + k] = sub(av[q], mul(l, wk[q]));
            if (k <= j && lane > j) X[lane * LDX + k] = sub(xv[q], mul(l, xj[q]));
        }
        if (warp == (j % NW) && lane == j) rdiag[j] = rdv;
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < CPW; ++q) {
        const int k = warp + q * NW;
        if (lane > k) A[lane * LDA + k] = mul(A[lane * LDA + k], rdiag[k]);
    }
    __syncthreads();
}

// cluster-wide barrier with release / acquire semantics (global-memory writes of the other CTAs become visible)
template <int CL>
__device__ __forceinline__ void diag_step_barrier() {
#ifndef DRE_SIMT_EMU
    if (CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        return;
    }
#endif
    __syncthreads();
}

// CL = CTAs per supernode (a thread-block cluster): every CTA of the cluster keeps its own copy of the block column and
// repeats the (cheap) 32x32 step and the solve of the rows below; the block products of the trailing update and of the
// inverse -- the bulk of the work, FP64-pipe bound on ONE SM for the fat supernodes near the root -- are dealt out over
// all warps of the cluster.  Only the CTAs exchange data through global memory (L2) behind the cluster barrier.
template <class T, int NW, int MINB, int CL>
__global__ void __launch_bounds__(32 * NW, MINB) k_diag2(DevSymbolic S, const int32_t* __restrict__ sns, T* L, T* Linv, T* dvec,
                                                         int32_t* errflag, int prow_cap) {
    constexpr int LDP = NB + 4;   // fragment loads (8 rows x 4 columns per half warp) hit 16 different banks
    DRE_DYN_SMEM_ALIGNED(unsigned char, dre_smem_raw);
    T* Pn = reinterpret_cast<T*>(dre_smem_raw);   // [prow_cap][LDP]: block column b, rows jb .. s-1
    T* Li0 = Pn + (size_t)prow_cap * LDP;         // 2 x [NB][LDP]: inverse of the unit-lower L_bb (steps b, b-1)
    T* dd = Li0 + 2 * NB * LDP;                   // [NB] pivots of block b, then [NB] their reciprocals
    T* rd = dd + NB;
    constexpr int LDD = NB + 1;
    T* Ds = rd + NB;                              // [NB][LDD]: the diagonal block while one warp factors it
    const int J = sns[blockIdx.x / CL];
    const int crank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;
    const int s = sn_s(S, J), f = s + sn_u(S, J);
    T* P = L + S.panel_off[J];
    T* LI = Linv + S.linv_off[J];
    T* dv = dvec + S.sn_first[J];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbk = (s + NB - 1) / NB;
    Blk32<T> blk;

    // block Jc of row block rb of the inverse: Linv[rb][Jc] = -Linv[rb][rb] sum_{K=Jc}^{rb-1} L[rb][K] Linv[K][Jc]
    auto inverse_item = [&](int rb, int Jc, const T* Lirb) {
        const int jr = rb * NB, nr = min(NB, s - jr);
        blk.clear();
        {
            const T* Arow = P + jr + (int64_t)(Jc * NB) * f;            // L[rb][Jc ..], k-th column at + k*f
            const T* Bcol = LI + Jc * NB + (int64_t)(Jc * NB) * s;      // Linv[Jc ..][Jc], k-th row at + k
            blk.gemm_stream((rb - Jc) * NB, lane,
                            [&](int i, int k) { return (i < nr) ? Arow[i + (int64_t)k * f] : zero<T>(); },
                            [&](int k, int j) { return Bcol[k + (int64_t)j * s]; });
        }
        blk.each(lane, [&](int i, int j, T v) {
            if (i < nr) LI[(int64_t)(jr + i) + (int64_t)(Jc * NB + j) * s] = v;
        });
        __syncwarp();
        blk.clear();
        blk.gemm(NB, lane, [&](int i, int k) { return Lirb[i * LDP + k]; },
                 [&](int k, int j) { return (k < nr) ? LI[(int64_t)(jr + k) + (int64_t)(Jc * NB + j) * s] : zero<T>(); });
        __syncwarp();
        blk.each(lane, [&](int i, int j, T v) {
            if (i < nr) LI[(int64_t)(jr + i) + (int64_t)(Jc * NB + j) * s] = sub(zero<T>(), v);
        });
    };

    for (int b = 0; b < nbk; ++b) {
        const int jb = b * NB, nb = min(NB, s - jb);
        const int mr = s - jb, mrb = (mr + NB - 1) / NB, mr32 = mrb * NB;
        T* Li = Li0 + (b & 1) * NB * LDP;
        // (1) block column -> shared memory (identity padding in the diagonal block, zero rows below the supernode)
        for (int col = warp; col < NB; col += NW) {
            const T* Pc = P + (int64_t)jb + (int64_t)(jb + col) * f;
            for (int row0 = lane; row0 < mr32; row0 += 32 * 8) {   // eight loads in flight before the first store
                T v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int row = row0 + 32 * q;
                    v[q] = (row == col) ? one<T>() : zero<T>();
                    if (row < mr && col < nb && (row >= NB || col <= row)) v[q] = Pc[row];
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int row = row0 + 32 * q;
                    if (row < mr32) Pn[row * LDP + col] = v[q];
                }
            }
        }
        __syncthreads();
        // (2) 32x32 LDL^T and the inverse of its unit-lower factor by the whole CTA, in shared memory
        for (int r = warp; r < NB; r += NW) Ds[r * LDD + lane] = Pn[r * LDP + lane];
        __syncthreads();
        cta_ldlt32_inv<T, NW, LDD, LDP>(Ds, Li, rd, lane, warp, errflag);
        if (warp == 0) dd[lane] = Ds[lane * LDD + lane];
        __syncthreads();
        // (3) rows below: L_Ib = A_Ib Linv_bb' D_b^-1 (both operands in shared memory, in place)
        for (int I = 1 + warp; I < mrb; I += NW) {
            blk.clear();
            blk.gemm(NB, lane, [&](int i, int k) { return Pn[(I * NB + i) * LDP + k]; },
                     [&](int k, int j) { return Li[j * LDP + k]; });
            __syncwarp();
            blk.each(lane, [&](int i, int j, T v) { Pn[(I * NB + i) * LDP + j] = mul(v, rd[j]); });
        }
        __syncthreads();
        // (4) the finished block column, the diagonal block of the inverse and the pivots go to global memory
        if (crank == 0) {
            for (int col = warp; col < nb; col += NW) {
                T* Pc = P + (int64_t)jb + (int64_t)(jb + col) * f;
                for (int row = lane; row < mr; row += 32)
                    if (row >= NB || col <= row) Pc[row] = (row < NB) ? Ds[row * LDD + col] : Pn[row * LDP + col];
                if (lane < nb) LI[(int64_t)(jb + lane) + (int64_t)(jb + col) * s] = Li[lane * LDP + col];
            }
            if (tid < nb) dv[jb + tid] = dd[tid];
        }
        // (5) work items of this step, dealt out over the warps of the cluster: trailing-update blocks (I, K),
        //     b < K <= I, and the blocks Jc < b-1 of row block b-1 of the inverse (one step behind: they need nothing
        //     of step b, and their longest item balances the shrinking trailing update)
        const int m = mrb - 1, npairs = m * (m + 1) / 2;
        const int ninv = b >= 2 ? b - 1 : 0;
        const T* Liprev = Li0 + ((b - 1) & 1) * NB * LDP;
        for (int p = crank * NW + warp; p < npairs + ninv; p += CL * NW) {
            if (p < ninv) {   // (the long items first)
                inverse_item(b - 1, p, Liprev);
                continue;
            }
            const int pp = p - ninv;
            int Ii = (int)((sqrtf(8.0f * pp + 1.0f) - 1.0f) * 0.5f);
            while ((Ii + 1) * (Ii + 2) / 2 <= pp) ++Ii;
            while (Ii * (Ii + 1) / 2 > pp) --Ii;
            const int Ki = pp - Ii * (Ii + 1) / 2;
            const int I = 1 + Ii, K = 1 + Ki;   // relative to block b
            blk.clear();
            blk.gemm(NB, lane, [&](int i, int k) { return Pn[(I * NB + i) * LDP + k]; },
                     [&](int k, int j) { return mul(Pn[(K * NB + j) * LDP + k], dd[k]); });
            blk.each(lane, [&](int i, int j, T v) {
                const int row = jb + I * NB + i, col = jb + K * NB + j;
                if (row < s && col < s) {
                    T* t = P + ((int64_t)row + (int64_t)col * f);
                    *t = sub(*t, v);
                }
            });
        }
        diag_step_barrier<CL>();
    }
    // row block nbk-1 of the inverse
    if (nbk >= 2) {
        const T* Lilast = Li0 + ((nbk - 1) & 1) * NB * LDP;
        for (int Jc = crank * NW + warp; Jc < nbk - 1; Jc += CL * NW) inverse_item(nbk - 1, Jc, Lilast);
    }
}

template <class T>
static int diag2_smem(int smax) {
    const int cap = (smax + NB - 1) / NB * NB;
    return (int)sizeof(T) * ((cap + 2 * NB) * (NB + 4) + 2 * NB + NB * (NB + 1));
}

// L21 = A21 Linv' D^-1.  CTA = (supernode, 64-row slab of L21), warp = 8 rows x all s columns, computed in
// place from the last column block to the first (block c only reads columns <= c of the warp's own rows).
template <class T>
__global__ void __launch_bounds__(256) k_l21(DevSymbolic S, const int2* __restrict__ items, T* L, const T* Linv,
                                             const T* dvec) {
    constexpr int NT = 4, CPN = MM<T>::CPN, PW = NT * CPN;
    const int2 item = items[blockIdx.x];
    const int J = item.x;
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    T* P = L + S.panel_off[J];
    const T* LI = Linv + S.linv_off[J];
    const T* dv = dvec + S.sn_first[J];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = item.y * 64 + warp * 8;
    if (r0 >= u) return;
    const int row = r0 + (lane >> 2), bc = MM<T>::bcol(lane);
    const T* Prow = P + (s + row);
    for (int cb = (s + PW - 1) / PW - 1; cb >= 0; --cb) {
        const int c0 = cb * PW;
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        strip_mma<T, NT>(acc, 0, min(s, c0 + PW), lane,
                         [&](int k) { return (row < u && k < s) ? Prow[(int64_t)k * f] : zero<T>(); },
                         [&](int k, int nt) {
                             const int col = c0 + nt * CPN + bc;
                             return (col < s && k < s) ? LI[(int64_t)col + (int64_t)k * s] : zero<T>();
                         });
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            MM<T>::each(acc[nt], lane, [&](int c, T v) {
                const int col = c0 + nt * CPN + c;
                if (row < u && col < s) P[(int64_t)(s + row) + (int64_t)col * f] = mul(v, recip(dv[col]));
            });
    }
}

// Schur complement U_J -= L21 D L21' (lower triangle).  CTA = 64x64 tile, 4 warps of 32x32.
template <class T>
__global__ void __launch_bounds__(128) k_schur(DevSymbolic S, const int4* __restrict__ items, const T* __restrict__ L,
                                               const T* __restrict__ dvec, T* U) {
    const int4 item = items[blockIdx.x];
    const int J = item.x;
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    const T* P = L + S.panel_off[J] + s;   // L21, ld f
    const T* dv = dvec + S.sn_first[J];
    T* UJ = U + S.upd_off[J];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i0 = item.y * 64 + (warp >> 1) * 32, j0 = item.z * 64 + (warp & 1) * 32;
    if (i0 >= u || j0 >= u || j0 > i0 + 31) return;
    Blk32<T> blk;
    blk.clear();
    blk.gemm(s, lane,
             [&](int i, int k) {
                 const int r = i0 + i;
                 return (r < u && k < s) ? mul(P[(int64_t)r + (int64_t)k * f], dv[k]) : zero<T>();
             },
             [&](int k, int j) {
                 const int c = j0 + j;
                 return (c < u && k < s) ? P[(int64_t)c + (int64_t)k * f] : zero<T>();
             });
    blk.each(lane, [&](int i, int j, T v) {
        const int r = i0 + i, c = j0 + j;
        if (r < u && c < u && r >= c) {
            T* t = UJ + ((int64_t)r + (int64_t)c * u);
            *t = sub(*t, v);
        }
    });
}

template <class T>
void launch_extend_add(const DevSymbolic& S, const int32_t* parents, int nparents, int gy, T* L, T* U,
                       cudaStream_t st, int64_t* launches) {
    if (nparents <= 0) return;
    dim3 grid(nparents, gy);
    DRE_LAUNCH((k_extend_add<T>), grid, 256, 0, st, S, parents, L, U);
    if (launches) *launches += 1;
}

template <class T>
void launch_diag(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, T* L, T* Linv, T* dvec, int32_t* errflag,
                 cudaStream_t st, int64_t* launches) {
    if (nsns <= 0) return;
    if (diag_variant == 2) {
        // real: 128 registers (2 x 256 or 4 x 128 threads per SM); complex: 255 registers (1 x 256 or 2 x 128).
        // Populous levels (the 1024 leaves) take 128-thread CTAs: the 32x32 steps of a supernode are serial work of
        // one warp, so what counts there is how many supernodes an SM holds at once.  Levels with a handful of fat
        // supernodes (near the root) give each supernode a cluster of 4 CTAs.
        constexpr int MB8 = sizeof(T) == sizeof(double) ? 2 : 1, MB4 = 2 * MB8;
        static bool done[DRE_MAX_DEVICES] = {};
        const int dev = current_device();
        if (!done[dev]) {
            cudaFuncSetAttribute(k_diag2<T, 8, MB8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag2_smem<T>(SN_MAX));
            cudaFuncSetAttribute(k_diag2<T, 4, MB4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag2_smem<T>(SN_MAX));
#ifndef DRE_SIMT_EMU
            cudaFuncSetAttribute(k_diag2<T, 8, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag2_smem<T>(SN_MAX));
#endif
            done[dev] = true;
        }
        const int cap = (smax + NB - 1) / NB * NB;
        bool launched = false;
#ifndef DRE_SIMT_EMU
        if (diag_cluster > 1 && nsns * 4 <= 148 && smax > 2 * NB) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(nsns * 4);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = diag2_smem<T>(smax);
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            launched = cudaLaunchKernelEx(&cfg, k_diag2<T, 8, 1, 4>, S, sns, L, Linv, dvec, errflag, cap) == cudaSuccess;
            if (!launched) cudaGetLastError();
        }
#endif
        if (launched) {
        } else if (nsns >= 2 * 148)
            DRE_LAUNCH((k_diag2<T, 4, MB4, 1>), nsns, 128, (diag2_smem<T>(smax)), st, S, sns, L, Linv, dvec, errflag, cap);
        else
            DRE_LAUNCH((k_diag2<T, 8, MB8, 1>), nsns, 256, (diag2_smem<T>(smax)), st, S, sns, L, Linv, dvec, errflag, cap);
    } else if (nsns >= diag_narrow_min) {
        DRE_LAUNCH((k_diag<T, 2>), nsns, 64, 0, st, S, sns, L, Linv, dvec, errflag);
    } else {
        DRE_LAUNCH((k_diag<T, 8>), nsns, 256, 0, st, S, sns, L, Linv, dvec, errflag);
    }
    if (launches) *launches += 1;
}

template <class T>
void launch_l21(const DevSymbolic& S, const int2* items, int nitems, T* L, const T* Linv, const T* dvec,
                cudaStream_t st, int64_t* launches) {
    if (nitems <= 0) return;
    DRE_LAUNCH((k_l21<T>), nitems, 256, 0, st, S, items, L, Linv, dvec);
    if (launches) *launches += 1;
}

template <class T>
void launch_schur(const DevSymbolic& S, const int4* items, int nitems, const T* L, const T* dvec, T* U,
                  cudaStream_t st, int64_t* launches) {
    if (nitems <= 0) return;
    DRE_LAUNCH((k_schur<T>), nitems, 128, 0, st, S, items, L, dvec, U);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// solves.  W is row-major (n x ldw); the update vector of supernode J is column-major
// (u_J contiguous per RHS column) at  t + rhs_off[J]*ldw.
// forward:  y_J = Linv (b_J + children),  t_J = children - L21 y_J
// backward: x_J = Linv' (D^-1 y_J - L21' x_struct)
// CTA = (supernode, chunk of CW = NT * CPN right-hand sides); 8 warps; warp w owns the 8-row strips
// w, w+8, w+16, w+24 of the supernode (s <= SN_MAX = 256) and keeps their results in registers, so the
// right-hand-side tile in shared memory is updated in place after one barrier.
// ------------------------------------------------------------------------------------------

// SWEEP_Q = strips per warp: 4 covers supernodes up to SN_MAX = 256 columns, 2 (levels whose widest supernode has
// <= 128 columns, i.e. the populous leaf levels) halves the accumulator registers -> one more CTA per SM
// NW = warps per CTA: 8 normally, 16 for the few fat supernodes near the root (half as many strips per warp
// -> half the dependent-load chain per level, which is what those levels cost)
template <class T, int NT, int SWEEP_Q, int NW>
__global__ void __launch_bounds__(32 * NW) k_fwd(DevSymbolic S, const int32_t* __restrict__ sns, const T* __restrict__ L,
                                             const T* __restrict__ Linv, T* W, int64_t ldw, int nrhs, T* tbuf,
                                             RhsSource src) {
    constexpr int CPN = MM<T>::CPN, CW = NT * CPN, LDB = RhsLd<T, CW>::value;
    DRE_DYN_SMEM_ALIGNED(unsigned char, dre_smem_raw);
    T* xs = reinterpret_cast<T*>(dre_smem_raw);   // [s8][LDB]
    const int J = sns[blockIdx.x];
    const int c0 = blockIdx.y * CW, ncw = min(CW, nrhs - c0);
    const int first = S.sn_first[J];
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u, s8 = (s + 7) & ~7;
    const T* P = L + S.panel_off[J];
    const T* LI = Linv + S.linv_off[J];
    T* tJ = tbuf + S.rhs_off[J] * ldw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, bc = MM<T>::bcol(lane);

    // the right-hand side [R, Vt] (real panels) is read where it lies: every row is first touched by the
    // supernode that owns it, so no staging copy into W is needed
    for (int idx = tid; idx < s8 * CW; idx += 32 * NW) {
        const int i = idx / CW, cc = idx - i * CW;
        T v = zero<T>();
        if (i < s && cc < ncw) {
            const int col = c0 + cc;
            const int64_t row = first + i;
            from_real(col < src.r ? src.R[row * src.ldr + col] : src.Vt[row * src.ldv + (col - src.r)], v);
        }
        xs[i * LDB + cc] = v;
    }
    for (int idx = tid; idx < u * ncw; idx += 32 * NW) {
        const int cc = idx / u, i = idx - cc * u;
        tJ[(int64_t)(c0 + cc) * u + i] = zero<T>();
    }
    __syncthreads();
    for (int ci = S.child_ptr[J]; ci < S.child_ptr[J + 1]; ++ci) {
        const int c = S.child_idx[ci];
        const int uc = sn_u(S, c);
        const int32_t* rel = S.relmap + S.sn_rowptr[c];
        const T* tch = tbuf + S.rhs_off[c] * ldw;
        for (int idx = tid; idx < uc * ncw; idx += 32 * NW) {
            const int cc = idx / uc, i = idx - cc * uc;
            const int pr = rel[i];
            const T val = tch[(int64_t)(c0 + cc) * uc + i];
            T* t = (pr < s) ? (xs + pr * LDB + cc) : (tJ + (int64_t)(c0 + cc) * u + (pr - s));
            *t = add(*t, val);
        }
        __syncthreads();
    }
    // y = Linv x (lower triangular): results stay in registers until every warp has read x
    double acc[SWEEP_Q][NT][2];
#pragma unroll
    for (int q = 0; q < SWEEP_Q; ++q) {
        const int i0 = (warp + NW * q) * 8;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;
        if (i0 < s) {
            const int row = i0 + r;
            const T* Lrow = LI + row;
            strip_mma<T, NT>(acc[q], 0, min(i0 + 8, s), lane,
                             [&](int k) { return (row < s && k < s) ? Lrow[(int64_t)k * s] : zero<T>(); },
                             [&](int k, int nt) { return xs[k * LDB + nt * CPN + bc]; });
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < SWEEP_Q; ++q) {
        const int row = (warp + NW * q) * 8 + r;
        if (row < s) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                MM<T>::each(acc[q][nt], lane, [&](int c, T v) { xs[row * LDB + nt * CPN + c] = v; });
        }
    }
    __syncthreads();
    for (int idx = tid; idx < s * CW; idx += 32 * NW) {
        const int i = idx / CW, cc = idx - i * CW;
        if (cc < ncw) W[(int64_t)(first + i) * ldw + c0 + cc] = xs[i * LDB + cc];
    }
    // t_J -= L21 y
    for (int i0 = warp * 8; i0 < u; i0 += 8 * NW) {
        const int row = i0 + r;
        const T* Lrow = P + (s + row);
        double a1[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) a1[nt][0] = a1[nt][1] = 0.0;
        strip_mma<T, NT>(a1, 0, s, lane,
                         [&](int k) { return (row < u && k < s) ? Lrow[(int64_t)k * f] : zero<T>(); },
                         [&](int k, int nt) { return xs[k * LDB + nt * CPN + bc]; });
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            MM<T>::each(a1[nt], lane, [&](int c, T v) {
                const int col = nt * CPN + c;
                if (row < u && col < ncw) {
                    T* t = tJ + (int64_t)(c0 + col) * u + row;
                    *t = sub(*t, v);
                }
            });
    }
}

template <class T, int NT, int SWEEP_Q, int NW>
__global__ void __launch_bounds__(32 * NW) k_bwd(DevSymbolic S, const int32_t* __restrict__ sns, const T* __restrict__ L,
                                             const T* __restrict__ Linv, const T* __restrict__ dvec, T* W, int64_t ldw,
                                             int nrhs, int srows) {
    constexpr int CPN = MM<T>::CPN, CW = NT * CPN, LDB = RhsLd<T, CW>::value;
    DRE_DYN_SMEM_ALIGNED(unsigned char, dre_smem_raw);
    T* xs = reinterpret_cast<T*>(dre_smem_raw);   // [srows][LDB]: z
    T* xt = xs + (size_t)srows * LDB;             // [64][LDB]: tile of x at the structure rows
    const int J = sns[blockIdx.x];
    const int c0 = blockIdx.y * CW, ncw = min(CW, nrhs - c0);
    const int first = S.sn_first[J];
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u, s8 = (s + 7) & ~7;
    const T* P = L + S.panel_off[J];
    const T* LI = Linv + S.linv_off[J];
    const T* dv = dvec + first;
    const int32_t* rows = S.sn_rows + S.sn_rowptr[J];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, bc = MM<T>::bcol(lane);

    double acc[SWEEP_Q][NT][2];
#pragma unroll
    for (int q = 0; q < SWEEP_Q; ++q)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;
    // acc = L21' x_struct, 64 structure rows at a time
    for (int r0 = 0; r0 < u; r0 += 64) {
        for (int idx = tid; idx < 64 * CW; idx += 32 * NW) {
            const int i = idx / CW, cc = idx - i * CW;
            xt[i * LDB + cc] = (r0 + i < u && cc < ncw) ? W[(int64_t)rows[r0 + i] * ldw + c0 + cc] : zero<T>();
        }
        __syncthreads();
        const int kt = min(64, u - r0);
#pragma unroll
        for (int q = 0; q < SWEEP_Q; ++q) {
            const int i0 = (warp + NW * q) * 8;
            if (i0 < s) {
                const int col = i0 + r;   // output row = column of L21
                const T* Lcol = P + (s + r0) + (int64_t)col * f;
                strip_mma<T, NT>(acc[q], 0, kt, lane,
                                 [&](int k) { return (col < s && k < kt) ? Lcol[k] : zero<T>(); },
                                 [&](int k, int nt) { return xt[k * LDB + nt * CPN + bc]; });
            }
        }
        __syncthreads();
    }
    // z = D^-1 y - acc
    for (int idx = tid; idx < (s8 - s) * CW; idx += 32 * NW) {
        const int i = s + idx / CW, cc = idx % CW;
        xs[i * LDB + cc] = zero<T>();
    }
#pragma unroll
    for (int q = 0; q < SWEEP_Q; ++q) {
        const int row = (warp + NW * q) * 8 + r;
        if (row < s) {
            const T rd = recip(dv[row]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                MM<T>::each(acc[q][nt], lane, [&](int c, T v) {
                    const int col = nt * CPN + c;
                    const T y = (col < ncw) ? W[(int64_t)(first + row) * ldw + c0 + col] : zero<T>();
                    xs[row * LDB + col] = sub(mul(y, rd), v);
                });
        }
    }
    __syncthreads();
    // x = Linv' z (upper triangular)
#pragma unroll
    for (int q = 0; q < SWEEP_Q; ++q) {
        const int i0 = (warp + NW * q) * 8;
        if (i0 < s) {
            const int row = i0 + r;
            const T* Lcol = LI + (int64_t)row * s;
            double a1[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) a1[nt][0] = a1[nt][1] = 0.0;
            strip_mma<T, NT>(a1, i0, s, lane,
                             [&](int k) { return (row < s && k < s) ? Lcol[k] : zero<T>(); },
                             [&](int k, int nt) { return xs[k * LDB + nt * CPN + bc]; });
            if (row < s) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    MM<T>::each(a1[nt], lane, [&](int c, T v) {
                        const int col = nt * CPN + c;
                        if (col < ncw) W[(int64_t)(first + row) * ldw + c0 + col] = v;
                    });
            }
        }
    }
}

template <class T, int NT>
static int fwd_smem(int smax) {
    return (int)sizeof(T) * ((smax + 7) & ~7) * RhsLd<T, NT * MM<T>::CPN>::value;
}
template <class T, int NT>
static int bwd_smem(int smax) {
    return (int)sizeof(T) * (((smax + 7) & ~7) + 64) * RhsLd<T, NT * MM<T>::CPN>::value;
}

template <class T, int NT, int Q, int NW>
static void set_sweep_attrs() {
    // function attributes are per device: one flag per device ordinal (several contexts of one process may live on
    // different GPUs)
    static bool done[DRE_MAX_DEVICES] = {};
    const int dev = current_device();
    if (done[dev]) return;
    cudaFuncSetAttribute(k_fwd<T, NT, Q, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem<T, NT>(8 * NW * Q));
    cudaFuncSetAttribute(k_bwd<T, NT, Q, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem<T, NT>(8 * NW * Q));
    done[dev] = true;
}

// widest chunk that still gives every SM a CTA; narrow chunks for the few fat supernodes near the root
template <class T>
static int pick_nt(int nsns, int nrhs, bool narrow) {
    const int cpn = MM<T>::CPN;
    // populous narrow levels: 64-column chunks halve the factor re-reads (each A fragment feeds 8 tiles)
    static const bool use_nt8 = getenv("DRE_NT8") != nullptr;   // experimental (164 registers: 1 CTA/SM)
    if (use_nt8 && narrow && (int64_t)nsns * ((nrhs + 8 * cpn - 1) / (8 * cpn)) >= 2 * 148) return 8;
    if ((int64_t)nsns * ((nrhs + 4 * cpn - 1) / (4 * cpn)) >= 148) return 4;
    if ((int64_t)nsns * ((nrhs + 2 * cpn - 1) / (2 * cpn)) >= 148) return 2;
    return 1;
}

template <class T, int NT, int Q, int NW>
static void launch_fwd_t(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, const T* L, const T* Linv, T* W,
                         int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, cudaStream_t st) {
    set_sweep_attrs<T, NT, Q, NW>();
    const int cw = NT * MM<T>::CPN;
    dim3 grid(nsns, (nrhs + cw - 1) / cw);
    DRE_LAUNCH((k_fwd<T, NT, Q, NW>), grid, 32 * NW, (fwd_smem<T, NT>(smax)), st, S, sns, L, Linv, W, ldw, nrhs, tbuf,
               src);
}

template <class T, int NT, int Q, int NW>
static void launch_bwd_t(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, const T* L, const T* Linv,
                         const T* dvec, T* W, int64_t ldw, int nrhs, cudaStream_t st) {
    set_sweep_attrs<T, NT, Q, NW>();
    const int cw = NT * MM<T>::CPN;
    dim3 grid(nsns, (nrhs + cw - 1) / cw);
    DRE_LAUNCH((k_bwd<T, NT, Q, NW>), grid, 32 * NW, (bwd_smem<T, NT>(smax)), st, S, sns, L, Linv, dvec, W, ldw, nrhs,
               (smax + 7) & ~7);
}

template <class T>
void launch_fwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, const T* L, const T* Linv,
                      T* W, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, cudaStream_t st, int64_t* launches) {
    if (nsns <= 0 || nrhs <= 0) return;
    const bool narrow = smax <= 128;
    const int nt = pick_nt<T>(nsns, nrhs, narrow);
#define DRE_FWD(NT_, Q_, NW_) launch_fwd_t<T, NT_, Q_, NW_>(S, sns, nsns, smax, L, Linv, W, ldw, nrhs, tbuf, src, st)
    if (nt == 8) DRE_FWD(8, 2, 8);
    else if (nt == 4) { if (narrow) DRE_FWD(4, 2, 8); else DRE_FWD(4, 4, 8); }
    else if (nt == 2) { if (narrow) DRE_FWD(2, 2, 8); else DRE_FWD(2, 2, 16); }
    else { if (narrow) DRE_FWD(1, 2, 8); else DRE_FWD(1, 2, 16); }
#undef DRE_FWD
    if (launches) *launches += 1;
}

template <class T>
void launch_bwd_level(const DevSymbolic& S, const int32_t* sns, int nsns, int smax, const T* L, const T* Linv,
                      const T* dvec, T* W, int64_t ldw, int nrhs, cudaStream_t st, int64_t* launches) {
    if (nsns <= 0 || nrhs <= 0) return;
    const bool narrow = smax <= 128;
    const int nt = pick_nt<T>(nsns, nrhs, narrow);
#define DRE_BWD(NT_, Q_, NW_) launch_bwd_t<T, NT_, Q_, NW_>(S, sns, nsns, smax, L, Linv, dvec, W, ldw, nrhs, st)
    if (nt == 8) DRE_BWD(8, 2, 8);
    else if (nt == 4) { if (narrow) DRE_BWD(4, 2, 8); else DRE_BWD(4, 4, 8); }
    else if (nt == 2) { if (narrow) DRE_BWD(2, 2, 8); else DRE_BWD(2, 2, 16); }
    else { if (narrow) DRE_BWD(1, 2, 8); else DRE_BWD(1, 2, 16); }
#undef DRE_BWD
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
    DRE_LAUNCH((k_load_rhs<T>), blocks, 256, 0, st, W, ldw, R, ldr, r, Vt, ldv, m, n);
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
    DRE_LAUNCH((k_smw_core<T>), 1, 256, 0, st, BtW, ldb, m, r, alpha, Sol, errflag);
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

// one warp per row: the m correction coefficients of the row are loaded once, lanes run over the r columns
template <class T, int MR>
__global__ void __launch_bounds__(256) k_smw_apply(const T* __restrict__ W, int64_t ldw, int r, int m,
                                                   const T* __restrict__ Sol, int mode, double d,
                                                   double* __restrict__ V1, int64_t ld1, double* __restrict__ V2,
                                                   int64_t ld2, int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n; row += nwarps) {
        const T* w = W + row * ldw;
        T y[MR];
#pragma unroll
        for (int j = 0; j < MR; ++j) y[j] = (j < m) ? w[r + j] : zero<T>();
        for (int c = lane; c < r; c += 32) {
            T v = w[c];
#pragma unroll
            for (int j = 0; j < MR; ++j)
                if (j < m) v = sub(v, mul(y[j], Sol[(int64_t)j * r + c]));
            emit(v, mode, d, V1, V2, row * ld1 + c, row * ld2 + c);
        }
    }
}

template <class T>
void launch_smw_apply(const T* W, int64_t ldw, int r, int m, const T* Sol, int mode, double d, double* V1,
                      int64_t ld1, double* V2, int64_t ld2, int64_t n, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || r <= 0) return;
    int blocks = (int)std::min<int64_t>((n + 7) / 8, 148 * 16);
    if (m <= 8) DRE_LAUNCH((k_smw_apply<T, 8>), blocks, 256, 0, st, W, ldw, r, m, Sol, mode, d, V1, ld1, V2, ld2, n);
    else DRE_LAUNCH((k_smw_apply<T, 32>), blocks, 256, 0, st, W, ldw, r, m, Sol, mode, d, V1, ld1, V2, ld2, n);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// Row-split sweeps ("sweep v2", DRE_SWEEP2=1).
// The per-level sweeps above give one CTA all rows of a supernode and run two dependent phases
// (y = Linv x, then t -= L21 y), so the few fat supernodes near the root of the tree occupy a handful of SMs for
// a long dependent chain.  With  M21 = L21 Linv  (k_m21, after the Schur complement of the level, in place
// in the panel) both products act on the same vector,
//   forward : [ y' ; t ] = [ D^-1 Linv ; -M21 ] (b + children)   (+ children's contributions to t)
//   backward: x = Linv' y' - M21' x_struct,
// every 8-row strip is independent, and a supernode's strips are spread over several CTAs (items (J, rb)).
// The forward sweep writes y' into a second block Y (the backward sweep of another row block of the same
// supernode must still find y' after this one has written its x into W).
// ------------------------------------------------------------------------------------------

// M21 = L21 Linv, in place.  CTA = (supernode, 64-row slab of L21), warp = 8 rows x all s columns, column blocks
// ascending (block c only reads columns >= c of the warp's own rows).
template <class T>
__global__ void __launch_bounds__(256) k_m21(DevSymbolic S, const int2* __restrict__ items, T* L, const T* Linv) {
    constexpr int NT = 4, CPN = MM<T>::CPN, PW = NT * CPN;
    const int2 item = items[blockIdx.x];
    const int J = item.x;
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u;
    T* P = L + S.panel_off[J];
    const T* LI = Linv + S.linv_off[J];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = item.y * 64 + warp * 8;
    if (r0 >= u) return;
    const int row = r0 + (lane >> 2), bc = MM<T>::bcol(lane);
    const T* Prow = P + (s + row);
    for (int c0 = 0; c0 < s; c0 += PW) {
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        strip_mma<T, NT>(acc, c0, s, lane,
                         [&](int k) { return (row < u && k < s) ? Prow[(int64_t)k * f] : zero<T>(); },
                         [&](int k, int nt) {
                             const int col = c0 + nt * CPN + bc;
                             return (col < s && k < s) ? LI[(int64_t)k + (int64_t)col * s] : zero<T>();
                         });
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
            MM<T>::each(acc[nt], lane, [&](int c, T v) {
                const int col = c0 + nt * CPN + c;
                if (row < u && col < s) P[(int64_t)(s + row) + (int64_t)col * f] = v;
            });
    }
}

template <class T>
void launch_m21(const DevSymbolic& S, const int2* items, int nitems, T* L, const T* Linv, cudaStream_t st,
                int64_t* launches) {
    if (nitems <= 0) return;
    DRE_LAUNCH((k_m21<T>), nitems, 256, 0, st, S, items, L, Linv);
    if (launches) *launches += 1;
}

// strips of supernode J: sy = ceil(s/8) strips of y rows, then ceil(u/8) strips of update rows
// CTA = (item (J, rb), chunk of CW right-hand sides); 8 warps x Q strips = strips [rb*8Q, (rb+1)*8Q)
template <class T, int NT, int Q>
__global__ void __launch_bounds__(256) k_fwd2(DevSymbolic S, const int2* __restrict__ items, const T* __restrict__ L,
                                              const T* __restrict__ Linv, const T* __restrict__ dvec, T* Y, int64_t ldw,
                                              int nrhs, T* tbuf, RhsSource src, int srows, int has_children) {
    constexpr int CPN = MM<T>::CPN, CW = NT * CPN, LDB = RhsLd<T, CW>::value, NW = 8, NS = NW * Q;
    DRE_DYN_SMEM_ALIGNED(unsigned char, dre_smem_raw);
    T* xs = reinterpret_cast<T*>(dre_smem_raw);   // [srows][LDB]: x = b + children
    T* ts = xs + (size_t)srows * LDB;             // [8*NS][LDB]: children's contributions to this block's update rows
    const int2 item = items[blockIdx.x];
    const int J = item.x;
    const int c0 = blockIdx.y * CW, ncw = min(CW, nrhs - c0);
    const int first = S.sn_first[J];
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u, s8 = (s + 7) & ~7, sy = s8 >> 3;
    const int strip0 = item.y * NS;                         // first strip of this block
    const int t0 = max(0, strip0 - sy) * 8;                 // first update row covered by this block
    const int t1 = min(u, max(0, strip0 + NS - sy) * 8);    // one past the last
    const T* P = L + S.panel_off[J];
    const T* LI = Linv + S.linv_off[J];
    T* tJ = tbuf + S.rhs_off[J] * ldw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, bc = MM<T>::bcol(lane);

    for (int idx = tid; idx < s8 * CW; idx += 256) {
        const int i = idx / CW, cc = idx - i * CW;
        T v = zero<T>();
        if (i < s && cc < ncw) {
            const int col = c0 + cc;
            const int64_t row = first + i;
            from_real(col < src.r ? src.R[row * src.ldr + col] : src.Vt[row * src.ldv + (col - src.r)], v);
        }
        xs[i * LDB + cc] = v;
    }
    if (has_children) {
        for (int idx = tid; idx < 8 * NS * CW; idx += 256) {
            const int i = idx / CW, cc = idx - i * CW;
            ts[i * LDB + cc] = zero<T>();
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
                if (pr < s) {
                    T* t = xs + pr * LDB + cc;
                    *t = add(*t, tch[(int64_t)(c0 + cc) * uc + i]);
                } else if (pr - s >= t0 && pr - s < t1) {
                    T* t = ts + (pr - s - t0) * LDB + cc;
                    *t = add(*t, tch[(int64_t)(c0 + cc) * uc + i]);
                }
            }
            __syncthreads();
        }
    } else {
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int strip = strip0 + warp + NW * q;
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        if (strip < sy) {
            // y' = D^-1 Linv x (lower triangular)
            const int i0 = strip * 8, row = i0 + r;
            const T* Lrow = LI + row;
            strip_mma<T, NT>(acc, 0, min(i0 + 8, s), lane,
                             [&](int k) { return (row < s && k < s) ? Lrow[(int64_t)k * s] : zero<T>(); },
                             [&](int k, int nt) { return xs[k * LDB + nt * CPN + bc]; });
            if (row < s) {
                const T rd = recip(dvec[first + row]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    MM<T>::each(acc[nt], lane, [&](int c, T v) {
                        const int col = nt * CPN + c;
                        if (col < ncw) Y[(int64_t)(first + row) * ldw + c0 + col] = mul(v, rd);
                    });
            }
        } else {
            // t = children - M21 x
            const int i0 = (strip - sy) * 8, row = i0 + r;
            if (i0 < u) {
                const T* Lrow = P + (s + row);
                strip_mma<T, NT>(acc, 0, s, lane,
                                 [&](int k) { return (row < u && k < s) ? Lrow[(int64_t)k * f] : zero<T>(); },
                                 [&](int k, int nt) { return xs[k * LDB + nt * CPN + bc]; });
                if (row < u) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        MM<T>::each(acc[nt], lane, [&](int c, T v) {
                            const int col = nt * CPN + c;
                            if (col < ncw) {
                                const T base = has_children ? ts[(row - t0) * LDB + col] : zero<T>();
                                tJ[(int64_t)(c0 + col) * u + row] = sub(base, v);
                            }
                        });
                }
            }
        }
    }
}

// CTA = (item (J, rb), chunk): strips [rb*8Q, (rb+1)*8Q) of the s rows of supernode J
template <class T, int NT, int Q>
__global__ void __launch_bounds__(256) k_bwd2(DevSymbolic S, const int2* __restrict__ items, const T* __restrict__ L,
                                              const T* __restrict__ Linv, const T* __restrict__ Y, T* W, int64_t ldw,
                                              int nrhs, int srows) {
    constexpr int CPN = MM<T>::CPN, CW = NT * CPN, LDB = RhsLd<T, CW>::value, NW = 8, NS = NW * Q;
    DRE_DYN_SMEM_ALIGNED(unsigned char, dre_smem_raw);
    T* xs = reinterpret_cast<T*>(dre_smem_raw);   // [srows][LDB]: y' (rows >= the first strip of this block)
    T* xt = xs + (size_t)srows * LDB;             // [64][LDB]: tile of x at the structure rows
    const int2 item = items[blockIdx.x];
    const int J = item.x;
    const int c0 = blockIdx.y * CW, ncw = min(CW, nrhs - c0);
    const int first = S.sn_first[J];
    const int s = sn_s(S, J), u = sn_u(S, J), f = s + u, s8 = (s + 7) & ~7;
    const int strip0 = item.y * NS;
    const T* P = L + S.panel_off[J];
    const T* LI = Linv + S.linv_off[J];
    const int32_t* rows = S.sn_rows + S.sn_rowptr[J];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = lane >> 2, bc = MM<T>::bcol(lane);

    double acc[Q][NT][2];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[q][nt][0] = acc[q][nt][1] = 0.0;
    // y' rows [8*strip0, s) (the upper-triangular product of a strip only needs the rows below its first one)
    for (int idx = tid + strip0 * 8 * CW; idx < s8 * CW; idx += 256) {
        const int i = idx / CW, cc = idx - i * CW;
        xs[i * LDB + cc] = (i < s && cc < ncw) ? Y[(int64_t)(first + i) * ldw + c0 + cc] : zero<T>();
    }
    // acc = M21' x_struct, 64 structure rows at a time
    for (int r0 = 0; r0 < u; r0 += 64) {
        if (r0) __syncthreads();
        for (int idx = tid; idx < 64 * CW; idx += 256) {
            const int i = idx / CW, cc = idx - i * CW;
            xt[i * LDB + cc] = (r0 + i < u && cc < ncw) ? W[(int64_t)rows[r0 + i] * ldw + c0 + cc] : zero<T>();
        }
        __syncthreads();
        const int kt = min(64, u - r0);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int i0 = (strip0 + warp + NW * q) * 8;
            if (i0 < s) {
                const int col = i0 + r;   // output row = column of M21
                const T* Lcol = P + (s + r0) + (int64_t)col * f;
                strip_mma<T, NT>(acc[q], 0, kt, lane,
                                 [&](int k) { return (col < s && k < kt) ? Lcol[k] : zero<T>(); },
                                 [&](int k, int nt) { return xt[k * LDB + nt * CPN + bc]; });
            }
        }
    }
    __syncthreads();
    // x = Linv' y' - acc
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int i0 = (strip0 + warp + NW * q) * 8;
        if (i0 < s) {
            const int row = i0 + r;
            const T* Lcol = LI + (int64_t)row * s;
            double a1[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) a1[nt][0] = a1[nt][1] = 0.0;
            strip_mma<T, NT>(a1, i0, s, lane,
                             [&](int k) { return (row < s && k < s) ? Lcol[k] : zero<T>(); },
                             [&](int k, int nt) { return xs[k * LDB + nt * CPN + bc]; });
            if (row < s) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    double d2[2] = {a1[nt][0] - acc[q][nt][0], a1[nt][1] - acc[q][nt][1]};
                    MM<T>::each(d2, lane, [&](int c, T v) {
                        const int col = nt * CPN + c;
                        if (col < ncw) W[(int64_t)(first + row) * ldw + c0 + col] = v;
                    });
                }
            }
        }
    }
}

template <class T, int NT, int Q>
static void launch_fwd2_t(const DevSymbolic& S, const int2* items, int nitems, int smax, const T* L, const T* Linv,
                          const T* dvec, T* Y, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, int has_children,
                          cudaStream_t st) {
    constexpr int LDB = RhsLd<T, NT * MM<T>::CPN>::value;
    const int srows = (smax + 7) & ~7;
    const int smem = (int)sizeof(T) * (srows + (has_children ? 64 * Q : 0)) * LDB;
    static int smem_set[DRE_MAX_DEVICES] = {};
    const int dev = current_device();
    if (smem > smem_set[dev]) {
        cudaFuncSetAttribute(k_fwd2<T, NT, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        smem_set[dev] = smem;
    }
    const int cw = NT * MM<T>::CPN;
    dim3 grid(nitems, (nrhs + cw - 1) / cw);
    DRE_LAUNCH((k_fwd2<T, NT, Q>), grid, 256, smem, st, S, items, L, Linv, dvec, Y, ldw, nrhs, tbuf, src, srows,
               has_children);
}

template <class T, int NT, int Q>
static void launch_bwd2_t(const DevSymbolic& S, const int2* items, int nitems, int smax, const T* L, const T* Linv,
                          const T* Y, T* W, int64_t ldw, int nrhs, cudaStream_t st) {
    constexpr int LDB = RhsLd<T, NT * MM<T>::CPN>::value;
    const int srows = (smax + 7) & ~7;
    const int smem = (int)sizeof(T) * (srows + 64) * LDB;
    static int smem_set[DRE_MAX_DEVICES] = {};
    const int dev = current_device();
    if (smem > smem_set[dev]) {
        cudaFuncSetAttribute(k_bwd2<T, NT, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        smem_set[dev] = smem;
    }
    const int cw = NT * MM<T>::CPN;
    dim3 grid(nitems, (nrhs + cw - 1) / cw);
    DRE_LAUNCH((k_bwd2<T, NT, Q>), grid, 256, smem, st, S, items, L, Linv, Y, W, ldw, nrhs, srows);
}

// q = strips per warp the item list was built for (schedule.h: Sweep2Level), wide = 8-tile chunks
template <class T>
void launch_fwd2_level(const DevSymbolic& S, const int2* items, int nitems, int q, int smax, const T* L, const T* Linv,
                       const T* dvec, T* Y, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, int has_children,
                       cudaStream_t st, int64_t* launches) {
    if (nitems <= 0 || nrhs <= 0) return;
    const int cpn = MM<T>::CPN;
    const bool nt4 = q == 3 || (int64_t)nitems * ((nrhs + 4 * cpn - 1) / (4 * cpn)) >= 148;
#define DRE_FWD2(NT_, Q_) \
    launch_fwd2_t<T, NT_, Q_>(S, items, nitems, smax, L, Linv, dvec, Y, ldw, nrhs, tbuf, src, has_children, st)
    if (q == 3) DRE_FWD2(4, 3);
    else if (nt4) DRE_FWD2(4, 1);
    else DRE_FWD2(2, 1);
#undef DRE_FWD2
    if (launches) *launches += 1;
}

template <class T>
void launch_bwd2_level(const DevSymbolic& S, const int2* items, int nitems, int q, int smax, const T* L, const T* Linv,
                       const T* Y, T* W, int64_t ldw, int nrhs, cudaStream_t st, int64_t* launches) {
    if (nitems <= 0 || nrhs <= 0) return;
    const int cpn = MM<T>::CPN;
    const bool nt4 = q == 2 || (int64_t)nitems * ((nrhs + 4 * cpn - 1) / (4 * cpn)) >= 148;
#define DRE_BWD2(NT_, Q_) launch_bwd2_t<T, NT_, Q_>(S, items, nitems, smax, L, Linv, Y, W, ldw, nrhs, st)
    if (q == 2) DRE_BWD2(4, 2);
    else if (nt4) DRE_BWD2(4, 1);
    else DRE_BWD2(2, 1);
#undef DRE_BWD2
    if (launches) *launches += 1;
}

// ---- explicit instantiations ----
#define DRE_INST(T)                                                                                                  \
    template void launch_assemble<T>(const DevSymbolic&, T*, double, T, cudaStream_t, int64_t*, const double*);                     \
    template void launch_extend_add<T>(const DevSymbolic&, const int32_t*, int, int, T*, T*, cudaStream_t,           \
                                       int64_t*);                                                                    \
    template void launch_diag<T>(const DevSymbolic&, const int32_t*, int, int, T*, T*, T*, int32_t*, cudaStream_t,        \
                                 int64_t*);                                                                          \
    template void launch_l21<T>(const DevSymbolic&, const int2*, int, T*, const T*, const T*, cudaStream_t,          \
                                int64_t*);                                                                           \
    template void launch_schur<T>(const DevSymbolic&, const int4*, int, const T*, const T*, T*, cudaStream_t,        \
                                  int64_t*);                                                                         \
    template void launch_fwd_level<T>(const DevSymbolic&, const int32_t*, int, int, const T*, const T*, T*, int64_t, \
                                      int, T*, const RhsSource&, cudaStream_t, int64_t*);                            \
    template void launch_bwd_level<T>(const DevSymbolic&, const int32_t*, int, int, const T*, const T*, const T*,    \
                                      T*, int64_t, int, cudaStream_t, int64_t*);                                     \
    template void launch_load_rhs<T>(T*, int64_t, const double*, int64_t, int, const double*, int64_t, int, int64_t, \
                                     cudaStream_t, int64_t*);                                                        \
    template void launch_smw_core<T>(const T*, int64_t, int, int, double, T*, int32_t*, cudaStream_t, int64_t*);     \
    template void launch_smw_apply<T>(const T*, int64_t, int, int, const T*, int, double, double*, int64_t, double*, \
                                      int64_t, int64_t, cudaStream_t, int64_t*);                                     \
    template void launch_m21<T>(const DevSymbolic&, const int2*, int, T*, const T*, cudaStream_t, int64_t*);         \
    template void launch_fwd2_level<T>(const DevSymbolic&, const int2*, int, int, int, const T*, const T*, const T*, \
                                       T*, int64_t, int, T*, const RhsSource&, int, cudaStream_t, int64_t*);         \
    template void launch_bwd2_level<T>(const DevSymbolic&, const int2*, int, int, int, const T*, const T*, const T*, \
                                       T*, int64_t, int, cudaStream_t, int64_t*);
DRE_INST(double)
DRE_INST(cplx)

}  // namespace dre
