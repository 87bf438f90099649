// Dense tall-skinny toolbox for the low-rank algebra of the ADI hot path:
//   Gram products X'Y on the FP64 tensor cores (DMMA m8n8k4), deterministic split reduction,
//   tall GEMM Y = beta Y + alpha X W (DMMA), rank-revealing pivoted-Cholesky panel selection,
//   diagonal-core Frobenius norm, layout helpers.
// Replaces LAPACK geqp3 / GEMM inside  norm(::LDLt) (src/LDLt.jl:77-89), compress! (:204-225),
// orthf (:237-245), orth (src/Stuff.jl:13-18), restrict (src/Stuff.jl:9) of the reference.
// Panels are row-major (row = state index in solver ordering, columns contiguous).
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace dre {

// launch_spmm: 1 = k_spmm (default), 2 = k_spmm2 (DRE_SPMM2=1)
int spmm_variant = (getenv("DRE_SPMM2") && atoi(getenv("DRE_SPMM2")) != 0) ? 2 : 1;


// ------------------------------------------------------------------------------------------
// Gram:  partial[s] (a x b) = sum_{rows in split s}  w[row] * X[row][:]^T Y[row][:]
// CTA = 256 threads = 8 warps, output tile 64 x 64, k-chunks of 16 rows.
// warp w owns output rows [8w, 8w+8) x 64 columns = 8 DMMA accumulators.
// ------------------------------------------------------------------------------------------
constexpr int GT = 64;    // tile edge
constexpr int GK = 16;    // rows per chunk
constexpr int GLD = 72;   // smem leading dimension (72 mod 16 == 8 -> 2 wavefronts per fragment load, the minimum)

__global__ void __launch_bounds__(256) k_gram(const double* __restrict__ X, int64_t ldx, int a,
                                              const double* __restrict__ Y, int64_t ldy, int b, int64_t n,
                                              const double* __restrict__ roww, double* __restrict__ partial,
                                              int tiles_b, int64_t rows_per_split) {
    __shared__ double Xs[2][GK][GLD];
    __shared__ double Ys[2][GK][GLD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ta = blockIdx.x / tiles_b, tb = blockIdx.x % tiles_b;
    const int a0 = ta * GT, b0 = tb * GT;
    const int64_t row_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t row_end = min(n, row_begin + rows_per_split);

    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;

    const int lc = tid & 63;  // column inside the tile handled by this thread when loading
    const int lr = tid >> 6;  // 0..3
    double xr[4], yr[4];
    auto gload = [&](int64_t row0) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int64_t row = row0 + lr + 4 * it;
            double xv = 0.0, yv = 0.0;
            if (row < row_end) {
                if (a0 + lc < a) {
                    xv = X[row * ldx + a0 + lc];
                    if (roww) xv *= roww[row];
                }
                if (b0 + lc < b) yv = Y[row * ldy + b0 + lc];
            }
            xr[it] = xv;
            yr[it] = yv;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            Xs[buf][lr + 4 * it][lc] = xr[it];
            Ys[buf][lr + 4 * it][lc] = yr[it];
        }
    };

    int buf = 0;
    if (row_begin < row_end) {
        gload(row_begin);
        sstore(0);
    }
    __syncthreads();
    for (int64_t row0 = row_begin; row0 < row_end; row0 += GK) {
        const bool more = row0 + GK < row_end;
        if (more) gload(row0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            const double af = Xs[buf][kk + (lane & 3)][8 * warp + (lane >> 2)];
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const double bf = Ys[buf][kk + (lane & 3)][8 * nb + (lane >> 2)];
                dmma884(acc[nb][0], acc[nb][1], af, bf);
            }
        }
        if (more) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    double* P = partial + (int64_t)blockIdx.y * a * b;
    const int i = a0 + 8 * warp + (lane >> 2);
    if (i < a) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int j = b0 + 8 * nb + 2 * (lane & 3);
            if (j < b) P[(int64_t)i * b + j] = acc[nb][0];
            if (j + 1 < b) P[(int64_t)i * b + j + 1] = acc[nb][1];
        }
    }
}

// Each warp reduces 4 consecutive output elements: lane = (split lane 0..7) x (element 0..3), so every
// load instruction reads full 32-byte sectors; fixed split assignment + fixed-shape shuffle tree keep the
// result deterministic, and hundreds of splits are no longer summed in one serial loop.
__global__ void __launch_bounds__(256) k_reduce_partials(const double* __restrict__ partial, int nsplit, int a,
                                                         int b, double* out1, int64_t ld1, double* out2,
                                                         int64_t ld2) {
    const int64_t total = (int64_t)a * b;
    const int lane = threadIdx.x & 31;
    const int e = lane & 3, kl = lane >> 2;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp0 * 4; base < total; base += nwarps * 4) {
        const int64_t idx = base + e;
        double s0 = 0.0, s1 = 0.0;
        if (idx < total) {
            int k = kl;
            for (; k + 8 < nsplit; k += 16) {
                s0 += partial[(int64_t)k * total + idx];
                s1 += partial[(int64_t)(k + 8) * total + idx];
            }
            if (k < nsplit) s0 += partial[(int64_t)k * total + idx];
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (kl == 0 && idx < total) {
            const int i = (int)(idx / b), j = (int)(idx % b);
            if (out1) out1[(int64_t)i * ld1 + j] = s;
            if (out2) out2[(int64_t)i * ld2 + j] += s;
        }
    }
}


// ------------------------------------------------------------------------------------------
// cp.async helpers (LDGSTS): 16-byte (.cg) and 8-byte (.ca) global -> shared copies with zero fill
// ------------------------------------------------------------------------------------------
#ifdef DRE_SIMT_EMU
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    simt::cp_async(smem, gmem, 16, src_bytes);
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    simt::cp_async(smem, gmem, 8, src_bytes);
}
__device__ __forceinline__ void cp_async_commit() { simt::cp_async_commit(); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { simt::cp_async_wait(N); }
#else
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif

// ------------------------------------------------------------------------------------------
// Gram, pipelined:  partial[split] (a x b) = sum_{rows in split} X[row][:]^T Y[row][:]
// CTA = 256 threads, output tile 64 (a) x 128 (b), 16-row chunks in a 3-stage cp.async pipeline;
// warp tile 32 x 32 (4 x 4 DMMA tiles: 8 fragment loads feed 16 DMMAs).
// ------------------------------------------------------------------------------------------
constexpr int G2A = 64, G2B = 128, G2K = 16, G2ST = 3;
constexpr int G2LDA = G2A + 8;   // = 8 mod 16: two-wavefront fragment loads

// TB = 128: 64 x 128 output tile (warp tile 32 x 32);  TB = 64: 64 x 64 tile for the many 64-column panels
template <int TB>
__global__ void __launch_bounds__(256) k_gram2(const double* __restrict__ X, int64_t ldx, int a,
                                               const double* __restrict__ Y, int64_t ldy, int b, int64_t n,
                                               double* __restrict__ partial, int tiles_b, int64_t rows_per_split,
                                               int aligned) {
    DRE_DYN_SMEM_ALIGNED(double, g2_smem);
    constexpr int G2B = TB, G2LDB = TB + 8, G2_STAGE = G2K * (G2LDA + G2LDB), NT = TB / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ta = blockIdx.x / tiles_b, tb = blockIdx.x % tiles_b;
    const int a0 = ta * G2A, b0 = tb * G2B;
    const int64_t row_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t row_end = min(n, row_begin + rows_per_split);
    const int nchunks = (int)((row_end - row_begin + G2K - 1) / G2K);

    auto load_stage = [&](int chunk, int st) {
        double* Xs = g2_smem + st * G2_STAGE;
        double* Ys = Xs + G2K * G2LDA;
        const int64_t r0 = row_begin + (int64_t)chunk * G2K;
        if (aligned) {
            for (int c = tid; c < G2K * (G2A / 2); c += 256) {      // 16 rows x 32 chunks of 2 doubles
                const int r = c / (G2A / 2), cc = (c % (G2A / 2)) * 2;
                const int64_t row = r0 + r;
                int nb = 0;
                if (row < row_end) nb = 8 * max(0, min(2, a - (a0 + cc)));
                cp_async16(Xs + r * G2LDA + cc, nb ? (const void*)(X + row * ldx + a0 + cc) : (const void*)X, nb);
            }
            for (int c = tid; c < G2K * (G2B / 2); c += 256) {
                const int r = c / (G2B / 2), cc = (c % (G2B / 2)) * 2;
                const int64_t row = r0 + r;
                int nb = 0;
                if (row < row_end) nb = 8 * max(0, min(2, b - (b0 + cc)));
                cp_async16(Ys + r * G2LDB + cc, nb ? (const void*)(Y + row * ldy + b0 + cc) : (const void*)Y, nb);
            }
        } else {
            for (int c = tid; c < G2K * G2A; c += 256) {
                const int r = c / G2A, cc = c % G2A;
                const int64_t row = r0 + r;
                const bool ok = row < row_end && a0 + cc < a;
                cp_async8(Xs + r * G2LDA + cc, ok ? (const void*)(X + row * ldx + a0 + cc) : (const void*)X, ok ? 8 : 0);
            }
            for (int c = tid; c < G2K * G2B; c += 256) {
                const int r = c / G2B, cc = c % G2B;
                const int64_t row = r0 + r;
                const bool ok = row < row_end && b0 + cc < b;
                cp_async8(Ys + r * G2LDB + cc, ok ? (const void*)(Y + row * ldy + b0 + cc) : (const void*)Y, ok ? 8 : 0);
            }
        }
    };

    double acc[4][NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int wa = (warp & 1) * 32, wb = (warp >> 1) * (NT * 8);
    const int kk = lane & 3, rr = lane >> 2;

#pragma unroll
    for (int st = 0; st < G2ST - 1; ++st) {
        if (st < nchunks) load_stage(st, st);
        cp_async_commit();
    }
    for (int it = 0; it < nchunks; ++it) {
        cp_async_wait<G2ST - 2>();
        __syncthreads();
        if (it + G2ST - 1 < nchunks) load_stage(it + G2ST - 1, (it + G2ST - 1) % G2ST);
        cp_async_commit();
        const double* Xs = g2_smem + (it % G2ST) * G2_STAGE;
        const double* Ys = Xs + G2K * G2LDA;
#pragma unroll
        for (int k0 = 0; k0 < G2K; k0 += 4) {
            double af[4], bf[NT];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[mt] = Xs[(k0 + kk) * G2LDA + wa + mt * 8 + rr];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bf[nt] = Ys[(k0 + kk) * G2LDB + wb + nt * 8 + rr];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();
    double* P = partial + (int64_t)blockIdx.y * a * b;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int i = a0 + wa + mt * 8 + rr;
        if (i >= a) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int j = b0 + wb + nt * 8 + 2 * kk;
            if (j < b) P[(int64_t)i * b + j] = acc[mt][nt][0];
            if (j + 1 < b) P[(int64_t)i * b + j + 1] = acc[mt][nt][1];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tall GEMM, pipelined:  Y[n x b] = beta*Y + alpha * X[n x a] * W
// CTA = 256 threads, output tile 128 rows x 64 cols, K chunks of 16 in a 3-stage cp.async pipeline;
// warp tile 32 x 32.
// ------------------------------------------------------------------------------------------
constexpr int T2M = 128, T2K = 16, T2ST = 3;
constexpr int T2LDX = T2K + 4;   // = 4 mod 16
template <int T2N> struct T2Cfg {
    static constexpr int LDW = ((T2N + 15) & ~15) + 8;   // = 8 mod 16
    static constexpr int STAGE = T2M * T2LDX + T2K * LDW;
    static constexpr int WC = T2N >= 64 ? 2 : 1, WR = 8 / WC;   // warps along columns / rows
    static constexpr int MT = T2M / (WR * 8), NT = T2N / (WC * 8);
};

// T2N = 64: warp tile 32 x 32;  T2N = 16 (a handful of candidate directions): warp tile 16 x 16
template <int T2N>
__global__ void __launch_bounds__(256) k_tall_gemm2(double alpha, const double* __restrict__ X, int64_t ldx, int a,
                                                    const double* __restrict__ W, int64_t ldw, int w_trans,
                                                    double beta, double* __restrict__ Y, int64_t ldy, int b,
                                                    int64_t n, int aligned_x, int aligned_w) {
    DRE_DYN_SMEM_ALIGNED(double, t2_smem);
    constexpr int T2LDW = T2Cfg<T2N>::LDW, T2_STAGE = T2Cfg<T2N>::STAGE;
    constexpr int MT = T2Cfg<T2N>::MT, NT = T2Cfg<T2N>::NT, WR = T2Cfg<T2N>::WR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * T2M;
    const int b0 = blockIdx.y * T2N;
    const int nchunks = (a + T2K - 1) / T2K;

    auto load_stage = [&](int chunk, int st) {
        double* Xs = t2_smem + st * T2_STAGE;
        double* Ws = Xs + T2M * T2LDX;
        const int k0 = chunk * T2K;
        if (aligned_x) {
            for (int c = tid; c < T2M * (T2K / 2); c += 256) {   // 128 rows x 8 chunks of 2 doubles
                const int r = c / (T2K / 2), kc = (c % (T2K / 2)) * 2;
                const int64_t row = row0 + r;
                int nb = 0;
                if (row < n) nb = 8 * max(0, min(2, a - (k0 + kc)));
                cp_async16(Xs + r * T2LDX + kc, nb ? (const void*)(X + row * ldx + k0 + kc) : (const void*)X, nb);
            }
        } else {
            for (int c = tid; c < T2M * T2K; c += 256) {
                const int r = c / T2K, kc = c % T2K;
                const int64_t row = row0 + r;
                const bool ok = row < n && k0 + kc < a;
                cp_async8(Xs + r * T2LDX + kc, ok ? (const void*)(X + row * ldx + k0 + kc) : (const void*)X, ok ? 8 : 0);
            }
        }
        if (!w_trans && aligned_w) {
            for (int c = tid; c < T2K * (T2N / 2); c += 256) {
                const int k = c / (T2N / 2), j = (c % (T2N / 2)) * 2;
                int nb = 0;
                if (k0 + k < a) nb = 8 * max(0, min(2, b - (b0 + j)));
                cp_async16(Ws + k * T2LDW + j, nb ? (const void*)(W + (int64_t)(k0 + k) * ldw + b0 + j) : (const void*)W, nb);
            }
        } else {
            for (int c = tid; c < T2K * T2N; c += 256) {
                int k, j;
                if (w_trans) { k = c % T2K; j = c / T2K; } else { k = c / T2N; j = c % T2N; }
                const bool ok = k0 + k < a && b0 + j < b;
                const double* src = w_trans ? W + (int64_t)(b0 + j) * ldw + k0 + k : W + (int64_t)(k0 + k) * ldw + b0 + j;
                cp_async8(Ws + k * T2LDW + j, ok ? (const void*)src : (const void*)W, ok ? 8 : 0);
            }
        }
    };

    double acc[MT][NT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int wr = (warp % WR) * (MT * 8), wc = (warp / WR) * (NT * 8);
    const int kk = lane & 3, rr = lane >> 2;

#pragma unroll
    for (int st = 0; st < T2ST - 1; ++st) {
        if (st < nchunks) load_stage(st, st);
        cp_async_commit();
    }
    for (int it = 0; it < nchunks; ++it) {
        cp_async_wait<T2ST - 2>();
        __syncthreads();
        if (it + T2ST - 1 < nchunks) load_stage(it + T2ST - 1, (it + T2ST - 1) % T2ST);
        cp_async_commit();
        const double* Xs = t2_smem + (it % T2ST) * T2_STAGE;
        const double* Ws = Xs + T2M * T2LDX;
#pragma unroll
        for (int k0 = 0; k0 < T2K; k0 += 4) {
            double af[MT], bf[NT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) af[mt] = Xs[(wr + mt * 8 + rr) * T2LDX + k0 + kk];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bf[nt] = Ws[(k0 + kk) * T2LDW + wc + nt * 8 + rr];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int64_t row = row0 + wr + mt * 8 + rr;
        if (row >= n) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int j = b0 + wc + nt * 8 + 2 * kk;
            double* y = Y + row * ldy + j;
            if (j < b) y[0] = (beta == 0.0 ? 0.0 : beta * y[0]) + alpha * acc[mt][nt][0];
            if (j + 1 < b) y[1] = (beta == 0.0 ? 0.0 : beta * y[1]) + alpha * acc[mt][nt][1];
        }
    }
}

GramPlan gram_plan(int64_t n, int a, int b, int sm_count, int waves_override) {
    GramPlan p;
    const int tbw = b <= 64 ? 64 : G2B;
    int tiles = ((a + G2A - 1) / G2A) * ((b + tbw - 1) / tbw);   // pipelined kernel: 64 x 128 (64 x 64) output tiles
    if (tiles < 1) tiles = 1;
    int64_t chunks = (n + GK - 1) / GK;
    // DRE_GRAM_WAVES=w (default 1): w waves of shorter CTAs -- only useful together with stream priorities (DRE_PRIO),
    // where a low-priority Gram kernel should give SMs back to the ADI chain every few tens of microseconds
    static const int waves_env = std::max(1, getenv("DRE_GRAM_WAVES") ? atoi(getenv("DRE_GRAM_WAVES")) : 1);
    const int waves = waves_override > 0 ? waves_override : waves_env;
    int want = std::max(1, (2 * sm_count * waves) / tiles);   // at most 2 CTAs per SM: ONE full wave (19 x 16 tiles = 304
                                                      // CTAs on 296 slots was measured at twice the time of 18 x 16)
    if (a <= 16 && b >= 64) want = std::max(1, 8 * sm_count / ((b + 255) / 256));  // skinny kernel: thread per column
    int64_t maxsplit = std::max<int64_t>(1, chunks / 8);          // at least 8 chunks (128 rows) per split
    int nsplit = (int)std::min<int64_t>(want, maxsplit);
    nsplit = std::min(nsplit, (a <= 16 && b >= 64) ? 1184 : 320 * waves);   // skinny kernel: bytes in flight need many CTAs
    int64_t cps = (chunks + nsplit - 1) / nsplit;
    p.rows_per_split = cps * GK;
    p.nsplit = (int)((n + p.rows_per_split - 1) / p.rows_per_split);
    if (p.nsplit < 1) p.nsplit = 1;
    p.partial_elems = (size_t)p.nsplit * (size_t)std::max(a, 1) * (size_t)std::max(b, 1);
    return p;
}

// Skinny variant (a <= MMAX columns on the X side, e.g. B' [Z, Y] of the SMW correction with m = 7):
// the 64x64 DMMA tile would be 9x padding, so each thread owns one Y column and keeps the a partial
// sums in registers; Y rows are read fully coalesced, X rows are warp-broadcast.
template <int MMAX>
__global__ void __launch_bounds__(256) k_gram_skinny(const double* __restrict__ X, int64_t ldx, int a,
                                                     const double* __restrict__ Y, int64_t ldy, int b, int64_t n,
                                                     double* __restrict__ partial, int64_t rows_per_split) {
    const int64_t row_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t row_end = min(n, row_begin + rows_per_split);
    const int j = blockIdx.x * 256 + threadIdx.x;
    double acc[MMAX];
#pragma unroll
    for (int i = 0; i < MMAX; ++i) acc[i] = 0.0;
    if (j < b) {
        int64_t row = row_begin;
        for (; row + 4 <= row_end; row += 4) {
            double y[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) y[q] = Y[(row + q) * ldy + j];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double* xr = X + (row + q) * ldx;
#pragma unroll
                for (int i = 0; i < MMAX; ++i)
                    if (i < a) acc[i] = fma(xr[i], y[q], acc[i]);
            }
        }
        for (; row < row_end; ++row) {
            const double y = Y[row * ldy + j];
            const double* xr = X + row * ldx;
#pragma unroll
            for (int i = 0; i < MMAX; ++i)
                if (i < a) acc[i] = fma(xr[i], y, acc[i]);
        }
        double* P = partial + (int64_t)blockIdx.y * a * b;
#pragma unroll
        for (int i = 0; i < MMAX; ++i)
            if (i < a) P[(int64_t)i * b + j] = acc[i];
    }
}

void launch_gram(const double* X, int64_t ldx, int a, const double* Y, int64_t ldy, int b, int64_t n,
                 const double* roww, double* partial, const GramPlan& plan, double* out1, int64_t ld1,
                 double* out2, int64_t ld2, cudaStream_t st, int64_t* launches) {
    if (a <= 0 || b <= 0) return;
    if (a <= 16 && roww == nullptr && b >= 64) {
        dim3 grid((b + 255) / 256, plan.nsplit);
        if (a <= 8) DRE_LAUNCH((k_gram_skinny<8>), grid, 256, 0, st, X, ldx, a, Y, ldy, b, n, partial,
                               plan.rows_per_split);
        else DRE_LAUNCH((k_gram_skinny<16>), grid, 256, 0, st, X, ldx, a, Y, ldy, b, n, partial, plan.rows_per_split);
        const int64_t total = (int64_t)a * b;
        int rb = (int)std::min<int64_t>((total + 31) / 32, 148 * 8);
        DRE_LAUNCH((k_reduce_partials), rb, 256, 0, st, partial, plan.nsplit, a, b, out1, ld1, out2, ld2);
        if (launches) *launches += 2;
        return;
    }
    const int64_t total = (int64_t)a * b;
    int rb = (int)std::min<int64_t>((total + 31) / 32, 148 * 8);  // 8 warps per CTA, 4 output elements per warp
    if (roww) {   // weighted variant (small core products only): simple 64x64 kernel
        const int tiles_a = (a + GT - 1) / GT, tiles_b = (b + GT - 1) / GT;
        dim3 grid(tiles_a * tiles_b, plan.nsplit);
        DRE_LAUNCH((k_gram), grid, 256, 0, st, X, ldx, a, Y, ldy, b, n, roww, partial, tiles_b, plan.rows_per_split);
    } else {
        static bool attr_set_dev[DRE_MAX_DEVICES] = {};
        bool& attr_set = attr_set_dev[current_device()];
        const int smem128 = G2ST * G2K * (G2LDA + 128 + 8) * (int)sizeof(double);
        const int smem64 = G2ST * G2K * (G2LDA + 64 + 8) * (int)sizeof(double);
        if (!attr_set) {
            cudaFuncSetAttribute(k_gram2<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128);
            cudaFuncSetAttribute(k_gram2<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
            attr_set = true;
        }
        const int tbw = b <= 64 ? 64 : G2B;
        const int tiles_a = (a + G2A - 1) / G2A, tiles_b = (b + tbw - 1) / tbw;
        dim3 grid(tiles_a * tiles_b, plan.nsplit);
        const int aligned = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) % 16 == 0) &&
                            ldx % 2 == 0 && ldy % 2 == 0;
        if (tbw == 64)
            DRE_LAUNCH((k_gram2<64>), grid, 256, smem64, st, X, ldx, a, Y, ldy, b, n, partial, tiles_b,
                       plan.rows_per_split, aligned);
        else
            DRE_LAUNCH((k_gram2<128>), grid, 256, smem128, st, X, ldx, a, Y, ldy, b, n, partial, tiles_b,
                       plan.rows_per_split, aligned);
    }
    DRE_LAUNCH((k_reduce_partials), rb, 256, 0, st, partial, plan.nsplit, a, b, out1, ld1, out2, ld2);
    if (launches) *launches += 2;
}

// Small-K variant (a <= 64, b <= 64 per CTA column block): the whole X tile (128 x 64) and W (64 x 64) are
// fetched with one cp.async burst -- one barrier instead of a 4-chunk pipeline whose prologue dominates.
constexpr int TSK = 64, TSLDX = TSK + 4, TSLDW = 64 + 8;
__global__ void __launch_bounds__(256) k_tall_gemm_smallk(double alpha, const double* __restrict__ X, int64_t ldx,
                                                          int a, const double* __restrict__ W, int64_t ldw,
                                                          int w_trans, double beta, double* __restrict__ Y,
                                                          int64_t ldy, int b, int64_t n, int aligned_x) {
    DRE_DYN_SMEM_ALIGNED(double, ts_smem);
    double* Xs = ts_smem;                 // [128][TSLDX]
    double* Ws = ts_smem + T2M * TSLDX;   // [64][TSLDW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * T2M;
    const int b0 = blockIdx.y * 64;
    const int a4 = (a + 3) & ~3;
    if (aligned_x) {
        for (int c = tid; c < T2M * (TSK / 2); c += 256) {
            const int r = c / (TSK / 2), kc = (c % (TSK / 2)) * 2;
            if (kc >= a4) continue;
            const int64_t row = row0 + r;
            int nb = 0;
            if (row < n) nb = 8 * max(0, min(2, a - kc));
            cp_async16(Xs + r * TSLDX + kc, nb ? (const void*)(X + row * ldx + kc) : (const void*)X, nb);
        }
    } else {
        for (int c = tid; c < T2M * TSK; c += 256) {
            const int r = c / TSK, kc = c % TSK;
            if (kc >= a4) continue;
            const int64_t row = row0 + r;
            const bool ok = row < n && kc < a;
            cp_async8(Xs + r * TSLDX + kc, ok ? (const void*)(X + row * ldx + kc) : (const void*)X, ok ? 8 : 0);
        }
    }
    for (int c = tid; c < TSK * 64; c += 256) {
        int k, j;
        if (w_trans) { k = c % TSK; j = c / TSK; } else { k = c / 64; j = c % 64; }
        if (k >= a4) continue;
        const bool ok = k < a && b0 + j < b;
        const double* src = w_trans ? W + (int64_t)(b0 + j) * ldw + k : W + (int64_t)k * ldw + b0 + j;
        cp_async8(Ws + k * TSLDW + j, ok ? (const void*)src : (const void*)W, ok ? 8 : 0);
    }
    cp_async_commit();
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int wr = (warp & 3) * 32, wc = (warp >> 2) * 32;
    const int kk = lane & 3, rr = lane >> 2;
    cp_async_wait<0>();
    __syncthreads();
    for (int k0 = 0; k0 < a4; k0 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) af[mt] = Xs[(wr + mt * 8 + rr) * TSLDX + k0 + kk];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) bf[nt] = Ws[(k0 + kk) * TSLDW + wc + nt * 8 + rr];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int64_t row = row0 + wr + mt * 8 + rr;
        if (row >= n) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int j = b0 + wc + nt * 8 + 2 * kk;
            double* y = Y + row * ldy + j;
            if (j < b) y[0] = (beta == 0.0 ? 0.0 : beta * y[0]) + alpha * acc[mt][nt][0];
            if (j + 1 < b) y[1] = (beta == 0.0 ? 0.0 : beta * y[1]) + alpha * acc[mt][nt][1];
        }
    }
}

template <int T2N>
static void launch_tall_gemm_t(double alpha, const double* X, int64_t ldx, int a, const double* W, int64_t ldw,
                               int w_trans, double beta, double* Y, int64_t ldy, int b, int64_t n, cudaStream_t st) {
    static bool attr_set_dev[DRE_MAX_DEVICES] = {};
        bool& attr_set = attr_set_dev[current_device()];
    const int smem = T2ST * T2Cfg<T2N>::STAGE * (int)sizeof(double);
    if (!attr_set) {
        cudaFuncSetAttribute(k_tall_gemm2<T2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    dim3 grid((unsigned)((n + T2M - 1) / T2M), (unsigned)((b + T2N - 1) / T2N));
    const int ax = reinterpret_cast<uintptr_t>(X) % 16 == 0 && ldx % 2 == 0;
    const int aw = reinterpret_cast<uintptr_t>(W) % 16 == 0 && ldw % 2 == 0;
    DRE_LAUNCH((k_tall_gemm2<T2N>), grid, 256, smem, st, alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n, ax,
               aw);
}

void launch_tall_gemm(double alpha, const double* X, int64_t ldx, int a, const double* W, int64_t ldw,
                      int w_trans, double beta, double* Y, int64_t ldy, int b, int64_t n, cudaStream_t st,
                      int64_t* launches) {
    if (b <= 0 || n <= 0) return;
    if (a <= TSK && b > 16) {
        static bool attr_set_dev[DRE_MAX_DEVICES] = {};
        bool& attr_set = attr_set_dev[current_device()];
        const int smem = (T2M * TSLDX + TSK * TSLDW) * (int)sizeof(double);
        if (!attr_set) {
            cudaFuncSetAttribute(k_tall_gemm_smallk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr_set = true;
        }
        dim3 grid((unsigned)((n + T2M - 1) / T2M), (unsigned)((b + 63) / 64));
        const int ax = reinterpret_cast<uintptr_t>(X) % 16 == 0 && ldx % 2 == 0;
        DRE_LAUNCH((k_tall_gemm_smallk), grid, 256, smem, st, alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n,
                   ax);
    } else if (b <= 16) {
        launch_tall_gemm_t<16>(alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n, st);
    } else {
        launch_tall_gemm_t<64>(alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n, st);
    }
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------
__global__ void k_copy_scale(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                             int64_t n, int cols, const double* __restrict__ colscale,
                             const double* __restrict__ rowscale) {
    const int64_t total = n * cols;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cols;
        const int c = (int)(idx % cols);
        double v = src[row * lds + c];
        if (colscale) v *= colscale[c];
        if (rowscale) v *= rowscale[row];
        dst[row * ldd + c] = v;
    }
}

void launch_copy_scale(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                       const double* colscale, cudaStream_t st, int64_t* launches, const double* rowscale) {
    if (n <= 0 || cols <= 0) return;
    int blocks = (int)std::min<int64_t>((n * cols + 255) / 256, 148 * 16);
    DRE_LAUNCH((k_copy_scale), blocks, 256, 0, st, dst, ldd, src, lds, n, cols, colscale, rowscale);
    if (launches) *launches += 1;
}

// partial[blk][c] = sum over the block's row range of P[row][c]^2   (cols <= 256, thread per column)
__global__ void __launch_bounds__(256) k_colnorm2(const double* __restrict__ P, int64_t ldp, int64_t n, int cols,
                                                  int64_t rows_per_blk, double* __restrict__ partial) {
    const int c = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_blk, r1 = min(n, r0 + rows_per_blk);
    double s0 = 0.0, s1 = 0.0;
    if (c < cols) {
        int64_t row = r0;
        for (; row + 1 < r1; row += 2) {
            const double v0 = P[row * ldp + c], v1 = P[(row + 1) * ldp + c];
            s0 = fma(v0, v0, s0);
            s1 = fma(v1, v1, s1);
        }
        if (row < r1) { const double v = P[row * ldp + c]; s0 = fma(v, v, s0); }
        partial[(int64_t)blockIdx.x * cols + c] = s0 + s1;
    }
}

// out[c] = sum_rows P[row][c]^2 for c < cols <= 256 (deterministic two-stage reduction); partial: nblk*cols doubles
void launch_colnorm2(const double* P, int64_t ldp, int64_t n, int cols, double* partial, int nblk, double* out,
                     cudaStream_t st, int64_t* launches) {
    if (n <= 0 || cols <= 0) return;
    const int64_t rpb = (n + nblk - 1) / nblk;
    DRE_LAUNCH((k_colnorm2), nblk, 256, 0, st, P, ldp, n, cols, rpb, partial);
    DRE_LAUNCH((k_reduce_partials), (cols + 31) / 32, 256, 0, st, partial, nblk, 1, cols, out, cols, nullptr, 0);
    if (launches) *launches += 2;
}

// dst[row][c] = w1[c] * src[row][c] + w2[c] * src[row][idx2[c]]   (two-term column combination)
__global__ void k_combine_cols(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                               int64_t n, int cols, const int32_t* __restrict__ idx2, const double* __restrict__ w1,
                               const double* __restrict__ w2) {
    const int64_t total = n * cols;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cols;
        const int c = (int)(idx % cols);
        const double* sr = src + row * lds;
        dst[row * ldd + c] = w1[c] * sr[c] + w2[c] * sr[idx2[c]];
    }
}

void launch_combine_cols(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                         const int32_t* idx2, const double* w1, const double* w2, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || cols <= 0) return;
    int blocks = (int)std::min<int64_t>((n * cols + 255) / 256, 148 * 16);
    DRE_LAUNCH((k_combine_cols), blocks, 256, 0, st, dst, ldd, src, lds, n, cols, idx2, w1, w2);
    if (launches) *launches += 1;
}

__global__ void k_axpby(double alpha, const double* __restrict__ X, int64_t ldx, double beta, double* __restrict__ Y,
                        int64_t ldy, int64_t n, int cols) {
    const int64_t total = n * cols;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / cols;
        const int c = (int)(idx % cols);
        double v = (beta == 0.0) ? 0.0 : beta * Y[row * ldy + c];
        if (alpha != 0.0) v += alpha * X[row * ldx + c];
        Y[row * ldy + c] = v;
    }
}

void launch_axpby(double alpha, const double* X, int64_t ldx, double beta, double* Y, int64_t ldy, int64_t n,
                  int cols, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || cols <= 0) return;
    int blocks = (int)std::min<int64_t>((n * cols + 255) / 256, 148 * 16);
    DRE_LAUNCH((k_axpby), blocks, 256, 0, st, alpha, X, ldx, beta, Y, ldy, n, cols);
    if (launches) *launches += 1;
}

// ---- Arnoldi step of the Heuristic shift strategy (heuristic.jl:111-125) on device-resident vectors ----
// The twice-repeated modified Gram-Schmidt sweep is a chain of 2(j+1) dependent dot products / updates on n-vectors.
// One launch per link: stage s first applies the update of the previous link, w -= g_{s-1} v_{s-1} (g_{s-1} is rebuilt by
// every CTA from the previous stage's per-CTA partial sums, in one fixed order, so all CTAs subtract the same number),
// then forms its partial of v_s . w (last stage: of w . w).  No host round trip inside the sweep; the coefficients are
// collected with one copy at the end.  Basis columns are strided (row-major panel) but the whole basis sits in L2.
constexpr int MGS_MAXB = 296;
__device__ __forceinline__ double mgs_sum_partials(const double* part, int nb, double* sh) {
    // warp 0: lane l adds part[l], part[l+32], ... in order, then a shuffle tree; result broadcast through sh[0]
    if (threadIdx.x < 32) {
        double a = 0.0;
        for (int i = threadIdx.x; i < nb; i += 32) a += part[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (threadIdx.x == 0) sh[0] = a;
    }
    __syncthreads();
    const double g = sh[0];
    __syncthreads();
    return g;
}

__global__ void __launch_bounds__(256) k_mgs_stage(const double* __restrict__ V, int64_t ldv, int ci_prev, int ci,
                                                   double* __restrict__ w, int64_t ldw, int64_t n,
                                                   const double* __restrict__ part_prev, int nb_prev,
                                                   double* __restrict__ part, double* __restrict__ g_out) {
    __shared__ double sh[8];
    double g = 0.0;
    if (ci_prev >= 0) {
        g = mgs_sum_partials(part_prev, nb_prev, sh);
        if (blockIdx.x == 0 && threadIdx.x == 0) *g_out = g;
    }
    double acc = 0.0;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        double wr = w[r * ldw];
        if (ci_prev >= 0) {
            wr -= V[r * ldv + ci_prev] * g;
            w[r * ldw] = wr;
        }
        acc = fma(ci >= 0 ? V[r * ldv + ci] : wr, wr, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += sh[i];
        part[blockIdx.x] = a;
    }
}

// beta = ||w|| from the partials of the last stage; v_next = (1/beta) w  (heuristic.jl:123-124)
__global__ void __launch_bounds__(256) k_mgs_finish(const double* __restrict__ w, int64_t ldw, double* __restrict__ vn,
                                                    int64_t ldvn, int64_t n, const double* __restrict__ part_prev,
                                                    int nb_prev, double* __restrict__ beta_out) {
    __shared__ double sh[8];
    const double beta = sqrt(mgs_sum_partials(part_prev, nb_prev, sh));
    if (blockIdx.x == 0 && threadIdx.x == 0) *beta_out = beta;
    const double inv = 1.0 / beta;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        vn[r * ldvn] = inv * w[r * ldw];
}

void launch_arnoldi_mgs(const double* V, int64_t ldv, int nbasis, double* w, int64_t ldw, double* vnext, int64_t ldvn,
                        int64_t n, double* partials, double* coef, cudaStream_t st, int64_t* launches) {
    // partials: 2 * MGS_MAXB doubles (ping-pong); coef: 2 * nbasis + 1 doubles (g of every link in order, then beta)
    const int nb = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, MGS_MAXB));
    const int links = 2 * nbasis;
    for (int s = 0; s <= links; ++s) {
        const int ci = s < links ? s % nbasis : -1;
        const int cp = s > 0 ? (s - 1) % nbasis : -1;
        DRE_LAUNCH((k_mgs_stage), nb, 256, 0, st, V, ldv, cp, ci, w, ldw, n, partials + ((s + 1) & 1) * MGS_MAXB, nb,
                   partials + (s & 1) * MGS_MAXB, coef + (s > 0 ? s - 1 : 0));
    }
    DRE_LAUNCH((k_mgs_finish), nb, 256, 0, st, w, ldw, vnext, ldvn, n, partials + (links & 1) * MGS_MAXB, nb,
               coef + links);
    if (launches) *launches += links + 2;
}

// column-major staging (original row order) <-> row-major panel (solver row order), 32x32 smem transpose
__global__ void k_cm2panel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                           int64_t n, int cols, const int32_t* __restrict__ iperm) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const int64_t r = r0 + threadIdx.x;
        if (r < n && c0 + cc < cols) tile[cc][threadIdx.x] = src[r + (int64_t)(c0 + cc) * lds];
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int64_t r = r0 + rr;
        const int c = c0 + threadIdx.x;
        if (r < n && c < cols) {
            const int64_t pr = iperm ? iperm[r] : r;
            dst[pr * ldd + c] = tile[threadIdx.x][rr];
        }
    }
}

__global__ void k_panel2cm(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds,
                           int64_t n, int cols, const int32_t* __restrict__ iperm) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int64_t r = r0 + rr;
        const int c = c0 + threadIdx.x;
        if (r < n && c < cols) {
            const int64_t pr = iperm ? iperm[r] : r;
            tile[rr][threadIdx.x] = src[pr * lds + c];
        }
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {
        const int64_t r = r0 + threadIdx.x;
        if (r < n && c0 + cc < cols) dst[r + (int64_t)(c0 + cc) * ldd] = tile[threadIdx.x][cc];
    }
}

void launch_colmajor_to_panel(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                              const int32_t* iperm, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || cols <= 0) return;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
    DRE_LAUNCH((k_cm2panel), grid, block, 0, st, dst, ldd, src, lds, n, cols, iperm);
    if (launches) *launches += 1;
}

void launch_panel_to_colmajor(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t n, int cols,
                              const int32_t* iperm, cudaStream_t st, int64_t* launches) {
    if (n <= 0 || cols <= 0) return;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((cols + 31) / 32)), block(32, 8);
    DRE_LAUNCH((k_panel2cm), grid, block, 0, st, dst, ldd, src, lds, n, cols, iperm);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// pivoted Cholesky panel selection (single CTA, pb <= 64)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pivchol(const double* __restrict__ G, int64_t ldg, int pb, double drop2,
                                                 double rel2, double* __restrict__ Wsel, int32_t* __restrict__ info,
                                                 double* __restrict__ dinfo) {
    DRE_DYN_SMEM(double, pc_smem);
    double (*A)[65] = reinterpret_cast<double (*)[65]>(pc_smem);            // Gram matrix
    double (*C)[65] = reinterpret_cast<double (*)[65]>(pc_smem + 64 * 65);  // Cholesky factor columns
    double (*Z)[65] = A;  // inverse of the permuted triangular factor (reuses A after the factorization)
    __shared__ double d[64];
    __shared__ int piv[64];
    __shared__ int s_p;
    __shared__ double s_thr, s_dfirst;
    __shared__ int s_nsel;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        A[i][j] = (i < pb && j < pb) ? G[(int64_t)i * ldg + j] : 0.0;
        C[i][j] = 0.0;
    }
    __syncthreads();
    if (tid < 64) d[tid] = (tid < pb) ? A[tid][tid] : -1.0;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < pb; ++i) m = fmax(m, d[i]);
        s_dfirst = m;
        s_thr = fmax(drop2, rel2 * m);
        s_nsel = 0;
    }
    __syncthreads();
    for (int j = 0; j < pb; ++j) {
        // pivot = arg max of the remaining Schur diagonal (warp 0; ties -> smallest index)
        if (tid < 32) {
            double v = (tid < pb) ? d[tid] : -1.0;
            int ix = tid;
            const double v2 = (tid + 32 < pb) ? d[tid + 32] : -1.0;
            if (v2 > v) { v = v2; ix = tid + 32; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, v, off);
                const int oi = __shfl_down_sync(0xffffffffu, ix, off);
                if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
            }
            if (tid == 0) s_p = (v >= s_thr && v > 0.0) ? ix : -1;
        }
        __syncthreads();
        const int p = s_p;
        if (p < 0) break;
        const double cjj = sqrt(d[p]);
        {   // column j of the factor: 4 threads per row split the dot product
            const int i = tid >> 2, q = tid & 3;
            double acc = 0.0;
            if (i < pb && i != p && d[i] >= 0.0)
                for (int t = q; t < j; t += 4) acc += C[i][t] * C[p][t];
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (q == 0 && i < pb) {
                if (i == p) C[i][j] = cjj;
                else if (d[i] >= 0.0) C[i][j] = (A[i][p] - acc) / cjj;
            }
        }
        __syncthreads();
        if (tid < pb) {
            const int i = tid;
            if (i == p) d[i] = -1.0;
            else if (d[i] >= 0.0) d[i] = fmax(d[i] - C[i][j] * C[i][j], 0.0);
        }
        if (tid == 0) { piv[j] = p; s_nsel = j + 1; }
        __syncthreads();
    }
    __syncthreads();
    const int nsel = s_nsel;
    for (int idx = tid; idx < 64 * 64; idx += 256) Z[idx >> 6][idx & 63] = 0.0;
    __syncthreads();
    // Z = C11^-1 (lower), C11[a][b] = C[piv[a]][b]
    if (tid < nsel) {
        const int bcol = tid;
        Z[bcol][bcol] = 1.0 / C[piv[bcol]][bcol];
        for (int a2 = bcol + 1; a2 < nsel; ++a2) {
            double v = 0.0;
            for (int t = bcol; t < a2; ++t) v += C[piv[a2]][t] * Z[t][bcol];
            Z[a2][bcol] = -v / C[piv[a2]][a2];
        }
    }
    __syncthreads();
    // Wsel[piv[a]][b] = Z[b][a]  (a <= b), zero elsewhere
    for (int idx = tid; idx < 64 * 64; idx += 256) Wsel[idx] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < nsel * nsel; idx += 256) {
        const int a2 = idx / nsel, b2 = idx % nsel;
        if (a2 <= b2) Wsel[piv[a2] * 64 + b2] = Z[b2][a2];
    }
    if (tid == 0) {
        info[0] = nsel;
        double m = 0.0;
        for (int i = 0; i < pb; ++i) m = fmax(m, d[i]);
        dinfo[0] = s_dfirst;
        dinfo[1] = m;
    }
}

void launch_pivchol(const double* G, int64_t ldg, int pb, double drop2, double rel2, double* Wsel, int32_t* info,
                    double* dinfo, cudaStream_t st, int64_t* launches) {
    static bool attr_set_dev[DRE_MAX_DEVICES] = {};
        bool& attr_set = attr_set_dev[current_device()];
    const int smem = 2 * 64 * 65 * (int)sizeof(double);
    if (!attr_set) {
        cudaFuncSetAttribute(k_pivchol, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set = true;
    }
    DRE_LAUNCH((k_pivchol), 1, 256, smem, st, G, ldg, pb, drop2, rel2, Wsel, info, dinfo);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// ||R diag(t) R'||_F^2 = sum_ij G_ij^2 t_i t_j,  G = R'R
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_norm_diag(const double* __restrict__ G, int64_t ldg, int r,
                                                    const double* __restrict__ t, double* __restrict__ out) {
    // one CTA of 32 warps; warp w sums rows w, w+32, ... (coalesced along the row), fixed-shape reduction
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double s = 0.0;
    for (int i = warp; i < r; i += 32) {
        const double ti = t[i];
        double si = 0.0;
        for (int j = lane; j < r; j += 32) {
            const double g = G[(int64_t)i * ldg + j];
            si = fma(g * g, t[j], si);
        }
        s = fma(si, ti, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        double v = red[lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) out[0] = v;
    }
}

void launch_norm_diag(const double* G, int64_t ldg, int r, const double* t, double* out, cudaStream_t st,
                      int64_t* launches) {
    DRE_LAUNCH((k_norm_diag), 1, 1024, 0, st, G, ldg, r, t, out);
    if (launches) *launches += 1;
}

__global__ void k_gather_rows(double* __restrict__ Wt, int64_t ldw, const double* __restrict__ V, int64_t ldv,
                              const int32_t* __restrict__ ids, int nsel, int len) {
    const int64_t total = (int64_t)nsel * len;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx / len), k = (int)(idx % len);
        Wt[(int64_t)j * ldw + k] = V[(int64_t)ids[j] * ldv + k];
    }
}

void launch_gather_rows(double* Wt, int64_t ldw, const double* V, int64_t ldv, const int32_t* ids, int nsel,
                        int len, cudaStream_t st, int64_t* launches) {
    if (nsel <= 0 || len <= 0) return;
    int blocks = (int)std::min<int64_t>(((int64_t)nsel * len + 255) / 256, 1024);
    DRE_LAUNCH((k_gather_rows), blocks, 256, 0, st, Wt, ldw, V, ldv, ids, nsel, len);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------
// CSR SpMM with fused axpby:  Y[row][:] = beta*Y[row][:] + alpha * sum_j val_j X[col_j][:]
// One warp per row, lanes over the (contiguous) panel columns: every X-row read and every
// Y-row write is a fully coalesced 256-byte segment; E / A values and indices are warp-broadcast.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spmm(const int32_t* __restrict__ ptr, const int32_t* __restrict__ col,
                                              const double* __restrict__ val, int64_t n, double alpha,
                                              const double* __restrict__ X, int64_t ldx, double beta,
                                              double* __restrict__ Y, int64_t ldy, int cols) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n; row += nwarps) {
        const int p0 = ptr[row], p1 = ptr[row + 1];
        for (int c0 = 0; c0 < cols; c0 += 128) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int p = p0; p < p1; ++p) {
                const double v = val[p];
                const double* xr = X + (int64_t)col[p] * ldx + c0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = lane + 32 * u;
                    if (c0 + c < cols) acc[u] = fma(v, xr[c], acc[u]);
                }
            }
            double* yr = Y + row * ldy + c0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = lane + 32 * u;
                if (c0 + c < cols) yr[c] = (beta == 0.0 ? 0.0 : beta * yr[c]) + alpha * acc[u];
            }
        }
    }
}

// Variant with the row's indices and values fetched by the lanes (one coalesced load each) and broadcast by
// shuffles: the X-row loads no longer wait for a dependent index load per nonzero, two nonzeros (8 row segments per
// lane) are in flight at a time, and the Y row (beta != 0) is requested before the accumulation starts.
// Opt-in (DRE_SPMM2=1) until measured: k_spmm reaches 49 % of the HBM peak by the byte model.
__global__ void __launch_bounds__(256) k_spmm2(const int32_t* __restrict__ ptr, const int32_t* __restrict__ col,
                                               const double* __restrict__ val, int64_t n, double alpha,
                                               const double* __restrict__ X, int64_t ldx, double beta,
                                               double* __restrict__ Y, int64_t ldy, int cols) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < n; row += nwarps) {
        const int p0 = ptr[row], nnz = ptr[row + 1] - p0;
        for (int c0 = 0; c0 < cols; c0 += 128) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            double yold[4] = {0.0, 0.0, 0.0, 0.0};
            double* yr = Y + row * ldy + c0;
            if (beta != 0.0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = lane + 32 * u;
                    if (c0 + c < cols) yold[u] = yr[c];
                }
            }
            for (int base = 0; base < nnz; base += 32) {
                const int cnt = min(32, nnz - base);
                const int myc = (lane < cnt) ? col[p0 + base + lane] : 0;
                const double myv = (lane < cnt) ? val[p0 + base + lane] : 0.0;
                int j = 0;
                for (; j + 2 <= cnt; j += 2) {
                    const double v0 = __shfl_sync(0xffffffffu, myv, j), v1 = __shfl_sync(0xffffffffu, myv, j + 1);
                    const double* x0 = X + (int64_t)__shfl_sync(0xffffffffu, myc, j) * ldx + c0;
                    const double* x1 = X + (int64_t)__shfl_sync(0xffffffffu, myc, j + 1) * ldx + c0;
                    double a0[4], a1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = lane + 32 * u;
                        const bool ok = c0 + c < cols;
                        a0[u] = ok ? x0[c] : 0.0;
                        a1[u] = ok ? x1[c] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = fma(v1, a1[u], fma(v0, a0[u], acc[u]));
                }
                if (j < cnt) {
                    const double v0 = __shfl_sync(0xffffffffu, myv, j);
                    const double* x0 = X + (int64_t)__shfl_sync(0xffffffffu, myc, j) * ldx + c0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = lane + 32 * u;
                        if (c0 + c < cols) acc[u] = fma(v0, x0[c], acc[u]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = lane + 32 * u;
                if (c0 + c < cols) yr[c] = (beta == 0.0 ? 0.0 : beta * yold[u]) + alpha * acc[u];
            }
        }
    }
}

void launch_spmm(const int32_t* ptr, const int32_t* col, const double* val, int64_t n, double alpha,
                 const double* X, int64_t ldx, double beta, double* Y, int64_t ldy, int cols, cudaStream_t st,
                 int64_t* launches) {
    if (n <= 0 || cols <= 0) return;
    int64_t blocks = (n + 7) / 8;  // 8 warps per CTA, one row per warp
    blocks = std::min<int64_t>(blocks, 148 * 32);
    if (spmm_variant == 2)
        DRE_LAUNCH((k_spmm2), (unsigned)blocks, 256, 0, st, ptr, col, val, n, alpha, X, ldx, beta, Y, ldy, cols);
    else
        DRE_LAUNCH((k_spmm), (unsigned)blocks, 256, 0, st, ptr, col, val, n, alpha, X, ldx, beta, Y, ldy, cols);
    if (launches) *launches += 1;
}

}  // namespace dre
