// In-tree symmetric eigensolver for the small projected core of compress! (the reference calls LAPACK through
// eigen(Symmetric(S)), src/LDLt.jl:214).  k is the numerical rank of hcat(Ls): a few hundred.
//
//   1. k_tridiag      Householder tridiagonalisation  S = Q T Q'  in ONE cooperative launch: the matrix stays in
//                     L2, the rows of the trailing block are spread over the warps of the whole grid, the rank-2
//                     update of step j is applied lazily while step j+1 reads the matrix (one grid barrier per
//                     column instead of two).
//   2. host           implicit QL iteration on the tridiagonal T (O(k^2) scalar work: 2 ms at k = 300, far below
//                     what a single GPU thread would need); only the plane rotations (c, s, i) are recorded.
//   3. k_form_q       explicit Q from the stored reflectors (backward accumulation, one thread per column);
//                     independent of step 2 and queued before the host starts it, so the two overlap.
//   4. k_apply_rot    eigenvectors V = Q G_1 G_2 ... : every row of Q is independent, one thread per row with the
//                     row in shared memory; written transposed and sorted (row j of the output = eigenvector of
//                     the j-th smallest eigenvalue).
// Backward stable (Householder + QL), orthogonal eigenvectors also for clustered / repeated / +-paired
// eigenvalues, which the cores of this path have (T = [0 D; D 0] blocks of the Lyapunov residual).
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace dre {

namespace {

constexpr int TD_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
    // deterministic block reduction (fixed tree); red: >= 32 doubles of shared memory
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}

// A: k x k symmetric, full storage, leading dimension k (row-major == column-major).  On exit d[0..k), e[0..k-1),
// reflector j (j = 0..k-3) in Vst[j*k + 0..m) with m = k-j-1 (v[0] = 1) and tau[j].  pbuf: k doubles of scratch.
__global__ void __launch_bounds__(TD_THREADS) k_tridiag(double* A, int k, double* d, double* e, double* Vst,
                                                        double* tau, double* pbuf) {
    extern __shared__ double sm[];
    double* v = sm;            // [k] current reflector
    double* vp = sm + k;       // [k] pending update: A22 -= vp wp' + wp vp'
    double* wp = sm + 2 * k;   // [k]
    double* red = sm + 3 * k;  // [40]
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x, lane = tid & 31;
    const int gwarp = (blockIdx.x * blockDim.x + tid) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    bool pending = false;
    for (int j = 0; j + 2 < k; ++j) {
        const int m = k - j - 1;
        // ---- step A (redundantly in every CTA): column j of the current matrix and its reflector ----
        // pending indices are relative to the previous trailing block, which started at row j: column j is its
        // index 0, row j+1+i its index i+1
        const double vp0 = pending ? vp[0] : 0.0, wp0 = pending ? wp[0] : 0.0;
        double part = 0.0;
        for (int i = tid; i < m; i += TD_THREADS) {
            double x = A[(size_t)j * k + (j + 1 + i)];
            if (pending) x -= vp[i + 1] * wp0 + wp[i + 1] * vp0;
            v[i] = x;
            if (i > 0) part += x * x;
        }
        const double xnorm2 = block_sum(part, red);   // (its barriers also publish v)
        const double alpha = v[0];
        double beta, t;
        if (xnorm2 == 0.0) {
            beta = alpha;
            t = 0.0;
        } else {
            beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
            t = (beta - alpha) / beta;
        }
        const double scal = (xnorm2 == 0.0) ? 0.0 : 1.0 / (alpha - beta);
        __syncthreads();
        for (int i = tid; i < m; i += TD_THREADS) v[i] = (i == 0) ? 1.0 : v[i] * scal;
        __syncthreads();
        if (blockIdx.x == 0) {
            for (int i = tid; i < m; i += TD_THREADS) Vst[(size_t)j * k + i] = v[i];
            if (tid == 0) {
                double djj = A[(size_t)j * k + j];
                if (pending) djj -= 2.0 * vp0 * wp0;
                d[j] = djj;
                e[j] = beta;
                tau[j] = t;
            }
        }
        // ---- step B: apply the pending update to the trailing block (rows over all warps) and p = tau A22 v ----
        for (int i = gwarp; i < m; i += nwarps) {
            double* row = A + (size_t)(j + 1 + i) * k + (j + 1);
            const double vpi = pending ? vp[i + 1] : 0.0, wpi = pending ? wp[i + 1] : 0.0;
            double acc = 0.0;
            for (int c = lane; c < m; c += 32) {
                double a = row[c];
                if (pending) {
                    a -= vpi * wp[c + 1] + wpi * vp[c + 1];
                    row[c] = a;
                }
                acc += a * v[c];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) pbuf[i] = t * acc;
        }
        grid.sync();
        // ---- step C (redundantly): w = p - (tau/2 p'v) v becomes the pending update ----
        part = 0.0;
        for (int i = tid; i < m; i += TD_THREADS) {
            const double p = pbuf[i];
            wp[i] = p;
            part += p * v[i];
        }
        const double K = 0.5 * t * block_sum(part, red);
        for (int i = tid; i < m; i += TD_THREADS) {
            wp[i] -= K * v[i];
            vp[i] = v[i];
        }
        pending = true;
        __syncthreads();
        // a fast CTA may already write p of the next column while a slow one still reads this one: two buffers
        pbuf += (j & 1) ? -k : k;
    }
    // ---- last 2 x 2 block (rows k-2, k-1) ----
    if (blockIdx.x == 0 && tid == 0) {
        if (k == 1) {
            d[0] = A[0];
        } else {
            const int j = k - 2;
            double a00 = A[(size_t)j * k + j], a10 = A[(size_t)j * k + j + 1], a11 = A[(size_t)(j + 1) * k + j + 1];
            if (pending) {
                a00 -= 2.0 * vp[0] * wp[0];
                a10 -= vp[1] * wp[0] + wp[1] * vp[0];
                a11 -= 2.0 * vp[1] * wp[1];
            }
            d[j] = a00;
            e[j] = a10;
            d[j + 1] = a11;
        }
    }
}

// Q = H_0 H_1 ... H_{k-3} (H_j = I - tau_j [0; v_j][0; v_j]' acting on rows j+1..k-1), stored TRANSPOSED:
// Qt[c*k + i] = Q[i][c].  Backward accumulation; every column of Q is independent: one warp per column, the column
// lives in registers (lane l owns the elements l, l+32, ...: NS slots cover k <= 32 NS), the reflector is read
// coalesced from L2 and the dot product is a shuffle tree -- no memory round trip inside the dependent chain.
template <int NS>
__global__ void __launch_bounds__(256) k_form_qt(double* Qt, int k, const double* __restrict__ Vst,
                                                 const double* __restrict__ tau) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= k) return;
    double col[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) col[q] = (lane + 32 * q == c) ? 1.0 : 0.0;
    for (int j = min(k - 3, c - 1); j >= 0; --j) {   // H_j leaves the columns c <= j untouched
        const double t = tau[j];
        if (t == 0.0) continue;
        const double* v = Vst + (size_t)j * k - (j + 1);   // v[idx] for the row index idx >= j+1
        double vv[NS];
        double z = 0.0;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const int idx = lane + 32 * q;
            vv[q] = (idx > j && idx < k) ? v[idx] : 0.0;
            z = fma(vv[q], col[q], z);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        z *= t;
#pragma unroll
        for (int q = 0; q < NS; ++q) col[q] = fma(-z, vv[q], col[q]);
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        const int idx = lane + 32 * q;
        if (idx < k) Qt[(size_t)c * k + idx] = col[q];
    }
}

// generic fallback for k > 1024 (not met on this path): column in global memory
__global__ void __launch_bounds__(256) k_form_qt_big(double* Qt, int k, const double* __restrict__ Vst,
                                                     const double* __restrict__ tau) {
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= k) return;
    double* col = Qt + (size_t)c * k;
    for (int idx = lane; idx < k; idx += 32) col[idx] = (idx == c) ? 1.0 : 0.0;
    __syncwarp();
    for (int j = min(k - 3, c - 1); j >= 0; --j) {
        const double t = tau[j];
        if (t == 0.0) continue;
        const double* v = Vst + (size_t)j * k - (j + 1);
        double z = 0.0;
        for (int idx = j + 1 + lane; idx < k; idx += 32) z = fma(v[idx], col[idx], z);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        z *= t;
        for (int idx = j + 1 + lane; idx < k; idx += 32) col[idx] = fma(-z, v[idx], col[idx]);
        __syncwarp();
    }
}

struct Rot {      // plane rotation of the QL iteration
    double c, s;
};
struct Sweep {    // (part of) one QL sweep: rotations rots[off .. off+len) act on the pairs (hi, hi+1), (hi-1, hi), ...
    int off, hi, len, pad;
};
constexpr int ROT_TILE = 1024;   // rotations per shared-memory tile; the host splits sweeps at tile boundaries

// Eigenvectors V = Q G_1 G_2 ...: rows [r0, r0 + nr) of Q, one thread of warp 0 per row with the row in shared
// memory.  A sweep is the chain
//   f = z[i+1];  z[i+1] = s z[i] + c f;  z[i] = c z[i] - s f      for i = hi, hi-1, ..., hi-len+1
// in which the new z[i] is the f of the next rotation: it is carried in a register, so the only dependent operation
// per rotation is one FMA, and the loads of z[i] are hoisted four rotations ahead of the stores (indices descend).
// The rotation parameters stream through two shared-memory tiles: warps 1..3 fetch tile t+1 from global memory while
// warp 0 runs the chains of tile t (a dependent global load per rotation cost 180 cycles per rotation before).
// Finally Out[rank[i]*k + row] = z[i] (transposed, sorted by eigenvalue).  Qt is Q transposed (k_form_qt).
__global__ void __launch_bounds__(128) k_apply_rot(const double* __restrict__ Qt, int k, const Rot* __restrict__ rots,
                                                   int nrot, const Sweep* __restrict__ sweeps,
                                                   const int32_t* __restrict__ tile_first, int ntiles,
                                                   const int32_t* __restrict__ rank, double* Out, int rows_per_cta) {
    extern __shared__ double zs[];   // [rows_per_cta][ldz], then two rotation tiles
    const int ldz = k + 1 + ((k & 1) ? 1 : 0);   // odd leading dimension: lanes of a warp hit different banks
    Rot* tiles = reinterpret_cast<Rot*>(zs + (size_t)rows_per_cta * ldz + (((size_t)rows_per_cta * ldz) & 1));
    const int r0 = blockIdx.x * rows_per_cta;
    const int nr = min(rows_per_cta, k - r0);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nr * k; idx += blockDim.x) {
        const int i = idx / nr, r = idx - i * nr;
        zs[r * ldz + i] = Qt[(size_t)i * k + (r0 + r)];
    }
    for (int g = tid; g < min(ROT_TILE, nrot); g += blockDim.x) tiles[g] = rots[g];
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const Rot* R = tiles + (size_t)(t & 1) * ROT_TILE;
        if (tid >= 32) {   // loader warps: next tile
            Rot* Rn = tiles + (size_t)((t + 1) & 1) * ROT_TILE;
            const int g0 = (t + 1) * ROT_TILE;
            for (int g = tid - 32; g < ROT_TILE && g0 + g < nrot; g += blockDim.x - 32) Rn[g] = rots[g0 + g];
        } else if (tid < nr) {
            double* z = zs + tid * ldz;
            const int base = t * ROT_TILE;
            for (int q = tile_first[t]; q < tile_first[t + 1]; ++q) {
                const Sweep sw = sweeps[q];
                const Rot* Rs = R + (sw.off - base);
                double carry = z[sw.hi + 1];
                int u = 0;
                for (; u + 4 <= sw.len; u += 4) {
                    const int i = sw.hi - u;
                    const double z0 = z[i], z1 = z[i - 1], z2 = z[i - 2], z3 = z[i - 3];
                    const Rot a = Rs[u], b = Rs[u + 1], c2 = Rs[u + 2], d = Rs[u + 3];
                    z[i + 1] = fma(a.s, z0, a.c * carry);
                    carry = fma(-a.s, carry, a.c * z0);
                    z[i] = fma(b.s, z1, b.c * carry);
                    carry = fma(-b.s, carry, b.c * z1);
                    z[i - 1] = fma(c2.s, z2, c2.c * carry);
                    carry = fma(-c2.s, carry, c2.c * z2);
                    z[i - 2] = fma(d.s, z3, d.c * carry);
                    carry = fma(-d.s, carry, d.c * z3);
                }
                for (; u < sw.len; ++u) {
                    const int i = sw.hi - u;
                    const double z0 = z[i];
                    const Rot a = Rs[u];
                    z[i + 1] = fma(a.s, z0, a.c * carry);
                    carry = fma(-a.s, carry, a.c * z0);
                }
                z[sw.hi - sw.len + 1] = carry;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nr * k; idx += blockDim.x) {
        const int i = idx / nr, r = idx - i * nr;
        Out[(size_t)rank[i] * k + (r0 + r)] = zs[r * ldz + i];
    }
}

// implicit QL with Wilkinson shifts on the symmetric tridiagonal (d, e); rotations are recorded in application order
bool tql_rotations(std::vector<double>& d, std::vector<double>& e, std::vector<Rot>& rots, std::vector<Sweep>& sweeps) {
    const int n = (int)d.size();
    e.resize(n, 0.0);
    e[n - 1] = 0.0;
    const double eps = 2.220446049250313e-16;
    // Deflation: the classical relative test, or an off-diagonal entry below 0.1 eps ||T|| -- an absolute perturbation
    // three orders of magnitude under the level at which compress! truncates eigenvalues (100 eps max|lambda|,
    // src/LDLt.jl:216-217); it keeps the iteration from grinding on round-off-sized trailing blocks of rank-deficient
    // cores, where the relative test compares noise with noise.
    double anorm = 0.0;
    for (int i = 0; i < n; ++i)
        anorm = std::max(anorm, std::fabs(d[i]) + std::fabs(e[i]) + (i > 0 ? std::fabs(e[i - 1]) : 0.0));
    if (!std::isfinite(anorm)) return false;
    const double tiny = 0.1 * eps * anorm;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        for (;;) {
            int m = l;
            for (; m < n - 1; ++m) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= eps * dd || std::fabs(e[m]) <= tiny) break;
            }
            if (m == l) break;
            if (++iter > 200) return false;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = std::hypot(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            const int off0 = (int)rots.size();
            for (; i >= l; --i) {
                double f = s * e[i];
                const double b = c * e[i];
                // (plain sqrt: f and g are entries of a matrix of moderate norm -- the core was formed from
                //  coefficients of unit-norm basis vectors -- so f*f + g*g neither overflows nor loses everything)
                r = std::sqrt(f * f + g * g);
                if (!(r > 1e-150) && r != 0.0) r = std::hypot(f, g);
                e[i + 1] = r;
                if (r == 0.0) {
                    d[i + 1] -= p;
                    e[m] = 0.0;
                    break;
                }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
                rots.push_back(Rot{c, s});
            }
            // recorded in pieces that do not cross a tile of ROT_TILE rotations: a piece ends by storing the carried
            // element, which is exactly what the next piece picks up as its first f
            for (int o = off0, hi = m - 1, end = (int)rots.size(); o < end;) {
                const int len = std::min(end - o, ROT_TILE - o % ROT_TILE);
                sweeps.push_back(Sweep{o, hi, len, 0});
                o += len;
                hi -= len;
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p;
            e[l] = g;
            e[m] = 0.0;
        }
    }
    return true;
}

}  // namespace

size_t eig_sym_work_doubles(int k) {
    // d, e, tau, 2 x pbuf (k each) + Vst (k*k) + Q (k*k)
    return (size_t)5 * k + 64 + (size_t)2 * k * k;
}

// S (k x k symmetric, device, ld k) is overwritten: row j = eigenvector of the j-th smallest eigenvalue.
// h_evals (host, k) receives the eigenvalues in ascending order; d_evals (device, k) too.
// work: eig_sym_work_doubles(k) doubles; rots_dev: device buffer for the rotations, grown through `grow_rots`
// (returns a device pointer with room for at least the requested number of bytes) -- also used for the rank table.
// Synchronises `st` (the tridiagonal crosses to the host).  Returns 0 ok, 1 CUDA error, 2 no convergence.
int eig_sym(double* S, int k, double* d_evals, double* h_evals, double* work, void* (*grow_rots)(void*, size_t),
            void* grow_ctx, int sm_count, cudaStream_t st, int64_t* launches) {
    if (k <= 0) return 0;
    double* d = work;
    double* e = d + k;
    double* tau = e + k;
    double* pbuf = tau + k;          // 2k (double buffered)
    double* Vst = pbuf + 2 * k;
    double* Qm = Vst + (size_t)k * k;
    if (k >= 3) {
        static int max_blocks_dev[DRE_MAX_DEVICES] = {};
        int& max_blocks = max_blocks_dev[current_device()];
        const size_t smem = (size_t)(3 * k + 40) * sizeof(double);
        if (max_blocks == 0 || smem > 40 * 1024) {
            cudaFuncSetAttribute(k_tridiag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024));
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tridiag, TD_THREADS, smem);
            max_blocks = std::max(1, per_sm) * sm_count;
        }
        // one warp per trailing row at the start; the barrier cost grows with the CTA count, the bandwidth of the
        // matrix-vector product too (tools/eig_grid_sweep.py on the B200: k = 420: 3.75 / 3.08 / 3.04 ms with 32 / 64 /
        // 96 CTAs, k = 690: 9.98 / 7.58 / 6.36 ms)
        int grid = std::min(std::min(max_blocks, k > 512 ? 96 : 64), (k + 7) / 8);
        grid = std::max(grid, 1);
        if (const char* ev = getenv("DRE_EIG_GRID")) grid = std::max(1, std::min(atoi(ev), max_blocks));
        void* args[] = {&S, &k, &d, &e, &Vst, &tau, &pbuf};
        if (cudaLaunchCooperativeKernel((void*)k_tridiag, dim3(grid), dim3(TD_THREADS), args, smem, st) != cudaSuccess)
            return 1;
        if (launches) *launches += 1;
    } else {
        // k = 1, 2: already tridiagonal
        if (cudaMemcpy2DAsync(d, sizeof(double), S, (size_t)(k + 1) * sizeof(double), sizeof(double), k,
                              cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return 1;
        if (k == 2 && cudaMemcpyAsync(e, S + 1, sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return 1;
    }
    const bool dbg = getenv("DRE_EIG_DEBUG") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(now() - t0).count();
    };
    auto t_start = now();
    double t_tri = 0.0, t_q = 0.0, t_ql = 0.0;
    if (dbg) { cudaStreamSynchronize(st); t_tri = ms_since(t_start); }
    std::vector<double> hd(k), he(std::max(k - 1, 1), 0.0);
    if (cudaMemcpyAsync(hd.data(), d, k * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (k > 1 && cudaMemcpyAsync(he.data(), e, (k - 1) * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return 1;
    // explicit Q (transposed), queued before the host works on the tridiagonal
    {
        const int blocks = (k * 32 + 255) / 256;
        if (k <= 256) k_form_qt<8><<<blocks, 256, 0, st>>>(Qm, k, Vst, tau);
        else if (k <= 512) k_form_qt<16><<<blocks, 256, 0, st>>>(Qm, k, Vst, tau);
        else if (k <= 768) k_form_qt<24><<<blocks, 256, 0, st>>>(Qm, k, Vst, tau);
        else if (k <= 1024) k_form_qt<32><<<blocks, 256, 0, st>>>(Qm, k, Vst, tau);
        else k_form_qt_big<<<blocks, 256, 0, st>>>(Qm, k, Vst, tau);
        if (launches) *launches += 1;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    if (dbg) t_q = ms_since(t_start);
    std::vector<Rot> rots;
    std::vector<Sweep> sweeps;
    rots.reserve((size_t)k * k + 64);
    sweeps.reserve((size_t)4 * k + 64);
    if (!tql_rotations(hd, he, rots, sweeps)) return 2;
    if (dbg) t_ql = ms_since(t_start);
    std::vector<int32_t> order(k), rank(k);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hd[a] < hd[b]; });
    for (int j = 0; j < k; ++j) {
        rank[order[j]] = j;
        h_evals[j] = hd[order[j]];
    }
    const int nrot = (int)rots.size(), ntiles = (nrot + ROT_TILE - 1) / ROT_TILE;
    std::vector<int32_t> tile_first(ntiles + 1, 0);
    {
        int q = 0;
        for (int t = 0; t < ntiles; ++t) {
            tile_first[t] = q;
            while (q < (int)sweeps.size() && sweeps[q].off < (t + 1) * ROT_TILE) ++q;
        }
        tile_first[ntiles] = (int)sweeps.size();
    }
    const size_t rot_bytes = rots.size() * sizeof(Rot), sw_off = (rot_bytes + 63) & ~(size_t)63;
    const size_t sw_bytes = sweeps.size() * sizeof(Sweep), tf_off = (sw_off + sw_bytes + 63) & ~(size_t)63;
    const size_t tf_bytes = tile_first.size() * sizeof(int32_t), rank_off = (tf_off + tf_bytes + 63) & ~(size_t)63;
    char* dev = (char*)grow_rots(grow_ctx, rank_off + (size_t)k * sizeof(int32_t) + 64);
    if (!dev) return 1;
    if (!rots.empty() && cudaMemcpyAsync(dev, rots.data(), rot_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (!sweeps.empty() &&
        cudaMemcpyAsync(dev + sw_off, sweeps.data(), sw_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return 1;
    if (cudaMemcpyAsync(dev + tf_off, tile_first.data(), tf_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (cudaMemcpyAsync(dev + rank_off, rank.data(), k * sizeof(int32_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return 1;
    if (cudaMemcpyAsync(d_evals, h_evals, k * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    {
        const int ldz = k + 1 + ((k & 1) ? 1 : 0);
        const size_t tile_bytes = 2 * ROT_TILE * sizeof(Rot);
        // the chains are latency bound and independent per row: few rows per CTA, so that they spread over many SMs
        int rows = (int)std::min<size_t>(8, (200 * 1024 - tile_bytes - 16) / ((size_t)ldz * sizeof(double)));
        if (rows < 1) return 1;   // k > 20 000: not a "small core" any more
        const size_t smem = ((size_t)rows * ldz + 1) * sizeof(double) + tile_bytes;
        static size_t attr_dev[DRE_MAX_DEVICES] = {};
        size_t& attr = attr_dev[current_device()];
        if (smem > attr && smem > 40 * 1024) {
            cudaFuncSetAttribute(k_apply_rot, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr = 200 * 1024;
        }
        k_apply_rot<<<(k + rows - 1) / rows, 128, smem, st>>>(Qm, k, (const Rot*)dev, nrot, (const Sweep*)(dev + sw_off),
                                                             (const int32_t*)(dev + tf_off), ntiles,
                                                             (const int32_t*)(dev + rank_off), S, rows);
        if (launches) *launches += 1;
    }
    // the host vectors (rots, rank) must outlive the asynchronous uploads
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    if (dbg)
        fprintf(stderr, "[dre eig] k %d: tridiag %.2f ms, +form Q %.2f, +host QL %.2f (%zu rotations, %zu sweeps), +apply %.2f\n",
                k, t_tri, t_q - t_tri, t_ql - t_q, rots.size(), sweeps.size(), ms_since(t_start) - t_ql);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dre
