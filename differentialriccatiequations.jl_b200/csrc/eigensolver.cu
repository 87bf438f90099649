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

// Q (k x k, row-major Qm[i*k + c]) = H_0 H_1 ... H_{k-3}, H_j = I - tau_j [0; v_j][0; v_j]' acting on rows j+1..k-1
__global__ void __launch_bounds__(128) k_form_q(double* Qm, int k, const double* Vst, const double* tau) {
    extern __shared__ double vs[];   // [k]
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < k; ++i)
        if (c < k) Qm[(size_t)i * k + c] = (i == c) ? 1.0 : 0.0;
    for (int j = k - 3; j >= 0; --j) {
        const int m = k - j - 1;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) vs[i] = Vst[(size_t)j * k + i];
        __syncthreads();
        const double t = tau[j];
        if (c < j + 1 || c >= k || t == 0.0) continue;
        double* col = Qm + (size_t)(j + 1) * k + c;
        double z = 0.0;
        for (int i = 0; i < m; ++i) z += vs[i] * col[(size_t)i * k];
        z *= t;
        for (int i = 0; i < m; ++i) col[(size_t)i * k] -= z * vs[i];
    }
}

struct Rot {
    double c, s;
    int i, pad;
};

// rows [r0, r0 + nr) of Q (row-major): z[i+1] = s z[i] + c f,  z[i] = c z[i] - s f  for every recorded rotation,
// then Out[rank[i]*k + row] = z[i]
__global__ void k_apply_rot(const double* Qm, int k, const Rot* rots, int nrot, const int32_t* rank, double* Out,
                            int rows_per_cta) {
    extern __shared__ double zs[];   // [rows_per_cta][k + 1]
    const int ldz = k + 1 + ((k & 1) ? 1 : 0);   // odd leading dimension: lanes of a warp hit different banks
    const int r0 = blockIdx.x * rows_per_cta;
    const int nr = min(rows_per_cta, k - r0);
    for (int idx = threadIdx.x; idx < nr * k; idx += blockDim.x) {
        const int r = idx / k, i = idx - r * k;
        zs[r * ldz + i] = Qm[(size_t)(r0 + r) * k + i];
    }
    __syncthreads();
    if ((int)threadIdx.x < nr) {
        double* z = zs + threadIdx.x * ldz;
        for (int q = 0; q < nrot; ++q) {
            const Rot R = rots[q];
            const double f = z[R.i + 1], g = z[R.i];
            z[R.i + 1] = R.s * g + R.c * f;
            z[R.i] = R.c * g - R.s * f;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nr * k; idx += blockDim.x) {
        const int i = idx / nr, r = idx - i * nr;
        Out[(size_t)rank[i] * k + (r0 + r)] = zs[r * ldz + i];
    }
}

// implicit QL with Wilkinson shifts on the symmetric tridiagonal (d, e); rotations are recorded in application order
bool tql_rotations(std::vector<double>& d, std::vector<double>& e, std::vector<Rot>& rots) {
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
            for (; i >= l; --i) {
                double f = s * e[i];
                const double b = c * e[i];
                r = std::hypot(f, g);
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
                rots.push_back(Rot{c, s, i, 0});
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
        // one warp per trailing row at the start; the barrier cost grows with the CTA count
        int grid = std::min(std::min(max_blocks, 64), (k + 7) / 8);
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
    std::vector<double> hd(k), he(std::max(k - 1, 1), 0.0);
    if (cudaMemcpyAsync(hd.data(), d, k * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
    if (k > 1 && cudaMemcpyAsync(he.data(), e, (k - 1) * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return 1;
    // explicit Q, queued before the host works on the tridiagonal
    {
        const size_t smem = (size_t)k * sizeof(double);
        static bool attr_dev[DRE_MAX_DEVICES] = {};
        bool& attr = attr_dev[current_device()];
        if (!attr && smem > 40 * 1024) {
            cudaFuncSetAttribute(k_form_q, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr = true;
        }
        k_form_q<<<(k + 127) / 128, 128, smem, st>>>(Qm, k, Vst, tau);
        if (launches) *launches += 1;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    std::vector<Rot> rots;
    rots.reserve((size_t)k * k);
    if (getenv("DRE_EIG_DEBUG")) {
        int bad = 0;
        double dmax = 0.0, emax = 0.0;
        for (int i = 0; i < k; ++i) { if (!std::isfinite(hd[i])) ++bad; else dmax = std::max(dmax, std::fabs(hd[i])); }
        for (int i = 0; i + 1 < k; ++i) { if (!std::isfinite(he[i])) ++bad; else emax = std::max(emax, std::fabs(he[i])); }
        fprintf(stderr, "[dre eig] k %d non-finite %d max|d| %.3e max|e| %.3e d0 %.6e e0 %.6e\n", k, bad, dmax, emax, hd[0],
                he[0]);
    }
    if (!tql_rotations(hd, he, rots)) return 2;
    std::vector<int32_t> order(k), rank(k);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return hd[a] < hd[b]; });
    for (int j = 0; j < k; ++j) {
        rank[order[j]] = j;
        h_evals[j] = hd[order[j]];
    }
    const size_t rot_bytes = rots.size() * sizeof(Rot), rank_off = (rot_bytes + 63) & ~(size_t)63;
    char* dev = (char*)grow_rots(grow_ctx, rank_off + (size_t)k * sizeof(int32_t) + 64);
    if (!dev) return 1;
    if (!rots.empty() && cudaMemcpyAsync(dev, rots.data(), rot_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    if (cudaMemcpyAsync(dev + rank_off, rank.data(), k * sizeof(int32_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
        return 1;
    if (cudaMemcpyAsync(d_evals, h_evals, k * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return 1;
    {
        const int ldz = k + 1 + ((k & 1) ? 1 : 0);
        int rows = (int)std::min<size_t>(32, (200 * 1024) / ((size_t)ldz * sizeof(double)));
        if (rows < 1) return 1;   // k > 25 000: not a "small core" any more
        const size_t smem = (size_t)rows * ldz * sizeof(double);
        static size_t attr_dev[DRE_MAX_DEVICES] = {};
        size_t& attr = attr_dev[current_device()];
        if (smem > attr && smem > 40 * 1024) {
            cudaFuncSetAttribute(k_apply_rot, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            attr = 200 * 1024;
        }
        const int threads = std::max(32, ((rows + 31) / 32) * 32) * 4;   // extra warps for the tile load / store
        k_apply_rot<<<(k + rows - 1) / rows, threads, smem, st>>>(Qm, k, (const Rot*)dev, (int)rots.size(),
                                                                 (const int32_t*)(dev + rank_off), S, rows);
        if (launches) *launches += 1;
    }
    // the host vectors (rots, rank) must outlive the asynchronous uploads
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace dre
