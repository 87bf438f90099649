// Shared device helpers: scalar traits for real / complex-symmetric arithmetic, FP64 tensor-core
// (DMMA m8n8k4) wrapper.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dre {

struct cplx {
    double x, y;
};

__host__ __device__ __forceinline__ cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }

// ---- arithmetic usable for T = double and T = cplx (no conjugation anywhere: complex SYMMETRIC) ----
__host__ __device__ __forceinline__ double zero_of(double*) { return 0.0; }
__host__ __device__ __forceinline__ cplx zero_of(cplx*) { return mk(0.0, 0.0); }
template <class T> __host__ __device__ __forceinline__ T zero() { return zero_of((T*)nullptr); }
__host__ __device__ __forceinline__ double one_of(double*) { return 1.0; }
__host__ __device__ __forceinline__ cplx one_of(cplx*) { return mk(1.0, 0.0); }
template <class T> __host__ __device__ __forceinline__ T one() { return one_of((T*)nullptr); }

__host__ __device__ __forceinline__ double add(double a, double b) { return a + b; }
__host__ __device__ __forceinline__ cplx add(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ double sub(double a, double b) { return a - b; }
__host__ __device__ __forceinline__ cplx sub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ double mul(double a, double b) { return a * b; }
__host__ __device__ __forceinline__ cplx mul(cplx a, cplx b) {
    return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ cplx mul(double a, cplx b) { return mk(a * b.x, a * b.y); }
// acc += a*b
__host__ __device__ __forceinline__ void fma_acc(double& acc, double a, double b) { acc = fma(a, b, acc); }
__host__ __device__ __forceinline__ void fma_acc(cplx& acc, cplx a, cplx b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
__host__ __device__ __forceinline__ double recip(double a) { return 1.0 / a; }
__host__ __device__ __forceinline__ cplx recip(cplx a) {
    // Smith's algorithm is not needed: |a| is O(pivot) and far from over/underflow
    double d = a.x * a.x + a.y * a.y;
    return mk(a.x / d, -a.y / d);
}
__host__ __device__ __forceinline__ bool is_bad(double a) { return !(a == a) || a == 0.0 || isinf(a); }
__host__ __device__ __forceinline__ bool is_bad(cplx a) {
    return !(a.x == a.x) || !(a.y == a.y) || (a.x == 0.0 && a.y == 0.0) || isinf(a.x) || isinf(a.y);
}
// real scalar -> T
__host__ __device__ __forceinline__ void from_real(double v, double& out) { out = v; }
__host__ __device__ __forceinline__ void from_real(double v, cplx& out) { out = mk(v, 0.0); }

// ---- FP64 tensor core: D(8x8) += A(8x4, row) * B(4x8, col) ----
// fragment layout (PTX ISA, mma.m8n8k4 .f64): lane l holds
//   a = A[l/4][l%4],  b = B[l%4][l/4],  c0,c1 = C[l/4][2*(l%4) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
#ifdef DRE_SIMT_EMU
    simt::dmma884(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

// Function attributes (cudaFuncSetAttribute) are per device; the "already set" flags of the launchers are arrays
// indexed by the device ordinal.
#define DRE_MAX_DEVICES 64
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < DRE_MAX_DEVICES) ? dev : 0;
}

// ---- kernel launch / dynamic shared memory spelling ----
// The kernel sources are also compiled by g++ against the host-side SIMT emulator of the CPU test tier
// (tests/simt/stub/cuda_runtime.h, -DDRE_SIMT_EMU), which supplies its own versions of these three macros;
// the product build sees plain CUDA.
#define DRE_UNPAREN(...) __VA_ARGS__
#ifndef DRE_SIMT_EMU
#define DRE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    DRE_UNPAREN kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define DRE_DYN_SMEM(type, name) extern __shared__ type name[]
#define DRE_DYN_SMEM_ALIGNED(type, name) extern __shared__ __align__(16) type name[]
#endif

}  // namespace dre
