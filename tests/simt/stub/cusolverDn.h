// TEST INFRASTRUCTURE ONLY -- stands in for <cusolverDn.h> under the host-side SIMT emulator (tests/simt/).
// The one cuSOLVER routine the product calls (Dsyevd on the small projected core of compress!) is replaced by a
// cyclic Jacobi eigensolver: slow, self-contained, accurate to a few ulp of the largest eigenvalue.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <numeric>
#include <vector>

typedef void* cusolverDnHandle_t;
typedef int cusolverStatus_t;
enum { CUSOLVER_STATUS_SUCCESS = 0 };
enum cusolverEigMode_t { CUSOLVER_EIG_MODE_NOVECTOR = 0, CUSOLVER_EIG_MODE_VECTOR = 1 };
enum cublasFillMode_t { CUBLAS_FILL_MODE_LOWER = 0, CUBLAS_FILL_MODE_UPPER = 1 };

static inline cusolverStatus_t cusolverDnCreate(cusolverDnHandle_t* h) { *h = (void*)1; return CUSOLVER_STATUS_SUCCESS; }
static inline cusolverStatus_t cusolverDnDestroy(cusolverDnHandle_t) { return CUSOLVER_STATUS_SUCCESS; }
static inline cusolverStatus_t cusolverDnSetStream(cusolverDnHandle_t, cudaStream_t) { return CUSOLVER_STATUS_SUCCESS; }
static inline cusolverStatus_t cusolverDnDsyevd_bufferSize(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int n,
                                                           const double*, int, const double*, int* lwork) {
    *lwork = n > 0 ? n : 1;
    return CUSOLVER_STATUS_SUCCESS;
}

// A (column-major, lda; the `uplo` triangle is read) -> eigenvectors in the columns of A, eigenvalues ascending in W
static inline cusolverStatus_t cusolverDnDsyevd(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t uplo, int n,
                                                double* A, int lda, double* W, double*, int, int* info) {
    std::vector<double> S((size_t)n * n), V((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            const bool lower = i >= j;
            const bool stored = (uplo == CUBLAS_FILL_MODE_LOWER) == lower || i == j;
            S[(size_t)i + (size_t)j * n] = stored ? A[(size_t)i + (size_t)j * lda] : A[(size_t)j + (size_t)i * lda];
        }
    for (int i = 0; i < n; ++i) V[(size_t)i + (size_t)i * n] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) (i == j ? diag : off) += S[(size_t)i + (size_t)j * n] * S[(size_t)i + (size_t)j * n];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = S[(size_t)p + (size_t)q * n];
                if (apq == 0.0) continue;
                const double app = S[(size_t)p + (size_t)p * n], aqq = S[(size_t)q + (size_t)q * n];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < n; ++k) {   // columns p, q of S
                    const double skp = S[(size_t)k + (size_t)p * n], skq = S[(size_t)k + (size_t)q * n];
                    S[(size_t)k + (size_t)p * n] = c * skp - sn * skq;
                    S[(size_t)k + (size_t)q * n] = sn * skp + c * skq;
                }
                for (int k = 0; k < n; ++k) {   // rows p, q of S
                    const double spk = S[(size_t)p + (size_t)k * n], sqk = S[(size_t)q + (size_t)k * n];
                    S[(size_t)p + (size_t)k * n] = c * spk - sn * sqk;
                    S[(size_t)q + (size_t)k * n] = sn * spk + c * sqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[(size_t)k + (size_t)p * n], vkq = V[(size_t)k + (size_t)q * n];
                    V[(size_t)k + (size_t)p * n] = c * vkp - sn * vkq;
                    V[(size_t)k + (size_t)q * n] = sn * vkp + c * vkq;
                }
            }
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(),
              [&](int a, int b) { return S[(size_t)a + (size_t)a * n] < S[(size_t)b + (size_t)b * n]; });
    for (int j = 0; j < n; ++j) {
        W[j] = S[(size_t)order[j] + (size_t)order[j] * n];
        for (int i = 0; i < n; ++i) A[(size_t)i + (size_t)j * lda] = V[(size_t)i + (size_t)order[j] * n];
    }
    if (info) *info = 0;
    return CUSOLVER_STATUS_SUCCESS;
}
