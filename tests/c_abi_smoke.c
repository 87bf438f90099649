/* Plain-C driver of the C ABI (include/dre_b200.h): what a non-Python host -- the Julia glue of
 * julia/DREB200.jl binds exactly these symbols with ccall -- does for one ADI step of
 * src/lyapunov/adi.jl:149-179 (/root/reference):
 *
 *   dre_create -> dre_set_pencil -> dre_mat_create/upload -> dre_set_operator -> dre_adi_step
 *   -> dre_ldlt_norm -> dre_mat_download -> dre_destroy
 *
 * and checks the result on the host with nothing but CSC mat-vecs:
 *   (A' - Vt U' + mu E') V = R_old      (closed-loop operator F = A + inv(alpha) U Vt', alpha = -1)
 *   R_new = R_old - 2 mu E V
 *   |alpha| ||R D R'||_F from dre_ldlt_norm == the same norm formed densely on the host
 * Pencil: 2D 5-point grid (nx x ny), E = diagonally dominant "mass" matrix, A = -(stiffness) - 0.1 E.
 * Test infrastructure: compiled by __graft_entry__.build(), run by tests/test_gpu_cabi_c.py (-m gpu).
 * Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dre_b200.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int32_t rc_ = (call);                                                                    \
        if (rc_ != DRE_OK) {                                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, (int)rc_, dre_last_error(ctx));       \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

/* y += a * M x, M in CSC (0-based here) */
static void csc_axpy(int64_t n, const int64_t* cp, const int64_t* ri, const double* nz, double a, const double* x,
                     double* y) {
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = cp[j]; p < cp[j + 1]; ++p) y[ri[p]] += a * nz[p] * x[j];
}

int main(void) {
    const int nx = 37, ny = 29;
    const int64_t n = (int64_t)nx * ny;
    const int r = 5, m = 2;
    dre_context* ctx = NULL;

    /* ---- pencil in CSC with 64-bit 0-based indices ---- */
    int64_t* cp = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t* ri = (int64_t*)malloc((size_t)n * 5 * sizeof(int64_t));
    double* ez = (double*)malloc((size_t)n * 5 * sizeof(double));
    double* az = (double*)malloc((size_t)n * 5 * sizeof(double));
    int64_t nnz = 0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const int64_t c = (int64_t)j * nx + i;
            /* rows in increasing order: (i, j-1), (i-1, j), (i, j), (i+1, j), (i, j+1) */
            const int di[5] = {0, -1, 0, 1, 0}, dj[5] = {-1, 0, 0, 0, 1};
            for (int k = 0; k < 5; ++k) {
                const int ii = i + di[k], jj = j + dj[k];
                if (ii < 0 || ii >= nx || jj < 0 || jj >= ny) continue;
                ri[nnz] = (int64_t)jj * nx + ii;
                if (k == 2) {
                    ez[nnz] = 1.0 + 0.01 * ((i * 7 + j * 3) % 5);
                    az[nnz] = -4.0 - 0.1 * ez[nnz];
                } else {
                    ez[nnz] = 0.1;
                    az[nnz] = 1.0 - 0.1 * 0.1;
                }
                ++nnz;
            }
            cp[c + 1] = nnz;
        }

    /* ---- right-hand side R (n x r), low-rank factors U = B (n x m), Vt = K' (n x m), column-major ---- */
    double* R = (double*)malloc((size_t)n * r * sizeof(double));
    double* U = (double*)malloc((size_t)n * m * sizeof(double));
    double* Vt = (double*)malloc((size_t)n * m * sizeof(double));
    unsigned s = 12345u;
    for (int64_t i = 0; i < n * r; ++i) { s = s * 1664525u + 1013904223u; R[i] = (double)(s >> 8) / 16777216.0 - 0.5; }
    for (int64_t i = 0; i < n * m; ++i) {
        s = s * 1664525u + 1013904223u; U[i] = ((double)(s >> 8) / 16777216.0 - 0.5) * 0.3;
        s = s * 1664525u + 1013904223u; Vt[i] = ((double)(s >> 8) / 16777216.0 - 0.5) * 0.3;
    }
    double* D = (double*)calloc((size_t)r * r, sizeof(double));
    for (int i = 0; i < r; ++i) D[i + i * r] = (i % 2) ? -1.0 - i : 1.0 + i;   /* indefinite diagonal core */

    if (dre_create(0, &ctx) != DRE_OK) {
        fprintf(stderr, "dre_create failed: %s\n", dre_last_error(NULL));
        return 1;
    }
    printf("%s\n", dre_version());
    CHECK(dre_set_pencil(ctx, n, cp, ri, ez, cp, ri, az, 0));
    dre_symbolic_info info;
    CHECK(dre_get_symbolic_info(ctx, &info));
    printf("n=%lld nnz(L)=%lld supernodes=%d levels=%d\n", (long long)info.n, (long long)info.nnz_L, info.nsupernodes,
           info.nlevels);

    int32_t idR, idV, idU, idK;
    CHECK(dre_mat_create(ctx, r, &idR));
    CHECK(dre_mat_create(ctx, r, &idV));
    CHECK(dre_mat_create(ctx, m, &idU));
    CHECK(dre_mat_create(ctx, m, &idK));
    const dre_view vR = {idR, 0, r}, vV = {idV, 0, r}, vU = {idU, 0, m}, vK = {idK, 0, m}, none = {-1, 0, 0};
    CHECK(dre_mat_upload(ctx, vR, R, n));
    CHECK(dre_mat_upload(ctx, vU, U, n));
    CHECK(dre_mat_upload(ctx, vK, Vt, n));

    /* norm(::LDLt) before the step, against the dense formula */
    double nrm = 0.0;
    CHECK(dre_ldlt_norm(ctx, vR, D, r, -2.0, &nrm));
    {
        /* || a R D R' ||_F^2 = a^2 tr(G D G D), G = R'R */
        double G[25], M[25], tr = 0.0;
        for (int a = 0; a < r; ++a)
            for (int b = 0; b < r; ++b) {
                double acc = 0.0;
                for (int64_t i = 0; i < n; ++i) acc += R[i + a * n] * R[i + b * n];
                G[a + b * r] = acc;
            }
        for (int a = 0; a < r; ++a)
            for (int b = 0; b < r; ++b) M[a + b * r] = G[a + b * r] * D[b + b * r];
        for (int a = 0; a < r; ++a)
            for (int b = 0; b < r; ++b) tr += M[a + b * r] * M[b + a * r];
        const double ref = 2.0 * sqrt(tr);
        printf("ldlt_norm %.15e host %.15e\n", nrm, ref);
        if (fabs(nrm - ref) > 1e-12 * ref) { fprintf(stderr, "dre_ldlt_norm mismatch\n"); return 1; }
    }

    /* one real ADI step: (F' + mu E') V = R with F = A + inv(alpha) U Vt', i.e. F' = A' + inv(alpha) Vt U' */
    const double mu = -0.75, alpha = -1.0;
    CHECK(dre_set_operator(ctx, 1.0, 0.0, alpha, vU, vK));
    CHECK(dre_prefactor(ctx, mu, 0.0));                      /* performance hint; results are unchanged */
    CHECK(dre_adi_step(ctx, mu, 0.0, vR, vV, none));
    double* V = (double*)malloc((size_t)n * r * sizeof(double));
    double* Rn = (double*)malloc((size_t)n * r * sizeof(double));
    CHECK(dre_mat_download(ctx, vV, V, n));
    CHECK(dre_mat_download(ctx, vR, Rn, n));
    CHECK(dre_sync(ctx));

    double worst_solve = 0.0, worst_upd = 0.0;
    double* y = (double*)malloc((size_t)n * sizeof(double));
    for (int c = 0; c < r; ++c) {
        const double* v = V + (int64_t)c * n;
        /* y = (A + mu E) v + inv(alpha) Vt (U' v) - R_old(:, c)    (pencil symmetric: A' = A, E' = E) */
        memset(y, 0, (size_t)n * sizeof(double));
        csc_axpy(n, cp, ri, az, 1.0, v, y);
        csc_axpy(n, cp, ri, ez, mu, v, y);
        for (int k = 0; k < m; ++k) {
            double dotk = 0.0;
            for (int64_t i = 0; i < n; ++i) dotk += U[i + (int64_t)k * n] * v[i];
            for (int64_t i = 0; i < n; ++i) y[i] += (1.0 / alpha) * Vt[i + (int64_t)k * n] * dotk;
        }
        double num = 0.0, den = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const double d = y[i] - R[i + (int64_t)c * n];
            num += d * d;
            den += R[i + (int64_t)c * n] * R[i + (int64_t)c * n];
        }
        if (sqrt(num / den) > worst_solve) worst_solve = sqrt(num / den);
        /* R_new = R_old - 2 mu E v */
        memset(y, 0, (size_t)n * sizeof(double));
        csc_axpy(n, cp, ri, ez, -2.0 * mu, v, y);
        num = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const double d = R[i + (int64_t)c * n] + y[i] - Rn[i + (int64_t)c * n];
            num += d * d;
        }
        if (sqrt(num / den) > worst_upd) worst_upd = sqrt(num / den);
    }
    printf("relative residual of the shifted closed-loop solve %.3e, of the residual update %.3e\n", worst_solve,
           worst_upd);
    if (!(worst_solve < 1e-11) || !(worst_upd < 1e-13)) { fprintf(stderr, "ADI step check failed\n"); return 1; }

    dre_stats st;
    CHECK(dre_stats_get(ctx, &st));
    printf("kernel launches %lld, factorizations %lld (prefactorized %lld), solves %lld\n", (long long)st.kernel_launches,
           (long long)st.factorizations, (long long)st.prefactor_hits, (long long)st.solves);
    if (st.kernel_launches <= 0 || st.factorizations < 1) { fprintf(stderr, "no GPU work recorded\n"); return 1; }
    CHECK(dre_mat_free(ctx, idR));
    CHECK(dre_mat_free(ctx, idV));
    CHECK(dre_mat_free(ctx, idU));
    CHECK(dre_mat_free(ctx, idK));
    if (dre_destroy(ctx) != DRE_OK) return 1;
    printf("C_ABI_SMOKE_OK\n");
    return 0;
}
