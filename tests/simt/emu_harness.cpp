// TEST INFRASTRUCTURE ONLY.  C entry points that run the *kernel sources* of differentialriccatiequations.jl_b200/csrc
// (sparse_kernels.cu, dense_kernels.cu, compiled by g++ against tests/simt/stub/cuda_runtime.h) on the host-side
// SIMT emulator, with the product's own launch schedule (csrc/schedule.h) and symbolic analysis
// (csrc/symbolic.cpp).  Loaded by tests/test_simt_kernels.py through ctypes; the product never links this.
#define SIMT_EMU_IMPL 1
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"
#include "schedule.h"
#include "symbolic.h"

using namespace dre;

namespace {

struct Emu {
    Symbolic S;
    DevSymbolic dS{};
    LevelLists lists;
    std::vector<int32_t> sn_first32;
    std::vector<double> L, Linv, dvec, U;   // sized for complex
    int32_t errflag = 0;
    int tw = 1;
    bool m21 = false;   // panels hold M21 = L21 Linv (row-split sweeps)
    std::string err;
};

DevSchedule schedule_of(const Emu* e) {
    return DevSchedule{e->lists.levels.data(), (int)e->lists.levels.size(), e->S.level_sn.data(),
                       e->lists.ea_parents.data(), e->lists.l21_items.data(), e->lists.schur_items.data(),
                       e->lists.fwd2_items.data(), e->lists.bwd2_items.data()};
}

}  // namespace

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API int emu_open(int64_t n, const int64_t* Ecp, const int64_t* Eri, const double* Enz, const int64_t* Acp,
                     const int64_t* Ari, const double* Anz, int base, int leaf, int maxsn, void** out) {
    Emu* e = new Emu();
    AnalyzeOptions opt;
    if (leaf > 0) opt.leaf_size = leaf;
    if (maxsn > 0) opt.max_snode = maxsn;
    opt.max_snode = std::min(opt.max_snode, (int)SN_MAX);
    e->err = analyze(n, Ecp, Eri, Enz, Acp, Ari, Anz, base, opt, e->S);
    if (!e->err.empty()) {
        fprintf(stderr, "emu_open: %s\n", e->err.c_str());
        delete e;
        return 1;
    }
    const Symbolic& S = e->S;
    DevSymbolic& D = e->dS;
    D.n = S.n; D.nsn = S.nsn; D.nlevels = S.nlevels;
    D.sn_first = S.sn_first.data(); D.sn_rowptr = S.sn_rowptr.data(); D.sn_rows = S.sn_rows.data();
    D.relmap = S.relmap.data(); D.child_ptr = S.child_ptr.data(); D.child_idx = S.child_idx.data();
    D.panel_off = S.panel_off.data(); D.linv_off = S.linv_off.data(); D.upd_off = S.upd_off.data();
    D.rhs_off = S.rhs_off.data();
    D.nasm = (int64_t)S.asm_dest.size(); D.asm_dest = S.asm_dest.data(); D.asm_a = S.asm_a.data(); D.asm_e = S.asm_e.data();
    build_level_lists(S, e->lists);
    e->L.assign((size_t)std::max<int64_t>(S.nnz_L, 1) * 2, 0.0);
    e->Linv.assign((size_t)std::max<int64_t>(S.linv_off[S.nsn], 1) * 2, 0.0);
    e->dvec.assign((size_t)S.n * 2, 0.0);
    e->U.assign((size_t)std::max<int64_t>(S.upd_total, 1) * 2, 0.0);
    *out = e;
    return 0;
}

EMU_API void emu_close(void* h) { delete (Emu*)h; }

// out: n, nsn, nlevels, nnz_L, linv_elems, upd_elems, sum_u, max_sn
EMU_API void emu_sizes(void* h, int64_t* out) {
    const Symbolic& S = ((Emu*)h)->S;
    out[0] = S.n; out[1] = S.nsn; out[2] = S.nlevels; out[3] = S.nnz_L; out[4] = S.linv_off[S.nsn];
    out[5] = S.upd_total; out[6] = S.rhs_total; out[7] = S.max_sn;
}

EMU_API void emu_perm(void* h, int32_t* perm) {
    const Symbolic& S = ((Emu*)h)->S;
    memcpy(perm, S.perm.data(), sizeof(int32_t) * (size_t)S.n);
}

template <class T>
static int factor_t(Emu* e, double a, T emu, bool m21) {
    const Symbolic& S = e->S;
    T* L = (T*)e->L.data();
    T* U = (T*)e->U.data();
    memset(L, 0, (size_t)S.nnz_L * sizeof(T));
    memset((void*)U, 0xFF, (size_t)S.upd_total * sizeof(T));   // the factorization zeroes its pool level by level
    // poison what the factorization must fully define itself
    memset(e->Linv.data(), 0xFF, e->Linv.size() * sizeof(double));
    memset(e->dvec.data(), 0xFF, e->dvec.size() * sizeof(double));
    e->errflag = 0;
    int64_t launches = 0;
    enqueue_factor<T>(e->dS, schedule_of(e), L, (T*)e->Linv.data(), (T*)e->dvec.data(), U, a, emu, &e->errflag, nullptr,
                      &launches, m21);
    e->tw = (int)(sizeof(T) / sizeof(double));
    e->m21 = m21;
    return e->errflag;
}

EMU_API int emu_factor(void* h, int is_cplx, double a, double emu_re, double emu_im, int m21) {
    Emu* e = (Emu*)h;
    return is_cplx ? factor_t<cplx>(e, a, mk(emu_re, emu_im), m21 != 0) : factor_t<double>(e, a, emu_re, m21 != 0);
}

// what: 0 = L (nnz_L), 1 = Linv (linv_elems), 2 = dvec (n); tw doubles per element
EMU_API void emu_get(void* h, int what, double* buf) {
    Emu* e = (Emu*)h;
    const Symbolic& S = e->S;
    const std::vector<double>& v = what == 0 ? e->L : what == 1 ? e->Linv : e->dvec;
    const int64_t cnt = what == 0 ? S.nnz_L : what == 1 ? S.linv_off[S.nsn] : S.n;
    memcpy(buf, v.data(), sizeof(double) * (size_t)cnt * e->tw);
}

template <class T>
static void sweeps_t(Emu* e, const RhsSource& src, T* W, int64_t ldw, int nrhs, int64_t* launches) {
    const Symbolic& S = e->S;
    std::vector<T> tbuf((size_t)std::max<int64_t>(S.rhs_total, 1) * ldw);
    memset((void*)tbuf.data(), 0xFF, tbuf.size() * sizeof(T));
    memset((void*)W, 0xFF, sizeof(T) * (size_t)S.n * ldw);
    const T* L = (const T*)e->L.data();
    const T* Linv = (const T*)e->Linv.data();
    const T* dvec = (const T*)e->dvec.data();
    if (e->m21) {
        std::vector<T> Y((size_t)S.n * ldw);
        memset((void*)Y.data(), 0xFF, Y.size() * sizeof(T));
        enqueue_sweeps2<T>(e->dS, schedule_of(e), L, Linv, dvec, W, Y.data(), ldw, nrhs, tbuf.data(), src, nullptr,
                           launches);
    } else {
        enqueue_sweeps<T>(e->dS, schedule_of(e), L, Linv, dvec, W, ldw, nrhs, tbuf.data(), src, nullptr, launches);
    }
}

// Block solve of the current factorization.  R (n x r, row-major ld ldr) and Vt (n x m, row-major ld ldv) are in
// SOLVER ordering; W (n x ldw elements of tw doubles) receives the solution of all r + m columns.
EMU_API int emu_sweeps(void* h, const double* R, int64_t ldr, int r, const double* Vt, int64_t ldv, int m, double* W,
                       int64_t ldw) {
    Emu* e = (Emu*)h;
    const Symbolic& S = e->S;
    const RhsSource src{R, ldr, r, m ? Vt : nullptr, m ? ldv : 0};
    const int nrhs = r + m;
    int64_t launches = 0;
    if (e->tw == 1) sweeps_t<double>(e, src, W, ldw, nrhs, &launches);
    else sweeps_t<cplx>(e, src, (cplx*)W, ldw, nrhs, &launches);
    return (int)launches;
}

// SMW core + epilogue on a solved block W (element type of the current factorization)
EMU_API void emu_smw(void* h, const double* BtW, int m, int r, double alpha, const double* W, int64_t ldw, int mode,
                     double d, double* V1, int64_t ld1, double* V2, int64_t ld2, int* errflag) {
    Emu* e = (Emu*)h;
    const int64_t n = e->S.n;
    int64_t launches = 0;
    int32_t ef = 0;
    if (e->tw == 1) {
        std::vector<double> Sol((size_t)std::max(1, m * r));
        launch_smw_core<double>(BtW, r + m, m, r, alpha, Sol.data(), &ef, nullptr, &launches);
        launch_smw_apply<double>(W, ldw, r, m, m ? Sol.data() : nullptr, mode, d, V1, ld1, V2, ld2, n, nullptr, &launches);
    } else {
        std::vector<cplx> Sol((size_t)std::max(1, m * r));
        launch_smw_core<cplx>((const cplx*)BtW, r + m, m, r, alpha, Sol.data(), &ef, nullptr, &launches);
        launch_smw_apply<cplx>((const cplx*)W, ldw, r, m, m ? Sol.data() : nullptr, mode, d, V1, ld1, V2, ld2, n, nullptr,
                               &launches);
    }
    *errflag = ef;
}

// SpMM with the permuted CSR copies of the pencil: which = 0 (A) or 1 (E)
EMU_API void emu_spmm(void* h, int which, double alpha, const double* X, int64_t ldx, double beta, double* Y,
                      int64_t ldy, int cols) {
    Emu* e = (Emu*)h;
    const Symbolic& S = e->S;
    int64_t launches = 0;
    launch_spmm(S.csr_ptr.data(), S.csr_col.data(), which ? S.csr_e.data() : S.csr_a.data(), S.n, alpha, X, ldx, beta,
                Y, ldy, cols, nullptr, &launches);
}

// ---------------- dense toolbox ----------------
EMU_API void emu_gram(const double* X, int64_t ldx, int a, const double* Y, int64_t ldy, int b, int64_t n,
                      const double* roww, double* out1, int64_t ld1, double* out2, int64_t ld2, int sm_count) {
    GramPlan plan = gram_plan(n, a, b, sm_count);
    std::vector<double> partial(std::max<size_t>(plan.partial_elems, 1));
    memset(partial.data(), 0xFF, partial.size() * sizeof(double));
    int64_t launches = 0;
    launch_gram(X, ldx, a, Y, ldy, b, n, roww, partial.data(), plan, out1, ld1, out2, ld2, nullptr, &launches);
}

EMU_API void emu_tall_gemm(double alpha, const double* X, int64_t ldx, int a, const double* W, int64_t ldw, int w_trans,
                           double beta, double* Y, int64_t ldy, int b, int64_t n) {
    int64_t launches = 0;
    launch_tall_gemm(alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n, nullptr, &launches);
}

EMU_API void emu_pivchol(const double* G, int64_t ldg, int pb, double drop2, double rel2, double* Wsel, int32_t* info,
                         double* dinfo) {
    int64_t launches = 0;
    launch_pivchol(G, ldg, pb, drop2, rel2, Wsel, info, dinfo, nullptr, &launches);
}

EMU_API void emu_norm_diag(const double* G, int64_t ldg, int r, const double* t, double* out) {
    int64_t launches = 0;
    launch_norm_diag(G, ldg, r, t, out, nullptr, &launches);
}

EMU_API void emu_colnorm2(const double* P, int64_t ldp, int64_t n, int cols, int nblk, double* out) {
    std::vector<double> partial((size_t)nblk * cols);
    int64_t launches = 0;
    launch_colnorm2(P, ldp, n, cols, partial.data(), nblk, out, nullptr, &launches);
}

EMU_API void emu_panel_transposes(double* panel, int64_t ldd, const double* cm, int64_t lds, int64_t n, int cols,
                                  const int32_t* iperm, double* back, int64_t ldb) {
    int64_t launches = 0;
    launch_colmajor_to_panel(panel, ldd, cm, lds, n, cols, iperm, nullptr, &launches);
    launch_panel_to_colmajor(back, ldb, panel, ldd, n, cols, iperm, nullptr, &launches);
}

// tuning knobs of the launchers (so that small test problems reach every kernel variant)
EMU_API void emu_set_diag_narrow_min(int v) { dre::diag_narrow_min = v; }
EMU_API void emu_set_diag_variant(int v) { dre::diag_variant = v; }
EMU_API void emu_set_spmm_variant(int v) { dre::spmm_variant = v; }

EMU_API void emu_counters(long* out) {
    out[0] = simt::M().launches;
    out[1] = simt::M().ctas;
    out[2] = simt::M().switches;
}
