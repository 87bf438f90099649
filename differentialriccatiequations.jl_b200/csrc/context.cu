// C ABI of libdre_b200.so (include/dre_b200.h): context, device memory, orchestration of the
// hand-written kernels.  Host-side C++17; no third-party device library: the small (rho x rho) projected core
// inside compress! (LAPACK syevr in the reference, src/LDLt.jl:214) goes through the in-tree eigensolver of
// eigensolver.cu.  (Only the host-side SIMT emulator of the CPU test tier, which cannot run a cooperative
// launch, substitutes its own Jacobi routine behind the cuSOLVER-shaped stub of tests/simt/stub/.)
#include <cuda_runtime.h>
#ifdef DRE_SIMT_EMU
#include <cusolverDn.h>
#endif

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#ifndef DRE_SIMT_EMU
#include <nvtx3/nvToolsExt.h>   // header-only (the injection library is looked up at run time; no link dependency)
#endif

#include "../../include/dre_b200.h"
#include "kernels.h"
#include "schedule.h"
#include "symbolic.h"

using namespace dre;

namespace {

thread_local std::string g_last_error;

// Arena allocator for the panels.  Everything that touches panels runs on ONE stream, so a block released by
// dre_mat_free can be handed out again immediately (its next use is stream-ordered behind the last one).
// cudaMalloc / cudaFree of the 0.1 - 2 GB panels of this workload were measured at 0.3 - 75 ms per call
// (tools/host_profile.py, DRE_TRACE=1) and the driver's stream-ordered pool at 25 ms, which made the ADI loop
// host-bound; the arena grabs a few large chunks (2, 4, 8, ... GB) once and sub-allocates first-fit with
// coalescing.  Chunks go back to the driver when the context dies or a new pencil is set.
struct Arena {
    struct Chunk {
        char* base = nullptr;
        size_t size = 0;
        std::map<size_t, size_t> free;   // offset -> length of the free ranges
    };
    std::vector<Chunk> chunks;
    size_t next_chunk = (size_t)2 << 30;
    static size_t align_up(size_t b) { return (b + 511) & ~(size_t)511; }
    void* alloc(size_t bytes, size_t* got) {
#ifdef DRE_SIMT_EMU
        // emulator test tier (tests/simt/): one NaN-filled heap block per request, exactly as large as asked for, so
        // that an address sanitizer sees every overrun and a read of never-written memory poisons the result
        *got = bytes;
        void* q = malloc(std::max<size_t>(bytes, 1));
        if (q) memset(q, 0xFF, bytes);
        return q;
#endif
        bytes = align_up(std::max<size_t>(bytes, 512));
        for (int pass = 0; pass < 2; ++pass) {
            for (Chunk& ch : chunks)
                for (auto it = ch.free.begin(); it != ch.free.end(); ++it)
                    if (it->second >= bytes) {
                        const size_t off = it->first, len = it->second;
                        ch.free.erase(it);
                        if (len > bytes) ch.free[off + bytes] = len - bytes;
                        *got = bytes;
                        return ch.base + off;
                    }
            if (pass == 1) break;
            Chunk ch;
            ch.size = std::max(next_chunk, align_up(bytes));
            const auto t0 = std::chrono::steady_clock::now();
            struct Report {
                const Chunk& ch; std::chrono::steady_clock::time_point t0;
                ~Report() {
                    if (getenv("DRE_TRACE_ALLOC"))
                        fprintf(stderr, "[dre alloc] new arena chunk %.2f GB in %.1f ms\n", ch.size / 1073741824.0,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
                }
            } report{ch, t0};
            if (cudaMalloc((void**)&ch.base, ch.size) != cudaSuccess) {
                cudaGetLastError();
                ch.size = align_up(bytes);   // memory is tight: exactly what is needed
                if (cudaMalloc((void**)&ch.base, ch.size) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            }
            ch.free[0] = ch.size;
            chunks.push_back(std::move(ch));
            next_chunk = std::min<size_t>(next_chunk * 2, (size_t)16 << 30);
        }
        return nullptr;
    }
    void release(void* p, size_t bytes) {
#ifdef DRE_SIMT_EMU
        (void)bytes;
        free(p);
        return;
#endif
        for (Chunk& ch : chunks) {
            char* q = (char*)p;
            if (q < ch.base || q >= ch.base + ch.size) continue;
            size_t off = (size_t)(q - ch.base), len = bytes;
            auto nx = ch.free.lower_bound(off);
            if (nx != ch.free.end() && off + len == nx->first) { len += nx->second; nx = ch.free.erase(nx); }
            if (nx != ch.free.begin()) {
                auto pv = std::prev(nx);
                if (pv->first + pv->second == off) { off = pv->first; len += pv->second; ch.free.erase(pv); }
            }
            ch.free[off] = len;
            return;
        }
    }
    void destroy() {
        for (Chunk& ch : chunks) cudaFree(ch.base);
        chunks.clear();
        next_chunk = (size_t)2 << 30;
    }
};

// Growable device workspace, sub-allocated from the context's arena (all users run on the main stream, so a
// released block may be handed out again immediately).  Growing through cudaFree + cudaMalloc cost up to 650 ms
// for the GB-sized Gram-Schmidt workspaces and showed up as outlier time steps.
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;       // elements
    size_t bytes = 0;     // size of the arena block
    Arena* arena = nullptr;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) { arena->release(p, bytes); p = nullptr; cap = 0; bytes = 0; }
#ifdef DRE_SIMT_EMU
        const size_t want = n;   // no slack under the emulator: overruns must be visible
#else
        const size_t want = n + n / 2 + 64;
#endif
        size_t got = 0;
        void* q = arena->alloc(want * sizeof(T), &got);
        if (!q) return cudaErrorMemoryAllocation;
        p = (T*)q;
        bytes = got;
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p && arena) arena->release(p, bytes); p = nullptr; cap = 0; bytes = 0; }
    void forget() { p = nullptr; cap = 0; bytes = 0; }   // the arena itself was destroyed
};

struct Panel {
    double* d = nullptr;
    int cols = 0;
    int64_t ld = 0;
    bool alive = false;
    size_t cap = 0;   // bytes of the underlying block (>= n * ld * 8)
    bool owned = true;   // false: memory of another context (dre_mat_wrap), never released here
};

}  // namespace

struct dre_symbolic {
    Symbolic sym;
};

namespace {
// state of a rank-revealing block Gram-Schmidt run (compress!, rrqr)
struct RRState {
    double* Q = nullptr;   // n x qcap row-major
    int64_t ldq = 0;
    int qcap = 0;
    int rho = 0;
    double* RT = nullptr;  // ktot x ldrt row-major: coefficients of every input column in the basis
    int64_t ldrt = 0;
    double scale2 = 0.0;   // largest squared column norm seen so far
    double drop_rel = 3e-15, drop_abs = 0.0;
    int rounds = 0;
    int skipped = 0;   // sub-panels skipped because their remainder was below the drop threshold
};
}  // namespace

struct dre_context {
    int device = 0;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    std::string err;
#ifdef DRE_SIMT_EMU
    cusolverDnHandle_t cusolver = nullptr;
#endif

    // pencil
    bool has_pencil = false;
    bool dense_only = false;   // dre_set_dense_only: a compression lane without a sparse solver
    Symbolic sym;
    DevSymbolic dS{};
    std::vector<void*> owned;  // device arrays of the symbolic structure
    int32_t* d_iperm = nullptr;
    int32_t* d_csr_ptr = nullptr;
    int32_t* d_csr_col = nullptr;
    double* d_csr_a = nullptr;
    double* d_csr_e = nullptr;
    int32_t* d_level_sn = nullptr;
    // per-level work lists (schedule.h)
    std::vector<LevelWork> levels;
    int32_t* d_ea_parents = nullptr;
    int2* d_l21_items = nullptr;
    int4* d_schur_items = nullptr;
    // Row-split sweeps (sparse_kernels.cu "sweep v2"): opt-in with DRE_SWEEP2=1 until measured on the GPU.
    // The factorization then leaves M21 = L21 Linv in the panels, so the flag is fixed per context.
    bool sweep2 = false;
    int2* d_fwd2_items = nullptr;
    int2* d_bwd2_items = nullptr;
    DBuf<unsigned char> Ybuf;
    int64_t linv_elems = 0, upd_elems = 0;

    // factor storage (sized for complex, reused for real)
    // Factor slots: while the sweeps of ADI step i read slot `cur` on the main stream, the numeric
    // factorizations of the NEXT shifts (dre_prefactor) run on side streams into the other slots -- one
    // factorization is a latency-bound chain of small launches that leaves most SMs idle; several of them and
    // the sweeps / Gram / compression kernels of the main stream fill the machine together.
    struct FactorSlot {
        void* L = nullptr;
        void* Linv = nullptr;
        void* dvec = nullptr;
        void* U = nullptr;           // update matrices (scratch of the factorization)
        bool valid = false;          // key below describes the factorization held (or in flight)
        double a = 0, re = 0, im = 0;
        int tw = 0;                  // 1 real, 2 complex
        cudaEvent_t ready = nullptr; // recorded behind the factorization
        cudaEvent_t released = nullptr;  // recorded on the main stream behind the last sweeps that read the slot
        bool has_reader = false;
        bool pending = false;        // queued by dre_prefactor and not yet adopted by a solve
        uint64_t queued_at = 0;      // value of solve_seq when dre_prefactor queued it (stale-slot reclamation)
        cudaStream_t st = nullptr;   // side stream of this slot (factorizations of different slots overlap)
        // The launch sequence of a numeric factorization only depends on the pencil's symbolic structure: it is
        // captured once per slot and scalar type ([0] real, [1] complex) as a CUDA graph and replayed for every shift;
        // the two shift-dependent scalars travel through the device parameter block d_prm (k_set_prm).
#ifndef DRE_SIMT_EMU
        cudaGraphExec_t fgraph[2] = {nullptr, nullptr};
#endif
        double* d_prm = nullptr;
    };
    static constexpr int NSLOT = 4;  // the slot in use + up to three prefactorizations in flight
    FactorSlot slot[NSLOT];
    int64_t factor_graph_nodes = 0;  // kernels in one captured factorization
    int cur = 0;
    uint64_t solve_seq = 0;          // block solves so far (acquire_factor calls)
    DBuf<unsigned char> tbuf;
    DBuf<unsigned char> Wbuf;
    DBuf<double> btw, sol;
    int32_t* d_errflag = nullptr;

    // operator F = a A + e E + inv(alpha) U Vt'
    double op_a = 1.0, op_e = 0.0, op_alpha = 1.0;
    dre_view op_U{-1, 0, 0}, op_Vt{-1, 0, 0};

    dre_view ortho_hint{-1, 0, 0};   // dre_hint_orthonormal: consumed by the next dre_ldlt_compress / dre_compress_add
    struct CompressJob {             // dre_compress_begin .. dre_compress_finish
        bool active = false;
        int kcap = 0, ktot = 0;
        double tol_factor = 100.0;
        RRState s;
        std::vector<double> signs;
        struct Dense { int row0 = 0, k = 0; std::vector<double> C; };
        std::vector<Dense> dense;
    } cjob;
    // asynchronous residual norm (dre_ldlt_norm_begin / _end): own stream, events, workspaces and pinned slot
    cudaStream_t norm_st = nullptr;
    cudaEvent_t norm_in = nullptr, norm_done = nullptr;
    DBuf<double> norm_partial, norm_g, norm_small;
    double* h_norm = nullptr;
    size_t h_norm_cap = 0;
    bool norm_pending = false;
    double norm_alpha = 0.0;
    int norm_k = 0;
    // panels
    std::vector<Panel> panels;
    Arena arena;

    // dense workspaces
    DBuf<double> gram_partial, gbuf, gbuf2, cbuf, wsel, wsel2, small, stage, qws, pws, qtmp, rt, rt2, tmp_panel, evals, cnorm;
    // look-ahead stage of the rank-revealing Gram-Schmidt (rr_process_chunks)
    cudaStream_t look_st = nullptr;
    cudaEvent_t look_done[2] = {nullptr, nullptr}, look_basis = nullptr;
    double* h_look = nullptr;
    DBuf<double> look_partial, look_cbuf, look_cnorm, cscale;
    DBuf<double> syevd_work;            // eigensolver workspace (d, e, tau, reflectors, Q)
    DBuf<unsigned char> eig_rots;       // plane rotations of the QL iteration + rank table
    DBuf<int32_t> ibuf;
    double* h_pinned = nullptr;
    size_t h_pinned_cap = 0;

    // stats
    dre_stats stats{};
    int64_t stats_stale_prefactors = 0;   // queued factorizations that were never adopted (reclaimed by pick_slot)
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tev0 = nullptr, tev1 = nullptr;
};

namespace {

template <class F>
void for_each_workspace(dre_context* c, F f) {
    f(c->tbuf); f(c->Wbuf); f(c->Ybuf); f(c->norm_partial); f(c->norm_g); f(c->norm_small);
    f(c->btw); f(c->sol); f(c->gram_partial); f(c->gbuf); f(c->gbuf2); f(c->cbuf);
    f(c->wsel); f(c->wsel2); f(c->small); f(c->stage); f(c->qws); f(c->pws); f(c->qtmp); f(c->rt); f(c->rt2);
    f(c->tmp_panel); f(c->evals); f(c->cnorm); f(c->syevd_work); f(c->eig_rots); f(c->ibuf);
    f(c->look_partial); f(c->look_cbuf); f(c->look_cnorm); f(c->cscale);
}

int fail(dre_context* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    g_last_error = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, DRE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

template <class T>
int upload_vec(dre_context* c, const std::vector<T>& h, T** out) {
    *out = nullptr;
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CU(cudaMalloc((void**)out, bytes));
    c->owned.push_back(*out);
    if (!h.empty()) CU(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return DRE_OK;
}

int ensure_pinned(dre_context* c, size_t doubles) {
    if (doubles <= c->h_pinned_cap) return DRE_OK;
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr;
    c->h_pinned_cap = 0;
    size_t want = doubles + doubles / 4 + 1024;
    CU(cudaMallocHost((void**)&c->h_pinned, want * sizeof(double)));
    c->h_pinned_cap = want;
    return DRE_OK;
}

int check_view(dre_context* c, const dre_view& v, const char* what, bool allow_empty = false) {
    if (v.ncols == 0 && allow_empty) return DRE_OK;
    if (v.id < 0 || v.id >= (int)c->panels.size() || !c->panels[v.id].alive)
        return fail(c, DRE_ERR_ARG, std::string(what) + ": invalid panel id");
    const Panel& p = c->panels[v.id];
    if (v.col0 < 0 || v.ncols < 0 || v.col0 + v.ncols > p.cols)
        return fail(c, DRE_ERR_ARG, std::string(what) + ": column range outside the panel");
    return DRE_OK;
}

inline double* vptr(dre_context* c, const dre_view& v) { return c->panels[v.id].d + v.col0; }
inline int64_t vld(dre_context* c, const dre_view& v) { return c->panels[v.id].ld; }

bool views_overlap(const dre_view& a, const dre_view& b) {
    if (a.ncols == 0 || b.ncols == 0 || a.id != b.id) return false;
    return a.col0 < b.col0 + b.ncols && b.col0 < a.col0 + a.ncols;
}

int check_errflag(dre_context* c) {
    int32_t flag = 0;
    CU(cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (flag != 0) {
        CU(cudaMemsetAsync(c->d_errflag, 0, sizeof(int32_t), c->st));
        if (flag == 1) {
            // the flag does not say which factorization broke down (it may be one queued on a side stream): none of
            // the held factors may be adopted again
            CU(cudaDeviceSynchronize());
            for (int i = 0; i < dre_context::NSLOT; ++i) c->slot[i].valid = c->slot[i].pending = false;
        }
        return fail(c, DRE_ERR_NUMERIC,
                    flag == 1 ? "numeric factorization broke down (zero or non-finite pivot)"
                              : "Sherman-Morrison-Woodbury core is singular");
    }
    return DRE_OK;
}

// host wall-clock trace of the C-ABI internals (DRE_TRACE=1): where does the launching thread block?
static const bool g_trace = getenv("DRE_TRACE") != nullptr;
// DRE_RR_STATS=1: totals of the rank-revealing Gram-Schmidt rounds, printed when a context dies
static const bool g_rr_stats = getenv("DRE_RR_STATS") != nullptr;
// fat Gram products of compress!'s look-ahead stream in that many waves of shorter-lived CTAs (see gram_dev_on);
// default 4 since run r02u (with stream priorities: 775 -> 748 ms per step; 8 and 16 waves add nothing)
static const int g_look_waves = getenv("DRE_LOOK_WAVES") ? std::max(0, atoi(getenv("DRE_LOOK_WAVES"))) : 4;
// DRE_RR_CHUNKPROJ=1: project the WHOLE chunk against the directions added since its look-ahead snapshot before its
// sub-panels are visited (default: every sub-panel does so itself, round by round).  Measured on the B200
// (profiles/r02_results.md, steady state, 4 compress! calls): selection 57.7 -> 34.4 ms but 33.6 ms for the extra
// chunk passes and one more sync per chunk: 104.4 vs 98.7 ms in total -- no gain, so it stays opt-in.
static const bool g_rr_legacy = getenv("DRE_RR_CHUNKPROJ") == nullptr || atoi(getenv("DRE_RR_CHUNKPROJ")) == 0;
// DRE_TIMELINE=<file>: device-side timeline of the streams (events recorded around the launch groups, resolved against
// a base event when the context dies; no synchronisation while the job runs).  Lines: "<ms begin> <ms end> <lane> <name>",
// lane 0 = main stream, 1..4 = factor slots' side streams, 5 = look-ahead / norm streams.
#ifndef DRE_SIMT_EMU
static const char* g_timeline = getenv("DRE_TIMELINE");
struct TlSpan { const char* name; int lane; cudaEvent_t e0, e1; };
static std::vector<TlSpan> g_tl;
static cudaEvent_t g_tl_base = nullptr;
struct TlScope {
    cudaStream_t st; size_t idx = (size_t)-1;
    TlScope(const char* name, int lane, cudaStream_t s) : st(s) {
        if (!g_timeline || g_tl.size() > 400000) return;
        if (!g_tl_base) { cudaEventCreate(&g_tl_base); cudaEventRecord(g_tl_base, s); }
        TlSpan sp{name, lane, nullptr, nullptr};
        cudaEventCreate(&sp.e0); cudaEventCreate(&sp.e1);
        cudaEventRecord(sp.e0, s);
        idx = g_tl.size();
        g_tl.push_back(sp);
    }
    ~TlScope() { if (idx != (size_t)-1) cudaEventRecord(g_tl[idx].e1, st); }
};
static void tl_dump() {
    if (!g_timeline || g_tl.empty()) return;
    cudaDeviceSynchronize();
    FILE* f = fopen(g_timeline, "w");
    if (!f) return;
    for (auto& sp : g_tl) {
        float a = 0, b = 0;
        if (cudaEventElapsedTime(&a, g_tl_base, sp.e0) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&b, g_tl_base, sp.e1) != cudaSuccess) continue;
        fprintf(f, "%.4f %.4f %d %s\n", a, b, sp.lane, sp.name);
        cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1);
    }
    fclose(f);
    g_tl.clear();
}
#else
struct TlScope { TlScope(const char*, int, cudaStream_t) {} };
static void tl_dump() {}
#endif

struct RRTotals {
    long calls = 0, blocks = 0, rounds = 0, productive = 0, skipped = 0, syncs = 0, rest_projections = 0,
         coef_only_passes = 0;
    double ms_rounds = 0, ms_core = 0, ms_eig = 0, ms_final = 0;   // host wall clock between the syncs compress! has anyway
    double ms_wait_stage1 = 0, ms_chunk_proj = 0, ms_select = 0, ms_extend = 0;   // inside the Gram-Schmidt phase
    double ms_entry_sync = 0, ms_setup = 0;   // waiting for the work queued before the call; argument checks + uploads
} g_rr;
static inline double wall_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct HostTrace {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    cudaStream_t sync_stream = nullptr;   // when set, the scope ends with a stream sync (GPU time attribution)
    explicit HostTrace(const char* n, cudaStream_t s = nullptr) : name(n), sync_stream(s) {
        if (g_trace) {
            if (sync_stream) cudaStreamSynchronize(sync_stream);
            t0 = std::chrono::steady_clock::now();
        }
    }
    ~HostTrace() {
        if (g_trace) {
            if (sync_stream) cudaStreamSynchronize(sync_stream);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms > 0.1) fprintf(stderr, "[dre trace] %-24s %9.3f ms\n", name, ms);
        }
    }
};

// NVTX ranges named after the reference's @timeit_debug sections (src/lyapunov/adi.jl:157,196, src/LDLt.jl:77,204,211,
// 214, src/blocklinear/sherman-morrison-woodbury.jl:10,19,36), so that a profiler timeline of this library lines up
// with the reference's TimerOutputs tree.  Zero cost unless a profiler is attached.
struct Range {
#ifndef DRE_SIMT_EMU
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
#else
    explicit Range(const char*) {}
#endif
};

struct Timer {
    dre_context* c;
    double* acc;
    Timer(dre_context* c_, double* acc_) : c(c_), acc(acc_) {
        if (c->timing) cudaEventRecord(c->ev0, c->st);
    }
    ~Timer() {
        if (c->timing) {
            cudaEventRecord(c->ev1, c->st);
            cudaEventSynchronize(c->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            *acc += ms;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// gram helper: out (a x b, row-major ld) = X' diag(w) Y   (device result)
// ---------------------------------------------------------------------------------------------
int gram_dev(dre_context* c, const double* X, int64_t ldx, int a, const double* Y, int64_t ldy, int b, int64_t n,
             const double* roww, double* out1, int64_t ld1, double* out2, int64_t ld2) {
    if (a <= 0 || b <= 0) return DRE_OK;
    GramPlan plan = gram_plan(n, a, b, c->sm_count);
    CU(c->gram_partial.ensure(plan.partial_elems));
    Timer t(c, &c->stats.ms_gram);
    launch_gram(X, ldx, a, Y, ldy, b, n, roww, c->gram_partial.p, plan, out1, ld1, out2, ld2, c->st,
                &c->stats.kernel_launches);
    CU(cudaGetLastError());
    c->stats.grams++;
    c->stats.flops_gram += 2.0 * (double)n * a * b;
    c->stats.bytes_gram += 8.0 * (double)n * ((X == Y && a == b) ? a : (a + b));
    return DRE_OK;
}

// the same on an explicit stream with its own (pre-sized) partial-sum workspace: the look-ahead stage of compress!
int gram_dev_on(dre_context* c, cudaStream_t st, DBuf<double>& partial, const double* X, int64_t ldx, int a,
                const double* Y, int64_t ldy, int b, int64_t n, const double* roww, double* out1, int64_t ld1,
                double* out2, int64_t ld2) {
    if (a <= 0 || b <= 0) return DRE_OK;
    // the look-ahead stream's fat Gram products run next to the latency-bound selection rounds of the main stream:
    // DRE_LOOK_WAVES > 1 cuts them into that many waves of shorter-lived CTAs, so that (with DRE_PRIO=1) a freed SM
    // goes to the waiting small kernel instead of staying with a one-wave kernel for its whole ~1 ms
    GramPlan plan = gram_plan(n, a, b, c->sm_count, st != c->st ? g_look_waves : 0);
    if (plan.partial_elems > partial.cap) {
        if (st != c->st) CU(cudaStreamSynchronize(st));   // (never on the sized-up-front path)
        CU(partial.ensure(plan.partial_elems));
    }
    launch_gram(X, ldx, a, Y, ldy, b, n, roww, partial.p, plan, out1, ld1, out2, ld2, st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    c->stats.grams++;
    c->stats.flops_gram += 2.0 * (double)n * a * b;
    c->stats.bytes_gram += 8.0 * (double)n * (a + b);
    return DRE_OK;
}

int tall_gemm_on(dre_context* c, cudaStream_t st, double alpha, const double* X, int64_t ldx, int a, const double* W,
                 int64_t ldw, int w_trans, double beta, double* Y, int64_t ldy, int b, int64_t n) {
    launch_tall_gemm(alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n, st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    c->stats.tallgemms++;
    c->stats.flops_tallgemm += 2.0 * (double)n * a * b;
    c->stats.bytes_tallgemm += 8.0 * (double)n * (a + (beta == 0.0 ? 1.0 : 2.0) * b);
    return DRE_OK;
}

int tall_gemm(dre_context* c, double alpha, const double* X, int64_t ldx, int a, const double* W, int64_t ldw,
              int w_trans, double beta, double* Y, int64_t ldy, int b, int64_t n) {
    Timer t(c, &c->stats.ms_tallgemm);
    launch_tall_gemm(alpha, X, ldx, a, W, ldw, w_trans, beta, Y, ldy, b, n, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    c->stats.tallgemms++;
    c->stats.flops_tallgemm += 2.0 * (double)n * a * b;
    c->stats.bytes_tallgemm += 8.0 * (double)n * (a + (beta == 0.0 ? 1.0 : 2.0) * b);
    return DRE_OK;
}

// ---------------------------------------------------------------------------------------------
// numeric factorization + solve
// ---------------------------------------------------------------------------------------------
inline DevSchedule dev_schedule(const dre_context* c) {
    return DevSchedule{c->levels.data(), (int)c->levels.size(), c->d_level_sn, c->d_ea_parents, c->d_l21_items,
                       c->d_schur_items, c->d_fwd2_items, c->d_bwd2_items};
}

#ifndef DRE_SIMT_EMU
static const bool g_graphs = !(getenv("DRE_GRAPHS") && atoi(getenv("DRE_GRAPHS")) == 0);
#else
static const bool g_graphs = false;
#endif

// Stream priorities (default on since run r02u: +3 % alone, +4 % with DRE_LOOK_WAVES; DRE_PRIO=0 turns them off): the
// ADI chain and the selection rounds of compress! (main stream) run at the highest priority, the prefactor side
// streams and the look-ahead stage one step below, a compression lane (dre_set_dense_only) at the lowest -- the block
// scheduler then hands a freed SM to the waiting CTA of the latency-bound chain instead of the next CTA of a fat
// Gram kernel that was launched earlier.  level: 0 highest, 1 middle, 2 lowest.
static const bool g_prio = !(getenv("DRE_PRIO") && atoi(getenv("DRE_PRIO")) == 0);
static cudaError_t make_stream(cudaStream_t* s, int level) {
    if (!g_prio) return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
    int least = 0, greatest = 0;   // numerically: greatest priority <= least priority
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
    if (e != cudaSuccess) return e;
    int pr = greatest + level * std::max(1, (least - greatest) / 2);
    if (pr > least) pr = least;
    return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, pr);
}

inline double re_of(double v) { return v; }
inline double im_of(double) { return 0.0; }
inline double re_of(cplx v) { return v.x; }
inline double im_of(cplx v) { return v.y; }

template <class T>
int factor(dre_context* c, dre_context::FactorSlot& fs, cudaStream_t st, T emu) {
    const Symbolic& S = c->sym;
    Timer t(c, &c->stats.ms_factor);
    TlScope tl(sizeof(T) == 8 ? "factor" : "factor(cplx)", st == c->st ? 0 : 1 + (int)(&fs - c->slot), st);
    T* L = (T*)fs.L;
    T* Linv = (T*)fs.Linv;
    T* dvec = (T*)fs.dvec;
    T* U = (T*)fs.U;
#ifndef DRE_SIMT_EMU
    if (g_graphs) {
        const int gi = sizeof(T) == sizeof(double) ? 0 : 1;
        if (!fs.d_prm) CU(cudaMalloc((void**)&fs.d_prm, 4 * sizeof(double)));
        if (!fs.fgraph[gi]) {
            // capture on a stream of its own (thread-local mode: the compression lane's thread may allocate meanwhile)
            cudaStream_t cs = nullptr;
            CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
            cudaGraph_t g = nullptr;
            int64_t nl = 0;
            CU(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
            cudaMemsetAsync(L, 0, (size_t)S.nnz_L * sizeof(T), cs);
            enqueue_factor<T>(c->dS, dev_schedule(c), L, Linv, dvec, U, 0.0, zero<T>(), c->d_errflag, cs, &nl, c->sweep2,
                              fs.d_prm);
            cudaError_t e = cudaStreamEndCapture(cs, &g);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaGraphInstantiate(&fs.fgraph[gi], g, 0);
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            if (e != cudaSuccess) {
                fs.fgraph[gi] = nullptr;
                return fail(c, DRE_ERR_CUDA, std::string("factor graph capture: ") + cudaGetErrorString(e));
            }
            c->factor_graph_nodes = nl;
        }
        launch_set_prm(fs.d_prm, c->op_a, re_of(emu), im_of(emu), st, &c->stats.kernel_launches);
        CU(cudaGraphLaunch(fs.fgraph[gi], st));
        c->stats.kernel_launches += c->factor_graph_nodes;   // kernels inside the graph (memsets not counted)
        c->stats.factorizations++;
        c->stats.flops_factor += S.flops * (sizeof(T) == sizeof(double) ? 1.0 : 4.0);
        return DRE_OK;
    }
#endif
    CU(cudaMemsetAsync(L, 0, (size_t)S.nnz_L * sizeof(T), st));
    enqueue_factor<T>(c->dS, dev_schedule(c), L, Linv, dvec, U, c->op_a, emu, c->d_errflag, st,
                      &c->stats.kernel_launches, c->sweep2);
    CU(cudaGetLastError());
    c->stats.factorizations++;
    c->stats.flops_factor += S.flops * (sizeof(T) == sizeof(double) ? 1.0 : 4.0);
    return DRE_OK;
}

template <class T>
int solve_sweeps(dre_context* c, T* W, int64_t ldw, int nrhs, const RhsSource& src) {
    const Symbolic& S = c->sym;
    Timer t(c, &c->stats.ms_solve);
    TlScope tl("sweeps", 0, c->st);
    {
        HostTrace tr("tbuf.ensure");
        CU(c->tbuf.ensure((size_t)std::max<int64_t>(S.rhs_total, 1) * ldw * sizeof(T)));
    }
    dre_context::FactorSlot& fs = c->slot[c->cur];
    const T* L = (const T*)fs.L;
    const T* Linv = (const T*)fs.Linv;
    const T* dvec = (const T*)fs.dvec;
    T* tb = (T*)c->tbuf.p;
    if (c->sweep2) {
        CU(c->Ybuf.ensure((size_t)S.n * ldw * sizeof(T)));
        enqueue_sweeps2<T>(c->dS, dev_schedule(c), L, Linv, dvec, W, (T*)c->Ybuf.p, ldw, nrhs, tb, src, c->st,
                           &c->stats.kernel_launches);
    } else {
        enqueue_sweeps<T>(c->dS, dev_schedule(c), L, Linv, dvec, W, ldw, nrhs, tb, src, c->st,
                          &c->stats.kernel_launches);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(fs.released, c->st));
    fs.has_reader = true;
    c->stats.solves++;
    {   // 2*nnz(L)*w (factor read once per sweep) + 4*n*r_tot*w (RHS read+write per sweep)
        const double w = (double)sizeof(T);
        c->stats.bytes_solve += 2.0 * (double)S.nnz_L * w + 4.0 * (double)S.n * nrhs * w;
        c->stats.flops_solve += 4.0 * (double)S.nnz_L * nrhs * (sizeof(T) == sizeof(double) ? 1.0 : 4.0);
    }
    return DRE_OK;
}

inline bool slot_matches(const dre_context* c, const dre_context::FactorSlot& fs, double mu_re, double mu_im, int tw) {
    return fs.valid && fs.a == c->op_a && fs.re == c->op_e + mu_re && fs.im == mu_im && fs.tw == tw;
}

// slot to (re)fill: an empty one, else a consumed one, else (only when `allow_pending`) a queued one
inline int pick_slot(dre_context* c, bool allow_pending) {
    // A queued factorization is adopted within NSLOT - 1 solves (the host queues the next shifts of the buffer in
    // order).  One that is still pending after more solves than that was queued for an ADI run that has ended
    // early (convergence) or for a buffer that was refilled: it is stale and its slot is free again -- without
    // this the pipeline depth would drop to one after the first converged solve.
    for (int i = 0; i < dre_context::NSLOT; ++i) {
        dre_context::FactorSlot& fs = c->slot[i];
        if (fs.pending && c->solve_seq > fs.queued_at + (uint64_t)dre_context::NSLOT) {
            fs.pending = false;
            c->stats_stale_prefactors++;
        }
    }
    for (int i = 0; i < dre_context::NSLOT; ++i)
        if (i != c->cur && !c->slot[i].valid) return i;
    for (int i = 0; i < dre_context::NSLOT; ++i)
        if (i != c->cur && !c->slot[i].pending) return i;
    if (!allow_pending) return -1;
    return (c->cur + 1) % dre_context::NSLOT;
}

// Make c->cur a slot holding the factorization for (op, mu) as seen from stream `st` (the main stream):
// a slot filled by dre_prefactor is adopted behind its `ready` event; otherwise a free slot is factored on `st`.
template <class T>
int acquire_factor(dre_context* c, double mu_re, double mu_im, T emu, cudaStream_t st) {
    const int tw = (int)(sizeof(T) / sizeof(double));
    c->solve_seq++;
    for (int i = 0; i < dre_context::NSLOT; ++i) {
        dre_context::FactorSlot& fs = c->slot[i];
        if (slot_matches(c, fs, mu_re, mu_im, tw)) {
            CU(cudaStreamWaitEvent(st, fs.ready, 0));
            if (fs.pending) c->stats.prefactor_hits++;
            fs.pending = false;
            c->cur = i;
            return DRE_OK;
        }
    }
    const int i = pick_slot(c, true);
    dre_context::FactorSlot& fs = c->slot[i];
    if (fs.valid) CU(cudaStreamWaitEvent(st, fs.ready, 0));   // an unused prefactorization may still be in flight
    fs.valid = fs.pending = false;
    HostTrace tr(sizeof(T) == 8 ? "factor<double> launch" : "factor<cplx> launch");
    int rc = factor<T>(c, fs, st, emu);
    if (rc) return rc;
    CU(cudaEventRecord(fs.ready, st));
    fs.valid = true;
    fs.a = c->op_a; fs.re = c->op_e + mu_re; fs.im = mu_im; fs.tw = tw;
    fs.has_reader = false;
    c->cur = i;
    return DRE_OK;
}

template <class T>
int prefactor_t(dre_context* c, double mu_re, double mu_im, T emu) {
    const int tw = (int)(sizeof(T) / sizeof(double));
    for (int k = 0; k < dre_context::NSLOT; ++k)
        if (slot_matches(c, c->slot[k], mu_re, mu_im, tw)) return DRE_OK;   // already held / in flight
    const int i = pick_slot(c, false);
    if (i < 0) return DRE_OK;                                               // every spare slot is queued
    dre_context::FactorSlot& fs = c->slot[i];
    if (fs.valid) CU(cudaStreamWaitEvent(fs.st, fs.ready, 0));
    if (fs.has_reader) CU(cudaStreamWaitEvent(fs.st, fs.released, 0));     // sweeps that still read the slot
    fs.valid = false;
    int rc = factor<T>(c, fs, fs.st, emu);
    if (rc) return rc;
    CU(cudaEventRecord(fs.ready, fs.st));
    fs.valid = true;
    fs.pending = true;
    fs.queued_at = c->solve_seq;
    fs.a = c->op_a; fs.re = c->op_e + mu_re; fs.im = mu_im; fs.tw = tw;
    fs.has_reader = false;
    c->stats.prefactors++;
    return DRE_OK;
}

inline void make_emu(double e, double mr, double mi, double& out) { out = e + mr; (void)mi; }
inline void make_emu(double e, double mr, double mi, cplx& out) { out = mk(e + mr, mi); }

// mode: 0 real, 1 complex raw, 2 complex ADI pair
template <class T>
int shifted_solve_t(dre_context* c, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2, int mode) {
    const int64_t n = c->sym.n;
    const int r = R.ncols, m = c->op_U.ncols;
    const int nrhs = r + m;
    const int64_t ldw = (nrhs + 3) & ~3;
    {
        HostTrace tr("Wbuf.ensure");
        CU(c->Wbuf.ensure((size_t)n * ldw * sizeof(T)));
    }
    T* W = (T*)c->Wbuf.p;
    const RhsSource rhs_src{vptr(c, R), vld(c, R), r, m ? vptr(c, c->op_Vt) : nullptr, m ? vld(c, c->op_Vt) : 0};
    T emu;
    make_emu(c->op_e, mu_re, mu_im, emu);
    const int tw = (int)(sizeof(T) / sizeof(double));  // 1 or 2 doubles per element
    int rc = DRE_OK;
    Range r_solve(sizeof(T) == 8 ? "solve (real)" : "solve (complex)");
    Range r_smw(m > 0 ? "Sherman-Morrison-Woodbury" : "Backslash");
    {
        Range r_sparse("solve (sparse)");
        rc = acquire_factor<T>(c, mu_re, mu_im, emu, c->st);
        if (rc) return rc;
        HostTrace tr(sizeof(T) == 8 ? "sweeps<double> launch" : "sweeps<cplx> launch");
        rc = solve_sweeps<T>(c, W, ldw, nrhs, rhs_src);
    }
    if (rc) return rc;
    T* Sol = nullptr;
    if (m > 0) {
        Range r_dense("solve (dense)");
        CU(c->btw.ensure((size_t)m * nrhs * tw));
        CU(c->sol.ensure((size_t)m * r * tw));
        rc = gram_dev(c, vptr(c, c->op_U), vld(c, c->op_U), m, (const double*)W, ldw * tw, nrhs * tw, n, nullptr,
                      c->btw.p, (int64_t)nrhs * tw, nullptr, 0);
        if (rc) return rc;
        Sol = (T*)c->sol.p;
        launch_smw_core<T>((const T*)c->btw.p, nrhs, m, r, c->op_alpha, Sol, c->d_errflag, c->st,
                           &c->stats.kernel_launches);
    }
    const double d = (mu_im != 0.0) ? mu_re / mu_im : 0.0;
    launch_smw_apply<T>(W, ldw, r, m, Sol, mode, d, vptr(c, V1), vld(c, V1), mode ? vptr(c, V2) : nullptr,
                        mode ? vld(c, V2) : 0, n, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    return DRE_OK;
}

int shifted_solve(dre_context* c, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2, bool adi_pair) {
    if (!c->has_pencil || c->dense_only) return fail(c, DRE_ERR_STATE, "no pencil set");
    int rc;
    if ((rc = check_view(c, R, "R"))) return rc;
    if ((rc = check_view(c, V1, "V1"))) return rc;
    if (V1.ncols != R.ncols) return fail(c, DRE_ERR_ARG, "V1 must have as many columns as R");
    if (views_overlap(R, V1)) return fail(c, DRE_ERR_ARG, "V1 must not alias R");
    if (c->op_U.ncols > 32) return fail(c, DRE_ERR_ARG, "low-rank update with more than 32 columns is not supported");
    if (mu_im != 0.0) {
        if ((rc = check_view(c, V2, "V2"))) return rc;
        if (V2.ncols != R.ncols) return fail(c, DRE_ERR_ARG, "V2 must have as many columns as R");
        if (views_overlap(R, V2) || views_overlap(V1, V2)) return fail(c, DRE_ERR_ARG, "V2 must not alias R or V1");
        return shifted_solve_t<cplx>(c, mu_re, mu_im, R, V1, V2, adi_pair ? 2 : 1);
    }
    return shifted_solve_t<double>(c, mu_re, 0.0, R, V1, V2, 0);
}

// ---------------------------------------------------------------------------------------------
// rank-revealing block Gram-Schmidt (shared by compress! and rrqr)
// ---------------------------------------------------------------------------------------------

struct RRChunk {
    const double* src;        // n x cols columns of an input term (row-major, leading dimension lds)
    int64_t lds;
    int cols;                 // <= 256
    const double* colscale;   // device, per column (nullptr: none)
    int rt_row;               // first row of this chunk's coefficients in RT
};

// Stage 1 of a chunk (stream `st`: the look-ahead stream or the main one): scaled copy into work buffer `slot`,
// original column norms, ONE projection pass against the basis columns [0, rho_snap) with the coefficients
// accumulated into the chunk's RT rows, remainder norms; both norm vectors go to the pinned slot.
int rr_stage1(dre_context* c, RRState& s, const RRChunk& ch, int slot, int rho_snap, cudaStream_t st, bool side) {
    const int64_t n = c->sym.n;
    constexpr int PBIG = 256, NBLK = 296;
    double* Pb = c->pws.p + (size_t)slot * n * PBIG;
    DBuf<double>& partial = side ? c->look_partial : c->gram_partial;
    double* cn = c->look_cnorm.p + (size_t)slot * 2 * PBIG;
    launch_copy_scale(Pb, PBIG, ch.src, ch.lds, n, ch.cols, ch.colscale, st, &c->stats.kernel_launches);
    launch_colnorm2(Pb, PBIG, n, ch.cols, partial.p, NBLK, cn, st, &c->stats.kernel_launches);
    if (rho_snap > 0) {
        double* cb = side ? c->look_cbuf.p : c->cbuf.p;
        int rc = gram_dev_on(c, st, partial, Pb, PBIG, ch.cols, s.Q, s.ldq, rho_snap, n, nullptr, cb, rho_snap,
                             s.RT + (int64_t)ch.rt_row * s.ldrt, s.ldrt);
        if (rc) return rc;
        rc = tall_gemm_on(c, st, -1.0, s.Q, s.ldq, rho_snap, cb, rho_snap, 1, 1.0, Pb, PBIG, ch.cols, n);
        if (rc) return rc;
        launch_colnorm2(Pb, PBIG, n, ch.cols, partial.p, NBLK, cn + PBIG, st, &c->stats.kernel_launches);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_look + (size_t)slot * 2 * PBIG, cn, 2 * PBIG * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(c->look_done[slot], st));
    return DRE_OK;
}

// Block Gram-Schmidt in two granularities over a list of chunks (<= 256 columns each, in input order).
//   stage 1 (fat, DMMA bound): the chunk is projected against the basis as it stood when the stage was queued;
//   stage 2 (latency bound):   a second pass if the chunk holds large columns, then its 64-column sub-panels are
//                              projected against the directions added since that snapshot, and pivoted-Cholesky
//                              selection / CholQR2 rounds extract the new directions of their remainders.
// LOOK-AHEAD: stage 1 of chunk i+1 is queued on a second stream BEFORE stage 2 of chunk i starts, against the
// basis columns that exist at that moment (they are never rewritten), in the other of two work buffers; the fat
// Gram / tall-GEMM kernels then fill the SMs the small selection kernels leave idle.  The directions chunk i adds
// are simply part of "added since the snapshot" for chunk i+1.  (Off while per-class event timing is on and under
// DRE_TRACE: both want one stream.  DRE_RR_LOOKAHEAD=0 disables it.)
int rr_process_chunks(dre_context* c, RRState& s, const std::vector<RRChunk>& chunks) {
    const int64_t n = c->sym.n;
    constexpr int PB = 64, PBIG = 256, NBLK = 296;
    const int nch = (int)chunks.size();
    if (nch == 0) return DRE_OK;
    static const bool look_env = !(getenv("DRE_RR_LOOKAHEAD") && atoi(getenv("DRE_RR_LOOKAHEAD")) == 0);
    const bool look = look_env && nch > 1 && !c->timing && !g_trace;
    int rc;
    if (!c->look_st) {
        CU(make_stream(&c->look_st, c->dense_only ? 2 : 1));
        for (int i = 0; i < 2; ++i) CU(cudaEventCreateWithFlags(&c->look_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->look_basis, cudaEventDisableTiming));
        CU(cudaMallocHost((void**)&c->h_look, 4 * PBIG * sizeof(double)));
    }
    // every buffer the look-ahead stream touches is sized once, up front (no reallocation while it is in flight)
    CU(c->pws.ensure((size_t)n * PBIG * 2));
    CU(c->look_cnorm.ensure(4 * PBIG));
    CU(c->look_cbuf.ensure((size_t)PBIG * s.qcap));
    {
        size_t pe = (size_t)NBLK * PBIG;
        for (int b = 64; b < s.qcap + 64; b += 64)
            pe = std::max(pe, gram_plan(n, PBIG, std::min(b, s.qcap), c->sm_count, g_look_waves).partial_elems);
        CU(c->look_partial.ensure(pe));
        CU(c->gram_partial.ensure((size_t)NBLK * PBIG));
    }
    CU(c->qtmp.ensure((size_t)n * PB));
    CU(c->gbuf.ensure(PB * PB));
    CU(c->gbuf2.ensure(PB * PB));
    CU(c->wsel.ensure(PB * PB));
    CU(c->wsel2.ensure(PB * PB));
    CU(c->ibuf.ensure(8));
    CU(c->small.ensure(8));
    CU(c->cnorm.ensure(PBIG));
    if ((rc = ensure_pinned(c, PBIG + 16))) return rc;
    double* Qt = c->qtmp.p;
    std::vector<int> snap(nch, 0);
    auto queue_stage1 = [&](int i) -> int {
        snap[i] = s.rho;
        if (look) {
            // behind everything queued on the main stream so far: the basis columns [0, rho) are complete and the
            // rounds that used this work buffer (chunk i - 2) are over
            CU(cudaEventRecord(c->look_basis, c->st));
            CU(cudaStreamWaitEvent(c->look_st, c->look_basis, 0));
            return rr_stage1(c, s, chunks[i], i & 1, snap[i], c->look_st, true);
        }
        CU(c->cbuf.ensure((size_t)PBIG * std::max(s.rho, PB)));
        return rr_stage1(c, s, chunks[i], i & 1, snap[i], c->st, false);
    };
    if (look && (rc = queue_stage1(0))) return rc;
    for (int i = 0; i < nch; ++i) {
        const RRChunk& ch = chunks[i];
        const int pbig = ch.cols;
        if (look) {
            if (i + 1 < nch && (rc = queue_stage1(i + 1))) return rc;
        } else if ((rc = queue_stage1(i))) return rc;
        const double t_s1 = wall_ms();
        CU(cudaEventSynchronize(c->look_done[i & 1]));
        if (look) CU(cudaStreamWaitEvent(c->st, c->look_done[i & 1], 0));
        g_rr.syncs++;
        const double t_s2 = wall_ms();
        g_rr.ms_wait_stage1 += t_s2 - t_s1;
        double* Pbig = c->pws.p + (size_t)(i & 1) * n * PBIG;
        const double* hn = c->h_look + (size_t)(i & 1) * 2 * PBIG;
        const int rho0 = snap[i];
        double block_max2 = 0.0;
        for (int j = 0; j < pbig; ++j) {   // the drop threshold is relative to the largest ORIGINAL column norm met so far
            s.scale2 = std::max(s.scale2, hn[j]);
            block_max2 = std::max(block_max2, hn[j]);
        }
        // remainder norms after the block passes: a sub-panel whose columns are all below the drop threshold
        // cannot contribute a direction (pivoted Cholesky would select nothing) and is skipped without its
        // Gram / selection / synchronisation
        std::vector<double> rem2(pbig, 0.0);
        const bool have_rem = rho0 > 0;
        if (have_rem)
            for (int j = 0; j < pbig; ++j) rem2[j] = hn[PBIG + j];
        // One projection pass leaves a basis component of ~30 eps |p| in a column p.  Columns below 5 % of the
        // global scale therefore stay an order of magnitude under the drop threshold after a single pass, and the
        // second ("twice is enough") pass is only spent on chunks with large columns.  The orthogonality of the
        // basis does not depend on it: candidates are re-orthogonalised below.
        const int npass = (block_max2 <= 0.0025 * s.scale2) ? 1 : 2;
        if (rho0 > 0 && npass == 2) {
            // If the first pass already left every column of the chunk below the drop threshold, the chunk adds no
            // direction and its remainder is discarded as a whole: a second pass would only move the part of that
            // remainder that lies in span(Q) -- at most drop * scale per column, no more than what is being
            // discarded anyway -- into the coefficients, so it is skipped.
            double m2 = 0.0;
            for (int j = 0; j < pbig; ++j) m2 = std::max(m2, rem2[j]);
            const double drop0 = std::max(s.drop_rel * std::sqrt(s.scale2), s.drop_abs);
            if (m2 < drop0 * drop0) {
                g_rr.coef_only_passes++;
            } else {
                CU(c->cbuf.ensure((size_t)PBIG * std::max(s.rho, PB)));
                rc = gram_dev(c, Pbig, PBIG, pbig, s.Q, s.ldq, rho0, n, nullptr, c->cbuf.p, rho0,
                              s.RT + (int64_t)ch.rt_row * s.ldrt, s.ldrt);
                if (rc) return rc;
                rc = tall_gemm(c, -1.0, s.Q, s.ldq, rho0, c->cbuf.p, rho0, 1, 1.0, Pbig, PBIG, pbig, n);
                if (rc) return rc;
                launch_colnorm2(Pbig, PBIG, n, pbig, c->gram_partial.p, NBLK, c->cnorm.p, c->st, &c->stats.kernel_launches);
                CU(cudaMemcpyAsync(c->h_pinned + 16, c->cnorm.p, pbig * sizeof(double), cudaMemcpyDeviceToHost, c->st));
                CU(cudaStreamSynchronize(c->st));
                g_rr.syncs++;
                for (int j = 0; j < pbig; ++j) rem2[j] = c->h_pinned[16 + j];
            }
        }
        g_rr.blocks++;
        // Directions added since the snapshot (by the previous chunk, whose stage 2 ran while this chunk's stage 1
        // was in flight): the WHOLE chunk is projected against them here, in one fat Gram / tall-GEMM pair per pass,
        // and its remainder norms are refreshed; the sub-panels below then only see the directions this chunk adds
        // itself.  (Before, every sub-panel re-projected against everything since the snapshot in every round:
        // 4 x npass x rounds narrow launches per chunk, most of them to find nothing.)
        const int rho_cs = s.rho;
        if (rho_cs > rho0 && !g_rr_legacy) {
            CU(c->cbuf.ensure((size_t)PBIG * std::max(s.rho, PB)));
            for (int pass = 0; pass < npass; ++pass) {
                rc = gram_dev(c, Pbig, PBIG, pbig, s.Q + rho0, s.ldq, rho_cs - rho0, n, nullptr, c->cbuf.p, rho_cs - rho0,
                              s.RT + (int64_t)ch.rt_row * s.ldrt + rho0, s.ldrt);
                if (rc) return rc;
                rc = tall_gemm(c, -1.0, s.Q + rho0, s.ldq, rho_cs - rho0, c->cbuf.p, rho_cs - rho0, 1, 1.0, Pbig, PBIG,
                               pbig, n);
                if (rc) return rc;
            }
            launch_colnorm2(Pbig, PBIG, n, pbig, c->gram_partial.p, NBLK, c->cnorm.p, c->st, &c->stats.kernel_launches);
            CU(cudaMemcpyAsync(c->h_pinned + 16, c->cnorm.p, pbig * sizeof(double), cudaMemcpyDeviceToHost, c->st));
            CU(cudaStreamSynchronize(c->st));
            g_rr.syncs++;
            g_rr.rest_projections++;
            for (int j = 0; j < pbig; ++j) rem2[j] = c->h_pinned[16 + j];
        }
        g_rr.ms_chunk_proj += wall_ms() - t_s2;
        const int rho_base = g_rr_legacy ? rho0 : rho_cs;   // what the sub-panels have been projected against already
        const bool have_rem2 = have_rem || (rho_cs > rho0 && !g_rr_legacy);
        for (int sc = 0; sc < pbig; sc += PB) {
            const int pb = std::min(PB, pbig - sc);
            if (have_rem2 && !g_trace) {
                double m2 = 0.0;
                for (int j = sc; j < sc + pb; ++j) m2 = std::max(m2, rem2[j]);
                const double drop0 = std::max(s.drop_rel * std::sqrt(s.scale2), s.drop_abs);
                // (projections against directions added since the snapshot can only shrink the columns further)
                if (m2 < drop0 * drop0) { s.skipped++; g_rr.skipped++; continue; }
            }
            double* Pw = Pbig + sc;
            double* RTrow = s.RT + (int64_t)(ch.rt_row + sc) * s.ldrt;
            double trace_d0 = 0.0;
            if (g_trace) {   // diagnostic only: largest squared column norm of the sub-panel's remainder so far
                rc = gram_dev(c, Pw, PBIG, pb, Pw, PBIG, pb, n, nullptr, c->gbuf.p, PB, nullptr, 0);
                if (rc) return rc;
                std::vector<double> hg((size_t)PB * PB);
                CU(cudaMemcpyAsync(hg.data(), c->gbuf.p, sizeof(double) * PB * PB, cudaMemcpyDeviceToHost, c->st));
                CU(cudaStreamSynchronize(c->st));
                for (int k = 0; k < pb; ++k) trace_d0 = std::max(trace_d0, hg[(size_t)k * PB + k]);
            }
            for (int round = 0; round < 8; ++round) {
                s.rounds++;
                g_rr.rounds++;
                const double t_r0 = wall_ms();
                // directions added since the snapshot (by earlier chunks / sub-panels / rounds) that this sub-panel
                // has not been projected against in this round
                const int nnew = s.rho - rho_base;
                if (nnew > 0) {
                    CU(c->cbuf.ensure((size_t)PBIG * std::max(s.rho, PB)));
                    // same rule as for the block passes: one pass unless the chunk holds large columns
                    for (int pass = 0; pass < npass; ++pass) {
                        rc = gram_dev(c, Pw, PBIG, pb, s.Q + rho_base, s.ldq, nnew, n, nullptr, c->cbuf.p, nnew,
                                      RTrow + rho_base, s.ldrt);
                        if (rc) return rc;
                        rc = tall_gemm(c, -1.0, s.Q + rho_base, s.ldq, nnew, c->cbuf.p, nnew, 1, 1.0, Pw, PBIG, pb, n);
                        if (rc) return rc;
                    }
                }
                rc = gram_dev(c, Pw, PBIG, pb, Pw, PBIG, pb, n, nullptr, c->gbuf.p, PB, nullptr, 0);
                if (rc) return rc;
                const double drop = std::max(s.drop_rel * std::sqrt(s.scale2), s.drop_abs);
                launch_pivchol(c->gbuf.p, PB, pb, drop * drop, 1e-12, c->wsel.p, c->ibuf.p, c->small.p, c->st,
                               &c->stats.kernel_launches);
                if (s.rho + PB > s.qcap) return fail(c, DRE_ERR_STATE, "rank-revealing QR: basis capacity exceeded");
                // the host needs nsel before it is worth orthonormalising candidates: most sub-panels of a
                // late ADI increment add nothing
                int32_t* hi = (int32_t*)(c->h_pinned + 8);
                CU(cudaMemcpyAsync(hi, c->ibuf.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
                CU(cudaMemcpyAsync(c->h_pinned, c->small.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
                CU(cudaStreamSynchronize(c->st));
                g_rr.syncs++;
                const int nsel = hi[0];
                const double t_r1 = wall_ms();
                g_rr.ms_select += t_r1 - t_r0;
                const double dfirst = c->h_pinned[0];
                const double remaining = c->h_pinned[1];
                if (round == 0) s.scale2 = std::max(s.scale2, dfirst);
                int nsel2 = 0;
                if (nsel > 0) {
                    rc = tall_gemm(c, 1.0, Pw, PBIG, pb, c->wsel.p, PB, 0, 0.0, Qt, PB, PB, n);
                    if (rc) return rc;
                    if (s.rho > 0) {
                        // Re-orthogonalise the (unit-norm) candidates against the WHOLE basis: Wsel combines columns
                        // with coefficients up to 1e6, which lifts the eps-level basis components of the large
                        // columns to ~1e-10 of a small selected direction; without this step the basis loses
                        // orthogonality and later panels "find" directions that are already in it.
                        CU(c->cbuf.ensure((size_t)PBIG * std::max(s.rho, PB)));
                        // (only the first nsel candidate columns are nonzero: narrow kernels when nsel <= 16)
                        rc = gram_dev(c, Qt, PB, nsel, s.Q, s.ldq, s.rho, n, nullptr, c->cbuf.p, s.rho, nullptr, 0);
                        if (rc) return rc;
                        rc = tall_gemm(c, -1.0, s.Q, s.ldq, s.rho, c->cbuf.p, s.rho, 1, 1.0, Qt, PB, nsel, n);
                        if (rc) return rc;
                    }
                    rc = gram_dev(c, Qt, PB, PB, Qt, PB, PB, n, nullptr, c->gbuf2.p, PB, nullptr, 0);
                    if (rc) return rc;
                    launch_pivchol(c->gbuf2.p, PB, PB, 0.01, 0.0, c->wsel2.p, c->ibuf.p + 2, c->small.p + 2, c->st,
                                   &c->stats.kernel_launches);
                    rc = tall_gemm(c, 1.0, Qt, PB, PB, c->wsel2.p, PB, 0, 0.0, s.Q + s.rho, s.ldq, PB, n);
                    if (rc) return rc;
                    // coefficients of the sub-panel in the new directions (candidate columns beyond nsel2 are
                    // zero) and removal of that part, so that the basis is complete even if this was the last round
                    CU(c->cbuf.ensure((size_t)PBIG * std::max(s.rho, PB)));
                    rc = gram_dev(c, Pw, PBIG, pb, s.Q + s.rho, s.ldq, PB, n, nullptr, c->cbuf.p, PB, RTrow + s.rho,
                                  s.ldrt);
                    if (rc) return rc;
                    rc = tall_gemm(c, -1.0, s.Q + s.rho, s.ldq, PB, c->cbuf.p, PB, 1, 1.0, Pw, PBIG, pb, n);
                    if (rc) return rc;
                    CU(cudaMemcpyAsync(hi + 2, c->ibuf.p + 2, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
                    CU(cudaStreamSynchronize(c->st));
                    nsel2 = hi[2];
                    g_rr.ms_extend += wall_ms() - t_r1;
                }
                if (g_trace)
                    fprintf(stderr,
                            "[dre rr] row0 %5d pb %2d round %d rho %4d nsel %2d nsel2 %2d rem/col %.2e left %.2e "
                            "scale %.2e\n",
                            ch.rt_row + sc, pb, round, s.rho, nsel, nsel2,
                            std::sqrt(dfirst / std::max(trace_d0, 1e-300)),
                            std::sqrt(remaining / std::max(trace_d0, 1e-300)), std::sqrt(s.scale2));
                if (nsel == 0 || nsel2 == 0) break;
                s.rho += nsel2;
                g_rr.productive++;
                // the remaining (unselected) columns are certified negligible when the Gram rounding noise
                // (~1e-13 * dfirst) is below the drop threshold and the remaining Schur diagonal is too
                const double drop_now = std::max(s.drop_rel * std::sqrt(s.scale2), s.drop_abs);
                if (nsel2 == nsel && remaining <= drop_now * drop_now && 1e-13 * dfirst <= drop_now * drop_now) break;
            }
        }
    }
    if (look) CU(cudaStreamSynchronize(c->look_st));   // (nothing is pending there; keeps the invariant explicit)
    return DRE_OK;
}

// chunks of one term (column blocks of at most 256 columns)
void rr_add_term(std::vector<RRChunk>& chunks, const double* src, int64_t lds, int cols, const double* colscale,
                 int rt_row0) {
    for (int c0 = 0; c0 < cols; c0 += 256)
        chunks.push_back(RRChunk{src + c0, lds, std::min(256, cols - c0), colscale ? colscale + c0 : nullptr, rt_row0 + c0});
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* dre_version(void) { return "dre_b200 0.1 (sm_100a)"; }

const char* dre_last_error(const dre_context* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int32_t dre_symbolic_create(int64_t n, const int64_t* Ecp, const int64_t* Eri, const double* Enz, const int64_t* Acp,
                            const int64_t* Ari, const double* Anz, int32_t base, int32_t leaf_size,
                            dre_symbolic** out) {
    if (!out || !Ecp || !Eri || !Enz || !Acp || !Ari || !Anz) return fail(nullptr, DRE_ERR_ARG, "null argument");
    dre_symbolic* s = new (std::nothrow) dre_symbolic();
    if (!s) return fail(nullptr, DRE_ERR_LIB, "out of host memory");
    AnalyzeOptions opt;   // defaults = what dre_set_pencil uses
    if ((leaf_size & 0xffff) > 0) opt.leaf_size = leaf_size & 0xffff;
    if ((leaf_size >> 16) > 0) opt.max_snode = std::min(leaf_size >> 16, (int)SN_MAX);  // upper half-word: width cap
    std::string e;
    try {
        e = analyze(n, Ecp, Eri, Enz, Acp, Ari, Anz, base, opt, s->sym);
    } catch (const std::exception& ex) {
        e = std::string("analyze: ") + ex.what();
    }
    if (!e.empty()) { delete s; return fail(nullptr, DRE_ERR_ARG, e); }
    *out = s;
    return DRE_OK;
}

void dre_symbolic_destroy(dre_symbolic* s) { delete s; }

static void fill_info(const Symbolic& S, dre_symbolic_info* info) {
    info->n = S.n;
    info->nnz_pattern = (int64_t)S.csr_col.size();
    info->nnz_L = S.nnz_L;
    info->sum_update_rows = S.sum_u;
    info->factor_flops = S.flops;
    info->nsupernodes = S.nsn;
    info->nlevels = S.nlevels;
    info->max_front = S.max_front;
    info->max_supernode = S.max_sn;
}

int32_t dre_symbolic_get_info(const dre_symbolic* s, dre_symbolic_info* info) {
    if (!s || !info) return fail(nullptr, DRE_ERR_ARG, "null argument");
    fill_info(s->sym, info);
    return DRE_OK;
}

int32_t dre_symbolic_export(const dre_symbolic* s, const char* what, void* buf, int64_t cap, int64_t* len) {
    if (!s || !what || !len) return fail(nullptr, DRE_ERR_ARG, "null argument");
    const Symbolic& S = s->sym;
    std::string w(what);
    auto put_i = [&](const auto& v) {
        *len = (int64_t)v.size();
        int64_t* o = (int64_t*)buf;
        for (int64_t i = 0; buf && i < std::min<int64_t>(cap, *len); ++i) o[i] = (int64_t)v[i];
        return DRE_OK;
    };
    auto put_d = [&](const std::vector<double>& v) {
        *len = (int64_t)v.size();
        double* o = (double*)buf;
        for (int64_t i = 0; buf && i < std::min<int64_t>(cap, *len); ++i) o[i] = v[i];
        return DRE_OK;
    };
    if (w == "perm") return put_i(S.perm);
    if (w == "iperm") return put_i(S.iperm);
    if (w == "sn_first") return put_i(S.sn_first);
    if (w == "sn_rowptr") return put_i(S.sn_rowptr);
    if (w == "sn_rows") return put_i(S.sn_rows);
    if (w == "sn_parent") return put_i(S.sn_parent);
    if (w == "sn_level") return put_i(S.sn_level);
    if (w == "linv_off") return put_i(S.linv_off);
    if (w == "level_ptr") return put_i(S.level_ptr);
    if (w == "level_sn") return put_i(S.level_sn);
    if (w == "panel_off") return put_i(S.panel_off);
    if (w == "upd_off") return put_i(S.upd_off);
    if (w == "rhs_off") return put_i(S.rhs_off);
    if (w == "child_ptr") return put_i(S.child_ptr);
    if (w == "child_idx") return put_i(S.child_idx);
    if (w == "relmap") return put_i(S.relmap);
    if (w == "asm_dest") return put_i(S.asm_dest);
    if (w == "asm_a") return put_d(S.asm_a);
    if (w == "asm_e") return put_d(S.asm_e);
    if (w == "csr_ptr") return put_i(S.csr_ptr);
    if (w == "csr_col") return put_i(S.csr_col);
    if (w == "csr_a") return put_d(S.csr_a);
    if (w == "csr_e") return put_d(S.csr_e);
    return fail(nullptr, DRE_ERR_ARG, "dre_symbolic_export: unknown array name " + w);
}


int32_t dre_create(int32_t device, dre_context** out) {
    if (!out) return fail(nullptr, DRE_ERR_ARG, "null argument");
    dre_context* c = new (std::nothrow) dre_context();
    if (!c) return fail(nullptr, DRE_ERR_LIB, "out of host memory");
    c->device = device;
    auto bail = [&](const std::string& m) { g_last_error = m; delete c; return DRE_ERR_CUDA; };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return bail(std::string("cudaSetDevice: ") + cudaGetErrorString(e) +
                                      " (libdre_b200 has no CPU fallback; a CUDA device is required)");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return bail(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major < 10) return bail("libdre_b200 is built for sm_100a (B200) only");
    c->sm_count = prop.multiProcessorCount;
    if (const char* ev = getenv("DRE_SWEEP2")) c->sweep2 = atoi(ev) != 0;
    // launcher variants (process-wide, re-read whenever a context is created so that tests can switch them)
    {
        const char* ev = getenv("DRE_SPMM2");
        spmm_variant = (ev && atoi(ev) != 0) ? 2 : 1;
        ev = getenv("DRE_DIAG_NARROW_MIN");
        diag_narrow_min = ev ? std::max(1, atoi(ev)) : (1 << 30);
    }
    e = make_stream(&c->st, 0);
    if (e != cudaSuccess) return bail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    for (auto& fs : c->slot) {
        e = make_stream(&fs.st, 1);
        if (e != cudaSuccess) return bail(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
        cudaEventCreateWithFlags(&fs.ready, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&fs.released, cudaEventDisableTiming);
    }
    for_each_workspace(c, [&](auto& w) { w.arena = &c->arena; });
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreate(&c->tev0);
    cudaEventCreate(&c->tev1);
#ifdef DRE_SIMT_EMU
    if (cusolverDnCreate(&c->cusolver) != CUSOLVER_STATUS_SUCCESS) return bail("cusolverDnCreate failed");
    cusolverDnSetStream(c->cusolver, c->st);
#endif
    e = cudaMalloc((void**)&c->d_errflag, sizeof(int32_t));
    if (e != cudaSuccess) return bail("cudaMalloc failed");
    cudaMemset(c->d_errflag, 0, sizeof(int32_t));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return DRE_OK;
}

static void release_pencil(dre_context* c) {
    for (auto& fs : c->slot) if (fs.st) cudaStreamSynchronize(fs.st);
    if (c->norm_st) cudaStreamSynchronize(c->norm_st);
    c->norm_pending = false;
    for (void* p : c->owned) cudaFree(p);
    c->owned.clear();
    for (auto& fs : c->slot) {
        if (fs.L) cudaFree(fs.L);
        if (fs.Linv) cudaFree(fs.Linv);
        if (fs.dvec) cudaFree(fs.dvec);
        if (fs.U) cudaFree(fs.U);
        fs.L = fs.Linv = fs.dvec = fs.U = nullptr;
        fs.valid = fs.has_reader = fs.pending = false;
#ifndef DRE_SIMT_EMU
        for (auto& g : fs.fgraph) {
            if (g) cudaGraphExecDestroy(g);
            g = nullptr;
        }
        if (fs.d_prm) cudaFree(fs.d_prm);
        fs.d_prm = nullptr;
#endif
    }
    c->cur = 0;
    c->has_pencil = false;
}

static void rr_print_totals(dre_context* c) {
    fprintf(stderr,
            "[dre rr totals] blocks %ld rounds %ld productive %ld skipped sub-panels %ld rest projections %ld "
            "skipped second passes %ld round syncs %ld kernel launches (context) %lld; compress!/rrqr calls %ld: "
            "Gram-Schmidt %.1f ms (wait stage 1 %.1f, chunk projections %.1f, selection %.1f, extension %.1f), core %.1f ms, "
            "eigen %.1f ms, L<-QV %.1f ms; entry sync %.1f ms, setup %.1f ms (host wall clock)\n",
            g_rr.blocks, g_rr.rounds, g_rr.productive, g_rr.skipped, g_rr.rest_projections, g_rr.coef_only_passes,
            g_rr.syncs, (long long)c->stats.kernel_launches, g_rr.calls, g_rr.ms_rounds, g_rr.ms_wait_stage1,
            g_rr.ms_chunk_proj, g_rr.ms_select, g_rr.ms_extend, g_rr.ms_core, g_rr.ms_eig, g_rr.ms_final,
            g_rr.ms_entry_sync, g_rr.ms_setup);
    g_rr = RRTotals{};
}

int32_t dre_destroy(dre_context* c) {
    if (g_rr_stats && c) rr_print_totals(c);
    if (!c) return DRE_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    tl_dump();
    c->arena.destroy();
    release_pencil(c);
    for_each_workspace(c, [](auto& w) { w.forget(); });   // the arena (destroyed above) owned their memory
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->h_norm) cudaFreeHost(c->h_norm);
    if (c->h_look) cudaFreeHost(c->h_look);
    for (int i = 0; i < 2; ++i) if (c->look_done[i]) cudaEventDestroy(c->look_done[i]);
    if (c->look_basis) cudaEventDestroy(c->look_basis);
    if (c->look_st) cudaStreamDestroy(c->look_st);
    if (c->norm_in) cudaEventDestroy(c->norm_in);
    if (c->norm_done) cudaEventDestroy(c->norm_done);
    if (c->norm_st) cudaStreamDestroy(c->norm_st);
    if (c->d_errflag) cudaFree(c->d_errflag);
#ifdef DRE_SIMT_EMU
    if (c->cusolver) cusolverDnDestroy(c->cusolver);
#endif
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tev0) cudaEventDestroy(c->tev0);
    if (c->tev1) cudaEventDestroy(c->tev1);
    for (auto& fs : c->slot) {
        if (fs.ready) cudaEventDestroy(fs.ready);
        if (fs.released) cudaEventDestroy(fs.released);
        if (fs.st) cudaStreamDestroy(fs.st);
    }
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return DRE_OK;
}

int32_t dre_sync(dre_context* c) {
    if (c) cudaSetDevice(c->device);   // may be the first CUDA call of a host thread (compression lane)
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    CU(cudaStreamSynchronize(c->st));
    for (auto& fs : c->slot) CU(cudaStreamSynchronize(fs.st));
    return DRE_OK;
}

int32_t dre_set_pencil(dre_context* c, int64_t n, const int64_t* Ecp, const int64_t* Eri, const double* Enz,
                       const int64_t* Acp, const int64_t* Ari, const double* Anz, int32_t base) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->st));
    release_pencil(c);
    c->dense_only = false;
    for (Panel& p : c->panels) { p.alive = false; p.d = nullptr; }
    for_each_workspace(c, [](auto& w) { w.forget(); });
    c->arena.destroy();
    c->op_U = c->op_Vt = dre_view{-1, 0, 0};
    AnalyzeOptions opt;
    if (const char* ev = getenv("DRE_LEAF_SIZE")) opt.leaf_size = std::max(8, atoi(ev));
    if (const char* ev = getenv("DRE_MAX_SNODE")) opt.max_snode = std::max(8, atoi(ev));
    opt.max_snode = std::min(opt.max_snode, (int)SN_MAX);
    std::string e;
    try {
        e = analyze(n, Ecp, Eri, Enz, Acp, Ari, Anz, base, opt, c->sym);
    } catch (const std::exception& ex) {
        e = std::string("analyze: ") + ex.what();
    }
    if (!e.empty()) return fail(c, DRE_ERR_ARG, e);
    const Symbolic& S = c->sym;
    if (S.max_sn > SN_MAX) return fail(c, DRE_ERR_STATE, "internal error: supernode wider than SN_MAX");
    c->linv_elems = S.linv_off[S.nsn];
    c->upd_elems = S.upd_total;

    int rc;
    int32_t *d_sn_first, *d_sn_rows, *d_relmap, *d_child_ptr, *d_child_idx;
    int64_t *d_sn_rowptr, *d_panel_off, *d_linv_off, *d_upd_off, *d_rhs_off, *d_asm_dest;
    double *d_asm_a, *d_asm_e;
    if ((rc = upload_vec(c, S.sn_first, &d_sn_first))) return rc;
    if ((rc = upload_vec(c, S.sn_rowptr, &d_sn_rowptr))) return rc;
    if ((rc = upload_vec(c, S.sn_rows, &d_sn_rows))) return rc;
    if ((rc = upload_vec(c, S.relmap, &d_relmap))) return rc;
    if ((rc = upload_vec(c, S.child_ptr, &d_child_ptr))) return rc;
    if ((rc = upload_vec(c, S.child_idx, &d_child_idx))) return rc;
    if ((rc = upload_vec(c, S.panel_off, &d_panel_off))) return rc;
    if ((rc = upload_vec(c, S.linv_off, &d_linv_off))) return rc;
    if ((rc = upload_vec(c, S.upd_off, &d_upd_off))) return rc;
    if ((rc = upload_vec(c, S.rhs_off, &d_rhs_off))) return rc;
    if ((rc = upload_vec(c, S.asm_dest, &d_asm_dest))) return rc;
    if ((rc = upload_vec(c, S.asm_a, &d_asm_a))) return rc;
    if ((rc = upload_vec(c, S.asm_e, &d_asm_e))) return rc;
    if ((rc = upload_vec(c, S.iperm, &c->d_iperm))) return rc;
    if ((rc = upload_vec(c, S.csr_ptr, &c->d_csr_ptr))) return rc;
    if ((rc = upload_vec(c, S.csr_col, &c->d_csr_col))) return rc;
    if ((rc = upload_vec(c, S.csr_a, &c->d_csr_a))) return rc;
    if ((rc = upload_vec(c, S.csr_e, &c->d_csr_e))) return rc;
    if ((rc = upload_vec(c, S.level_sn, &c->d_level_sn))) return rc;
    DevSymbolic& D = c->dS;
    D.n = S.n; D.nsn = S.nsn; D.nlevels = S.nlevels;
    D.sn_first = d_sn_first; D.sn_rowptr = d_sn_rowptr; D.sn_rows = d_sn_rows; D.relmap = d_relmap;
    D.child_ptr = d_child_ptr; D.child_idx = d_child_idx; D.panel_off = d_panel_off; D.linv_off = d_linv_off;
    D.upd_off = d_upd_off; D.rhs_off = d_rhs_off;
    D.nasm = (int64_t)S.asm_dest.size(); D.asm_dest = d_asm_dest; D.asm_a = d_asm_a; D.asm_e = d_asm_e;

    // per-level work lists
    LevelLists lists;
    build_level_lists(S, lists);
    c->levels = lists.levels;
    const std::vector<int32_t>& ea_parents = lists.ea_parents;
    const std::vector<int2>& l21_items = lists.l21_items;
    const std::vector<int4>& schur_items = lists.schur_items;
    if ((rc = upload_vec(c, ea_parents, &c->d_ea_parents))) return rc;
    if ((rc = upload_vec(c, l21_items, &c->d_l21_items))) return rc;
    if ((rc = upload_vec(c, schur_items, &c->d_schur_items))) return rc;
    if ((rc = upload_vec(c, lists.fwd2_items, &c->d_fwd2_items))) return rc;
    if ((rc = upload_vec(c, lists.bwd2_items, &c->d_bwd2_items))) return rc;

    // factor storage, sized for complex
    for (auto& fs : c->slot) {
        CU(cudaMalloc(&fs.L, (size_t)std::max<int64_t>(S.nnz_L, 1) * sizeof(cplx)));
        CU(cudaMalloc(&fs.Linv, (size_t)std::max<int64_t>(c->linv_elems, 1) * sizeof(cplx)));
        CU(cudaMalloc(&fs.dvec, (size_t)S.n * sizeof(cplx)));
        CU(cudaMalloc(&fs.U, (size_t)std::max<int64_t>(c->upd_elems, 1) * sizeof(cplx)));
    }
    c->has_pencil = true;
    c->op_a = 1.0; c->op_e = 0.0; c->op_alpha = 1.0;
    return DRE_OK;
}

int32_t dre_get_symbolic_info(const dre_context* c, dre_symbolic_info* info) {
    if (!c || !info) return fail(nullptr, DRE_ERR_ARG, "null argument");
    if (!c->has_pencil) return fail(const_cast<dre_context*>(c), DRE_ERR_STATE, "no pencil set");
    fill_info(c->sym, info);
    return DRE_OK;
}

// ---- panels ----
int32_t dre_mat_create(dre_context* c, int32_t cols, int32_t* id) {
    if (c) cudaSetDevice(c->device);   // may be the first CUDA call of a host thread (compression lane)
    if (!c || !id) return fail(c, DRE_ERR_ARG, "null argument");
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    if (cols < 0) return fail(c, DRE_ERR_ARG, "negative column count");
    Panel p;
    p.cols = cols;
    p.ld = std::max(1, (cols + 1) & ~1);  // even leading dimension: 16-byte aligned rows
    {
        const size_t want = (size_t)c->sym.n * (size_t)p.ld * sizeof(double);
        size_t got = 0;
        void* blk;
        {
            HostTrace tr("panel arena alloc");
            blk = c->arena.alloc(want, &got);
        }
        if (!blk) return fail(c, DRE_ERR_CUDA, "out of device memory for a panel");
        p.d = (double*)blk;
        p.cap = got;
    }
    p.alive = true;
    for (size_t i = 0; i < c->panels.size(); ++i)
        if (!c->panels[i].alive) { c->panels[i] = p; *id = (int32_t)i; return DRE_OK; }
    c->panels.push_back(p);
    *id = (int32_t)c->panels.size() - 1;
    return DRE_OK;
}

int32_t dre_mat_free(dre_context* c, int32_t id) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (id < 0 || id >= (int)c->panels.size() || !c->panels[id].alive) return fail(c, DRE_ERR_ARG, "invalid panel id");
    if (c->op_U.id == id || c->op_Vt.id == id) c->op_U = c->op_Vt = dre_view{-1, 0, 0};
    if (c->panels[id].owned) c->arena.release(c->panels[id].d, c->panels[id].cap);
    c->panels[id].alive = false;
    c->panels[id].d = nullptr;
    return DRE_OK;
}

int32_t dre_mat_upload(dre_context* c, dre_view dst, const double* host, int64_t ld) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    TlScope tl_api("upload", 0, c ? c->st : nullptr);
    if (c->dense_only) return fail(c, DRE_ERR_STATE, "dense-only context: no row permutation to upload through");
    int rc;
    if ((rc = check_view(c, dst, "dst", true))) return rc;
    if (dst.ncols == 0) return DRE_OK;
    const int64_t n = c->sym.n;
    if (!host || ld < n) return fail(c, DRE_ERR_ARG, "bad host matrix");
    CU(c->stage.ensure((size_t)n * dst.ncols));
    CU(cudaMemcpy2DAsync(c->stage.p, n * sizeof(double), host, ld * sizeof(double), n * sizeof(double), dst.ncols,
                         cudaMemcpyHostToDevice, c->st));
    launch_colmajor_to_panel(vptr(c, dst), vld(c, dst), c->stage.p, n, n, dst.ncols, c->d_iperm, c->st,
                             &c->stats.kernel_launches);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->st));  // the host buffer is only borrowed for the duration of the call
    return DRE_OK;
}

int32_t dre_mat_download(dre_context* c, dre_view src, double* host, int64_t ld) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    TlScope tl_api("download", 0, c ? c->st : nullptr);
    if (c->dense_only) return fail(c, DRE_ERR_STATE, "dense-only context: no row permutation to download through");
    int rc;
    if ((rc = check_view(c, src, "src", true))) return rc;
    if (src.ncols == 0) return DRE_OK;
    const int64_t n = c->sym.n;
    if (!host || ld < n) return fail(c, DRE_ERR_ARG, "bad host matrix");
    CU(c->stage.ensure((size_t)n * src.ncols));
    launch_panel_to_colmajor(c->stage.p, n, vptr(c, src), vld(c, src), n, src.ncols, c->d_iperm, c->st,
                             &c->stats.kernel_launches);
    CU(cudaGetLastError());
    CU(cudaMemcpy2DAsync(host, ld * sizeof(double), c->stage.p, n * sizeof(double), n * sizeof(double), src.ncols,
                         cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return check_errflag(c);
}

int32_t dre_mat_copy(dre_context* c, dre_view dst, dre_view src) {
    if (c) cudaSetDevice(c->device);   // may be the first CUDA call of a host thread (compression lane)
    TlScope tl_api("mat_copy", 0, c ? c->st : nullptr);
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    int rc;
    if ((rc = check_view(c, dst, "dst", true)) || (rc = check_view(c, src, "src", true))) return rc;
    if (dst.ncols != src.ncols) return fail(c, DRE_ERR_ARG, "copy: column counts differ");
    if (views_overlap(dst, src)) return fail(c, DRE_ERR_ARG, "copy: views overlap");
    if (dst.ncols == 0) return DRE_OK;
    launch_copy_scale(vptr(c, dst), vld(c, dst), vptr(c, src), vld(c, src), c->sym.n, dst.ncols, nullptr, c->st,
                      &c->stats.kernel_launches);
    CU(cudaGetLastError());
    return DRE_OK;
}

int32_t dre_mat_axpby(dre_context* c, double alpha, dre_view X, double beta, dre_view Y) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    TlScope tl_api("axpby", 0, c ? c->st : nullptr);
    int rc;
    if ((rc = check_view(c, Y, "Y", true))) return rc;
    if (Y.ncols == 0) return DRE_OK;
    if (alpha != 0.0) {
        if ((rc = check_view(c, X, "X"))) return rc;
        if (X.ncols != Y.ncols) return fail(c, DRE_ERR_ARG, "axpby: column counts differ");
    }
    launch_axpby(alpha, alpha != 0.0 ? vptr(c, X) : nullptr, alpha != 0.0 ? vld(c, X) : 0, beta, vptr(c, Y),
                 vld(c, Y), c->sym.n, Y.ncols, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    return DRE_OK;
}

// ---- Heuristic shifts: one Arnoldi orthogonalisation step on device-resident vectors (heuristic.jl:111-125) ----
int32_t dre_arnoldi_orth(dre_context* c, dre_view V, dre_view w, dre_view vnext, double* h) {
    if (!c || !h) return fail(c, DRE_ERR_ARG, "null argument");
    TlScope tl_api("arnoldi_orth", 0, c->st);
    int rc;
    if ((rc = check_view(c, V, "V")) || (rc = check_view(c, w, "w", true)) || (rc = check_view(c, vnext, "vnext", true)))
        return rc;
    if (V.ncols < 1 || w.ncols != 1 || vnext.ncols != 1)
        return fail(c, DRE_ERR_ARG, "arnoldi_orth: V needs >= 1 column, w and vnext exactly one");
    if (views_overlap(V, w) || views_overlap(V, vnext) || views_overlap(w, vnext))
        return fail(c, DRE_ERR_ARG, "arnoldi_orth: views overlap");
    if (c->norm_pending) return fail(c, DRE_ERR_STATE, "arnoldi_orth: an asynchronous norm is pending (shared scratch)");
    const int nb = V.ncols, ncoef = 2 * nb + 1;
    CU(c->norm_small.ensure((size_t)2 * 296 + ncoef + 8));
    if ((size_t)ncoef + 4 > c->h_norm_cap) {
        if (c->h_norm) cudaFreeHost(c->h_norm);
        c->h_norm = nullptr;
        c->h_norm_cap = 0;
        CU(cudaMallocHost((void**)&c->h_norm, ((size_t)ncoef + 4 + 256) * sizeof(double)));
        c->h_norm_cap = (size_t)ncoef + 4 + 256;
    }
    double* part = c->norm_small.p;
    double* coef = c->norm_small.p + 2 * 296;
    launch_arnoldi_mgs(vptr(c, V), vld(c, V), nb, vptr(c, w), vld(c, w), vptr(c, vnext), vld(c, vnext), c->sym.n, part,
                       coef, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_norm, coef, ncoef * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    for (int i = 0; i < nb; ++i) h[i] = (0.0 + c->h_norm[i]) + c->h_norm[nb + i];   // H[i,j] += g, twice
    h[nb] = c->h_norm[2 * nb];
    if (!(h[nb] > 0.0) || !std::isfinite(h[nb]))
        return fail(c, DRE_ERR_NUMERIC, "arnoldi_orth: the Krylov space is exhausted (zero or non-finite remainder)");
    return DRE_OK;
}

// ---- products ----
int32_t dre_spmm(dre_context* c, int32_t op, double alpha, dre_view X, double beta, dre_view Y) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (!c->has_pencil || c->dense_only) return fail(c, DRE_ERR_STATE, "no pencil set");
    int rc;
    if ((rc = check_view(c, X, "X", true)) || (rc = check_view(c, Y, "Y", true))) return rc;
    if (X.ncols != Y.ncols) return fail(c, DRE_ERR_ARG, "spmm: column counts differ");
    if (views_overlap(X, Y)) return fail(c, DRE_ERR_ARG, "spmm: X and Y overlap");
    if (X.ncols == 0) return DRE_OK;
    const double* val = (op == 'E' || op == 'e') ? c->d_csr_e : (op == 'A' || op == 'a') ? c->d_csr_a : nullptr;
    if (!val) return fail(c, DRE_ERR_ARG, "spmm: op must be 'E' or 'A'");
    Timer t(c, &c->stats.ms_spmm);
    TlScope tl("spmm", 0, c->st);
    launch_spmm(c->d_csr_ptr, c->d_csr_col, val, c->sym.n, alpha, vptr(c, X), vld(c, X), beta, vptr(c, Y), vld(c, Y),
                X.ncols, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    c->stats.spmms++;
    c->stats.bytes_spmm += (double)c->sym.csr_col.size() * 12.0 + (double)(c->sym.n + 1) * 4.0 +
                           8.0 * (double)c->sym.n * X.ncols * (beta == 0.0 ? 2.0 : 3.0);
    return DRE_OK;
}

int32_t dre_gemm_tn(dre_context* c, dre_view X, dre_view Y, double* out, int64_t ld) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    TlScope tl_api("gemm_tn", 0, c ? c->st : nullptr);
    int rc;
    if ((rc = check_view(c, X, "X", true)) || (rc = check_view(c, Y, "Y", true))) return rc;
    const int a = X.ncols, b = Y.ncols;
    if (a == 0 || b == 0) return DRE_OK;
    if (!out || ld < a) return fail(c, DRE_ERR_ARG, "gemm_tn: bad output");
    CU(c->gbuf.ensure((size_t)a * b));
    if ((rc = gram_dev(c, vptr(c, X), vld(c, X), a, vptr(c, Y), vld(c, Y), b, c->sym.n, nullptr, c->gbuf.p, b, nullptr,
                       0)))
        return rc;
    if ((rc = ensure_pinned(c, (size_t)a * b))) return rc;
    CU(cudaMemcpyAsync(c->h_pinned, c->gbuf.p, (size_t)a * b * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j) out[i + (int64_t)j * ld] = c->h_pinned[(int64_t)i * b + j];
    return check_errflag(c);
}

int32_t dre_gemm_nn(dre_context* c, double alpha, dre_view X, const double* W, int64_t ldw, double beta, dre_view Y) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    TlScope tl_api("gemm_nn", 0, c ? c->st : nullptr);
    int rc;
    if ((rc = check_view(c, X, "X", true)) || (rc = check_view(c, Y, "Y", true))) return rc;
    const int a = X.ncols, b = Y.ncols;
    if (b == 0) return DRE_OK;
    if (views_overlap(X, Y)) return fail(c, DRE_ERR_ARG, "gemm_nn: X and Y overlap");
    if (a == 0) return dre_mat_axpby(c, 0.0, X, beta, Y);
    if (!W || ldw < a) return fail(c, DRE_ERR_ARG, "gemm_nn: bad W");
    if ((rc = ensure_pinned(c, (size_t)a * b))) return rc;
    CU(cudaStreamSynchronize(c->st));  // pinned staging may still be in flight
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j) c->h_pinned[(int64_t)i * b + j] = W[i + (int64_t)j * ldw];
    CU(c->gbuf2.ensure((size_t)a * b));
    CU(cudaMemcpyAsync(c->gbuf2.p, c->h_pinned, (size_t)a * b * sizeof(double), cudaMemcpyHostToDevice, c->st));
    rc = tall_gemm(c, alpha, vptr(c, X), vld(c, X), a, c->gbuf2.p, b, 0, beta, vptr(c, Y), vld(c, Y), b, c->sym.n);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->st));  // h_pinned / gbuf2 are reused by the next call
    return DRE_OK;
}

// ---- operator and solves ----
int32_t dre_set_operator(dre_context* c, double a, double e, double alpha, dre_view U, dre_view Vt) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    int rc;
    if ((rc = check_view(c, U, "U", true)) || (rc = check_view(c, Vt, "Vt", true))) return rc;
    if (U.ncols != Vt.ncols) return fail(c, DRE_ERR_ARG, "set_operator: U and Vt must have the same column count");
    if (U.ncols > 0 && alpha == 0.0) return fail(c, DRE_ERR_ARG, "set_operator: alpha must be nonzero");
    c->op_a = a; c->op_e = e; c->op_alpha = alpha;
    c->op_U = U.ncols ? U : dre_view{-1, 0, 0};
    c->op_Vt = Vt.ncols ? Vt : dre_view{-1, 0, 0};
    return DRE_OK;
}

int32_t dre_prefactor(dre_context* c, double mu_re, double mu_im) {
    if (c && c->dense_only) return fail(c, DRE_ERR_STATE, "no pencil set");
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    if (c->timing) return DRE_OK;   // per-class event timing serialises everything: no overlap to gain
    if (mu_im != 0.0) {
        cplx emu;
        make_emu(c->op_e, mu_re, mu_im, emu);
        return prefactor_t<cplx>(c, mu_re, mu_im, emu);
    }
    double emu;
    make_emu(c->op_e, mu_re, 0.0, emu);
    return prefactor_t<double>(c, mu_re, 0.0, emu);
}

int32_t dre_shift_solve(dre_context* c, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    return shifted_solve(c, mu_re, mu_im, R, V1, V2, false);
}

int32_t dre_adi_solve(dre_context* c, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    return shifted_solve(c, mu_re, mu_im, R, V1, V2, true);
}

int32_t dre_get_stream(dre_context* c, void** stream) {
    if (!c || !stream) return fail(c, DRE_ERR_ARG, "null argument");
    *stream = (void*)c->st;
    return DRE_OK;
}

int32_t dre_set_dense_only(dre_context* c, int64_t n) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (n <= 0) return fail(c, DRE_ERR_ARG, "dense-only context: bad row count");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->st));
    release_pencil(c);
    for (Panel& p : c->panels) { p.alive = false; p.d = nullptr; }
    for_each_workspace(c, [](auto& w) { w.forget(); });
    c->arena.destroy();
    c->op_U = c->op_Vt = dre_view{-1, 0, 0};
    c->sym = Symbolic();
    c->sym.n = n;
    c->levels.clear();
    c->has_pencil = true;   // panels and the dense toolbox work; the sparse entry points find an empty schedule
    c->dense_only = true;
    if (g_prio) {   // a lane yields to the ADI chain of the main context
        cudaStream_t low = nullptr;
        CU(make_stream(&low, 2));
        cudaStreamDestroy(c->st);
        c->st = low;
        if (c->look_st) { cudaStreamDestroy(c->look_st); c->look_st = nullptr; CU(make_stream(&c->look_st, 2)); }
    }
    return DRE_OK;
}

int32_t dre_mat_wrap(dre_context* c, void* device_ptr, int64_t ld, int32_t cols, int32_t* id) {
    if (!c || !id || !device_ptr) return fail(c, DRE_ERR_ARG, "null argument");
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    if (cols < 0 || ld < std::max(cols, 1)) return fail(c, DRE_ERR_ARG, "wrap: bad shape");
    Panel p;
    p.d = (double*)device_ptr;
    p.cols = cols;
    p.ld = ld;
    p.cap = 0;
    p.owned = false;
    p.alive = true;
    for (size_t i = 0; i < c->panels.size(); ++i)
        if (!c->panels[i].alive) { c->panels[i] = p; *id = (int32_t)i; return DRE_OK; }
    c->panels.push_back(p);
    *id = (int32_t)c->panels.size() - 1;
    return DRE_OK;
}

int32_t dre_mat_devptr(dre_context* c, dre_view v, void** ptr, int64_t* ld) {
    if (!c || !ptr || !ld) return fail(c, DRE_ERR_ARG, "null argument");
    int rc;
    if ((rc = check_view(c, v, "view"))) return rc;
    *ptr = vptr(c, v);
    *ld = vld(c, v);
    return DRE_OK;
}

int32_t dre_adi_step(dre_context* c, double mu_re, double mu_im, dre_view R, dre_view V1, dre_view V2) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    int rc = shifted_solve(c, mu_re, mu_im, R, V1, V2, true);
    if (rc) return rc;
    const double coef = (mu_im != 0.0) ? -2.0 * 1.4142135623730951 * mu_re : -2.0 * mu_re;
    Timer t(c, &c->stats.ms_spmm);
    launch_spmm(c->d_csr_ptr, c->d_csr_col, c->d_csr_e, c->sym.n, coef, vptr(c, V1), vld(c, V1), 1.0, vptr(c, R),
                vld(c, R), R.ncols, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    c->stats.spmms++;
    c->stats.bytes_spmm += (double)c->sym.csr_col.size() * 12.0 + (double)(c->sym.n + 1) * 4.0 +
                           8.0 * (double)c->sym.n * R.ncols * 3.0;
    return DRE_OK;
}

// ---- low-rank algebra ----
int32_t dre_ldlt_norm(dre_context* c, dre_view L, const double* D, int64_t ldd, double alpha, double* out) {
    if (!c || !out) return fail(c, DRE_ERR_ARG, "null argument");
    int rc;
    if ((rc = check_view(c, L, "L", true))) return rc;
    const int k = L.ncols;
    if (k == 0) { *out = 0.0; return DRE_OK; }
    if (!D || ldd < k) return fail(c, DRE_ERR_ARG, "norm: bad core matrix");
    Range r_norm("norm(::LDLt)");
    TlScope tl("norm", 0, c->st);
    bool diag = true;
    for (int j = 0; j < k && diag; ++j)
        for (int i = 0; i < k; ++i)
            if (i != j && D[i + (int64_t)j * ldd] != 0.0) { diag = false; break; }
    CU(c->gbuf.ensure((size_t)k * k));
    if ((rc = gram_dev(c, vptr(c, L), vld(c, L), k, vptr(c, L), vld(c, L), k, c->sym.n, nullptr, c->gbuf.p, k, nullptr,
                       0)))
        return rc;
    if ((rc = ensure_pinned(c, (size_t)k * k + k + 2))) return rc;
    if (diag) {
        // One host synchronisation per call (this runs once per ADI iteration): the pinned staging buffer is never
        // in flight when a C-ABI call starts (every asynchronous copy that touches it is followed by a
        // synchronisation inside the call that issued it), and the error flag rides on the same read-back.
        for (int i = 0; i < k; ++i) c->h_pinned[i] = D[i + (int64_t)i * ldd];
        CU(c->small.ensure(k + 8));
        CU(cudaMemcpyAsync(c->small.p + 8, c->h_pinned, k * sizeof(double), cudaMemcpyHostToDevice, c->st));
        launch_norm_diag(c->gbuf.p, k, k, c->small.p + 8, c->small.p, c->st, &c->stats.kernel_launches);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c->h_pinned + k, c->small.p, sizeof(double), cudaMemcpyDeviceToHost, c->st));
        int32_t* hflag = reinterpret_cast<int32_t*>(c->h_pinned + k + 1);
        CU(cudaMemcpyAsync(hflag, c->d_errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        const double v2 = c->h_pinned[k];
        *out = std::fabs(alpha) * std::sqrt(std::max(v2, 0.0));
        return *hflag != 0 ? check_errflag(c) : DRE_OK;   // (check_errflag resets the flag and words the message)
    } else {
        CU(cudaMemcpyAsync(c->h_pinned, c->gbuf.p, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
        // ||L D L'||_F^2 = tr(G D G D),  G = L'L  (host, rare path: dense core)
        std::vector<double> M((size_t)k * k, 0.0);
        const double* G = c->h_pinned;  // row-major, symmetric
        for (int i = 0; i < k; ++i)
            for (int t = 0; t < k; ++t) {
                const double g = G[(int64_t)i * k + t];
                if (g == 0.0) continue;
                for (int j = 0; j < k; ++j) M[(size_t)i * k + j] += g * D[t + (int64_t)j * ldd];
            }
        double tr = 0.0;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) tr += M[(size_t)i * k + j] * M[(size_t)j * k + i];
        *out = std::fabs(alpha) * std::sqrt(std::max(tr, 0.0));
    }
    return check_errflag(c);
}

int32_t dre_ldlt_norm_begin(dre_context* c, dre_view L, const double* d, double alpha) {
    if (!c || !d) return fail(c, DRE_ERR_ARG, "null argument");
    int rc;
    if ((rc = check_view(c, L, "L"))) return rc;
    if (c->norm_pending) return fail(c, DRE_ERR_STATE, "norm_begin: the previous asynchronous norm was not collected");
    const int k = L.ncols;
    if (!c->norm_st) {
        CU(make_stream(&c->norm_st, 0));
        CU(cudaEventCreateWithFlags(&c->norm_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->norm_done, cudaEventDisableTiming));
    }
    if ((size_t)k + 4 > c->h_norm_cap) {
        if (c->h_norm) cudaFreeHost(c->h_norm);
        c->h_norm = nullptr;
        c->h_norm_cap = 0;
        CU(cudaMallocHost((void**)&c->h_norm, ((size_t)k + 4 + 256) * sizeof(double)));
        c->h_norm_cap = (size_t)k + 4 + 256;
    }
    const GramPlan plan = gram_plan(c->sym.n, k, k, c->sm_count);
    // (growing a workspace is safe: no asynchronous norm is in flight here)
    CU(c->norm_partial.ensure(plan.partial_elems));
    CU(c->norm_g.ensure((size_t)k * k));
    CU(c->norm_small.ensure((size_t)k + 8));
    for (int i = 0; i < k; ++i) c->h_norm[i] = d[i];
    CU(cudaEventRecord(c->norm_in, c->st));                 // behind everything queued on the main stream so far
    CU(cudaStreamWaitEvent(c->norm_st, c->norm_in, 0));
    launch_gram(vptr(c, L), vld(c, L), k, vptr(c, L), vld(c, L), k, c->sym.n, nullptr, c->norm_partial.p, plan,
                c->norm_g.p, k, nullptr, 0, c->norm_st, &c->stats.kernel_launches);
    CU(cudaMemcpyAsync(c->norm_small.p + 8, c->h_norm, k * sizeof(double), cudaMemcpyHostToDevice, c->norm_st));
    launch_norm_diag(c->norm_g.p, k, k, c->norm_small.p + 8, c->norm_small.p, c->norm_st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_norm + k, c->norm_small.p, sizeof(double), cudaMemcpyDeviceToHost, c->norm_st));
    CU(cudaMemcpyAsync(c->h_norm + k + 1, c->d_errflag, sizeof(int32_t), cudaMemcpyDeviceToHost, c->norm_st));
    CU(cudaEventRecord(c->norm_done, c->norm_st));
    c->stats.grams++;
    c->stats.flops_gram += 2.0 * (double)c->sym.n * k * k;
    c->stats.bytes_gram += 8.0 * (double)c->sym.n * k;
    c->norm_pending = true;
    c->norm_alpha = alpha;
    c->norm_k = k;
    return DRE_OK;
}

int32_t dre_ldlt_norm_end(dre_context* c, double* out) {
    if (!c || !out) return fail(c, DRE_ERR_ARG, "null argument");
    if (!c->norm_pending) return fail(c, DRE_ERR_STATE, "norm_end without norm_begin");
    CU(cudaEventSynchronize(c->norm_done));
    CU(cudaStreamWaitEvent(c->st, c->norm_done, 0));        // later writers of L on the main stream are ordered behind it
    c->norm_pending = false;
    const int k = c->norm_k;
    *out = std::fabs(c->norm_alpha) * std::sqrt(std::max(c->h_norm[k], 0.0));
    const int32_t flag = *reinterpret_cast<const int32_t*>(c->h_norm + k + 1);
    return flag != 0 ? check_errflag(c) : DRE_OK;
}

// eigen-decomposition of the symmetric k x k matrix S (device, ld k): on return row j of S (row-major view) is the
// eigenvector of the j-th smallest eigenvalue, d_evals[j]
static int eig_sym_dev(dre_context* c, double* S, int k, double* d_evals) {
    Range r_eigen("eigen");
    HostTrace tr("eig_sym_dev", c->st);
#ifdef DRE_SIMT_EMU
    int lwork = 0;
    if (cusolverDnDsyevd_bufferSize(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, k, S, k, d_evals,
                                    &lwork) != CUSOLVER_STATUS_SUCCESS)
        return fail(c, DRE_ERR_LIB, "emulator eigensolver: bufferSize failed");
    CU(c->syevd_work.ensure((size_t)lwork + 8));
    CU(c->ibuf.ensure(8));
    if (cusolverDnDsyevd(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, k, S, k, d_evals,
                         c->syevd_work.p, lwork, c->ibuf.p + 4) != CUSOLVER_STATUS_SUCCESS)
        return fail(c, DRE_ERR_LIB, "emulator eigensolver failed");
    return DRE_OK;
#else
    CU(c->syevd_work.ensure(eig_sym_work_doubles(k)));
    std::vector<double> hev((size_t)k);
    auto grow = [](void* ctx, size_t bytes) -> void* {
        dre_context* cc = (dre_context*)ctx;
        return cc->eig_rots.ensure(bytes) == cudaSuccess ? (void*)cc->eig_rots.p : nullptr;
    };
    const int rc = eig_sym(S, k, d_evals, hev.data(), c->syevd_work.p, grow, c, c->sm_count, c->st,
                           &c->stats.kernel_launches);
    if (rc == 2) return fail(c, DRE_ERR_NUMERIC, "symmetric eigensolver: QL iteration did not converge");
    if (rc != 0) return fail(c, DRE_ERR_CUDA, std::string("symmetric eigensolver: ") + cudaGetErrorString(cudaGetLastError()));
    return DRE_OK;
#endif
}

static int rr_setup(dre_context* c, RRState& s, int ktot, double drop_rel, double drop_abs) {
    const int64_t n = c->sym.n;
    s.qcap = (int)std::min<int64_t>(ktot, n) + 64;
    s.ldq = (s.qcap + 1) & ~1;
    CU(c->qws.ensure((size_t)n * s.ldq));
    s.Q = c->qws.p;
    s.ldrt = s.ldq;
    CU(c->rt.ensure((size_t)ktot * s.ldrt));
    s.RT = c->rt.p;
    CU(cudaMemsetAsync(s.RT, 0, (size_t)ktot * s.ldrt * sizeof(double), c->st));
    s.rho = 0;
    s.scale2 = 0.0;
    s.drop_rel = drop_rel;
    s.drop_abs = drop_abs;
    return DRE_OK;
}

// compress! in three phases (dre_compress_begin / _add / _finish; dre_ldlt_compress = all three in one call).  The job
// (basis, coefficient rows, signs, dense cores) lives in the context between the calls, so that a caller who receives
// the terms one by one -- the compression lane of the multi-GPU pipeline mode -- can orthogonalise each increment while
// the next one is still being computed, and only the core / eigen / L <- QV tail is left when compress! is due.
int32_t dre_compress_begin(dre_context* c, int32_t max_cols, double tol_factor) {
    if (c) cudaSetDevice(c->device);   // may be the first CUDA call of a host thread (compression lane)
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    if (max_cols < 0) return fail(c, DRE_ERR_ARG, "compress: negative capacity");
    dre_context::CompressJob& job = c->cjob;
    job = dre_context::CompressJob{};
    job.kcap = max_cols;
    job.tol_factor = tol_factor;
    if (max_cols == 0) { job.active = true; return DRE_OK; }
    // Basis directions are dropped at HALF the relative level at which compress! truncates the eigenvalues of the
    // projected core below (tol_factor * eps, src/LDLt.jl:216-217): a direction whose coefficients are below
    // sigma * scale changes X by at most O(sigma) ||X|| through its cross terms with the large directions, i.e. by
    // no more than the truncation the reference applies itself.  (A tighter threshold -- 3e-15 was used before --
    // sits inside the round-off left by the block projections (~30 eps per pass over n rows), so that late ADI
    // increments "found" one to three noise directions per sub-panel: two thirds of all selection rounds.)
    static const double drop_env = getenv("DRE_RR_DROP") ? atof(getenv("DRE_RR_DROP")) : 0.0;
    const double drop_rel = drop_env > 0.0 ? drop_env : 0.5 * tol_factor * 2.220446049250313e-16;
    int rc;
    if ((rc = rr_setup(c, job.s, max_cols, drop_rel, 0.0))) return rc;
    CU(c->cscale.ensure((size_t)max_cols));
    job.signs.reserve(max_cols);
    job.active = true;   // (only now: a failed allocation leaves no half-open job behind)
    return DRE_OK;
}

int32_t dre_compress_scale_hint(dre_context* c, double scale) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (!c->cjob.active) return fail(c, DRE_ERR_STATE, "dre_compress_scale_hint without dre_compress_begin");
    if (!(scale >= 0.0) || !std::isfinite(scale)) return fail(c, DRE_ERR_ARG, "compress: bad scale hint");
    c->cjob.s.scale2 = std::max(c->cjob.s.scale2, scale * scale);
    return DRE_OK;
}

int32_t dre_compress_add(dre_context* c, int32_t nterms, const dre_view* Ls, const double* const* Ds,
                         const int64_t* ldds, const double* alphas) {
    if (c) cudaSetDevice(c->device);
    if (!c || !Ls || !Ds || !ldds || !alphas) return fail(c, DRE_ERR_ARG, "null argument");
    dre_context::CompressJob& job = c->cjob;
    if (!job.active) return fail(c, DRE_ERR_STATE, "dre_compress_add without dre_compress_begin");
    int rc;
    int kadd = 0;
    for (int t = 0; t < nterms; ++t) {
        if ((rc = check_view(c, Ls[t], "Ls[i]", true))) return rc;
        kadd += Ls[t].ncols;
    }
    const dre_view hint = c->ortho_hint;
    c->ortho_hint = dre_view{-1, 0, 0};
    if (kadd == 0) return DRE_OK;
    if (job.ktot + kadd > job.kcap) return fail(c, DRE_ERR_ARG, "compress: more columns than dre_compress_begin reserved");
    TlScope tl("compress", 0, c->st);
    const int64_t n = c->sym.n;
    Range r_compress("compress!(::LDLt)");
    RRState& s = job.s;
    // column scalings |alpha d_j|^(1/2) of the diagonal-core terms (1 for the columns of dense-core terms, whose core
    // enters after the basis is built, exactly as in the reference, src/LDLt.jl:206-213), uploaded once
    if ((rc = ensure_pinned(c, (size_t)kadd + 64))) return rc;
    const double t_e0 = wall_ms();
    CU(cudaStreamSynchronize(c->st));
    const double t_e1 = wall_ms();
    std::vector<char> is_diag(nterms, 0);
    const int base = job.ktot;
    job.signs.resize((size_t)base + kadd, 1.0);
    {
        int row0 = 0;
        for (int t = 0; t < nterms; ++t) {
            const int k = Ls[t].ncols;
            if (k == 0) continue;
            const double* D = Ds[t];
            const int64_t ldd = ldds[t];
            if (!D || ldd < k) return fail(c, DRE_ERR_ARG, "compress: bad core matrix");
            bool diag = true;
            for (int j = 0; j < k && diag; ++j)
                for (int i = 0; i < k; ++i)
                    if (i != j && D[i + (int64_t)j * ldd] != 0.0) { diag = false; break; }
            is_diag[t] = diag;
            for (int j = 0; j < k; ++j) {
                if (diag) {
                    const double v = alphas[t] * D[j + (int64_t)j * ldd];
                    c->h_pinned[row0 + j] = std::sqrt(std::fabs(v));
                    job.signs[base + row0 + j] = (v < 0.0) ? -1.0 : 1.0;
                } else {
                    c->h_pinned[row0 + j] = 1.0;
                }
            }
            if (!diag) {   // the symmetrised core alpha (D + D') / 2 is kept until dre_compress_finish
                dre_context::CompressJob::Dense dt;
                dt.row0 = base + row0;
                dt.k = k;
                dt.C.resize((size_t)k * k);
                for (int j = 0; j < k; ++j)
                    for (int i = 0; i < k; ++i)
                        dt.C[i + (size_t)j * k] = 0.5 * alphas[t] * (D[i + (int64_t)j * ldd] + D[j + (int64_t)i * ldd]);
                job.dense.push_back(std::move(dt));
            }
            row0 += k;
        }
    }
    CU(cudaMemcpyAsync(c->cscale.p + base, c->h_pinned, kadd * sizeof(double), cudaMemcpyHostToDevice, c->st));
    std::vector<RRChunk> chunks;
    {
        int row0 = 0;
        for (int t = 0; t < nterms; ++t) {
            const int k = Ls[t].ncols;
            if (k == 0) continue;
            const bool hinted = is_diag[t] && s.rho == 0 && job.ktot == 0 && chunks.empty() && hint.id == Ls[t].id &&
                                hint.col0 == Ls[t].col0 && hint.ncols == Ls[t].ncols && k + 64 <= s.qcap;
            if (hinted) {
                // the caller vouches that these columns are orthonormal (the outer factor a previous compress!
                // produced): they ARE the first k basis vectors, their coefficients are the column scalings
                launch_copy_scale(s.Q, s.ldq, vptr(c, Ls[t]), vld(c, Ls[t]), n, k, nullptr, c->st,
                                  &c->stats.kernel_launches);
                CU(cudaMemcpy2DAsync(s.RT + (int64_t)(base + row0) * s.ldrt, (size_t)(s.ldrt + 1) * sizeof(double),
                                     c->cscale.p + base + row0, sizeof(double), sizeof(double), k,
                                     cudaMemcpyDeviceToDevice, c->st));
                for (int j = 0; j < k; ++j) s.scale2 = std::max(s.scale2, c->h_pinned[row0 + j] * c->h_pinned[row0 + j]);
                s.rho = k;
            } else {
                rr_add_term(chunks, vptr(c, Ls[t]), vld(c, Ls[t]), k, is_diag[t] ? c->cscale.p + base + row0 : nullptr,
                            base + row0);
            }
            row0 += k;
        }
    }
    job.ktot += kadd;
    CU(cudaStreamSynchronize(c->st));   // (h_pinned is reused by the rounds)
    const double t_w0 = wall_ms();
    {
        Range r_orthf("orthf");
        HostTrace tr("compress: rank-revealing Gram-Schmidt", c->st);
        if ((rc = rr_process_chunks(c, s, chunks))) return rc;
    }
    if (g_rr_stats) {
        g_rr.ms_entry_sync += t_e1 - t_e0;
        g_rr.ms_setup += t_w0 - t_e1;
        g_rr.ms_rounds += wall_ms() - t_w0;
    }
    return check_errflag(c);
}

int32_t dre_compress_finish(dre_context* c, dre_view out, double* lam, int32_t* newrank) {
    if (c) cudaSetDevice(c->device);
    if (!c || !lam || !newrank) return fail(c, DRE_ERR_ARG, "null argument");
    dre_context::CompressJob& job = c->cjob;
    if (!job.active) return fail(c, DRE_ERR_STATE, "dre_compress_finish without dre_compress_begin");
    job.active = false;
    int rc;
    if ((rc = check_view(c, out, "out", true))) return rc;
    *newrank = 0;
    const int ktot = job.ktot;
    if (ktot == 0) return DRE_OK;
    TlScope tl("compress", 0, c->st);
    Range r_compress("compress!(::LDLt)");
    const int64_t n = c->sym.n;
    RRState& s = job.s;
    const double tol_factor = job.tol_factor;
    const std::vector<double>& signs = job.signs;
    const int rho = s.rho;
    if (rho == 0) return check_errflag(c);
    // S = RT' C RT (rho x rho), C = blockdiag(diag(signs) for the scaled diagonal-core terms, alpha_t D_t for the
    // dense-core terms):  M = C RT block by block, then one Gram product RT' M over the ktot coefficient rows.
    if ((rc = ensure_pinned(c, (size_t)ktot + rho + 64))) return rc;
    CU(cudaStreamSynchronize(c->st));
    const double t_w1 = wall_ms();
    for (int i = 0; i < ktot; ++i) c->h_pinned[i] = signs[i];
    CU(c->evals.ensure((size_t)ktot + rho));
    CU(cudaMemcpyAsync(c->evals.p, c->h_pinned, ktot * sizeof(double), cudaMemcpyHostToDevice, c->st));
    CU(c->gbuf.ensure((size_t)rho * rho));
    // M = C RT: rows of the diagonal-core terms scaled by their signs (exact), dense-core blocks replaced by
    // C_t RT_block; then S = RT' M is an ordinary Gram product over the ktot coefficient rows on the DMMA path
    CU(c->rt2.ensure((size_t)ktot * s.ldrt));
    launch_copy_scale(c->rt2.p, s.ldrt, s.RT, s.ldrt, ktot, rho, nullptr, c->st, &c->stats.kernel_launches, c->evals.p);
    CU(cudaGetLastError());
    for (const dre_context::CompressJob::Dense& dt : job.dense) {
        const int k = dt.k;
        if ((rc = ensure_pinned(c, (size_t)k * k + 64))) return rc;
        CU(cudaStreamSynchronize(c->st));
        std::memcpy(c->h_pinned, dt.C.data(), (size_t)k * k * sizeof(double));
        CU(c->gbuf2.ensure((size_t)k * k));
        CU(cudaMemcpyAsync(c->gbuf2.p, c->h_pinned, (size_t)k * k * sizeof(double), cudaMemcpyHostToDevice, c->st));
        // M_block (k x rho) = C_t (k x k, symmetric) * RT_block (k x rho)
        if ((rc = tall_gemm(c, 1.0, c->gbuf2.p, k, k, s.RT + (int64_t)dt.row0 * s.ldrt, s.ldrt, 0, 0.0,
                            c->rt2.p + (int64_t)dt.row0 * s.ldrt, s.ldrt, rho, k)))
            return rc;
        CU(cudaStreamSynchronize(c->st));   // h_pinned / gbuf2 are reused
    }
    if ((rc = gram_dev(c, s.RT, s.ldrt, rho, c->rt2.p, s.ldrt, rho, ktot, nullptr, c->gbuf.p, rho, nullptr, 0))) return rc;
    double* d_ev = c->evals.p + ktot;
    double t_w2 = 0.0;
    if (g_rr_stats) { CU(cudaStreamSynchronize(c->st)); t_w2 = wall_ms(); }
    if ((rc = eig_sym_dev(c, c->gbuf.p, rho, d_ev))) return rc;
    CU(cudaMemcpyAsync(c->h_pinned, d_ev, rho * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    const double t_w3 = wall_ms();
    double lmax = 0.0;
    for (int i = 0; i < rho; ++i) lmax = std::max(lmax, std::fabs(c->h_pinned[i]));
    const double eps_ = tol_factor * lmax * 2.220446049250313e-16;
    std::vector<int32_t> ids;
    for (int i = 0; i < rho; ++i)
        if (std::fabs(c->h_pinned[i]) >= eps_) ids.push_back(i);
    const int k2 = (int)ids.size();
    if (k2 > out.ncols) return fail(c, DRE_ERR_ARG, "compress: output panel too small for the compressed rank");
    for (int j = 0; j < k2; ++j) lam[j] = c->h_pinned[ids[j]];
    *newrank = k2;
    if (k2 > 0) {
        CU(c->ibuf.ensure((size_t)k2 + 16));
        int32_t* hi = (int32_t*)(c->h_pinned + rho);
        for (int j = 0; j < k2; ++j) hi[j] = ids[j];
        CU(cudaMemcpyAsync(c->ibuf.p + 16, hi, k2 * sizeof(int32_t), cudaMemcpyHostToDevice, c->st));
        CU(c->gbuf2.ensure((size_t)k2 * rho));
        launch_gather_rows(c->gbuf2.p, rho, c->gbuf.p, rho, c->ibuf.p + 16, k2, rho, c->st, &c->stats.kernel_launches);
        CU(cudaGetLastError());
        if ((rc = tall_gemm(c, 1.0, s.Q, s.ldq, rho, c->gbuf2.p, rho, 1, 0.0, vptr(c, out), vld(c, out), k2, n)))
            return rc;
        CU(cudaStreamSynchronize(c->st));
    }
    if (g_rr_stats) {
        g_rr.calls++;
        g_rr.ms_core += t_w2 - t_w1;
        g_rr.ms_eig += t_w3 - t_w2;
        g_rr.ms_final += wall_ms() - t_w3;
    }
    return check_errflag(c);
}

int32_t dre_ldlt_compress(dre_context* c, int32_t nterms, const dre_view* Ls, const double* const* Ds,
                          const int64_t* ldds, const double* alphas, double tol_factor, dre_view out, double* lam,
                          int32_t* newrank) {
    if (c) cudaSetDevice(c->device);   // may be the first CUDA call of a host thread (compression lane)
    if (!c || !Ls || !Ds || !ldds || !alphas || !lam || !newrank) return fail(c, DRE_ERR_ARG, "null argument");
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    int rc;
    int ktot = 0;
    for (int t = 0; t < nterms; ++t) {
        if ((rc = check_view(c, Ls[t], "Ls[i]", true))) return rc;
        ktot += Ls[t].ncols;
    }
    if ((rc = check_view(c, out, "out", true))) return rc;
    *newrank = 0;
    const dre_view hint = c->ortho_hint;   // (dre_compress_begin leaves it alone; _add consumes it)
    if ((rc = dre_compress_begin(c, ktot, tol_factor))) return rc;
    c->ortho_hint = hint;
    if ((rc = dre_compress_add(c, nterms, Ls, Ds, ldds, alphas))) { c->cjob.active = false; return rc; }
    return dre_compress_finish(c, out, lam, newrank);
}

int32_t dre_hint_orthonormal(dre_context* c, dre_view v) {
    if (c) cudaSetDevice(c->device);   // may be the first CUDA call of a host thread (compression lane)
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    int rc;
    if ((rc = check_view(c, v, "view", true))) return rc;
    c->ortho_hint = v;
    return DRE_OK;
}

int32_t dre_rrqr(dre_context* c, int32_t nviews, const dre_view* views, double drop_rel, double drop_abs, dre_view Q,
                 double* Rt, int64_t ldr, int32_t* rho_out) {
    if (!c || !views || !rho_out) return fail(c, DRE_ERR_ARG, "null argument");
    TlScope tl_api("rrqr", 0, c ? c->st : nullptr);
    if (!c->has_pencil) return fail(c, DRE_ERR_STATE, "no pencil set");
    int rc;
    int ktot = 0;
    for (int t = 0; t < nviews; ++t) {
        if ((rc = check_view(c, views[t], "views[i]", true))) return rc;
        ktot += views[t].ncols;
    }
    if ((rc = check_view(c, Q, "Q", true))) return rc;
    *rho_out = 0;
    if (ktot == 0) return DRE_OK;
    if (!Rt || ldr < ktot) return fail(c, DRE_ERR_ARG, "rrqr: bad Rt");
    // the basis and coefficient workspaces are shared with an open dre_compress_begin .. _finish job
    if (c->cjob.active)
        return fail(c, DRE_ERR_STATE, "rrqr: a compress job is open on this context (dre_compress_finish it first)");
    Range r_shifts("shifts");
    RRState s;
    if ((rc = rr_setup(c, s, ktot, drop_rel, drop_abs))) return rc;
    int row0 = 0;
    std::vector<RRChunk> chunks;
    for (int t = 0; t < nviews; ++t) {
        const int k = views[t].ncols;
        if (k == 0) continue;
        rr_add_term(chunks, vptr(c, views[t]), vld(c, views[t]), k, nullptr, row0);
        row0 += k;
    }
    if ((rc = rr_process_chunks(c, s, chunks))) return rc;
    const int rho = s.rho;
    if (rho > Q.ncols) return fail(c, DRE_ERR_ARG, "rrqr: Q panel too small");
    *rho_out = rho;
    if (rho == 0) return DRE_OK;
    launch_copy_scale(vptr(c, Q), vld(c, Q), s.Q, s.ldq, c->sym.n, rho, nullptr, c->st, &c->stats.kernel_launches);
    CU(cudaGetLastError());
    if ((rc = ensure_pinned(c, (size_t)ktot * s.ldrt))) return rc;
    CU(cudaMemcpyAsync(c->h_pinned, s.RT, (size_t)ktot * s.ldrt * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    for (int i = 0; i < ktot; ++i)
        for (int j = 0; j < rho; ++j) Rt[i + (int64_t)j * ldr] = c->h_pinned[(int64_t)i * s.ldrt + j];
    return check_errflag(c);
}

int32_t dre_debug_export(dre_context* c, const char* what, void* buf, int64_t cap_bytes, int64_t* len_bytes) {
    if (!c || !what || !len_bytes) return fail(c, DRE_ERR_ARG, "null argument");
    const std::string w(what);
    if (w == "timeline") {   // DRE_TIMELINE=<file>: resolve and write the recorded spans now
        tl_dump();
        *len_bytes = 0;
        return DRE_OK;
    }
    if (!c->has_pencil || !c->slot[c->cur].valid) return fail(c, DRE_ERR_STATE, "no numeric factorization held");
    const dre_context::FactorSlot& fs = c->slot[c->cur];
    const size_t tw = (size_t)fs.tw * sizeof(double);
    const void* src = nullptr;
    size_t bytes = 0;
    if (w == "L") { src = fs.L; bytes = (size_t)c->sym.nnz_L * tw; }
    else if (w == "Linv") { src = fs.Linv; bytes = (size_t)c->linv_elems * tw; }
    else if (w == "dvec") { src = fs.dvec; bytes = (size_t)c->sym.n * tw; }
    else if (w == "U") { src = fs.U; bytes = (size_t)c->upd_elems * tw; }
    else return fail(c, DRE_ERR_ARG, "dre_debug_export: unknown array name " + w);
    *len_bytes = (int64_t)bytes;
    if (buf && cap_bytes > 0) {
        CU(cudaStreamSynchronize(c->st));
        CU(cudaMemcpy(buf, src, std::min<size_t>(bytes, (size_t)cap_bytes), cudaMemcpyDeviceToHost));
    }
    return check_errflag(c);
}

int32_t dre_debug_eigh(dre_context* c, int32_t k, const double* A, double* evals, double* evecs) {
    if (!c || !A || !evals || !evecs || k <= 0) return fail(c, DRE_ERR_ARG, "dre_debug_eigh: bad argument");
    cudaSetDevice(c->device);
    CU(c->gbuf.ensure((size_t)k * k));
    CU(c->evals.ensure((size_t)k));
    CU(cudaMemcpyAsync(c->gbuf.p, A, (size_t)k * k * sizeof(double), cudaMemcpyHostToDevice, c->st));
    int rc = eig_sym_dev(c, c->gbuf.p, k, c->evals.p);
    if (rc) return rc;
    // row j of the row-major result = eigenvector j = column j of the column-major output
    CU(cudaMemcpyAsync(evecs, c->gbuf.p, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(evals, c->evals.p, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return DRE_OK;
}

int32_t dre_timer_start(dre_context* c) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    CU(cudaStreamSynchronize(c->st));
    CU(cudaEventRecord(c->tev0, c->st));
    return DRE_OK;
}

int32_t dre_timer_stop(dre_context* c, double* ms) {
    if (!c || !ms) return fail(c, DRE_ERR_ARG, "null argument");
    CU(cudaEventRecord(c->tev1, c->st));
    CU(cudaEventSynchronize(c->tev1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, c->tev0, c->tev1));
    *ms = f;
    return DRE_OK;
}

int32_t dre_stats_reset(dre_context* c, int32_t enable_event_timing) {
    if (!c) return fail(nullptr, DRE_ERR_ARG, "null context");
    if (g_rr_stats && g_rr.calls) rr_print_totals(c);   // (totals since the last reset; a reset starts a new window)
    c->stats = dre_stats{};
    c->timing = enable_event_timing != 0;
    return DRE_OK;
}

int32_t dre_stats_get(dre_context* c, dre_stats* out) {
    if (!c || !out) return fail(nullptr, DRE_ERR_ARG, "null argument");
    *out = c->stats;
    return DRE_OK;
}

}  // extern "C"
