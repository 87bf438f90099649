// Launch schedule of the supernodal solver: per-level work lists built once per pencil from the symbolic
// analysis, and the launch sequences of one numeric factorization / one pair of sweeps.
// Shared by context.cu (product: device pointers, CUDA streams) and by the SIMT-emulator harness of the CPU
// test tier (tests/simt/emu_harness.cpp: host pointers), so that both run the same schedule.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "kernels.h"
#include "symbolic.h"

namespace dre {

// work of one level (levels = heights in the supernodal tree, leaves first)
struct LevelWork {
    int sn_begin = 0, sn_count = 0, smax = 0;
    int ea_begin = 0, ea_count = 0, ea_gy = 1;
    int l21_begin = 0, l21_count = 0;
    int schur_begin = 0, schur_count = 0;
    // row-split sweeps (sweep v2): items (J, row block of 8*q strips)
    int fwd2_begin = 0, fwd2_count = 0, fwd2_q = 1;
    int bwd2_begin = 0, bwd2_count = 0, bwd2_q = 1;
    int has_children = 0;
    int64_t upd_off = 0, upd_size = 0;   // this level's segment of the update-matrix pool (zeroed before the level runs)
};

struct LevelLists {
    std::vector<LevelWork> levels;
    std::vector<int32_t> ea_parents;   // supernodes with children, grouped by level
    std::vector<int2> l21_items;       // (J, 64-row slab of L21)
    std::vector<int4> schur_items;     // (J, ti, tj) lower-triangular 64x64 tiles of the update matrix
    std::vector<int2> fwd2_items;      // (J, rb): strips [rb*8q, (rb+1)*8q) of the ceil(s/8) + ceil(u/8) front strips
    std::vector<int2> bwd2_items;      // (J, rb): strips [rb*8q, (rb+1)*8q) of the ceil(s/8) supernode strips
};

// levels with at least this many supernodes count as populous for the row-split sweeps (37 x 8 chunks of 32
// right-hand sides = two CTAs per SM at ~250 columns); DRE_SWEEP2_POPULOUS_MIN overrides it (tests)
inline int sweep2_populous_min() {
    const char* ev = getenv("DRE_SWEEP2_POPULOUS_MIN");
    return ev ? std::max(1, atoi(ev)) : 37;
}

inline void build_level_lists(const Symbolic& S, LevelLists& out) {
    out.levels.assign(S.nlevels, LevelWork());
    out.ea_parents.clear();
    out.l21_items.clear();
    out.schur_items.clear();
    out.fwd2_items.clear();
    out.bwd2_items.clear();
    for (int l = 0; l < S.nlevels; ++l) {
        LevelWork& lw = out.levels[l];
        lw.sn_begin = S.level_ptr[l];
        lw.sn_count = S.level_ptr[l + 1] - S.level_ptr[l];
        lw.ea_begin = (int)out.ea_parents.size();
        lw.l21_begin = (int)out.l21_items.size();
        lw.schur_begin = (int)out.schur_items.size();
        int max_f = 0;
        for (int p = S.level_ptr[l]; p < S.level_ptr[l + 1]; ++p) {
            const int J = S.level_sn[p];
            const int u = S.sn_nrows(J);
            lw.smax = std::max(lw.smax, S.sn_size(J));
            max_f = std::max(max_f, S.front(J));
            if (S.child_ptr[J + 1] > S.child_ptr[J]) out.ea_parents.push_back(J);
            for (int sl = 0; sl * 64 < u; ++sl) out.l21_items.push_back(make_int2(J, sl));
            const int nt = (u + 63) / 64;
            for (int ti = 0; ti < nt; ++ti)
                for (int tj = 0; tj <= ti; ++tj) out.schur_items.push_back(make_int4(J, ti, tj, 0));
        }
        lw.ea_count = (int)out.ea_parents.size() - lw.ea_begin;
        lw.l21_count = (int)out.l21_items.size() - lw.l21_begin;
        lw.schur_count = (int)out.schur_items.size() - lw.schur_begin;
        lw.ea_gy = std::min(64, std::max(1, max_f / 8));
        lw.upd_off = S.upd_level_off[l];
        lw.upd_size = S.upd_level_size[l];
        // Row-split sweeps.  Populous levels (enough supernodes to fill the machine with one CTA each at ~250
        // right-hand sides) keep a supernode in as few CTAs as possible; the sparse levels near the root spread
        // every supernode over one CTA per 64 rows.
        const bool populous = lw.sn_count >= sweep2_populous_min();
        lw.fwd2_q = populous ? 3 : 1;
        lw.bwd2_q = populous ? 2 : 1;
        lw.has_children = lw.ea_count > 0;
        lw.fwd2_begin = (int)out.fwd2_items.size();
        lw.bwd2_begin = (int)out.bwd2_items.size();
        for (int p = S.level_ptr[l]; p < S.level_ptr[l + 1]; ++p) {
            const int J = S.level_sn[p];
            const int sy = (S.sn_size(J) + 7) / 8, st = (S.sn_nrows(J) + 7) / 8;
            for (int rb = 0; rb * 8 * lw.fwd2_q < sy + st; ++rb) out.fwd2_items.push_back(make_int2(J, rb));
            for (int rb = 0; rb * 8 * lw.bwd2_q < sy; ++rb) out.bwd2_items.push_back(make_int2(J, rb));
        }
        lw.fwd2_count = (int)out.fwd2_items.size() - lw.fwd2_begin;
        lw.bwd2_count = (int)out.bwd2_items.size() - lw.bwd2_begin;
    }
}

// the lists as the kernels see them (device pointers in the product, host pointers under the emulator)
struct DevSchedule {
    const LevelWork* levels;   // host array, nlevels entries
    int nlevels;
    const int32_t* level_sn;
    const int32_t* ea_parents;
    const int2* l21_items;
    const int4* schur_items;
    const int2* fwd2_items;    // null unless the context runs the row-split sweeps
    const int2* bwd2_items;
};

// numeric LDL^T of  a*A + emu*E  into (L, Linv, dvec); L must be zeroed by the caller, the pooled update matrices U
// are zeroed level by level here (a level's segment recycles the space of levels that are already consumed).
// m21: leave M21 = L21 Linv instead of L21 in the panels (what the row-split sweeps read).
template <class T>
inline void enqueue_factor(const DevSymbolic& dS, const DevSchedule& sch, T* L, T* Linv, T* dvec, T* U, double a, T emu,
                           int32_t* errflag, cudaStream_t st, int64_t* launches, bool m21 = false,
                           const double* prm = nullptr) {
    launch_assemble<T>(dS, L, a, emu, st, launches, prm);
    for (int l = 0; l < sch.nlevels; ++l) {
        const LevelWork& lw = sch.levels[l];
        if (lw.upd_size > 0) cudaMemsetAsync(U + lw.upd_off, 0, (size_t)lw.upd_size * sizeof(T), st);
        if (lw.ea_count > 0)
            launch_extend_add<T>(dS, sch.ea_parents + lw.ea_begin, lw.ea_count, lw.ea_gy, L, U, st, launches);
        launch_diag<T>(dS, sch.level_sn + lw.sn_begin, lw.sn_count, lw.smax, L, Linv, dvec, errflag, st, launches);
        if (lw.l21_count > 0) launch_l21<T>(dS, sch.l21_items + lw.l21_begin, lw.l21_count, L, Linv, dvec, st, launches);
        if (lw.schur_count > 0)
            launch_schur<T>(dS, sch.schur_items + lw.schur_begin, lw.schur_count, L, dvec, U, st, launches);
        if (m21 && lw.l21_count > 0) launch_m21<T>(dS, sch.l21_items + lw.l21_begin, lw.l21_count, L, Linv, st, launches);
    }
}

// forward + backward sweep of the row-major block W (n x ldw, nrhs columns); tbuf holds the update vectors
template <class T>
inline void enqueue_sweeps(const DevSymbolic& dS, const DevSchedule& sch, const T* L, const T* Linv, const T* dvec,
                           T* W, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, cudaStream_t st,
                           int64_t* launches) {
    for (int l = 0; l < sch.nlevels; ++l) {
        const LevelWork& lw = sch.levels[l];
        launch_fwd_level<T>(dS, sch.level_sn + lw.sn_begin, lw.sn_count, lw.smax, L, Linv, W, ldw, nrhs, tbuf, src, st,
                            launches);
    }
    for (int l = sch.nlevels - 1; l >= 0; --l) {
        const LevelWork& lw = sch.levels[l];
        launch_bwd_level<T>(dS, sch.level_sn + lw.sn_begin, lw.sn_count, lw.smax, L, Linv, dvec, W, ldw, nrhs, st,
                            launches);
    }
}

// row-split sweeps on a factorization made with m21 = true: y' = D^-1 Linv (b + children) goes to Y, x to W
template <class T>
inline void enqueue_sweeps2(const DevSymbolic& dS, const DevSchedule& sch, const T* L, const T* Linv, const T* dvec,
                            T* W, T* Y, int64_t ldw, int nrhs, T* tbuf, const RhsSource& src, cudaStream_t st,
                            int64_t* launches) {
    for (int l = 0; l < sch.nlevels; ++l) {
        const LevelWork& lw = sch.levels[l];
        launch_fwd2_level<T>(dS, sch.fwd2_items + lw.fwd2_begin, lw.fwd2_count, lw.fwd2_q, lw.smax, L, Linv, dvec, Y, ldw,
                             nrhs, tbuf, src, lw.has_children, st, launches);
    }
    for (int l = sch.nlevels - 1; l >= 0; --l) {
        const LevelWork& lw = sch.levels[l];
        launch_bwd2_level<T>(dS, sch.bwd2_items + lw.bwd2_begin, lw.bwd2_count, lw.bwd2_q, lw.smax, L, Linv, Y, W, ldw,
                             nrhs, st, launches);
    }
}

}  // namespace dre
