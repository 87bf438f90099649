// Launch schedule of the supernodal solver: per-level work lists built once per pencil from the symbolic
// analysis, and the launch sequences of one numeric factorization / one pair of sweeps.
// Shared by context.cu (product: device pointers, CUDA streams) and by the SIMT-emulator harness of the CPU
// test tier (tests/simt/emu_harness.cpp: host pointers), so that both run the same schedule.
#pragma once
#include <algorithm>
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
};

struct LevelLists {
    std::vector<LevelWork> levels;
    std::vector<int32_t> ea_parents;   // supernodes with children, grouped by level
    std::vector<int2> l21_items;       // (J, 64-row slab of L21)
    std::vector<int4> schur_items;     // (J, ti, tj) lower-triangular 64x64 tiles of the update matrix
};

inline void build_level_lists(const Symbolic& S, LevelLists& out) {
    out.levels.assign(S.nlevels, LevelWork());
    out.ea_parents.clear();
    out.l21_items.clear();
    out.schur_items.clear();
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
};

// numeric LDL^T of  a*A + emu*E  into (L, Linv, dvec); L and U must be zeroed by the caller
template <class T>
inline void enqueue_factor(const DevSymbolic& dS, const DevSchedule& sch, T* L, T* Linv, T* dvec, T* U, double a, T emu,
                           int32_t* errflag, cudaStream_t st, int64_t* launches) {
    launch_assemble<T>(dS, L, a, emu, st, launches);
    for (int l = 0; l < sch.nlevels; ++l) {
        const LevelWork& lw = sch.levels[l];
        if (lw.ea_count > 0)
            launch_extend_add<T>(dS, sch.ea_parents + lw.ea_begin, lw.ea_count, lw.ea_gy, L, U, st, launches);
        launch_diag<T>(dS, sch.level_sn + lw.sn_begin, lw.sn_count, L, Linv, dvec, errflag, st, launches);
        if (lw.l21_count > 0) launch_l21<T>(dS, sch.l21_items + lw.l21_begin, lw.l21_count, L, Linv, dvec, st, launches);
        if (lw.schur_count > 0)
            launch_schur<T>(dS, sch.schur_items + lw.schur_begin, lw.schur_count, L, dvec, U, st, launches);
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

}  // namespace dre
