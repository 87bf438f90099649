// Host-side symbolic analysis for the per-shift sparse factorization of  M(mu) = a*A + (e+mu)*E.
//
// Replaces (for the hot path) what SuiteSparse does behind the reference's
//   factorize(A)            /root/reference/src/blocklinear/types.jl:41-42, backslash.jl:13
//   factorize(::LowRankUpdate)  /root/reference/src/LowRankUpdate.jl:88-91
// pattern(A) u pattern(E) is constant over all shifts and time steps, so
// this runs ONCE per pencil: nested-dissection ordering, supernode partition (= the dissection
// blocks), supernodal elimination tree, row structures, level schedule (by height, leaves first),
// extend-add index maps and the scatter map that assembles a*A+(e+mu)*E into the supernodal panels
// with two scalars per shift.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dre {

struct Symbolic {
    int64_t n = 0;
    // permutation: perm[new] = old, iperm[old] = new
    std::vector<int32_t> perm, iperm;

    // supernodes (in elimination order; children always precede parents)
    int32_t nsn = 0;
    std::vector<int32_t> sn_first;   // nsn+1: columns [sn_first[J], sn_first[J+1])
    std::vector<int64_t> sn_rowptr;  // nsn+1: offsets into sn_rows
    std::vector<int32_t> sn_rows;    // below-supernode row structure (sorted, new indices)
    std::vector<int32_t> sn_parent;  // supernodal etree (-1 = root)
    std::vector<int32_t> sn_level;   // level = height in the tree (leaves 0): level l only depends on levels < l
    int32_t nlevels = 0;
    std::vector<int32_t> level_ptr;  // nlevels+1 offsets into level_sn
    std::vector<int32_t> level_sn;   // supernodes grouped by level

    std::vector<int64_t> panel_off;  // nsn+1: offset of the f_J x s_J column-major panel (ld = f_J) in L storage
    std::vector<int64_t> linv_off;   // nsn+1: offset of the s_J x s_J column-major inverse of the unit-lower diagonal block
    // Pooled by liveness (written on level(J), last read on level(parent(J))); see pool_by_level in symbolic.cpp
    std::vector<int64_t> upd_off;    // nsn+1: offset of the u_J x u_J column-major update matrix; [nsn] = upd_total
    std::vector<int64_t> rhs_off;    // nsn: row offset of the u_J-row update vector of the forward sweep
    int64_t upd_total = 0;           // entries of the update-matrix pool
    int64_t rhs_total = 0;           // rows of the update-vector pool
    std::vector<int64_t> upd_level_off, upd_level_size;   // nlevels: the pool segment of every level (zeroed per level)

    // children lists (CSR by parent) and relative maps child-struct-row -> parent front local index
    std::vector<int32_t> child_ptr, child_idx;
    std::vector<int32_t> relmap;     // same indexing as sn_rows: local row in the parent's front

    // scatter map for assembly: lower-triangular union pattern of A and E in the new ordering
    std::vector<int64_t> asm_dest;   // destination offset in L storage
    std::vector<double> asm_a, asm_e;

    // full symmetric CSR copies in the new ordering (for SpMM with E, A)
    std::vector<int32_t> csr_ptr;    // n+1   (pattern union, shared by A and E)
    std::vector<int32_t> csr_col;
    std::vector<double> csr_a, csr_e;

    // statistics
    int64_t nnz_L = 0;       // sum f_J*s_J (stored panel entries, incl. relaxed zeros)
    double flops = 0;        // real multiply-add pairs*2 of the supernodal LDL^T
    int32_t max_front = 0, max_sn = 0;
    int64_t sum_u = 0;       // sum of u_J (statistics; the update vectors occupy rhs_total <= sum_u rows)

    int32_t sn_size(int32_t J) const { return sn_first[J + 1] - sn_first[J]; }
    int32_t sn_nrows(int32_t J) const { return (int32_t)(sn_rowptr[J + 1] - sn_rowptr[J]); }
    int32_t front(int32_t J) const { return sn_size(J) + sn_nrows(J); }
};

struct AnalyzeOptions {
    int32_t leaf_size = 96;       // dissection stops below this many vertices (a leaf is ONE dense supernode)
    int32_t max_snode = 256;      // dissection blocks wider than this are split into chains of supernodes
};

// E and A: CSC, index_base 0 or 1, 64-bit indices (Julia SparseMatrixCSC{Float64,Int64} zero-copy).
// Both must be square n x n with symmetric values (checked; the nonsymmetric case is SURVEY 8f rank 3).
// Returns empty string on success, otherwise an error message.
std::string analyze(int64_t n, const int64_t* Ecolptr, const int64_t* Erowval, const double* Enz,
                    const int64_t* Acolptr, const int64_t* Arowval, const double* Anz, int index_base,
                    const AnalyzeOptions& opt, Symbolic& out);

}  // namespace dre
