// Host-side symbolic analysis for the per-shift sparse factorization of  M(mu) = a*A + (e+mu)*E.
//
// Replaces (for the hot path) what SuiteSparse does behind the reference's
//   factorize(A)            /root/reference/src/blocklinear/types.jl:41-42, backslash.jl:13
//   factorize(::LowRankUpdate)  /root/reference/src/LowRankUpdate.jl:88-91
// pattern(A) u pattern(E) is constant over all shifts and time steps, so
// this runs ONCE per pencil: nested-dissection ordering, supernode partition (= the dissection
// blocks), supernodal elimination tree, row structures, level schedule (by depth, deepest first),
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
    std::vector<int32_t> sn_level;   // level index, 0 = deepest level (processed first in factor/forward)
    int32_t nlevels = 0;
    std::vector<int32_t> level_ptr;  // nlevels+1 offsets into level_sn
    std::vector<int32_t> level_sn;   // supernodes grouped by level

    // bottom subtrees: maximal subtrees with at most `subtree_cols` columns.  One CTA walks a whole
    // subtree, so the (many, tiny) lower levels of the tree need a single kernel launch; only the "top"
    // supernodes above the subtree roots are processed level by level.
    std::vector<int32_t> sn_subtree;   // nsn: subtree id, -1 for top supernodes
    int32_t nsubtrees = 0;
    std::vector<int32_t> st_ptr;       // nsubtrees+1 offsets into st_sn
    std::vector<int32_t> st_sn;        // supernodes of each subtree, ascending (children before parents)
    int32_t ntoplevels = 0;
    std::vector<int32_t> top_level_ptr;  // ntoplevels+1 offsets into top_level_sn (deepest top level first)
    std::vector<int32_t> top_level_sn;

    std::vector<int64_t> panel_off;  // nsn+1: offset of the f_J x s_J column-major panel (ld = f_J) in L storage
    // update matrices (u_J x u_J): absolute offsets into one buffer laid out as
    //   [ all bottom supernodes | top levels of even parity | top levels of odd parity ]
    std::vector<int64_t> upd_off;
    int64_t upd_bottom_elems = 0;
    int64_t max_upd_level[2] = {0, 0};  // top part, even / odd (top-)levels
    std::vector<int64_t> top_level_upd_begin, top_level_upd_elems;  // per top level: absolute range to clear
    // update vectors of the solves: row offset (cumulative over ALL supernodes, no buffer reuse)
    std::vector<int64_t> rhs_off;

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
    int64_t sum_u = 0;       // sum of u_J (rows of update vectors in a solve)

    int32_t sn_size(int32_t J) const { return sn_first[J + 1] - sn_first[J]; }
    int32_t sn_nrows(int32_t J) const { return (int32_t)(sn_rowptr[J + 1] - sn_rowptr[J]); }
    int32_t front(int32_t J) const { return sn_size(J) + sn_nrows(J); }
};

struct AnalyzeOptions {
    int32_t leaf_size = 48;       // dissection stops below this many vertices
    int32_t max_snode = 1 << 30;  // dissection blocks wider than this are split into chains of supernodes
    int32_t subtree_cols = 0;     // bottom-subtree size cap; 0 = automatic (n/400 clamped to [64, 1024])
};

// E and A: CSC, index_base 0 or 1, 64-bit indices (Julia SparseMatrixCSC{Float64,Int64} zero-copy).
// Both must be square n x n with symmetric values (checked; the nonsymmetric case is SURVEY 8f rank 3).
// Returns empty string on success, otherwise an error message.
std::string analyze(int64_t n, const int64_t* Ecolptr, const int64_t* Erowval, const double* Enz,
                    const int64_t* Acolptr, const int64_t* Arowval, const double* Anz, int index_base,
                    const AnalyzeOptions& opt, Symbolic& out);

}  // namespace dre
