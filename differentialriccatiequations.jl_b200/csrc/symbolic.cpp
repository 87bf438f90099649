// Host-side symbolic analysis (see symbolic.h).  Plain C++17, no third-party ordering library:
// the image has no METIS/AMD, so the nested dissection is written here (BFS level-set and
// two-ended "bisector" vertex separators, trimmed, best of both).
#include "symbolic.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstring>
#include <numeric>
#include <queue>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace dre {
namespace {

struct Graph {
    int64_t n;
    std::vector<int64_t> ptr;
    std::vector<int32_t> adj;
};

// ------------------------------------------------------------------------------------------
// nested dissection
// ------------------------------------------------------------------------------------------
struct Dissector {
    const Graph& g;
    int32_t leaf;
    std::vector<int32_t> stamp;    // region membership stamp
    std::vector<int32_t> d1, d2;   // BFS distances
    std::vector<int8_t> part;      // 0 / 1 / 2 (=separator)
    std::atomic<int32_t> cur_stamp{0};
    // The two halves of a dissected region are independent, so the top levels of the recursion run as parallel
    // tasks (par_depth levels -> up to 2^par_depth tasks).  The per-vertex arrays above are shared: a task only
    // touches vertices of its own region (every neighbour access is guarded by the region stamp, and the stamps of
    // the separators above it do not change while it runs); the BFS queue and the output are per task, and the
    // outputs are concatenated in the sequential order (left half, right half, separator), so the ordering does not
    // depend on the number of threads.
    int par_depth = 0;
    struct Out {
        std::vector<int32_t> order;      // perm: new -> old
        std::vector<int32_t> block_end;  // end offsets (in `order`) of supernode blocks
        void emit_block(const std::vector<int32_t>& verts) {
            if (verts.empty()) return;
            for (int32_t v : verts) order.push_back(v);
            block_end.push_back((int32_t)order.size());
        }
        void append(const Out& o) {
            const int32_t base = (int32_t)order.size();
            order.insert(order.end(), o.order.begin(), o.order.end());
            for (int32_t e : o.block_end) block_end.push_back(base + e);
        }
    };
    struct Scratch {
        std::vector<int32_t> queue_;
    };
    // output
    Out out;

    Dissector(const Graph& g_, int32_t leaf_) : g(g_), leaf(leaf_) {
        stamp.assign(g.n, -1);
        d1.assign(g.n, -1);
        d2.assign(g.n, -1);
        part.assign(g.n, 0);
        out.order.reserve(g.n);
        int threads = (int)std::thread::hardware_concurrency();
        if (const char* ev = getenv("DRE_SYMBOLIC_THREADS")) threads = atoi(ev);
        threads = std::max(1, std::min(threads, 64));
        while ((1 << par_depth) < threads) ++par_depth;
    }

    void run(std::vector<int32_t>& verts) {
        Scratch sc;
        sc.queue_.reserve(g.n);
        dissect(verts, out, sc, 0);
    }

    // BFS inside the region marked with `st`; fills dist for reached vertices, returns visit order in queue_
    int32_t bfs(int32_t src, int32_t st, std::vector<int32_t>& dist, const std::vector<int32_t>& verts,
                std::vector<int32_t>& queue_) {
        for (int32_t v : verts) dist[v] = -1;
        queue_.clear();
        queue_.push_back(src);
        dist[src] = 0;
        size_t head = 0;
        while (head < queue_.size()) {
            int32_t v = queue_[head++];
            for (int64_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                int32_t w = g.adj[p];
                if (stamp[w] == st && dist[w] < 0) {
                    dist[w] = dist[v] + 1;
                    queue_.push_back(w);
                }
            }
        }
        return (int32_t)queue_.size();
    }

    struct Cand {
        bool ok = false;
        int64_t s = 0, p0 = 0, p1 = 0;
        double cost = 1e300;
        std::vector<int8_t> lab;  // label per position in verts
    };

    // trim a vertex separator: move separator vertices that touch only one side into that side
    void trim(const std::vector<int32_t>& verts, int32_t st, Cand& c) {
        for (size_t i = 0; i < verts.size(); ++i) part[verts[i]] = c.lab[i];
        for (size_t i = 0; i < verts.size(); ++i) {
            int32_t v = verts[i];
            if (part[v] != 2) continue;
            int n0 = 0, n1 = 0;
            for (int64_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                int32_t w = g.adj[p];
                if (stamp[w] != st) continue;
                if (part[w] == 0) ++n0;
                else if (part[w] == 1) ++n1;
            }
            if (n1 == 0 && n0 == 0) {
                int8_t side = (c.p0 <= c.p1) ? 0 : 1;
                part[v] = side;
                (side == 0 ? c.p0 : c.p1)++;
                c.s--;
            } else if (n1 == 0) {
                part[v] = 0; c.p0++; c.s--;
            } else if (n0 == 0) {
                part[v] = 1; c.p1++; c.s--;
            }
        }
        for (size_t i = 0; i < verts.size(); ++i) c.lab[i] = part[verts[i]];
        score(verts.size(), c);
    }

    // balanced cuts (smaller side >= 40 %) compete on separator size and balance; unbalanced ones only win when
    // no balanced cut exists and are then ranked by balance first
    static void score(size_t nv, Cand& c) {
        c.ok = c.p0 > 0 && c.p1 > 0;
        if (!c.ok) { c.cost = 1e300; return; }
        double mn = (double)std::min(c.p0, c.p1), tot = (double)nv;
        double imb = std::fabs((double)c.p0 - (double)c.p1) / tot;
        // the imbalance weight buys a perfectly balanced tree (12 levels instead of 15 at n = 79 841, every level
        // one launch per sweep) at no cost in fill: nnz(L) 8.41 M vs 8.84 M with the weaker weight 0.5
        if (mn >= 0.4 * tot) c.cost = (double)c.s * (1.0 + 2.0 * imb);
        else c.cost = 1e12 * (1.0 - mn / tot) + (double)c.s;
    }

    // choose threshold on an integer key: S = {key in [th, th+width)}, P0 = {key < th}, P1 = rest
    Cand best_threshold(const std::vector<int32_t>& verts, const std::vector<int32_t>& key, int32_t kmin,
                        int32_t kmax, int32_t width) {
        int32_t range = kmax - kmin + 1;
        std::vector<int64_t> pre(range + 1, 0);  // pre[k] = #{key - kmin < k}
        for (size_t i = 0; i < verts.size(); ++i) pre[key[i] - kmin + 1]++;
        for (int32_t k = 0; k < range; ++k) pre[k + 1] += pre[k];
        Cand best;
        int32_t best_th = -1;
        for (int32_t th = 1; th + width <= range - 1; ++th) {
            Cand c;
            c.p0 = pre[th];
            c.s = pre[th + width] - pre[th];
            c.p1 = (int64_t)verts.size() - pre[th + width];
            score(verts.size(), c);
            if (c.ok && c.cost < best.cost) { best = c; best_th = th; }
        }
        if (best_th >= 0) {
            best.lab.resize(verts.size());
            for (size_t i = 0; i < verts.size(); ++i) {
                int32_t k = key[i] - kmin;
                best.lab[i] = k < best_th ? 0 : (k < best_th + width ? 2 : 1);
            }
        }
        return best;
    }

    void dissect(std::vector<int32_t>& verts, Out& out, Scratch& sc, int depth) {
        std::vector<int32_t>& queue_ = sc.queue_;
        if ((int32_t)verts.size() <= leaf) { out.emit_block(verts); return; }
        int32_t st = ++cur_stamp;
        for (int32_t v : verts) stamp[v] = st;

        // connected components: small ones are packed into shared blocks, large ones dissected
        int32_t reached = bfs(verts[0], st, d1, verts, queue_);
        if (reached < (int32_t)verts.size()) {
            std::vector<std::vector<int32_t>> big;
            std::vector<int32_t> pending;
            for (int32_t v : verts) d2[v] = -1;
            for (int32_t v0 : verts) {
                if (d2[v0] >= 0) continue;
                // BFS of this component using d2 as the visited marker
                queue_.clear();
                queue_.push_back(v0);
                d2[v0] = 0;
                size_t head = 0;
                while (head < queue_.size()) {
                    int32_t v = queue_[head++];
                    for (int64_t p = g.ptr[v]; p < g.ptr[v + 1]; ++p) {
                        int32_t w = g.adj[p];
                        if (stamp[w] == st && d2[w] < 0) { d2[w] = 0; queue_.push_back(w); }
                    }
                }
                if ((int32_t)queue_.size() > leaf) {
                    big.emplace_back(queue_.begin(), queue_.end());
                } else {
                    if ((int32_t)(pending.size() + queue_.size()) > leaf) { out.emit_block(pending); pending.clear(); }
                    pending.insert(pending.end(), queue_.begin(), queue_.end());
                }
            }
            out.emit_block(pending);
            verts.clear(); verts.shrink_to_fit();
            for (auto& c : big) dissect(c, out, sc, depth);
            return;
        }
        // pseudo-peripheral pair (s, t)
        int32_t s = verts[0], ecc = -1;
        for (int it = 0; it < 5; ++it) {
            bfs(s, st, d1, verts, queue_);
            int32_t far = queue_.back(), e = d1[far];
            // among the last level pick the minimum-degree vertex
            int64_t bestdeg = INT64_MAX;
            for (size_t i = queue_.size(); i-- > 0;) {
                int32_t v = queue_[i];
                if (d1[v] != e) break;
                int64_t dg = g.ptr[v + 1] - g.ptr[v];
                if (dg < bestdeg) { bestdeg = dg; far = v; }
            }
            if (e <= ecc) break;
            ecc = e;
            s = far;
        }
        bfs(s, st, d1, verts, queue_);
        int32_t t = queue_.back();
        ecc = d1[t];
        if (ecc < 2) { out.emit_block(verts); return; }  // clique-like: keep as one dense supernode
        bfs(t, st, d2, verts, queue_);

        std::vector<int32_t> key(verts.size());
        // candidate A: BFS level set from s
        for (size_t i = 0; i < verts.size(); ++i) key[i] = d1[verts[i]];
        Cand ca = best_threshold(verts, key, 0, ecc, 1);
        // candidate B: bisector of (s, t), two layers thick before trimming
        int32_t kmin = INT32_MAX, kmax = INT32_MIN;
        for (size_t i = 0; i < verts.size(); ++i) {
            key[i] = d1[verts[i]] - d2[verts[i]];
            kmin = std::min(kmin, key[i]);
            kmax = std::max(kmax, key[i]);
        }
        Cand cb = best_threshold(verts, key, kmin, kmax, 2);
        if (ca.ok) trim(verts, st, ca);
        if (cb.ok) trim(verts, st, cb);
        Cand* best = nullptr;
        if (ca.ok) best = &ca;
        if (cb.ok && (!best || cb.cost < best->cost)) best = &cb;
        if (!best) { out.emit_block(verts); return; }

        std::vector<int32_t> p0, p1, sep;
        p0.reserve(best->p0); p1.reserve(best->p1); sep.reserve(best->s);
        for (size_t i = 0; i < verts.size(); ++i) {
            (best->lab[i] == 0 ? p0 : best->lab[i] == 1 ? p1 : sep).push_back(verts[i]);
        }
        verts.clear(); verts.shrink_to_fit();
        ca.lab.clear(); cb.lab.clear();
        if (depth < par_depth && p0.size() + p1.size() >= 2048) {
            Out o0, o1;
            std::thread left([&]() {
                Scratch s0;
                s0.queue_.reserve(p0.size());
                dissect(p0, o0, s0, depth + 1);
            });
            dissect(p1, o1, sc, depth + 1);
            left.join();
            out.append(o0);
            out.append(o1);
        } else {
            dissect(p0, out, sc, depth + 1);
            dissect(p1, out, sc, depth + 1);
        }
        out.emit_block(sep);
    }
};

// Offsets of per-supernode buffers that live from level(J) to level(parent(J)): every level gets one contiguous
// segment (members in level_sn order), placed first-fit into the gaps left by the segments that are already dead.
// Returns the pool size; lvl_off / lvl_size (optional) receive the segments.
int64_t pool_by_level(const Symbolic& S, const std::vector<int64_t>& size, std::vector<int64_t>& off,
                      std::vector<int64_t>* lvl_off, std::vector<int64_t>* lvl_size) {
    struct Seg { int64_t off, size; int32_t death; };
    std::vector<Seg> live;   // sorted by offset
    off.assign(S.nsn, 0);
    if (lvl_off) lvl_off->assign(S.nlevels, 0);
    if (lvl_size) lvl_size->assign(S.nlevels, 0);
    int64_t total = 0;
    for (int32_t l = 0; l < S.nlevels; ++l) {
        int64_t need = 0;
        int32_t death = l;
        for (int32_t p = S.level_ptr[l]; p < S.level_ptr[l + 1]; ++p) {
            const int32_t J = S.level_sn[p], P = S.sn_parent[J];
            need += size[J];
            if (P >= 0) death = std::max(death, S.sn_level[P]);
        }
        // a segment may be reused on level l only if its last reader ran on a level before l
        live.erase(std::remove_if(live.begin(), live.end(), [&](const Seg& g) { return g.death < l; }), live.end());
        int64_t at = 0;
        size_t ins = 0;
        for (; ins < live.size(); ++ins) {
            if (live[ins].off - at >= need) break;
            at = live[ins].off + live[ins].size;
        }
        if (need > 0) live.insert(live.begin() + ins, Seg{at, need, death});
        total = std::max(total, at + need);
        if (lvl_off) (*lvl_off)[l] = at;
        if (lvl_size) (*lvl_size)[l] = need;
        for (int32_t p = S.level_ptr[l]; p < S.level_ptr[l + 1]; ++p) {
            const int32_t J = S.level_sn[p];
            off[J] = at;
            at += size[J];
        }
    }
    return total;
}

}  // namespace

std::string analyze(int64_t n, const int64_t* Ecp, const int64_t* Eri, const double* Enz,
                    const int64_t* Acp, const int64_t* Ari, const double* Anz, int base,
                    const AnalyzeOptions& opt, Symbolic& S) {
    if (n <= 0) return "analyze: n must be positive";
    if (n >= (int64_t)1 << 31) return "analyze: n too large for 32-bit device indices";
    if (base != 0 && base != 1) return "analyze: index_base must be 0 or 1";
    S = Symbolic();
    S.n = n;
    const bool trace = getenv("DRE_TRACE_SYMBOLIC") != nullptr;
    auto tprev = std::chrono::steady_clock::now();
    auto phase = [&](const char* name) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[dre symbolic] %-28s %7.1f ms\n", name, std::chrono::duration<double, std::milli>(now - tprev).count());
        tprev = now;
    };
    const int64_t nnzE = Ecp[n] - base, nnzA = Acp[n] - base;
    if (Ecp[0] != base || Acp[0] != base) return "analyze: colptr[0] != index_base";

    // ---- union triplets in OLD indices (full pattern, both triangles as given), sorted by (column, row) ----
    struct Trip { int64_t key; double a, e; };
    std::vector<Trip> trip;
    trip.reserve(nnzE + nnzA);
    bool sorted_input = true;   // SparseMatrixCSC keeps row indices sorted per column: merge instead of sorting
    for (int64_t j = 0; j < n && sorted_input; ++j) {
        for (int64_t p = Ecp[j] - base + 1; p < Ecp[j + 1] - base; ++p)
            if (Eri[p] <= Eri[p - 1]) { sorted_input = false; break; }
        for (int64_t p = Acp[j] - base + 1; p < Acp[j + 1] - base && sorted_input; ++p)
            if (Ari[p] <= Ari[p - 1]) { sorted_input = false; break; }
    }
    if (sorted_input) {
        for (int64_t j = 0; j < n; ++j) {
            int64_t pe = Ecp[j] - base, pe1 = Ecp[j + 1] - base, pa = Acp[j] - base, pa1 = Acp[j + 1] - base;
            while (pe < pe1 || pa < pa1) {
                const int64_t ie = pe < pe1 ? Eri[pe] - base : INT64_MAX, ia = pa < pa1 ? Ari[pa] - base : INT64_MAX;
                const int64_t i = std::min(ie, ia);
                if (i < 0 || i >= n) return ie <= ia ? "analyze: E row index out of range" : "analyze: A row index out of range";
                Trip t{j * n + i, 0.0, 0.0};
                if (ie == i) t.e = Enz[pe++];
                if (ia == i) t.a = Anz[pa++];
                trip.push_back(t);
            }
        }
    } else {
        for (int64_t j = 0; j < n; ++j) {
            for (int64_t p = Ecp[j] - base; p < Ecp[j + 1] - base; ++p) {
                int64_t i = Eri[p] - base;
                if (i < 0 || i >= n) return "analyze: E row index out of range";
                trip.push_back({j * n + i, 0.0, Enz[p]});
            }
            for (int64_t p = Acp[j] - base; p < Acp[j + 1] - base; ++p) {
                int64_t i = Ari[p] - base;
                if (i < 0 || i >= n) return "analyze: A row index out of range";
                trip.push_back({j * n + i, Anz[p], 0.0});
            }
        }
        std::sort(trip.begin(), trip.end(), [](const Trip& x, const Trip& y) { return x.key < y.key; });
        size_t w = 0;
        for (size_t r = 0; r < trip.size(); ++r) {
            if (w > 0 && trip[w - 1].key == trip[r].key) {
                trip[w - 1].a += trip[r].a;
                trip[w - 1].e += trip[r].e;
            } else {
                trip[w++] = trip[r];
            }
        }
        trip.resize(w);
    }
    // symmetry check of values (the symmetric pencil is the scope of this build): (i, j) is looked up in column i
    {
        double amax = 0, emax = 0;
        for (auto& t : trip) { amax = std::max(amax, std::fabs(t.a)); emax = std::max(emax, std::fabs(t.e)); }
        std::vector<int64_t> colstart(n + 1, 0);
        for (auto& t : trip) colstart[t.key / n + 1]++;
        for (int64_t j = 0; j < n; ++j) colstart[j + 1] += colstart[j];
        auto find = [&](int64_t col, int64_t row) -> const Trip* {
            auto b = trip.begin() + colstart[col], e = trip.begin() + colstart[col + 1];
            auto it = std::lower_bound(b, e, col * n + row, [](const Trip& x, int64_t k) { return x.key < k; });
            return (it != e && it->key == col * n + row) ? &*it : nullptr;
        };
        for (auto& t : trip) {
            int64_t j = t.key / n, i = t.key % n;
            if (i == j) continue;
            const Trip* o = find(i, j);
            double oa = o ? o->a : 0.0, oe = o ? o->e : 0.0;
            if (std::fabs(t.a - oa) > 1e-12 * amax || std::fabs(t.e - oe) > 1e-12 * emax)
                return "analyze: E and A must be symmetric (nonsymmetric pencils are not supported by this build)";
        }
    }

    phase("triplets + symmetry check");
    // ---- adjacency graph (symmetrized pattern, no diagonal) ----
    Graph g;
    g.n = n;
    g.ptr.assign(n + 1, 0);
    for (auto& t : trip) {
        int64_t j = t.key / n, i = t.key % n;
        if (i == j) continue;
        g.ptr[i + 1]++;
        g.ptr[j + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
    g.adj.resize(g.ptr[n]);
    {
        std::vector<int64_t> pos(g.ptr.begin(), g.ptr.end() - 1);
        for (auto& t : trip) {
            int64_t j = t.key / n, i = t.key % n;
            if (i == j) continue;
            g.adj[pos[i]++] = (int32_t)j;
            g.adj[pos[j]++] = (int32_t)i;
        }
        // sort + unique each row, compact
        std::vector<int64_t> nptr(n + 1, 0);
        int64_t w = 0;
        for (int64_t i = 0; i < n; ++i) {
            int64_t b = g.ptr[i], e = g.ptr[i + 1];
            std::sort(g.adj.begin() + b, g.adj.begin() + e);
            int64_t start = w;
            for (int64_t p = b; p < e; ++p)
                if (w == start || g.adj[w - 1] != g.adj[p]) g.adj[w++] = g.adj[p];
            nptr[i + 1] = w;
        }
        g.adj.resize(w);
        g.ptr.swap(nptr);
    }

    phase("adjacency graph");
    // ---- nested dissection -> permutation + supernode blocks ----
    Dissector dis(g, std::max(4, opt.leaf_size));
    {
        std::vector<int32_t> all(n);
        std::iota(all.begin(), all.end(), 0);
        dis.run(all);
    }
    if ((int64_t)dis.out.order.size() != n) return "analyze: internal error (ordering incomplete)";
    S.perm = dis.out.order;
    S.iperm.assign(n, -1);
    for (int64_t k = 0; k < n; ++k) S.iperm[S.perm[k]] = (int32_t)k;
    // supernodes = dissection blocks, split into chains of at most max_snode columns (a chain link's
    // parent is the next link: the within-separator dependency becomes tree levels, so every front has
    // a single block column and the wide parts of its update parallelise over rows)
    {
        const int32_t cap = std::max(1, opt.max_snode);
        S.sn_first.clear();
        S.sn_first.push_back(0);
        int32_t b0 = 0;
        for (int32_t b1 : dis.out.block_end) {
            const int32_t size = b1 - b0;
            const int32_t nchunk = (size + cap - 1) / cap;
            const int32_t csz = (size + nchunk - 1) / nchunk;
            for (int32_t c = b0 + csz; c < b1; c += csz) S.sn_first.push_back(c);
            S.sn_first.push_back(b1);
            b0 = b1;
        }
        S.nsn = (int32_t)S.sn_first.size() - 1;
    }
    std::vector<int32_t> sn_of(n);
    for (int32_t J = 0; J < S.nsn; ++J)
        for (int32_t c = S.sn_first[J]; c < S.sn_first[J + 1]; ++c) sn_of[c] = J;

    phase("nested dissection");
    // ---- supernodal symbolic factorization ----
    S.sn_rowptr.assign(S.nsn + 1, 0);
    S.sn_parent.assign(S.nsn, -1);
    std::vector<std::vector<int32_t>> children(S.nsn);
    {
        std::vector<int32_t> mark(n, -1), rows;
        for (int32_t J = 0; J < S.nsn; ++J) {
            rows.clear();
            int32_t last = S.sn_first[J + 1] - 1;
            for (int32_t c = S.sn_first[J]; c <= last; ++c) {
                int32_t vo = S.perm[c];
                for (int64_t p = g.ptr[vo]; p < g.ptr[vo + 1]; ++p) {
                    int32_t i = S.iperm[g.adj[p]];
                    if (i > last && mark[i] != J) { mark[i] = J; rows.push_back(i); }
                }
            }
            for (int32_t ch : children[J]) {
                for (int64_t p = S.sn_rowptr[ch]; p < S.sn_rowptr[ch + 1]; ++p) {
                    int32_t i = S.sn_rows[p];
                    if (i > last && mark[i] != J) { mark[i] = J; rows.push_back(i); }
                }
            }
            std::sort(rows.begin(), rows.end());
            S.sn_rows.insert(S.sn_rows.end(), rows.begin(), rows.end());
            S.sn_rowptr[J + 1] = (int64_t)S.sn_rows.size();
            if (!rows.empty()) {
                int32_t P = sn_of[rows[0]];
                S.sn_parent[J] = P;
                children[P].push_back(J);
            }
        }
    }
    // children CSR
    S.child_ptr.assign(S.nsn + 1, 0);
    for (int32_t J = 0; J < S.nsn; ++J) S.child_ptr[J + 1] = S.child_ptr[J] + (int32_t)children[J].size();
    S.child_idx.resize(S.child_ptr[S.nsn]);
    for (int32_t J = 0; J < S.nsn; ++J)
        std::copy(children[J].begin(), children[J].end(), S.child_idx.begin() + S.child_ptr[J]);

    phase("symbolic factorization");
    // ---- levels by height (leaves = level 0; children precede parents in the numbering) ----
    {
        S.sn_level.assign(S.nsn, 0);
        int32_t maxl = 0;
        for (int32_t J = 0; J < S.nsn; ++J) {
            const int32_t P = S.sn_parent[J];
            if (P >= 0) S.sn_level[P] = std::max(S.sn_level[P], S.sn_level[J] + 1);
            maxl = std::max(maxl, S.sn_level[J]);
        }
        S.nlevels = maxl + 1;
        S.level_ptr.assign(S.nlevels + 1, 0);
        for (int32_t J = 0; J < S.nsn; ++J) S.level_ptr[S.sn_level[J] + 1]++;
        for (int32_t l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
        S.level_sn.resize(S.nsn);
        std::vector<int32_t> pos(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (int32_t J = 0; J < S.nsn; ++J) S.level_sn[pos[S.sn_level[J]]++] = J;
    }

    // ---- storage offsets, statistics ----
    S.panel_off.assign(S.nsn + 1, 0);
    S.linv_off.assign(S.nsn + 1, 0);
    S.upd_off.assign(S.nsn + 1, 0);
    S.rhs_off.assign(S.nsn, 0);
    {
        for (int32_t J = 0; J < S.nsn; ++J) {
            int64_t s = S.sn_size(J), u = S.sn_nrows(J), f = s + u;
            S.panel_off[J + 1] = S.panel_off[J] + f * s;
            S.linv_off[J + 1] = S.linv_off[J] + s * s;
            S.flops += 2.0 * ((double)s * s * s / 3.0 + (double)u * s * s + (double)u * u * s);
            S.max_front = std::max<int32_t>(S.max_front, (int32_t)f);
            S.max_sn = std::max<int32_t>(S.max_sn, (int32_t)s);
            S.sum_u += u;
        }
        S.nnz_L = S.panel_off[S.nsn];
    }
    // Update matrices (factorization) and update vectors (forward sweep) are pooled: the data of supernode J is
    // written on level(J) and last read when its parent gathers it on level(parent(J)), so a level's segment is
    // recycled as soon as every parent of its members has run.  Laid out without reuse the update matrices of a
    // 3D pencil outgrow the factor itself by an order of magnitude (n = 216 000: 6.6e8 entries against
    // nnz(L) = 8.9e7); pooled, two or three levels are alive at a time.
    {
        std::vector<int64_t> usz(S.nsn), rsz(S.nsn);
        for (int32_t J = 0; J < S.nsn; ++J) {
            const int64_t u = S.sn_nrows(J);
            usz[J] = u * u;
            rsz[J] = u;
        }
        std::vector<int64_t> uoff, roff;
        S.upd_total = pool_by_level(S, usz, uoff, &S.upd_level_off, &S.upd_level_size);
        S.rhs_total = pool_by_level(S, rsz, roff, nullptr, nullptr);
        for (int32_t J = 0; J < S.nsn; ++J) {
            S.upd_off[J] = uoff[J];
            S.rhs_off[J] = roff[J];
        }
        S.upd_off[S.nsn] = S.upd_total;
    }

    // ---- relative maps child struct row -> parent front local index ----
    S.relmap.assign(S.sn_rows.size(), -1);
    for (int32_t J = 0; J < S.nsn; ++J) {
        int32_t P = S.sn_parent[J];
        if (P < 0) continue;
        int32_t pf = S.sn_first[P], pl = S.sn_first[P + 1] - 1, ps = pl - pf + 1;
        int64_t q = S.sn_rowptr[P], qe = S.sn_rowptr[P + 1];
        for (int64_t p = S.sn_rowptr[J]; p < S.sn_rowptr[J + 1]; ++p) {
            int32_t i = S.sn_rows[p];
            if (i <= pl) {
                if (i < pf) return "analyze: internal error (child row below parent)";
                S.relmap[p] = i - pf;
            } else {
                while (q < qe && S.sn_rows[q] < i) ++q;
                if (q >= qe || S.sn_rows[q] != i) return "analyze: internal error (child row missing in parent)";
                S.relmap[p] = ps + (int32_t)(q - S.sn_rowptr[P]);
            }
        }
    }

    phase("levels, offsets, relmaps");
    // ---- permuted matrices: assembly scatter map (lower triangle) and full CSR ----
    {
        struct PT { int32_t r, c; double a, e; };
        std::vector<PT> pt;
        pt.reserve(trip.size() * 2);
        // symmetrized union pattern (an entry present in only one triangle is mirrored with its value)
        for (auto& t : trip) {
            int64_t jo = t.key / n, io = t.key % n;
            pt.push_back({S.iperm[io], S.iperm[jo], t.a, t.e});
        }
        S.csr_ptr.assign(n + 1, 0);
        for (auto& p : pt) S.csr_ptr[p.r + 1]++;
        for (int64_t i = 0; i < n; ++i) S.csr_ptr[i + 1] += S.csr_ptr[i];
        {   // counting sort by row, then the handful of entries of every row by column
            std::vector<PT> sorted(pt.size());
            std::vector<int32_t> pos(S.csr_ptr.begin(), S.csr_ptr.end() - 1);
            for (auto& p : pt) sorted[pos[p.r]++] = p;
            pt.swap(sorted);
            for (int64_t i = 0; i < n; ++i)
                std::sort(pt.begin() + S.csr_ptr[i], pt.begin() + S.csr_ptr[i + 1],
                          [](const PT& x, const PT& y) { return x.c < y.c; });
        }
        S.csr_col.resize(pt.size());
        S.csr_a.resize(pt.size());
        S.csr_e.resize(pt.size());
        for (size_t k = 0; k < pt.size(); ++k) {
            S.csr_col[k] = pt[k].c; S.csr_a[k] = pt[k].a; S.csr_e[k] = pt[k].e;
        }
        for (auto& p : pt) {
            if (p.r < p.c) continue;
            int32_t J = sn_of[p.c];
            int32_t fJ = S.sn_first[J], lJ = S.sn_first[J + 1] - 1, sJ = lJ - fJ + 1;
            int64_t f = S.front(J);
            int64_t lr;
            if (p.r <= lJ) {
                lr = p.r - fJ;
            } else {
                auto b = S.sn_rows.begin() + S.sn_rowptr[J], e = S.sn_rows.begin() + S.sn_rowptr[J + 1];
                auto it = std::lower_bound(b, e, p.r);
                if (it == e || *it != p.r) return "analyze: internal error (entry outside supernode structure)";
                lr = sJ + (it - b);
            }
            S.asm_dest.push_back(S.panel_off[J] + lr + (int64_t)(p.c - fJ) * f);
            S.asm_a.push_back(p.a);
            S.asm_e.push_back(p.e);
        }
    }
    phase("permuted CSR + scatter map");
    return "";
}

}  // namespace dre
