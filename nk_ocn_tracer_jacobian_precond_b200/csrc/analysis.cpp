// analysis.cpp -- host analysis phase of the B200 tracer-Jacobian solver.
//
// Replaces what SuperLU_DIST does inside the first pdgssvx* call of the reference
// drivers before numeric factorisation (src/SuperLU_brief_tree.txt:4-9: column
// ordering, symbfact, ddistribute; options at src/solve_ABglobal.c:332-334):
//
//   1. structure of A + A^T
//   2. fill-reducing nested-dissection ordering.  Every unknown of the reference's
//      operand is an ocean cell with integer coordinates (i,j,k) stored in the matrix
//      file (src/matrix.c:322-329), so the dissection is geometric: median cuts along
//      the coordinate direction that yields the smallest *actual* vertex separator
//      (the boundary of one half towards the other, computed from the graph, which
//      handles the periodic i direction of src/matrix.c:795-798 and the +-2 reach of
//      upwind3 without special cases).  Without coordinates a BFS level-structure
//      dissection is used.
//   3. assembly tree = relaxed supernodes of the elimination tree of that ordering
//      (supernodes_from_etree below: etree, postorder, column counts, path supernodes;
//      NKP_SUPERNODES=0 keeps the dissection nodes themselves as the fronts),
//      symbolic front structures by a bottom-up merge
//   4. memory plan (factor arena + two ping-pong pools of update matrices)
//   5. static task lists for every kernel launch of the numeric phase and solves
//
// The result depends only on the sparsity pattern and is reused across numeric
// refactorisations (BASELINE.json config 5).
#include "nkp_internal.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>

#include <unistd.h>

namespace nkp {

namespace {

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Graph {
    int n = 0;
    std::vector<int64_t> xadj;
    std::vector<int> adj;
};

// symmetric structure of A + A^T without the diagonal
void build_graph(int n, const int* rowptr, const int* colind, Graph& g) {
    g.n = n;
    std::vector<int64_t> cnt(n + 1, 0);
    for (int i = 0; i < n; i++)
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            int j = colind[p];
            if (j == i) continue;
            cnt[i + 1]++;
            cnt[j + 1]++;
        }
    for (int i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    std::vector<int> tmp(cnt[n]);
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int i = 0; i < n; i++)
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            int j = colind[p];
            if (j == i) continue;
            tmp[pos[i]++] = j;
            tmp[pos[j]++] = i;
        }
    // sort + unique each list (rows are independent), then compact
    std::vector<int64_t> ucnt(n + 1, 0);
#pragma omp parallel for schedule(dynamic, 4096)
    for (int i = 0; i < n; i++) {
        int64_t b = cnt[i], e = cnt[i + 1];
        std::sort(tmp.begin() + b, tmp.begin() + e);
        ucnt[i + 1] = std::unique(tmp.begin() + b, tmp.begin() + e) - (tmp.begin() + b);
    }
    g.xadj.assign(n + 1, 0);
    for (int i = 0; i < n; i++) g.xadj[i + 1] = g.xadj[i] + ucnt[i + 1];
    g.adj.resize(g.xadj[n]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) std::copy(tmp.begin() + cnt[i], tmp.begin() + cnt[i] + ucnt[i + 1], g.adj.begin() + g.xadj[i]);
}

struct TreeNode {
    std::vector<int> verts;     // original vertex ids in elimination order
    std::vector<int> children;
    int parent = -1;
};

struct Dissector {
    const Graph& g;
    const int* const* coords;   // may be null
    Options opt;
    std::vector<int> order;     // vertex list, partitioned in place
    std::vector<int> regid;     // region id per vertex
    std::vector<signed char> side;
    std::vector<TreeNode> nodes;    // result: the dissection tree in postorder
    std::atomic<int> next_rid{1};

    // The two halves of a dissection are independent and are built by OpenMP tasks.  Every call
    // returns its subtree with LOCAL node numbers (postorder) and the caller concatenates
    // [first half | second half | separator], so the numbering does not depend on scheduling: all
    // ranks of a multi-GPU run, and runs with any thread count, get the identical tree.  Tasks write
    // regid / side / order only for vertices of their own region; a neighbour outside the region is
    // only ever compared against this region's id, which nobody else writes.
    struct Sub {
        std::vector<TreeNode> nodes;
        std::vector<int> roots;
    };

    Dissector(const Graph& g_, const int* const* c, const Options& o) : g(g_), coords(c), opt(o) {
        order.resize(g.n);
        std::iota(order.begin(), order.end(), 0);
        regid.assign(g.n, 0);
        side.assign(g.n, 0);
    }

    int make_node(Sub& sub, int begin, int end, const std::vector<int>& children) {
        TreeNode nd;
        nd.verts.assign(order.begin() + begin, order.begin() + end);
        std::sort(nd.verts.begin(), nd.verts.end());
        nd.children = children;
        int id = (int)sub.nodes.size();
        sub.nodes.push_back(std::move(nd));
        for (int c : children) sub.nodes[c].parent = id;
        return id;
    }

    // append `b` to `a`, shifting b's local node numbers
    static void append(Sub& a, Sub&& b) {
        const int off = (int)a.nodes.size();
        a.nodes.reserve(a.nodes.size() + b.nodes.size());
        for (TreeNode& nd : b.nodes) {
            if (nd.parent >= 0) nd.parent += off;
            for (int& c : nd.children) c += off;
            a.nodes.push_back(std::move(nd));
        }
        for (int r : b.roots) a.roots.push_back(r + off);
    }

    // Evaluate the split "side[v] in {0,1}" of region rid: count the one-sided separators.
    // Returns sizes of S_a (side-0 vertices touching side 1) and S_b.
    void count_sep(int begin, int end, int rid, int& sa, int& sb) const {
        sa = sb = 0;
        for (int p = begin; p < end; p++) {
            int v = order[p];
            int sv = side[v];
            bool touch = false;
            for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
                int u = g.adj[q];
                if (regid[u] == rid && side[u] != sv) {
                    touch = true;
                    break;
                }
            }
            if (touch) (sv == 0 ? sa : sb)++;
        }
    }

    struct Cand {
        double score = 1e300;
        int dim = -1;
        int cut = 0;       // side 0: coord < cut
        int take = 0;      // 0: separator from side 0, 1: from side 1
        int sep = 0;
    };

    void try_dim(int begin, int end, int rid, int d, Cand& best) {
        const int* x = coords[d];
        int nv = end - begin;
        std::vector<int> scratch(nv);
        for (int p = 0; p < nv; p++) scratch[p] = x[order[begin + p]];
        std::nth_element(scratch.begin(), scratch.begin() + nv / 2, scratch.end());
        int med = scratch[nv / 2];
        int lo = *std::min_element(scratch.begin(), scratch.end());
        int hi = *std::max_element(scratch.begin(), scratch.end());
        if (lo == hi) return;
        int cut = med;
        if (cut == lo) cut = lo + 1;  // side 0 must be non-empty
        int n0 = 0;
        for (int p = begin; p < end; p++) {
            int v = order[p];
            side[v] = (x[v] < cut) ? 0 : 1;
            n0 += (side[v] == 0);
        }
        int n1 = nv - n0;
        if (n0 == 0 || n1 == 0) return;
        int sa, sb;
        count_sep(begin, end, rid, sa, sb);
        for (int take = 0; take < 2; take++) {
            int sep = take == 0 ? sa : sb;
            int r0 = n0 - (take == 0 ? sa : 0);
            int r1 = n1 - (take == 1 ? sb : 0);
            double minfrac = (double)std::min(r0, r1) / nv;
            double score = (double)sep * (1.0 + 4.0 * std::max(0.0, 0.30 - minfrac)) + 1e-3 * std::abs(r0 - r1);
            if (score < best.score) {
                best.score = score;
                best.dim = d;
                best.cut = cut;
                best.take = take;
                best.sep = sep;
            }
        }
    }

    // BFS level-structure split when no coordinates are available
    bool bfs_split(int begin, int end, int rid) {
        int nv = end - begin;
        // distance labels: one n-sized array per thread, entries reset to -1 after use (no task
        // scheduling point inside this function, so a thread runs one bfs_split at a time)
        static thread_local std::vector<int> dist;
        if (dist.size() != (size_t)g.n) dist.assign(g.n, -1);
        auto bfs = [&](int src, std::vector<int>& out) {
            out.clear();
            out.push_back(src);
            dist[src] = 0;
            for (size_t h = 0; h < out.size(); h++) {
                int v = out[h];
                for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
                    int u = g.adj[q];
                    if (regid[u] == rid && dist[u] < 0) {
                        dist[u] = dist[v] + 1;
                        out.push_back(u);
                    }
                }
            }
        };
        std::vector<int> reach, reach2;
        bfs(order[begin], reach);
        int far = reach.back();
        for (int v : reach) dist[v] = -1;
        if ((int)reach.size() < nv) {
            // disconnected: reached component vs the rest, empty separator
            for (int p = begin; p < end; p++) side[order[p]] = 1;
            for (int v : reach) side[v] = 0;
            return true;
        }
        bfs(far, reach2);
        int maxd = dist[reach2.back()];
        if (maxd < 2) {
            for (int v : reach2) dist[v] = -1;
            return false;
        }
        std::vector<int> cntl(maxd + 1, 0);
        for (int v : reach2) cntl[dist[v]]++;
        // choose the smallest level whose cumulative position lies in the middle 40 %
        int best = -1;
        int cum = 0;
        for (int l = 0; l <= maxd; l++) {
            int before = cum;
            cum += cntl[l];
            if (l == 0 || l == maxd) continue;
            double fb = (double)before / nv, fa = (double)(nv - cum) / nv;
            if (fb < 0.25 || fa < 0.25) continue;
            if (best < 0 || cntl[l] < cntl[best]) best = l;
        }
        if (best < 0) best = maxd / 2 == 0 ? 1 : maxd / 2;
        for (int v : reach2) {
            side[v] = dist[v] < best ? 0 : (dist[v] > best ? 1 : 2);
            dist[v] = -1;
        }
        return true;
    }

    Sub build(int begin, int end) {
        Sub out;
        int nv = end - begin;
        if (nv == 0) return out;
        if (nv <= opt.leaf) {
            out.roots.push_back(make_node(out, begin, end, {}));
            return out;
        }
        int rid = next_rid.fetch_add(1);
        for (int p = begin; p < end; p++) regid[order[p]] = rid;

        bool ok = false;
        if (coords) {
            Cand best;
            for (int d = 0; d < 3; d++)
                if (coords[d]) try_dim(begin, end, rid, d, best);
            if (best.dim >= 0) {
                const int* x = coords[best.dim];
                for (int p = begin; p < end; p++) {
                    int v = order[p];
                    side[v] = (x[v] < best.cut) ? 0 : 1;
                }
                // mark the separator (one-sided boundary)
                std::vector<int> sepv;
                for (int p = begin; p < end; p++) {
                    int v = order[p];
                    if (side[v] != best.take) continue;
                    for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
                        int u = g.adj[q];
                        if (regid[u] == rid && side[u] == 1 - best.take) {
                            sepv.push_back(v);
                            break;
                        }
                    }
                }
                for (int v : sepv) side[v] = 2;
                ok = true;
            }
        }
        if (!ok) ok = bfs_split(begin, end, rid);
        if (!ok) {
            out.roots.push_back(make_node(out, begin, end, {}));
            return out;
        }
        // partition order[begin:end) into [side0 | side1 | sep]
        int n0 = 0, n1 = 0, n2 = 0;
        for (int p = begin; p < end; p++) {
            int sv = side[order[p]];
            n0 += sv == 0;
            n1 += sv == 1;
            n2 += sv == 2;
        }
        if ((n0 == 0 && n2 == 0) || (n1 == 0 && n2 == 0) || n2 == nv) {
            out.roots.push_back(make_node(out, begin, end, {}));
            return out;
        }
        {
            std::vector<int> tmp(order.begin() + begin, order.begin() + end);
            int p0 = begin, p1 = begin + n0, p2 = begin + n0 + n1;
            for (int v : tmp) {
                int sv = side[v];
                if (sv == 0) order[p0++] = v;
                else if (sv == 1) order[p1++] = v;
                else order[p2++] = v;
            }
        }
        Sub second;
        const bool par = std::min(n0, n1) > 20000;   // small halves are not worth a task
#pragma omp task shared(out) if (par)
        out = build(begin, begin + n0);
#pragma omp task shared(second) if (par)
        second = build(begin + n0, begin + n0 + n1);
#pragma omp taskwait
        append(out, std::move(second));
        if (n2 == 0) return out;
        const std::vector<int> child_roots = out.roots;
        out.roots.assign(1, make_node(out, begin + n0 + n1, end, child_roots));
        return out;
    }

    void run(std::vector<int>& roots_out) {
        Sub all;
#pragma omp parallel
#pragma omp single
        all = build(0, g.n);
        nodes = std::move(all.nodes);
        roots_out = std::move(all.roots);
    }
};

// ---- on-disk cache of the ordering ------------------------------------
// The nested dissection is the expensive, purely pattern-dependent part of the analysis (about two
// thirds of it at gx1v6-shape).  The reference refactors from scratch on every process start
// (src/solve_ABglobal.c:350-353); with a cache directory set (nkp_set_analysis_cache or the
// NKP_ANALYSIS_CACHE environment variable) the ordering and its assembly tree are stored under a key derived from
// the pattern, the coordinates and the ordering options, and later processes read it back.
// File: magic, key, n, nf, nroots, node sizes[nf], node parents[nf], roots[nroots], iperm[n]
// (= the vertices of the nodes in node order).  Any mismatch or short read means "recompute".

std::string g_cache_dir;
bool g_cache_dir_set = false;

uint64_t fnv1a(uint64_t h, const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < bytes; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

uint64_t ordering_key(int n, const int* rowptr, const int* colind, const int* const* coords, const Options& opt) {
    uint64_t h = 1469598103934665603ull;
    const int tag[6] = {0x4e4b5032 /* format 2 */, n, opt.leaf, opt.period_i, opt.etree_supernodes, opt.relax_small};
    h = fnv1a(h, tag, sizeof tag);
    h = fnv1a(h, &opt.relax_frac, sizeof opt.relax_frac);
    h = fnv1a(h, rowptr, sizeof(int) * ((size_t)n + 1));
    h = fnv1a(h, colind, sizeof(int) * (size_t)rowptr[n]);
    for (int d = 0; d < 3; d++) {
        const int present = coords && coords[d] ? 1 : 0;
        h = fnv1a(h, &present, sizeof present);
        if (present) h = fnv1a(h, coords[d], sizeof(int) * (size_t)n);
    }
    return h;
}

std::string cache_dir() {
    if (g_cache_dir_set) return g_cache_dir;
    const char* e = getenv("NKP_ANALYSIS_CACHE");
    return e ? std::string(e) : std::string();
}

std::string cache_path(const std::string& dir, uint64_t key) {
    char name[64];
    snprintf(name, sizeof name, "/nkp_order_%016llx.bin", (unsigned long long)key);
    return dir + name;
}

const char CACHE_MAGIC[8] = {'N', 'K', 'P', 'O', 'R', 'D', '1', 0};

bool load_ordering(const std::string& path, uint64_t key, int n, std::vector<TreeNode>& nodes, std::vector<int>& roots) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    bool ok = false;
    char magic[8];
    uint64_t k = 0;
    int hdr[3] = {0, 0, 0};
    std::vector<int> sizes, parents, iperm;
    do {
        if (fread(magic, 1, 8, f) != 8 || memcmp(magic, CACHE_MAGIC, 8) != 0) break;
        if (fread(&k, sizeof k, 1, f) != 1 || k != key) break;
        if (fread(hdr, sizeof(int), 3, f) != 3 || hdr[0] != n || hdr[1] <= 0 || hdr[1] > n || hdr[2] <= 0 || hdr[2] > hdr[1]) break;
        const int nf = hdr[1], nr = hdr[2];
        sizes.resize(nf);
        parents.resize(nf);
        roots.resize(nr);
        iperm.resize(n);
        if (fread(sizes.data(), sizeof(int), nf, f) != (size_t)nf) break;
        if (fread(parents.data(), sizeof(int), nf, f) != (size_t)nf) break;
        if (fread(roots.data(), sizeof(int), nr, f) != (size_t)nr) break;
        if (fread(iperm.data(), sizeof(int), n, f) != (size_t)n) break;
        // consistency: sizes sum to n, parents come later (postorder), iperm is a permutation
        int64_t tot = 0;
        bool good = true;
        for (int t = 0; t < nf && good; t++) {
            good = sizes[t] > 0 && (parents[t] == -1 || (parents[t] > t && parents[t] < nf));
            tot += sizes[t];
        }
        if (!good || tot != n) break;
        std::vector<char> seen(n, 0);
        for (int i = 0; i < n && good; i++) {
            good = iperm[i] >= 0 && iperm[i] < n && !seen[iperm[i]];
            if (good) seen[iperm[i]] = 1;
        }
        for (int r : roots) good = good && r >= 0 && r < nf && parents[r] == -1;
        if (!good) break;
        nodes.assign(nf, TreeNode());
        int pos = 0;
        for (int t = 0; t < nf; t++) {
            nodes[t].verts.assign(iperm.begin() + pos, iperm.begin() + pos + sizes[t]);
            pos += sizes[t];
            nodes[t].parent = parents[t];
            if (parents[t] >= 0) nodes[parents[t]].children.push_back(t);   // creation order == index order
        }
        ok = true;
    } while (false);
    fclose(f);
    return ok;
}

void save_ordering(const std::string& path, uint64_t key, int n, const std::vector<TreeNode>& nodes, const std::vector<int>& roots) {
    char tmp[32];
    snprintf(tmp, sizeof tmp, ".tmp%ld", (long)getpid());
    const std::string tpath = path + tmp;
    FILE* f = fopen(tpath.c_str(), "wb");
    if (!f) return;
    const int nf = (int)nodes.size();
    const int hdr[3] = {n, nf, (int)roots.size()};
    std::vector<int> sizes(nf), parents(nf);
    for (int t = 0; t < nf; t++) {
        sizes[t] = (int)nodes[t].verts.size();
        parents[t] = nodes[t].parent;
    }
    bool ok = fwrite(CACHE_MAGIC, 1, 8, f) == 8 && fwrite(&key, sizeof key, 1, f) == 1 && fwrite(hdr, sizeof(int), 3, f) == 3 &&
              fwrite(sizes.data(), sizeof(int), nf, f) == (size_t)nf && fwrite(parents.data(), sizeof(int), nf, f) == (size_t)nf &&
              fwrite(roots.data(), sizeof(int), roots.size(), f) == roots.size();
    for (int t = 0; t < nf && ok; t++)
        ok = fwrite(nodes[t].verts.data(), sizeof(int), nodes[t].verts.size(), f) == nodes[t].verts.size();
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tpath.c_str(), path.c_str()) == 0;   // atomic: concurrent ranks write the same content
    if (!ok) remove(tpath.c_str());
}

}  // namespace

void set_analysis_cache_dir(const char* dir) {
    g_cache_dir_set = true;
    g_cache_dir = dir ? dir : "";
}


namespace {

// ---- assembly tree from the elimination tree of the dissection ordering ---------------------------------------
// The dissection tree is a good ORDERING but a poor assembly tree for this operator family: ocean subdomains cut by
// a coordinate plane are often disconnected (basins behind ridges, land), and a separator treated as one dense front
// then carries the union of the boundaries of all the pieces it touches -- measured at gx3v7-shape: 2.08x the flops
// and 1.44x the factor entries that the same ordering needs (independent symbolic factorisation,
// oracle/plan_sim.cpp::nkp_true_colcounts).  So the fronts are re-derived the textbook way: elimination tree of
// pattern(A + A^T) in the dissection numbering (Liu), postorder, column counts (Gilbert, Ng & Peyton 1994: skeleton
// leaves + least common ancestors, O(nnz alpha)), supernodes = maximal paths of the tree whose columns share their
// structure, relaxed: a path is joined with its parent path while the explicit zeros this stores stay below
// `relax_frac` of the joined panel (or the joined panel has at most `relax_small` columns).  The result replaces
// `nodes` (same TreeNode format, postorder, verts in elimination order), so everything downstream -- boundary sets,
// levels, memory plan, task lists, the multi-GPU mapping, the ordering cache -- is unchanged.
void supernodes_from_etree(const Graph& g, std::vector<TreeNode>& nodes, std::vector<int>& roots, const Options& opt) {
    const int n = g.n;
    // numbering of the dissection tree: nodes in index order (postorder), vertices in stored order
    std::vector<int> num((size_t)n), inv((size_t)n);
    {
        int next = 0;
        for (const TreeNode& t : nodes)
            for (int v : t.verts) {
                num[v] = next;
                inv[next] = v;
                next++;
            }
    }
    // elimination tree (path compression over the "virtual ancestor")
    std::vector<int> parent((size_t)n, -1);
    {
        std::vector<int> anc((size_t)n, -1);
        for (int i = 0; i < n; i++) {
            const int v = inv[i];
            for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
                int k = num[g.adj[q]];
                while (k != -1 && k < i) {
                    const int nx = anc[k];
                    anc[k] = i;
                    if (nx == -1) parent[k] = i;
                    k = nx;
                }
            }
        }
    }
    // postorder (children in increasing order: the highest-numbered child directly precedes its parent)
    std::vector<int> post((size_t)n), ipost((size_t)n);
    {
        std::vector<int> head((size_t)n, -1), next((size_t)n, -1), stack;
        std::vector<int> root_list;
        for (int j = n - 1; j >= 0; j--) {   // reversed, so that the lists come out ascending
            if (parent[j] < 0) {
                root_list.push_back(j);
                continue;
            }
            next[j] = head[parent[j]];
            head[parent[j]] = j;
        }
        std::reverse(root_list.begin(), root_list.end());
        int k = 0;
        for (int r : root_list) {
            stack.push_back(r);
            while (!stack.empty()) {
                const int v = stack.back();
                const int c = head[v];
                if (c == -1) {
                    stack.pop_back();
                    post[k] = v;
                    ipost[v] = k;
                    k++;
                } else {
                    head[v] = next[c];
                    stack.push_back(c);
                }
            }
        }
    }
    // final numbering fin = ipost o num; etree in it (a postordered tree: every subtree is a contiguous range)
    std::vector<int> fin((size_t)n), finv((size_t)n), par((size_t)n, -1);
    for (int v = 0; v < n; v++) {
        fin[v] = ipost[num[v]];
        finv[fin[v]] = v;
    }
    for (int j = 0; j < n; j++)
        if (parent[j] >= 0) par[ipost[j]] = ipost[parent[j]];
    // column counts (diagonal included)
    std::vector<int64_t> cc((size_t)n, 0);
    std::vector<int> first((size_t)n, -1);   // first (lowest) descendant: the subtree of j is the range [first[j], j]
    {
        std::vector<int> maxfirst((size_t)n, -1), prevleaf((size_t)n, -1), anc((size_t)n);
        for (int k = 0; k < n; k++) {
            cc[k] = first[k] == -1 ? 1 : 0;   // leaf of the tree
            for (int j = k; j != -1 && first[j] == -1; j = par[j]) first[j] = k;
        }
        for (int i = 0; i < n; i++) anc[i] = i;
        for (int j = 0; j < n; j++) {
            if (par[j] != -1) cc[par[j]]--;
            const int v = finv[j];
            for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
                const int i = fin[g.adj[q]];
                if (i <= j || first[j] <= maxfirst[i]) continue;   // j is not a leaf of the row subtree of i
                maxfirst[i] = first[j];
                const int jprev = prevleaf[i];
                prevleaf[i] = j;
                cc[j]++;
                if (jprev != -1) {   // subsequent leaf: the overlap with the previous one ends at their common ancestor
                    int q2 = jprev;
                    while (q2 != anc[q2]) q2 = anc[q2];
                    for (int s2 = jprev; s2 != q2;) {
                        const int sp = anc[s2];
                        anc[s2] = q2;
                        s2 = sp;
                    }
                    cc[q2]--;
                }
            }
            if (par[j] != -1) anc[j] = par[j];
        }
        for (int j = 0; j < n; j++)
            if (par[j] != -1) cc[par[j]] += cc[j];
    }
    // supernodes: left to right; the supernode that ends at column f - 1 is joined with the one starting at f when it
    // hangs under f and the padding is acceptable.  zeros[] = explicit zeros already accepted inside a supernode.
    struct SN {
        int first, s;
        int64_t r, zeros;
    };
    std::vector<SN> sn;
    // Whole subtrees of at most `leaf` columns become ONE front, as the leaves of the dissection were: inside them the
    // tree is a thicket of short paths (dozens of levels of tiny supernodes) whose exact structure saves a percent of
    // the flops and costs the sweeps a launch pair per level.  sub_root[f] = root of the maximal such subtree that
    // starts at column f.
    std::vector<int> sub_root((size_t)n, -1);
    for (int j = 0; j < n; j++) {
        const int size = j - first[j] + 1;
        if (size > opt.leaf || size < 2) continue;
        if (par[j] != -1 && par[j] - first[par[j]] + 1 <= opt.leaf) continue;   // not maximal
        sub_root[first[j]] = j;
    }
    for (int j = 0; j < n; j++) {
        SN cur{j, 1, cc[j] - 1, 0};
        if (sub_root[j] >= 0) {
            const int root = sub_root[j];
            const int64_t sz = root - j + 1;
            int64_t have = 0;
            for (int c = j; c <= root; c++) have += cc[c];
            cur = SN{j, (int)sz, cc[root] - 1, sz * (sz + 1) / 2 + sz * (cc[root] - 1) - have};
            j = root;
        }
        while (!sn.empty()) {
            const SN& pr = sn.back();
            if (par[pr.first + pr.s - 1] != cur.first) break;
            // joined: pr.s + cur.s pivots with boundary cur.r; the columns of pr get cur.s + cur.r - pr.r rows they lack
            const int64_t z = pr.zeros + cur.zeros + (int64_t)pr.s * (cur.s + cur.r - pr.r);
            const int64_t sj = (int64_t)pr.s + cur.s;
            const int64_t tot = sj * (sj + 1) / 2 + sj * cur.r;
            if (!(sj <= opt.relax_small || (double)z <= opt.relax_frac * (double)tot)) break;
            cur.first = pr.first;
            cur.s = (int)sj;
            cur.zeros = z;
            sn.pop_back();
        }
        sn.push_back(cur);   // the next column may still join it
    }
    // (a column arrives as a supernode of its own; every time it extends the supernode to its left, the supernode
    //  before that one is looked at again -- the accepted zeros are cumulative, so the bound holds for the result)
    const int ns = (int)sn.size();
    std::vector<int> sn_of((size_t)n);
    for (int q = 0; q < ns; q++)
        for (int a = 0; a < sn[q].s; a++) sn_of[sn[q].first + a] = q;
    std::vector<TreeNode> out((size_t)ns);
    roots.clear();
    for (int q = 0; q < ns; q++) {
        out[q].verts.resize((size_t)sn[q].s);
        for (int a = 0; a < sn[q].s; a++) out[q].verts[a] = finv[sn[q].first + a];
        const int p = par[sn[q].first + sn[q].s - 1];
        out[q].parent = p < 0 ? -1 : sn_of[p];
        if (p < 0) roots.push_back(q);
    }
    for (int q = 0; q < ns; q++)
        if (out[q].parent >= 0) out[out[q].parent].children.push_back(q);
    nodes.swap(out);
}

}  // namespace

int analyse(int n, const int* rowptr_in, const int* colind_in, const int* const coords[3],
            const Options& opt, Plan& plan, const int* rowmap) {
    plan = Plan();
    plan.n = n;
    plan.nnz = rowptr_in[n];
    plan.opt = opt;
    const int nb = opt.nb;
    if (opt.tn > nb || nb % opt.tn != 0) return -2;

    double t0 = now_s();
    // Static row permutation (rowperm.cpp): row i of A becomes row rowmap[i] of the matrix that is ordered and
    // factored.  Ordering, symbolic phase and the ordering cache see the permuted pattern; the scatter map at the
    // end is indexed by the caller's CRS, so the values never have to be reordered.
    std::vector<int> rp2, ci2;
    const int* rowptr = rowptr_in;
    const int* colind = colind_in;
    if (rowmap) {
        std::vector<int> inv((size_t)n, -1);
        for (int i = 0; i < n; i++) {
            if (rowmap[i] < 0 || rowmap[i] >= n || inv[rowmap[i]] >= 0) return -8;   // not a permutation
            inv[rowmap[i]] = i;
        }
        rp2.assign((size_t)n + 1, 0);
        ci2.resize((size_t)plan.nnz);
        for (int j = 0; j < n; j++) rp2[j + 1] = rp2[j] + (rowptr_in[inv[j] + 1] - rowptr_in[inv[j]]);
        for (int j = 0; j < n; j++)
            std::copy(colind_in + rowptr_in[inv[j]], colind_in + rowptr_in[inv[j] + 1], ci2.begin() + rp2[j]);
        rowptr = rp2.data();
        colind = ci2.data();
    }
    Graph g;
    build_graph(n, rowptr, colind, g);

    bool have_coords = coords && (coords[0] || coords[1] || coords[2]);
    Dissector ds(g, have_coords ? coords : nullptr, opt);
    std::vector<int> roots;
    const std::string cdir = cache_dir();
    uint64_t key = 0;
    if (!cdir.empty()) {
        key = ordering_key(n, rowptr, colind, have_coords ? coords : nullptr, opt);
        plan.order_cached = load_ordering(cache_path(cdir, key), key, n, ds.nodes, roots);
    }
    if (!plan.order_cached) {
        roots.clear();        // a rejected cache file may have left partial data behind
        ds.nodes.clear();
        ds.run(roots);
        if (opt.etree_supernodes) supernodes_from_etree(g, ds.nodes, roots, opt);
        if (!cdir.empty()) save_ordering(cache_path(cdir, key), key, n, ds.nodes, roots);
    }
    std::vector<TreeNode>& nodes = ds.nodes;
    int nf = (int)nodes.size();

    // nodes were created in postorder (children before parents)
    plan.fronts.assign(nf, Front());
    plan.perm.assign(n, -1);
    plan.iperm.assign(n, -1);
    {
        int next = 0;
        for (int t = 0; t < nf; t++) {
            Front& f = plan.fronts[t];
            f.first = next;
            f.s = (int)nodes[t].verts.size();
            f.parent = nodes[t].parent;
            f.nchild = (int)nodes[t].children.size();
            for (size_t c = 0; c < nodes[t].children.size(); c++) plan.fronts[nodes[t].children[c]].child_rank = (int)c;
            for (int v : nodes[t].verts) {
                plan.perm[v] = next;
                plan.iperm[next] = v;
                next++;
            }
        }
        if (next != n) return -3;
    }
    plan.roots = roots;
    plan.t_order = now_s() - t0;

    // ---- symbolic: boundary index sets --------------------------------------------------
    t0 = now_s();
    std::vector<int64_t> boff(nf + 1, 0);
    {
        std::vector<int> mark(n, -1);
        std::vector<int> cand;
        plan.bidx.clear();
        for (int t = 0; t < nf; t++) {
            Front& f = plan.fronts[t];
            int last = f.first + f.s - 1;
            cand.clear();
            for (int v : nodes[t].verts)
                for (int64_t q = g.xadj[v]; q < g.xadj[v + 1]; q++) {
                    int pu = plan.perm[g.adj[q]];
                    if (pu > last && mark[pu] != t) {
                        mark[pu] = t;
                        cand.push_back(pu);
                    }
                }
            for (int c : nodes[t].children) {
                const Front& fc = plan.fronts[c];
                for (int64_t q = 0; q < fc.r; q++) {
                    int pu = plan.bidx[fc.bidx_off + q];
                    if (pu > last && mark[pu] != t) {
                        mark[pu] = t;
                        cand.push_back(pu);
                    }
                }
            }
            std::sort(cand.begin(), cand.end());
            f.r = (int)cand.size();
            f.m = f.s + f.r;
            f.ld = f.m + (f.m & 1);
            f.bidx_off = (int64_t)plan.bidx.size();
            plan.bidx.insert(plan.bidx.end(), cand.begin(), cand.end());
            boff[t + 1] = (int64_t)plan.bidx.size();
            plan.max_front = std::max(plan.max_front, f.m);
        }
    }
    // roots must have empty boundaries
    for (int t : roots)
        if (plan.fronts[t].r != 0) return -4;

    // levels (depth from root); fronts are in postorder so parents come later
    for (int t = nf - 1; t >= 0; t--) {
        Front& f = plan.fronts[t];
        f.level = f.parent < 0 ? 0 : plan.fronts[f.parent].level + 1;
        plan.nlevels = std::max(plan.nlevels, f.level + 1);
    }

    // parent-local indices
    plan.rel.assign(plan.bidx.size(), -1);
    for (int t = 0; t < nf; t++) {
        Front& f = plan.fronts[t];
        f.rel_off = f.bidx_off;
        if (f.parent < 0) continue;
        const Front& p = plan.fronts[f.parent];
        const int* pb = plan.bidx.data() + p.bidx_off;
        int64_t q = 0;  // merge pointer into the parent's boundary
        for (int a = 0; a < f.r; a++) {
            int x = plan.bidx[f.bidx_off + a];
            int loc;
            if (x < p.first + p.s) {
                if (x < p.first) return -5;
                loc = x - p.first;
            } else {
                while (q < p.r && pb[q] < x) q++;
                if (q >= p.r || pb[q] != x) return -6;
                loc = p.s + (int)q;
            }
            plan.rel[f.rel_off + a] = loc;
        }
    }
    plan.t_symbolic = now_s() - t0;

    // ---- multi-GPU partition: rank-private subtrees + a shared top ----------------------------
    //  Subtrees are independent: their factorisation and their part of the sweeps
    // need no communication.  The heaviest candidates are split until there is one subtree per
    // rank; the fronts that were split form the top of the tree.  A top front is owned by the
    // owner of its heaviest child; update matrices / vectors of children owned elsewhere
    // travel over NCCL (Plan::xfers).
    t0 = now_s();
    const int P = std::max(1, opt.nranks);
    plan.rank = opt.rank;
    plan.nranks = P;
    plan.owner.assign(nf, 0);
    plan.is_top.assign(nf, 0);
    std::vector<double> fl(nf, 0.0), flsub(nf, 0.0);
    for (int t = 0; t < nf; t++) {
        double s = plan.fronts[t].s, r = plan.fronts[t].r;
        fl[t] = 2.0 / 3.0 * s * s * s + 2.0 * s * s * r + 2.0 * s * r * r;
        flsub[t] += fl[t];
        if (plan.fronts[t].parent >= 0) flsub[plan.fronts[t].parent] += flsub[t];
    }
    // Proportional mapping onto a binary hierarchy of rank ranges ([0,P) -> halves -> ... -> single ranks).
    // mapping(cands, [lo,hi)): the candidate subtrees handed to the range [lo,hi) are distributed over its two halves
    // by longest-processing-time (a half of k ranks works k times as fast); while the halves are unbalanced beyond
    // `split_tol`, the heaviest candidate is split: the split front becomes a TOP front factored jointly by all ranks
    // of [lo,hi), its children are new candidates.  Then each half maps its candidates the same way; a range of one
    // rank owns its candidates' whole subtrees.  Every group is a node of the hierarchy: at most P - 1 communicators,
    // nested, and the imbalance is bounded at every level instead of compounding.
    std::vector<int> top_lo(nf, 0), top_hi(nf, 0);
    {
        std::vector<int> cand_owner(nf, -1);
        struct Job {
            std::vector<int> cands;
            int lo, hi;
        };
        std::vector<Job> jobs;
        jobs.push_back(Job{std::vector<int>(roots.begin(), roots.end()), 0, P});
        while (!jobs.empty()) {
            Job job = std::move(jobs.back());
            jobs.pop_back();
            std::vector<int>& cand = job.cands;
            if (job.hi - job.lo == 1) {
                for (int t : cand) cand_owner[t] = job.lo;
                continue;
            }
            const int mid = job.lo + (job.hi - job.lo + 1) / 2;
            const double sp[2] = {(double)(mid - job.lo), (double)(job.hi - mid)};
            std::vector<int> bin[2];
            for (;;) {
                std::sort(cand.begin(), cand.end(), [&](int x, int y) { return flsub[x] > flsub[y] || (flsub[x] == flsub[y] && x < y); });
                double load[2] = {0.0, 0.0};
                bin[0].clear();
                bin[1].clear();
                for (int t : cand) {
                    const int b = (load[0] + flsub[t]) / sp[0] <= (load[1] + flsub[t]) / sp[1] ? 0 : 1;
                    bin[b].push_back(t);
                    load[b] += flsub[t];
                }
                const double mean = (load[0] + load[1]) / (sp[0] + sp[1]);
                const double imb = mean > 0 ? std::max(load[0] / sp[0], load[1] / sp[1]) / mean : 1.0;
                if (!bin[0].empty() && !bin[1].empty() && (imb <= 1.0 + opt.split_tol || (int)cand.size() >= opt.split_max)) break;
                int best = -1;
                for (size_t q = 0; q < cand.size(); q++)
                    if (!nodes[cand[q]].children.empty() && (best < 0 || flsub[cand[q]] > flsub[cand[best]])) best = (int)q;
                if (best < 0) break;
                const int t = cand[best];
                plan.is_top[t] = 1;
                top_lo[t] = job.lo;
                top_hi[t] = job.hi;
                cand.erase(cand.begin() + best);
                for (int c : nodes[t].children) cand.push_back(c);
            }
            jobs.push_back(Job{bin[0], job.lo, mid});
            jobs.push_back(Job{bin[1], mid, job.hi});
        }
        // owners: subtree members inherit top-down
        for (int t = nf - 1; t >= 0; t--) {
            if (plan.is_top[t]) continue;
            if (cand_owner[t] >= 0) plan.owner[t] = cand_owner[t];
            else plan.owner[t] = plan.owner[plan.fronts[t].parent];
        }
        // the rank-private subtrees: maximal subtrees without a top front
        plan.subtree_roots.clear();
        for (int t = 0; t < nf; t++)
            if (!plan.is_top[t] && (plan.fronts[t].parent < 0 || plan.is_top[plan.fronts[t].parent])) plan.subtree_roots.push_back(t);
        // first permuted index of every subtree (postorder => contiguous ranges)
        std::vector<int> lo(nf);
        for (int t = 0; t < nf; t++) {
            lo[t] = plan.fronts[t].first;
            for (int c : nodes[t].children) lo[t] = std::min(lo[t], lo[c]);
        }
        plan.subtree_lo.clear();
        for (int t : plan.subtree_roots) plan.subtree_lo.push_back(lo[t]);
    }
    // groups of the top fronts: the rank range that factors them (children come first in postorder)
    plan.group_of.assign(nf, -1);
    for (int t = 0; t < nf; t++) {
        if (!plan.is_top[t]) continue;
        std::vector<int> g;
        for (int r = top_lo[t]; r < top_hi[t]; r++) g.push_back(r);
        int gi = -1;
        for (size_t q = 0; q < plan.groups.size(); q++)
            if (plan.groups[q] == g) gi = (int)q;
        if (gi < 0) {
            gi = (int)plan.groups.size();
            plan.groups.push_back(g);
        }
        plan.group_of[t] = gi;
        plan.owner[t] = g[0];   // the member that publishes the front's part of the solution
    }
    auto in_group = [&](int t, int r) {
        const std::vector<int>& g = plan.groups[plan.group_of[t]];
        return std::binary_search(g.begin(), g.end(), r);
    };
    // stores(t): this rank keeps the factors of front t (its own subtrees; every top front whose group it is in)
    auto stores_r = [&](int t, int r) { return plan.is_top[t] ? in_group(t, r) : plan.owner[t] == r; };
    auto stores = [&](int t) { return stores_r(t, plan.rank); };
    auto mine = [&](int t) { return !plan.is_top[t] && plan.owner[t] == plan.rank; };
    auto ghost = [&](int t) {
        return !stores(t) && plan.fronts[t].parent >= 0 && stores(plan.fronts[t].parent);
    };

    // ---- memory plan (this rank's fronts only) ---------------------------------------------------
    const int64_t AL = 16;  // 128-byte alignment in doubles
    int64_t off = 0;
    double flops = 0;
    int64_t nnz_lu = 0;
    for (int t = 0; t < nf; t++) {
        Front& f = plan.fronts[t];
        flops += fl[t];
        nnz_lu += (int64_t)f.s * f.s + 2 * (int64_t)f.s * f.r;
        if (!stores(t)) {
            f.Loff = f.UToff = -1;
            continue;
        }
        plan.flops_local += plan.is_top[t] ? fl[t] / (double)plan.groups[plan.group_of[t]].size() : fl[t];
        plan.nnz_lu_local += (int64_t)f.s * f.s + 2 * (int64_t)f.s * f.r;
        f.Loff = off;
        off = align_up(off + (int64_t)f.ld * f.s, AL);
        f.UToff = off;
        off = align_up(off + (int64_t)f.ld * f.s, AL);
    }
    plan.factor_len = off;
    plan.flops = flops;
    plan.nnz_lu = nnz_lu;

    plan.levels.assign(plan.nlevels, LevelPlan());
    for (int l = 0; l < plan.nlevels; l++) plan.levels[l].level = l;
    for (int t = 0; t < nf; t++) {
        LevelPlan& L = plan.levels[plan.fronts[t].level];
        L.fronts.push_back(t);
        if (mine(t)) L.mine.push_back(t);
        if (stores(t)) L.stored.push_back(t);
        else if (ghost(t)) L.ghosts.push_back(t);
        // forward sweep: members of the parent's group that do not sweep this child receive its update vector
        int pa = plan.fronts[t].parent;
        if (pa >= 0 && plan.is_top[pa]) {
            std::vector<int> have;
            if (plan.is_top[t]) have = plan.groups[plan.group_of[t]];
            else have.push_back(plan.owner[t]);
            int k = 0;
            for (int p : plan.groups[plan.group_of[pa]]) {
                if (std::binary_search(have.begin(), have.end(), p)) continue;
                L.xfers.push_back((int)plan.xfers.size());
                plan.xfers.push_back(Xfer{t, have[k % have.size()], p, plan.fronts[t].level});
                k++;
            }
        }
    }
    // who publishes which part of the solution after the backward sweep
    for (size_t q = 0; q < plan.subtree_roots.size(); q++) {
        int t = plan.subtree_roots[q];
        plan.pub.push_back(PubRange{plan.subtree_lo[q], plan.fronts[t].first + plan.fronts[t].s, plan.owner[t]});
    }
    for (int t = 0; t < nf; t++)
        if (plan.is_top[t]) plan.pub.push_back(PubRange{plan.fronts[t].first, plan.fronts[t].first + plan.fronts[t].s, plan.owner[t]});
    // update-matrix pools: level l uses pool l & 1 (own fronts and ghost children)
    {
        int64_t len[2] = {0, 0};
        for (int l = 0; l < plan.nlevels; l++) {
            int64_t o = 0;
            for (int t : plan.levels[l].fronts) {
                Front& f = plan.fronts[t];
                if (!stores(t) && !ghost(t)) {
                    f.F22off = -1;
                    continue;
                }
                f.F22off = o;  // relative to the pool start, fixed up below
                o = align_up(o + (int64_t)f.r * f.r, AL);
            }
            plan.levels[l].f22_zero_len = o;
            len[l & 1] = std::max(len[l & 1], o);
        }
        plan.pool_len[0] = len[0];
        plan.pool_len[1] = len[1];
        plan.pool_off[0] = plan.factor_len;
        plan.pool_off[1] = plan.factor_len + len[0];
        plan.heap_len = plan.factor_len + len[0] + len[1];
        for (int l = 0; l < plan.nlevels; l++) {
            plan.levels[l].f22_zero_off = plan.pool_off[l & 1];
            for (int t : plan.levels[l].fronts)
                if (plan.fronts[t].F22off >= 0) plan.fronts[t].F22off += plan.pool_off[l & 1];
        }
    }
    // solve work vectors (own fronts and ghost children)
    {
        int64_t o = 0;
        for (int t = 0; t < nf; t++) {
            if (!stores(t) && !ghost(t)) {
                plan.fronts[t].woff = -1;
                continue;
            }
            plan.fronts[t].woff = o;
            o += plan.fronts[t].m;
        }
        plan.solve_pool_len = o;
    }

    // ---- scatter map: CRS entry -> heap offset ---------------------------------------------
    {
        // front of each permuted index
        std::vector<int> front_of(n);
        for (int t = 0; t < nf; t++)
            for (int a = 0; a < plan.fronts[t].s; a++) front_of[plan.fronts[t].first + a] = t;
        plan.scatter.resize(plan.nnz);
        bool bad = false;
#pragma omp parallel for schedule(dynamic, 8192) reduction(|| : bad)
        for (int i = 0; i < n; i++) {
            int pi = plan.perm[rowmap ? rowmap[i] : i];
            for (int p = rowptr_in[i]; p < rowptr_in[i + 1]; p++) {
                int pj = plan.perm[colind_in[p]];
                int t = front_of[std::min(pi, pj)];
                const Front& f = plan.fronts[t];
                if (!stores(t)) {
                    plan.scatter[p] = -1;
                    continue;
                }
                auto local = [&](int x) -> int {
                    if (x < f.first + f.s) return x - f.first;
                    const int* b = plan.bidx.data() + f.bidx_off;
                    const int* e = b + f.r;
                    const int* it = std::lower_bound(b, e, x);
                    if (it == e || *it != x) return -1;
                    return f.s + (int)(it - b);
                };
                int a = local(pi), b = local(pj);
                if (a < 0 || b < 0) {
                    bad = true;
                    continue;
                }
                plan.scatter[p] = front_entry(a, b, f.s, f.m, f.ld, nb, f.Loff, f.UToff, f.F22off);
            }
        }
        if (bad) return -7;
    }

    // ---- task lists ------------------------------------------------------------------------------
    // one Schur-update task: C[M x N] -= A[M x K] B[N x K]^T; tile0 counts the tiles of the launch it belongs to
    auto push_gemm = [&](int& tile0, int64_t A, int64_t B, int64_t C, int M, int N, int K, int lda, int ldb, int ldc, int skip) {
        if (M <= 0 || N <= 0) return;
        GemmTask gt;
        gt.Aoff = A;
        gt.Boff = B;
        gt.Coff = C;
        gt.M = M;
        gt.N = N;
        gt.K = K;
        gt.lda = lda;
        gt.ldb = ldb;
        gt.ldc = ldc;
        gt.skip = skip;
        gt.tile0 = tile0;
        gt.tiles_m = (M + opt.tm - 1) / opt.tm;
        gt.pad = 0;
        tile0 += gt.tiles_m * ((N + opt.tn - 1) / opt.tn);
        plan.gemm_tasks.push_back(gt);
        // useful (algorithmic) entries of this update
        double area = 0;
        if (skip == 0) area = (double)M * N;
        else
            for (int c0 = 0; c0 < N; c0 += nb) {
                int w = std::min(nb, N - c0);
                int first_row = skip == 1 ? c0 : std::min(c0 + nb, N);
                area += (double)(M - first_row) * w;
            }
        plan.gemm_flops += 2.0 * K * area;
    };
    // inner panel [k0, k0 + kb) of front f: diagonal block, both panel solves, and -- while the outer block
    // [ko0, ke) is not finished -- the narrow update of the rest of the outer block (K = kb).
    // Two-level blocking: G inner panels (nb columns each) form an outer block.  After every inner panel only
    // the rest of the outer block is updated; after the last one everything beyond the outer block receives ONE
    // wide update with K = width of the outer block, so the big trailing matrices are read and written once per
    // G panels.
    auto push_panel = [&](const Front& f, int k0, int ke, int& cta0, int& tile0) {
        const int kb = std::min(nb, f.s - k0), k1 = k0 + kb;
        const int64_t ld = f.ld;
        int64_t Dblk = f.Loff + k0 + (int64_t)k0 * ld;
        DiagTask d;
        d.Doff = Dblk;
        d.UTDoff = f.UToff + k0 + (int64_t)k0 * ld;
        d.ld = f.ld;
        d.kb = kb;
        plan.diag_tasks.push_back(d);
        const int below = f.m - k1;
        if (below <= 0) return;
        for (int which = 0; which < 2; which++) {
            TrsmTask tt;
            tt.Xoff = (which == 0 ? f.Loff : f.UToff) + k1 + (int64_t)k0 * ld;
            tt.Toff = Dblk;
            tt.ld = f.ld;
            tt.nrows = below;
            tt.kb = kb;
            tt.unit = which;
            tt.cta0 = cta0;
            tt.pad = 0;
            cta0 += (below + opt.trsm_rows - 1) / opt.trsm_rows;
            plan.trsm_tasks.push_back(tt);
        }
        if (k1 < ke) {
            // narrow: block columns / rows [k1, ke) of the outer block, all rows below
            int64_t Lpan = f.Loff + k1 + (int64_t)k0 * ld;    // L[k1.., k0:k1]
            int64_t UTpan = f.UToff + k1 + (int64_t)k0 * ld;  // U^T[k1.., k0:k1]
            push_gemm(tile0, Lpan, UTpan, f.Loff + k1 + (int64_t)k1 * ld, below, ke - k1, kb, f.ld, f.ld, f.ld, 1);
            push_gemm(tile0, UTpan, Lpan, f.UToff + k1 + (int64_t)k1 * ld, below, ke - k1, kb, f.ld, f.ld, f.ld, 2);
        }
    };
    auto push_add = [&](const Front& fc, const Front& fp, int& tile0) {
        AddTask a;
        a.Coff = fc.F22off;
        a.rel_off = fc.rel_off;
        a.Loff = fp.Loff;
        a.UToff = fp.UToff;
        a.F22off = fp.F22off;
        a.rc = fc.r;
        a.sp = fp.s;
        a.mp = fp.m;
        a.tile0 = tile0;
        a.tiles_m = (fc.r + opt.add_tile - 1) / opt.add_tile;
        a.ldp = fp.ld;
        tile0 += a.tiles_m * a.tiles_m;
        plan.add_tasks.push_back(a);
    };
    const int G = std::max(1, opt.outer);
    const int W = G * nb;   // outer block width of the rank-private fronts
    const int Wt = std::max(1, opt.top_outer) * nb;   // outer block width = distribution block of the top fronts
    for (int l = plan.nlevels - 1; l >= 0; l--) {
        LevelPlan& L = plan.levels[l];
        int maxs = 0;
        for (int t : L.mine) maxs = std::max(maxs, plan.fronts[t].s);
        L.nsteps = (maxs + nb - 1) / nb;

        // extend-add passes: children (level l+1) grouped by child_rank
        int npass = 0;
        for (int t : L.mine) npass = std::max(npass, plan.fronts[t].nchild);
        L.add_begin.assign(npass + 1, 0);
        L.add_tiles.assign(npass, 0);
        if (l + 1 < plan.nlevels) {
            for (int pass = 0; pass < npass; pass++) {
                L.add_begin[pass] = (int)plan.add_tasks.size();
                int tile0 = 0;
                for (int c : plan.levels[l + 1].fronts) {
                    const Front& fc = plan.fronts[c];
                    if (fc.child_rank != pass || fc.r == 0 || !mine(fc.parent)) continue;
                    push_add(fc, plan.fronts[fc.parent], tile0);
                }
                L.add_tiles[pass] = tile0;
            }
        }
        L.add_begin[npass] = (int)plan.add_tasks.size();

        L.diag_begin.assign(L.nsteps + 1, 0);
        L.trsm_begin.assign(L.nsteps + 1, 0);
        L.gemm_begin.assign(L.nsteps + 1, 0);
        L.trsm_ctas.assign(L.nsteps, 0);
        L.gemm_tiles.assign(L.nsteps, 0);
        for (int step = 0; step < L.nsteps; step++) {
            L.diag_begin[step] = (int)plan.diag_tasks.size();
            L.trsm_begin[step] = (int)plan.trsm_tasks.size();
            L.gemm_begin[step] = (int)plan.gemm_tasks.size();
            int cta0 = 0, tile0 = 0;
            int k0 = step * nb;
            for (int t : L.mine) {
                const Front& f = plan.fronts[t];
                if (k0 >= f.s) continue;
                const int k1 = k0 + std::min(nb, f.s - k0);
                const int ko0 = (step / G) * W;                     // first column of the outer block
                const int ke = std::min(f.s, ko0 + W);              // one past its last column
                push_panel(f, k0, ke, cta0, tile0);
                if (k1 == ke && f.m > k1) {
                    // wide: the outer block [ko0, ke) is completely factored
                    const int64_t ld = f.ld;
                    const int K = ke - ko0;
                    const int rest = f.m - ke;
                    const int trail_s = f.s - ke;
                    int64_t Lpan = f.Loff + ke + (int64_t)ko0 * ld;    // L[ke.., ko0:ke]
                    int64_t UTpan = f.UToff + ke + (int64_t)ko0 * ld;  // U^T[ke.., ko0:ke]
                    if (trail_s > 0) {
                        push_gemm(tile0, Lpan, UTpan, f.Loff + ke + (int64_t)ke * ld, rest, trail_s, K, f.ld, f.ld, f.ld, 1);
                        push_gemm(tile0, UTpan, Lpan, f.UToff + ke + (int64_t)ke * ld, rest, trail_s, K, f.ld, f.ld, f.ld, 2);
                    }
                    if (f.r > 0) {
                        int64_t Lb = f.Loff + f.s + (int64_t)ko0 * ld;
                        int64_t UTb = f.UToff + f.s + (int64_t)ko0 * ld;
                        push_gemm(tile0, Lb, UTb, f.F22off, f.r, f.r, K, f.ld, f.ld, f.r, 0);
                    }
                }
            }
            L.trsm_ctas[step] = cta0;
            L.gemm_tiles[step] = tile0;
        }
        L.diag_begin[L.nsteps] = (int)plan.diag_tasks.size();
        L.trsm_begin[L.nsteps] = (int)plan.trsm_tasks.size();
        L.gemm_begin[L.nsteps] = (int)plan.gemm_tasks.size();

        // ---- the top fronts of this level (identical list on every rank; tasks only where this rank takes part)
        for (int t : L.fronts) {
            if (!plan.is_top[t]) continue;
            const Front& f = plan.fronts[t];
            const std::vector<int>& grp = plan.groups[plan.group_of[t]];
            const int g = (int)grp.size();
            TopFront tf;
            tf.front = t;
            tf.group = plan.group_of[t];
            tf.member = stores(t) ? 1 : 0;
            const int nK = (f.s + Wt - 1) / Wt;
            auto f22_owner = [&](const Front& ff, const std::vector<int>& gg, int jc) {
                return gg[(((ff.s + Wt - 1) / Wt) + jc) % (int)gg.size()];
            };
            // update matrices of the children -> every member (complete copies; the extend-add is done by all)
            tf.cb_begin = (int)plan.top_child_bcasts.size();
            for (int c : nodes[t].children) {
                const Front& fc = plan.fronts[c];
                if (fc.r == 0) continue;
                if (!plan.is_top[c]) {
                    plan.top_child_bcasts.push_back(TopBcast{plan.owner[c], 0, tf.member ? fc.F22off : -1, (int64_t)fc.r * fc.r});
                } else {
                    const std::vector<int>& gc = plan.groups[plan.group_of[c]];
                    for (int jc = 0; jc * Wt < fc.r; jc++) {
                        const int w = std::min(Wt, fc.r - jc * Wt);
                        plan.top_child_bcasts.push_back(TopBcast{f22_owner(fc, gc, jc), 0,
                                                                 tf.member ? fc.F22off + (int64_t)jc * Wt * fc.r : -1,
                                                                 (int64_t)w * fc.r});
                    }
                }
            }
            tf.cb_end = (int)plan.top_child_bcasts.size();
            tf.block_begin = (int)plan.top_blocks.size();
            if (tf.member) {
                tf.f22_off = f.F22off;
                tf.f22_len = (int64_t)f.r * f.r;
                for (int c : nodes[t].children) {   // one pass per child: no two tasks of a launch touch the same entries
                    const Front& fc = plan.fronts[c];
                    tf.add_begin.push_back((int)plan.add_tasks.size());
                    int tile0 = 0;
                    if (fc.r > 0) push_add(fc, f, tile0);
                    tf.add_tiles.push_back(tile0);
                }
                tf.add_begin.push_back((int)plan.add_tasks.size());
            }
            for (int K = 0; K < nK; K++) {
                TopBlock tb;
                const int K0 = K * Wt, ke = std::min(f.s, K0 + Wt);
                tb.owner = grp[K % g];
                tb.step_begin = tb.step_end = (int)plan.top_steps.size();
                tb.next_begin = tb.next_end = tb.rest_begin = tb.rest_end = (int)plan.gemm_tasks.size();
                if (tf.member) {
                    const int64_t ld = f.ld;
                    // the factored column blocks, as flat ranges from (K0, K0) to (m - 1, ke - 1)
                    const int64_t cnt = (int64_t)(ke - 1 - K0) * ld + (f.m - K0);
                    tb.bl = TopBcast{tb.owner, 0, f.Loff + K0 + (int64_t)K0 * ld, cnt};
                    tb.bu = TopBcast{tb.owner, 0, f.UToff + K0 + (int64_t)K0 * ld, cnt};
                    if (tb.owner == plan.rank) {
                        for (int k0 = K0; k0 < ke; k0 += nb) {
                            TopStep ts;
                            ts.diag_begin = (int)plan.diag_tasks.size();
                            ts.trsm_begin = (int)plan.trsm_tasks.size();
                            ts.gemm_begin = (int)plan.gemm_tasks.size();
                            int cta0 = 0, tile0 = 0;
                            push_panel(f, k0, ke, cta0, tile0);
                            ts.diag_end = (int)plan.diag_tasks.size();
                            ts.trsm_end = (int)plan.trsm_tasks.size();
                            ts.gemm_end = (int)plan.gemm_tasks.size();
                            ts.trsm_ctas = cta0;
                            ts.gemm_tiles = tile0;
                            plan.top_steps.push_back(ts);
                        }
                        tb.step_end = (int)plan.top_steps.size();
                    }
                    // wide update from block K: first the block that is factored next, then the rest of what this rank owns
                    const int Kw = ke - K0;
                    auto wide_pivot_block = [&](int J, int& tile0) {
                        const int J0 = J * Wt, wJ = std::min(f.s, J0 + Wt) - J0;
                        int64_t Lrow = f.Loff + J0 + (int64_t)K0 * ld;     // L[J0.., K0:ke]
                        int64_t UTrow = f.UToff + J0 + (int64_t)K0 * ld;   // U^T[J0.., K0:ke]
                        push_gemm(tile0, Lrow, UTrow, f.Loff + J0 + (int64_t)J0 * ld, f.m - J0, wJ, Kw, f.ld, f.ld, f.ld, 1);
                        push_gemm(tile0, UTrow, Lrow, f.UToff + J0 + (int64_t)J0 * ld, f.m - J0, wJ, Kw, f.ld, f.ld, f.ld, 2);
                    };
                    tb.next_begin = (int)plan.gemm_tasks.size();
                    int tile0 = 0;
                    if (K + 1 < nK && grp[(K + 1) % g] == plan.rank) wide_pivot_block(K + 1, tile0);
                    tb.next_end = (int)plan.gemm_tasks.size();
                    tb.next_tiles = tile0;
                    tb.rest_begin = (int)plan.gemm_tasks.size();
                    tile0 = 0;
                    for (int J = K + 2; J < nK; J++)
                        if (grp[J % g] == plan.rank) wide_pivot_block(J, tile0);
                    for (int jc = 0; jc * Wt < f.r; jc++) {
                        if (f22_owner(f, grp, jc) != plan.rank) continue;
                        const int jc0 = jc * Wt, wj = std::min(Wt, f.r - jc0);
                        push_gemm(tile0, f.Loff + f.s + (int64_t)K0 * ld, f.UToff + f.s + jc0 + (int64_t)K0 * ld,
                                  f.F22off + (int64_t)jc0 * f.r, f.r, wj, Kw, f.ld, f.ld, f.r, 0);
                    }
                    tb.rest_end = (int)plan.gemm_tasks.size();
                    tb.rest_tiles = tile0;
                }
                plan.top_blocks.push_back(tb);
            }
            tf.block_end = (int)plan.top_blocks.size();
            L.tops.push_back((int)plan.top_fronts.size());
            plan.top_fronts.push_back(tf);
        }
    }

    // solve tasks grouped by level, deepest first
    for (int l = plan.nlevels - 1; l >= 0; l--) {
        LevelPlan& L = plan.levels[l];
        L.solve_begin = (int)plan.solve_tasks.size();
        for (int t : L.stored) {
            const Front& f = plan.fronts[t];
            SolveTask st;
            st.Loff = f.Loff;
            st.UToff = f.UToff;
            st.bidx_off = f.bidx_off;
            st.rel_off = f.rel_off;
            st.woff = f.woff;
            st.child_list = (int64_t)plan.solve_children.size();
            st.first = f.first;
            st.s = f.s;
            st.r = f.r;
            st.m = f.m;
            st.ld = f.ld;
            st.nchild = f.nchild;
            st.big = 0;
            for (int c : nodes[t].children) {
                const Front& fc = plan.fronts[c];
                SolveChild sc;
                sc.woff = fc.woff;
                sc.rel_off = fc.rel_off;
                sc.s = fc.s;
                sc.r = fc.r;
                plan.solve_children.push_back(sc);
            }
            plan.solve_tasks.push_back(st);
        }
        L.solve_end = (int)plan.solve_tasks.size();
        // split into single-CTA fronts and big (multi-CTA dataflow) fronts
        L.small_begin = (int)plan.solve_small.size();
        L.big_begin = (int)plan.big_fronts.size();
        int64_t part_slots = 0;
        L.inv_begin = (int)plan.inv_tasks.size();
        for (int q = L.solve_begin; q < L.solve_end; q++) {
            SolveTask& st = plan.solve_tasks[q];
            // every front keeps its 64 x 64 diagonal blocks inverted for the sweeps
            for (int k0 = 0; k0 < st.s; k0 += 64) {
                DiagTask d;
                d.Doff = st.Loff + k0 + (int64_t)k0 * st.ld;
                d.UTDoff = st.UToff + k0 + (int64_t)k0 * st.ld;
                d.ld = st.ld;
                d.kb = std::min(64, st.s - k0);
                plan.inv_tasks.push_back(d);
            }
            if ((int64_t)st.m * st.s >= opt.big_entries || st.m >= opt.big_rows) {
                st.big = 1;
                BigFront bf;
                bf.Loff = st.Loff;
                bf.UToff = st.UToff;
                bf.bidx_off = st.bidx_off;
                bf.woff = st.woff;
                bf.first = st.first;
                bf.s = st.s;
                bf.r = st.r;
                bf.m = st.m;
                bf.npiv = (st.s + 63) / 64;
                bf.nslab = bf.npiv + (st.r + 63) / 64;
                bf.pad0 = 0;
                bf.nchild = st.nchild;
                bf.child_list = st.child_list;
                bf.nchunk = ((st.r + 63) / 64 + BWD_CHUNK - 1) / BWD_CHUNK;
                bf.part_off = part_slots;
                bf.ld = st.ld;
                bf.clo_off = (int64_t)plan.child_lo.size();
                for (int c = 0; c < st.nchild; c++) {
                    const SolveChild& sc = plan.solve_children[st.child_list + c];
                    const int* rb = plan.rel.data() + sc.rel_off;
                    for (int j = 0; j < bf.nslab; j++) {
                        int row0 = j < bf.npiv ? 64 * j : st.s + 64 * (j - bf.npiv);
                        plan.child_lo.push_back((int)(std::lower_bound(rb, rb + sc.r, row0) - rb));
                    }
                }
                part_slots += (int64_t)bf.npiv * bf.nchunk;
                plan.big_fronts.push_back(bf);
            } else {
                plan.solve_small.push_back(st);
            }
        }
        L.small_end = (int)plan.solve_small.size();
        L.big_end = (int)plan.big_fronts.size();
        L.inv_end = (int)plan.inv_tasks.size();
        L.fwd_item_begin = (int)plan.big_fwd_items.size();
        L.bwd_item_begin = (int)plan.big_bwd_items.size();
        {
            int maxslab = 0, maxpiv = 0;
            for (int b = L.big_begin; b < L.big_end; b++) {
                maxslab = std::max(maxslab, plan.big_fronts[b].nslab);
                maxpiv = std::max(maxpiv, plan.big_fronts[b].npiv);
            }
            // forward: slabs in increasing order, interleaved over the fronts of the level
            for (int i = 0; i < maxslab; i++)
                for (int b = L.big_begin; b < L.big_end; b++)
                    if (i < plan.big_fronts[b].nslab) plan.big_fwd_items.push_back(BigItem{b, i});
            // backward: panels from the last one down, interleaved
            for (int d = 0; d < maxpiv; d++)
                for (int b = L.big_begin; b < L.big_end; b++)
                    if (d < plan.big_fronts[b].npiv) plan.big_bwd_items.push_back(BigItem{b, plan.big_fronts[b].npiv - 1 - d});
        }
        L.fwd_item_end = (int)plan.big_fwd_items.size();
        L.bwd_item_end = (int)plan.big_bwd_items.size();
        // rectangular (boundary) part of the backward sweep: independent (panel, row chunk) items
        plan.bwd_part_slots = std::max(plan.bwd_part_slots, part_slots);
        L.rect_item_begin = (int)plan.big_rect_items.size();
        for (int b = L.big_begin; b < L.big_end; b++) {
            const BigFront& bf = plan.big_fronts[b];
            for (int q = 0; q < bf.npiv * bf.nchunk; q++) plan.big_rect_items.push_back(BigItem{b, q});
        }
        L.rect_item_end = (int)plan.big_rect_items.size();
    }
    plan.t_plan = now_s() - t0;

    if (opt.verbose) {
        fprintf(stderr,
                "[nkp] analysis: n=%d nnz=%lld fronts=%d levels=%d max_front=%d nnz(LU)=%lld (%.2f GB) "
                "heap=%.2f GB flops=%.3e  t(order,symb,plan)=%.2f,%.2f,%.2f s\n",
                n, (long long)plan.nnz, nf, plan.nlevels, plan.max_front, (long long)plan.nnz_lu,
                plan.nnz_lu * 8e-9, plan.heap_len * 8e-9, plan.flops, plan.t_order, plan.t_symbolic, plan.t_plan);
    }
    if (opt.verbose && plan.nranks > 1) {
        std::vector<int> hist(plan.nranks + 1, 0);
        for (const auto& g : plan.groups) hist[g.size()]++;
        fprintf(stderr, "[nkp] partition: %d top fronts, %d groups (by size:", (int)plan.top_fronts.size(), (int)plan.groups.size());
        for (int q = 1; q <= plan.nranks; q++)
            if (hist[q]) fprintf(stderr, " %dx%d", hist[q], q);
        double tops = 0;
        for (int t = 0; t < nf; t++)
            if (plan.is_top[t]) tops += fl[t];
        fprintf(stderr, "), %.1f %% of the flops in top fronts, %d rank-private subtrees\n", 100.0 * tops / plan.flops,
                (int)plan.subtree_roots.size());
    }
    if (opt.verbose >= 2) {
        // per level: fronts, big fronts, largest pivot count / front size, factor bytes (L + U^T panels)
        for (int l = 0; l < plan.nlevels; l++) {
            const LevelPlan& L = plan.levels[l];
            int maxs = 0, maxm = 0;
            double bytes = 0;
            for (int t : L.mine) {
                const Front& f = plan.fronts[t];
                maxs = std::max(maxs, f.s);
                maxm = std::max(maxm, f.m);
                bytes += 8.0 * ((double)f.s * f.s + 2.0 * (double)f.r * f.s);
            }
            fprintf(stderr, "[nkp] level %2d: fronts %6d big %5d max_s %6d max_m %6d factor bytes %8.3f GB\n", l,
                    (int)L.mine.size(), L.big_end - L.big_begin, maxs, maxm, bytes * 1e-9);
        }
    }
    return 0;
}

}  // namespace nkp
