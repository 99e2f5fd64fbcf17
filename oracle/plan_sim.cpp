// plan_sim.cpp -- TEST INFRASTRUCTURE (oracle/): a CPU interpreter of the solver's
// static plan.  It executes exactly the task lists the CUDA kernels execute
// (nkp_internal.hpp: DiagTask / TrsmTask / GemmTask / AddTask / SolveTask) with plain
// loops, so the host analysis (ordering, symbolic structure, memory plan, scatter map,
// tile-skip rules) can be validated in a container without a GPU, and so that GPU
// results can be compared front-by-front when debugging.  It is never linked into the
// product library and never called by it.
//
// Algorithm restated: multifrontal LU with static pivoting as performed by
// SuperLU_DIST's pdgstrf for the reference (src/SuperLU_brief_tree.txt:11-15), followed
// by the forward/backward sweeps of pdgstrs (src/SuperLU_brief_tree.txt:17-18).
// Parity of this interpreter is pinned against scipy's SuperLU in tests/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../nk_ocn_tracer_jacobian_precond_b200/csrc/nkp_internal.hpp"

using namespace nkp;

namespace {

void sim_diag(double* H, const DiagTask& t, double tiny, int* nrepl) {
    double* D = H + t.Doff;
    int ld = t.ld, kb = t.kb;
    for (int k = 0; k < kb; k++) {
        double p = D[k + (int64_t)k * ld];
        if (std::fabs(p) < tiny) {
            p = (p < 0 ? -tiny : tiny);
            D[k + (int64_t)k * ld] = p;
            (*nrepl)++;
        }
        double inv = 1.0 / p;
        for (int i = k + 1; i < kb; i++) D[i + (int64_t)k * ld] *= inv;
        for (int j = k + 1; j < kb; j++) {
            double u = D[k + (int64_t)j * ld];
            for (int i = k + 1; i < kb; i++) D[i + (int64_t)j * ld] -= D[i + (int64_t)k * ld] * u;
        }
    }
    // U_kk^T into the UT diagonal block (lower triangle incl. diagonal)
    double* UD = H + t.UTDoff;
    for (int a = 0; a < kb; a++)
        for (int b = 0; b <= a; b++) UD[a + (int64_t)b * ld] = D[b + (int64_t)a * ld];
}

void sim_trsm(double* H, const TrsmTask& t) {
    double* X = H + t.Xoff;
    const double* T = H + t.Toff;
    int ld = t.ld, kb = t.kb;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < t.nrows; i++) {
        for (int j = 0; j < kb; j++) {
            double x = X[i + (int64_t)j * ld];
            if (t.unit == 0) {
                for (int p = 0; p < j; p++) x -= X[i + (int64_t)p * ld] * T[p + (int64_t)j * ld];
                x /= T[j + (int64_t)j * ld];
            } else {
                for (int p = 0; p < j; p++) x -= X[i + (int64_t)p * ld] * T[j + (int64_t)p * ld];
            }
            X[i + (int64_t)j * ld] = x;
        }
    }
}

void sim_gemm(double* H, const GemmTask& t, const Options& o) {
    const double* A = H + t.Aoff;
    const double* B = H + t.Boff;
    double* C = H + t.Coff;
    int tiles_n = (t.N + o.tn - 1) / o.tn;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int tj = 0; tj < tiles_n; tj++)
        for (int ti = 0; ti < t.tiles_m; ti++) {
            int rend = (ti + 1) * o.tm;
            if (getenv("NKP_SIM_NOSKIP")) rend = 1 << 30;
            if (t.skip == 1 && rend <= (tj * o.tn / o.nb) * o.nb) continue;
            if (t.skip == 2 && rend <= std::min((tj * o.tn / o.nb + 1) * o.nb, t.N)) continue;
            int i0 = ti * o.tm, i1 = std::min(t.M, i0 + o.tm);
            int j0 = tj * o.tn, j1 = std::min(t.N, j0 + o.tn);
            for (int j = j0; j < j1; j++)
                for (int p = 0; p < t.K; p++) {
                    double b = B[j + (int64_t)p * t.ldb];
                    const double* a = A + (int64_t)p * t.lda;
                    double* c = C + (int64_t)j * t.ldc;
                    for (int i = i0; i < i1; i++) c[i] -= a[i] * b;
                }
            if (getenv("NKP_SIM_POISON") && t.skip) {
                for (int j = j0; j < j1; j++) {
                    int lim = t.skip == 1 ? (j / o.nb) * o.nb : std::min((j / o.nb + 1) * o.nb, t.N);
                    for (int i = i0; i < i1 && i < lim; i++) C[i + (int64_t)j * t.ldc] = NAN;
                }
            }
        }
}

void sim_add(double* H, const AddTask& t, const int* rel, int nb) {
    const double* C = H + t.Coff;
    const int* rl = rel + t.rel_off;
    for (int b = 0; b < t.rc; b++)
        for (int a = 0; a < t.rc; a++) {
            int64_t dst = front_entry(rl[a], rl[b], t.sp, t.mp, t.ldp, nb, t.Loff, t.UToff, t.F22off);
            H[dst] += C[a + (int64_t)b * t.rc];
        }
}

// In-place inversion of the nb x nb diagonal blocks of the fronts swept by the multi-CTA dataflow
// kernels (k_invert_diag): the strictly lower part of the Larr diagonal block receives the strictly
// lower part of L_kk^-1 (unit diagonal implied), the lower part of the UTarr diagonal block
// receives (U_kk^-1)^T.
void sim_invert_diag(double* H, const DiagTask& t) {
    int ld = t.ld, kb = t.kb;
    double* D = H + t.Doff;
    double* UD = H + t.UTDoff;
    std::vector<double> X((size_t)kb * kb, 0.0), V((size_t)kb * kb, 0.0);
    for (int j = 0; j < kb; j++) {   // column j of L^-1
        X[j + (size_t)j * kb] = 1.0;
        for (int i = j + 1; i < kb; i++) {
            double acc = 0;
            for (int p = j; p < i; p++) acc += D[i + (int64_t)p * ld] * X[p + (size_t)j * kb];
            X[i + (size_t)j * kb] = -acc;
        }
    }
    for (int j = 0; j < kb; j++) {   // column j of U^-1 (upper), U(p,q) = UD[q + p*ld]
        V[j + (size_t)j * kb] = 1.0 / UD[j + (int64_t)j * ld];
        for (int i = j - 1; i >= 0; i--) {
            double acc = 0;
            for (int p = i + 1; p <= j; p++) acc += UD[p + (int64_t)i * ld] * V[p + (size_t)j * kb];
            V[i + (size_t)j * kb] = -acc / UD[i + (int64_t)i * ld];
        }
    }
    for (int j = 0; j < kb; j++)
        for (int i = j + 1; i < kb; i++) D[i + (int64_t)j * ld] = X[i + (size_t)j * kb];
    for (int b = 0; b < kb; b++)
        for (int a = b; a < kb; a++) UD[a + (int64_t)b * ld] = V[b + (size_t)a * kb];
}

}  // namespace

static int g_last_order_cached = 0;

// assembly-tree options from the environment (the same variables csrc/solver.cu reads)
static void tree_env(Options& opt) {
    if (getenv("NKP_SUPERNODES")) opt.etree_supernodes = atoi(getenv("NKP_SUPERNODES"));
    if (getenv("NKP_RELAX_FRAC")) opt.relax_frac = atof(getenv("NKP_RELAX_FRAC"));
    if (getenv("NKP_RELAX_SMALL")) opt.relax_small = atoi(getenv("NKP_RELAX_SMALL"));
}

extern "C" {

// ordering cache of the analysis (csrc/analysis.cpp): directory, and whether the last analysis used it
void nkp_sim_set_cache_dir(const char* dir) { set_analysis_cache_dir(dir); }
int nkp_sim_last_order_cached(void) { return g_last_order_cached; }

// Interprets the plans of `nranks` ranks in lockstep inside one process: every rank has its own
// heap / work vectors; Plan::xfers and the top-front broadcasts are carried out as memcpys.
// With nranks == 1 this is the plain single-GPU plan.  Returns the solution seen by rank 0.
// stats_out[0..7]: n_fronts, n_levels, max_front, nnz_lu, heap_len(rank 0), flops, tiny_pivots, analysis seconds
// part_out (may be null, 4 ints per rank): fronts owned, xfers as src, xfers as dst, local flops / 1e6
// rowmap / Rs / Cs (each may be null): static row permutation and row / column scalings of the factored matrix
// (csrc/rowperm.cpp, nkp_create_rowperm): heap gets Rs[i] a_ij Cs[j] at the position of (rowmap[i], j); the
// right-hand side enters as Rs[i] b[i] at row rowmap[i], the solution leaves as Cs[j] y[j].
static int sim_run_impl(int n, const int* rowptr, const int* colind, const double* val, const int* ci,
                        const int* cj, const int* ck, int nb, int leaf, const double* B, int nrhs,
                        double* X, double* stats_out, int* perm_out, int analysis_only, int nranks,
                        double* part_out, const int* rowmap, const double* Rs, const double* Cs) {
    Options opt;
    opt.nb = nb;
    opt.leaf = leaf;
    if (opt.tn > nb) opt.tn = nb;
    if (getenv("NKP_OUTER")) opt.outer = opt.top_outer = atoi(getenv("NKP_OUTER"));
    if (getenv("NKP_TOP_OUTER")) opt.top_outer = atoi(getenv("NKP_TOP_OUTER"));
    if (getenv("NKP_SPLIT_TOL")) opt.split_tol = atof(getenv("NKP_SPLIT_TOL"));
    if (getenv("NKP_SPLIT_MAX")) opt.split_max = atoi(getenv("NKP_SPLIT_MAX"));
    tree_env(opt);
    if (getenv("NKP_SIM_TM")) opt.tm = atoi(getenv("NKP_SIM_TM"));
    if (getenv("NKP_SIM_TN")) opt.tn = atoi(getenv("NKP_SIM_TN"));
    opt.verbose = getenv("NKP_SIM_VERBOSE") ? atoi(getenv("NKP_SIM_VERBOSE")) : 0;
    opt.nranks = nranks;
    const int* coords[3] = {ci, cj, ck};
    std::vector<Plan> plans(nranks);
    for (int r = 0; r < nranks; r++) {
        opt.rank = r;
        int rc = analyse(n, rowptr, colind, (ci || cj || ck) ? coords : nullptr, opt, plans[r], rowmap);
        if (rc) return rc;
    }
    Plan& P0 = plans[0];
    g_last_order_cached = P0.order_cached ? 1 : 0;
    if (stats_out) {
        stats_out[0] = (double)P0.fronts.size();
        stats_out[1] = P0.nlevels;
        stats_out[2] = P0.max_front;
        stats_out[3] = (double)P0.nnz_lu;
        stats_out[4] = (double)P0.heap_len;
        stats_out[5] = P0.flops;
        stats_out[7] = P0.t_order + P0.t_symbolic + P0.t_plan;
    }
    if (perm_out) memcpy(perm_out, P0.perm.data(), sizeof(int) * n);
    if (part_out)
        for (int r = 0; r < nranks; r++) {
            int owned = 0, as_src = 0, as_dst = 0;
            for (int o : plans[r].owner) owned += (o == r);
            for (const Xfer& x : plans[r].xfers) as_src += (x.src == r), as_dst += (x.dst == r);
            part_out[4 * r + 0] = owned;
            part_out[4 * r + 1] = as_src;
            part_out[4 * r + 2] = as_dst;
            part_out[4 * r + 3] = plans[r].flops_local;
        }
    // the partition must be identical on all ranks
    for (int r = 1; r < nranks; r++) {
        if (plans[r].owner != P0.owner || plans[r].xfers.size() != P0.xfers.size()) return -20;
        for (size_t q = 0; q < P0.xfers.size(); q++)
            if (plans[r].xfers[q].front != P0.xfers[q].front || plans[r].xfers[q].src != P0.xfers[q].src ||
                plans[r].xfers[q].dst != P0.xfers[q].dst)
                return -21;
    }
    if (analysis_only) return 0;

    std::vector<double> sval(val, val + P0.nnz);
    if (Rs && Cs)
        for (int i = 0; i < n; i++)
            for (int p = rowptr[i]; p < rowptr[i + 1]; p++) sval[p] = val[p] * Rs[i] * Cs[colind[p]];
    double amax = 0;
    for (int64_t p = 0; p < P0.nnz; p++) amax = std::max(amax, std::fabs(sval[p]));
    double tiny = std::sqrt(2.220446049250313e-16) * amax;
    std::vector<std::vector<double>> heaps(nranks);
    for (int r = 0; r < nranks; r++) {
        heaps[r].assign((size_t)plans[r].heap_len, 0.0);
        for (int64_t p = 0; p < plans[r].nnz; p++)
            if (plans[r].scatter[p] >= 0) heaps[r][plans[r].scatter[p]] += sval[p];
    }

    int nrepl = 0;
    for (int l = P0.nlevels - 1; l >= 0; l--) {
        for (int r = 0; r < nranks; r++) {
            const LevelPlan& L = plans[r].levels[l];
            memset(heaps[r].data() + L.f22_zero_off, 0, sizeof(double) * (size_t)L.f22_zero_len);
        }
        for (int r = 0; r < nranks; r++) {
            Plan& P = plans[r];
            double* H = heaps[r].data();
            const LevelPlan& L = P.levels[l];
            int npass = (int)L.add_tiles.size();
            for (int pass = 0; pass < npass; pass++) {
#pragma omp parallel for schedule(dynamic)
                for (int q = L.add_begin[pass]; q < L.add_begin[pass + 1]; q++) sim_add(H, P.add_tasks[q], P.rel.data(), nb);
            }
            for (int step = 0; step < L.nsteps; step++) {
#pragma omp parallel for schedule(dynamic) reduction(+ : nrepl)
                for (int q = L.diag_begin[step]; q < L.diag_begin[step + 1]; q++) sim_diag(H, P.diag_tasks[q], tiny, &nrepl);
                for (int q = L.trsm_begin[step]; q < L.trsm_begin[step + 1]; q++) sim_trsm(H, P.trsm_tasks[q]);
                for (int q = L.gemm_begin[step]; q < L.gemm_begin[step + 1]; q++) sim_gemm(H, P.gemm_tasks[q], P.opt);
            }
        }
        // the top fronts of this level, in postorder: factored by their groups (nkp_internal.hpp, TopFront).
        // A broadcast is a memcpy from the root's heap into every other member's heap.
        for (int ti : P0.levels[l].tops) {
            const std::vector<int>& grp = P0.groups[P0.top_fronts[ti].group];
            auto bcast = [&](int which /* 0 child, 1 bl, 2 bu */, int idx) -> int {
                auto get = [&](int r) -> const TopBcast& {
                    return which == 0 ? plans[r].top_child_bcasts[idx] : (which == 1 ? plans[r].top_blocks[idx].bl : plans[r].top_blocks[idx].bu);
                };
                const int root = get(grp[0]).root;
                if (!std::binary_search(grp.begin(), grp.end(), root)) return -24;
                for (int r : grp) {
                    if (r == root) continue;
                    if (get(r).root != root || get(r).count != get(root).count || get(r).off < 0 || get(root).off < 0) return -25;
                    memcpy(heaps[r].data() + get(r).off, heaps[root].data() + get(root).off, sizeof(double) * (size_t)get(root).count);
                }
                return 0;
            };
            const TopFront& tf0 = plans[grp[0]].top_fronts[ti];
            for (int q = tf0.cb_begin; q < tf0.cb_end; q++)
                if (int rc = bcast(0, q)) return rc;
            for (int r : grp) {
                Plan& P = plans[r];
                const TopFront& tf = P.top_fronts[ti];
                if (!tf.member) return -26;
                for (size_t pass = 0; pass + 1 < tf.add_begin.size(); pass++)
                    for (int q = tf.add_begin[pass]; q < tf.add_begin[pass + 1]; q++) sim_add(heaps[r].data(), P.add_tasks[q], P.rel.data(), nb);
            }
            auto panel = [&](int r, const TopBlock& tb) {
                Plan& P = plans[r];
                double* H = heaps[r].data();
                for (int st = tb.step_begin; st < tb.step_end; st++) {
                    const TopStep& ts = P.top_steps[st];
                    for (int q = ts.diag_begin; q < ts.diag_end; q++) sim_diag(H, P.diag_tasks[q], tiny, &nrepl);
                    for (int q = ts.trsm_begin; q < ts.trsm_end; q++) sim_trsm(H, P.trsm_tasks[q]);
                    for (int q = ts.gemm_begin; q < ts.gemm_end; q++) sim_gemm(H, P.gemm_tasks[q], P.opt);
                }
            };
            const int nK = tf0.block_end - tf0.block_begin;
            for (int K = 0; K < nK; K++) {
                const int bi = tf0.block_begin + K;
                const int owner = plans[grp[0]].top_blocks[bi].owner;
                if (K == 0) panel(owner, plans[owner].top_blocks[bi]);
                if (int rc = bcast(1, bi)) return rc;
                if (int rc = bcast(2, bi)) return rc;
                for (int r : grp) {
                    Plan& P = plans[r];
                    const TopBlock& tb = P.top_blocks[bi];
                    if (tb.owner != owner) return -27;
                    for (int q = tb.next_begin; q < tb.next_end; q++) sim_gemm(heaps[r].data(), P.gemm_tasks[q], P.opt);
                    if (K + 1 < nK && P.top_blocks[bi + 1].owner == r) panel(r, P.top_blocks[bi + 1]);
                    for (int q = tb.rest_begin; q < tb.rest_end; q++) sim_gemm(heaps[r].data(), P.gemm_tasks[q], P.opt);
                }
            }
        }
    }
    if (stats_out) stats_out[6] = nrepl;
    // the sweeps use inverted 64 x 64 diagonal blocks (k_invert_diag)
    const bool invdiag = !getenv("NKP_SIM_NO_INVDIAG");
    if (invdiag)
        for (int r = 0; r < nranks; r++) {
            Plan& P = plans[r];
            double* H = heaps[r].data();
#pragma omp parallel for schedule(dynamic)
            for (size_t q = 0; q < P.inv_tasks.size(); q++) sim_invert_diag(H, P.inv_tasks[q]);
        }
    if (nrhs <= 0) return 0;

    // ---- solves: y = permuted rhs (replicated); forward deepest -> root, backward root -> deepest
    std::vector<std::vector<double>> Ws(nranks), ys(nranks);
    for (int r = 0; r < nranks; r++) {
        Ws[r].assign((size_t)plans[r].solve_pool_len, 0.0);
        ys[r].assign(n, 0.0);
    }
    int nref = getenv("NKP_SIM_REFINE") ? atoi(getenv("NKP_SIM_REFINE")) : 0;
    std::vector<double> rhs(n), xacc(n);
    for (int c = 0; c < nrhs; c++)
      for (int it = 0; it <= nref; it++) {
        if (it == 0) {
            for (int i = 0; i < n; i++) rhs[i] = B[i + (int64_t)c * n];
            std::fill(xacc.begin(), xacc.end(), 0.0);
        } else {
            // residual r = b - A x  (src/SuperLU_brief_tree.txt:20-24, pdgsrfs)
            double rn = 0, bn = 0;
            for (int i = 0; i < n; i++) {
                long double acc = B[i + (int64_t)c * n];
                for (int p = rowptr[i]; p < rowptr[i + 1]; p++) acc -= (long double)val[p] * xacc[colind[p]];
                rhs[i] = (double)acc;
                rn += rhs[i] * rhs[i];
                bn += B[i + (int64_t)c * n] * B[i + (int64_t)c * n];
            }
            fprintf(stderr, "[sim] rhs %d refine it %d: relres before = %.3e\n", c, it, std::sqrt(rn / bn));
        }
        for (int r = 0; r < nranks; r++)
            for (int i = 0; i < n; i++) ys[r][P0.perm[rowmap ? rowmap[i] : i]] = (Rs ? Rs[i] : 1.0) * rhs[i];
        for (int l = P0.nlevels - 1; l >= 0; l--) {
            // update vectors of children owned elsewhere
            if (l + 1 < P0.nlevels)
                for (int q : P0.levels[l + 1].xfers) {
                    const Xfer& x = P0.xfers[q];
                    const Front& fs = plans[x.src].fronts[x.front];
                    const Front& fd = plans[x.dst].fronts[x.front];
                    if (fs.woff < 0 || fd.woff < 0) return -23;
                    memcpy(Ws[x.dst].data() + fd.woff, Ws[x.src].data() + fs.woff, sizeof(double) * (size_t)fs.m);
                }
            for (int r = 0; r < nranks; r++) {
                Plan& P = plans[r];
                const double* H = heaps[r].data();
                std::vector<double>& W = Ws[r];
                std::vector<double>& y = ys[r];
                const LevelPlan& L = P.levels[l];
#pragma omp parallel for schedule(dynamic)
                for (int q = L.solve_begin; q < L.solve_end; q++) {
                    const SolveTask& t = P.solve_tasks[q];
                    double* w = W.data() + t.woff;
                    for (int a = 0; a < t.s; a++) w[a] = y[t.first + a];
                    for (int a = t.s; a < t.m; a++) w[a] = 0;
                    for (int c2 = 0; c2 < t.nchild; c2++) {
                        const SolveChild& sc = P.solve_children[t.child_list + c2];
                        const double* wc = W.data() + sc.woff + sc.s;
                        const int* rl = P.rel.data() + sc.rel_off;
                        for (int a = 0; a < sc.r; a++) w[rl[a]] += wc[a];
                    }
                    const double* Lr = H + t.Loff;
                    if (invdiag) {
                        // block forward substitution with inverted 64 x 64 diagonal blocks (k_sweep_big<FWD>, k_fwd_small)
                        double tmp[64];
                        for (int k0 = 0; k0 < t.s; k0 += 64) {
                            int kb = std::min(64, t.s - k0);
                            for (int a = 0; a < kb; a++) {
                                double acc = w[k0 + a];
                                for (int p = 0; p < a; p++) acc += Lr[k0 + a + (int64_t)(k0 + p) * t.ld] * w[k0 + p];
                                tmp[a] = acc;
                            }
                            for (int a = 0; a < kb; a++) w[k0 + a] = tmp[a];
                            for (int p = 0; p < kb; p++) {
                                double yp = w[k0 + p];
                                for (int a = k0 + kb; a < t.m; a++) w[a] -= Lr[a + (int64_t)(k0 + p) * t.ld] * yp;
                            }
                        }
                    } else {
                        for (int p = 0; p < t.s; p++) {
                            double yp = w[p];
                            for (int a = p + 1; a < t.m; a++) w[a] -= Lr[a + (int64_t)p * t.ld] * yp;
                        }
                    }
                    for (int a = 0; a < t.s; a++) y[t.first + a] = w[a];
                }
            }
        }
        for (int l = 0; l < P0.nlevels; l++) {
            for (int r = 0; r < nranks; r++) {
                Plan& P = plans[r];
                const double* H = heaps[r].data();
                std::vector<double>& W = Ws[r];
                std::vector<double>& y = ys[r];
                const LevelPlan& L = P.levels[l];
#pragma omp parallel for schedule(dynamic)
                for (int q = L.solve_begin; q < L.solve_end; q++) {
                    const SolveTask& t = P.solve_tasks[q];
                    double* w = W.data() + t.woff;
                    const int* bi = P.bidx.data() + t.bidx_off;
                    for (int a = 0; a < t.r; a++) w[t.s + a] = y[bi[a]];
                    const double* UT = H + t.UToff;
                    if (invdiag) {
                        // block back substitution, diagonal blocks hold (U_kk^-1)^T (k_sweep_big<BWD_*>, k_bwd_small)
                        double z[64];
                        for (int k0 = (t.s - 1) / 64 * 64; k0 >= 0; k0 -= 64) {
                            int kb = std::min(64, t.s - k0);
                            for (int p = 0; p < kb; p++) {
                                double acc = y[t.first + k0 + p];
                                for (int a = k0 + kb; a < t.m; a++) acc -= UT[a + (int64_t)(k0 + p) * t.ld] * w[a];
                                z[p] = acc;
                            }
                            for (int p = 0; p < kb; p++) {
                                double acc = 0;   // x_p = sum_{q >= p} Uinv(p,q) z_q, Uinv(p,q) = UT[k0+q + (k0+p) m]
                                for (int q2 = p; q2 < kb; q2++) acc += UT[k0 + q2 + (int64_t)(k0 + p) * t.ld] * z[q2];
                                w[k0 + p] = acc;
                                y[t.first + k0 + p] = acc;
                            }
                        }
                    } else {
                        for (int p = t.s - 1; p >= 0; p--) {
                            double z = y[t.first + p];
                            for (int a = p + 1; a < t.m; a++) z -= UT[a + (int64_t)p * t.ld] * w[a];
                            w[p] = z / UT[p + (int64_t)p * t.ld];
                            y[t.first + p] = w[p];
                        }
                    }
                }
            }
        }
        // every part of the solution is published by one rank that holds it (Plan::pub)
        for (const PubRange& pr : P0.pub)
            for (int r = 0; r < nranks; r++)
                if (r != pr.root) memcpy(ys[r].data() + pr.lo, ys[pr.root].data() + pr.lo, sizeof(double) * (size_t)(pr.hi - pr.lo));
        for (int i = 0; i < n; i++) {
            xacc[i] += (Cs ? Cs[i] : 1.0) * ys[0][P0.perm[i]];
            X[i + (int64_t)c * n] = xacc[i];
        }
      }
    return 0;
}

int nkp_sim_run_dist(int n, const int* rowptr, const int* colind, const double* val, const int* ci,
                     const int* cj, const int* ck, int nb, int leaf, const double* B, int nrhs,
                     double* X, double* stats_out, int* perm_out, int analysis_only, int nranks,
                     double* part_out) {
    return sim_run_impl(n, rowptr, colind, val, ci, cj, ck, nb, leaf, B, nrhs, X, stats_out, perm_out, analysis_only,
                        nranks, part_out, nullptr, nullptr, nullptr);
}

int nkp_sim_run_rowperm(int n, const int* rowptr, const int* colind, const double* val, const int* ci,
                        const int* cj, const int* ck, int nb, int leaf, const double* B, int nrhs,
                        double* X, double* stats_out, int nranks, const int* rowmap, const double* Rs,
                        const double* Cs) {
    return sim_run_impl(n, rowptr, colind, val, ci, cj, ck, nb, leaf, B, nrhs, X, stats_out, nullptr, 0, nranks, nullptr,
                        rowmap, Rs, Cs);
}

// Partition as seen by ONE rank (for the world_size-2 gloo test): owner_out[n_fronts],
// xfer_out[3 * n_xfers] = (front, src, dst); returns n_fronts, or a negative error.
int nkp_sim_partition(int n, const int* rowptr, const int* colind, const int* ci, const int* cj, const int* ck,
                      int nb, int leaf, int rank, int nranks, int* owner_out, int owner_cap, int* xfer_out,
                      int xfer_cap, int* n_xfers, double* local_out) {
    Options opt;
    opt.nb = nb;
    opt.leaf = leaf;
    tree_env(opt);
    if (opt.tn > nb) opt.tn = nb;
    opt.rank = rank;
    opt.nranks = nranks;
    if (getenv("NKP_OUTER")) opt.outer = opt.top_outer = atoi(getenv("NKP_OUTER"));
    if (getenv("NKP_TOP_OUTER")) opt.top_outer = atoi(getenv("NKP_TOP_OUTER"));
    if (getenv("NKP_SPLIT_TOL")) opt.split_tol = atof(getenv("NKP_SPLIT_TOL"));
    if (getenv("NKP_SPLIT_MAX")) opt.split_max = atoi(getenv("NKP_SPLIT_MAX"));
    const int* coords[3] = {ci, cj, ck};
    Plan P;
    int rc = analyse(n, rowptr, colind, (ci || cj || ck) ? coords : nullptr, opt, P);
    if (rc) return rc;
    int nf = (int)P.fronts.size();
    if (nf > owner_cap || (int)P.xfers.size() > xfer_cap) return -30;
    for (int t = 0; t < nf; t++) owner_out[t] = P.owner[t];
    for (size_t q = 0; q < P.xfers.size(); q++) {
        xfer_out[3 * q + 0] = P.xfers[q].front;
        xfer_out[3 * q + 1] = P.xfers[q].src;
        xfer_out[3 * q + 2] = P.xfers[q].dst;
    }
    *n_xfers = (int)P.xfers.size();
    if (local_out) {
        local_out[0] = P.flops_local;
        local_out[1] = P.flops;
        local_out[2] = (double)P.heap_len;
        local_out[3] = (double)P.nnz_lu_local;
        // distributed top of the tree: number of top fronts, largest number of outer blocks of one of them,
        // number of distinct groups, largest group
        int maxblk = 0, maxgrp = 0;
        for (const TopFront& tf : P.top_fronts) maxblk = std::max(maxblk, tf.block_end - tf.block_begin);
        for (const std::vector<int>& g : P.groups) maxgrp = std::max(maxgrp, (int)g.size());
        local_out[4] = (double)P.top_fronts.size();
        local_out[5] = maxblk;
        local_out[6] = (double)P.groups.size();
        local_out[7] = maxgrp;
    }
    return nf;
}

// Structural invariants of the plan that the CUDA kernels rely on (checked on the CPU, no numerics):
// returns 0 when all hold, otherwise a positive code naming the first violated invariant.
int nkp_sim_check_plan(int n, const int* rowptr, const int* colind, const int* ci, const int* cj, const int* ck,
                       int nb, int leaf, int nranks) {
    Options opt;
    opt.nb = nb;
    opt.leaf = leaf;
    tree_env(opt);
    if (opt.tn > nb) opt.tn = nb;
    opt.nranks = nranks;
    const int* coords[3] = {ci, cj, ck};
    for (int rank = 0; rank < nranks; rank++) {
        opt.rank = rank;
        Plan P;
        int rc = analyse(n, rowptr, colind, (ci || cj || ck) ? coords : nullptr, opt, P);
        if (rc) return 1;
        // 16-byte alignment of every panel column (cp.async 16 in the sweeps)
        for (const Front& f : P.fronts) {
            if (f.Loff < 0) continue;
            if ((f.ld & 1) || f.ld < f.m || f.ld > f.m + 1 || (f.Loff & 1) || (f.UToff & 1)) return 2;
        }
        // every solve task is either small or big, never both; big fronts mirror their task
        size_t nbig = 0, nsmall = 0;
        for (const SolveTask& t : P.solve_tasks) (t.big ? nbig : nsmall)++;
        if (nbig != P.big_fronts.size() || nsmall != P.solve_small.size()) return 3;
        // the diagonal blocks of every front are inverted exactly once
        size_t ninv = 0;
        for (const SolveTask& t : P.solve_tasks) ninv += (size_t)(t.s + 63) / 64;
        if (ninv != P.inv_tasks.size()) return 4;
        for (const DiagTask& d : P.inv_tasks)
            if (d.kb < 1 || d.kb > 64 || (d.ld & 1)) return 5;
        for (const LevelPlan& L : P.levels) {
            // forward items: every slab of every big front of the level exactly once, pivot slabs of a
            // front in increasing order (a CTA only waits on items listed before its own)
            std::vector<int> seen;
            std::vector<int> last_piv(P.big_fronts.size(), -1);
            int64_t nfwd = 0, nbwd = 0, nrect = 0, slots = 0;
            for (int b = L.big_begin; b < L.big_end; b++) {
                const BigFront& bf = P.big_fronts[b];
                if (bf.npiv != (bf.s + 63) / 64 || bf.nslab != bf.npiv + (bf.r + 63) / 64) return 6;
                if (bf.nchunk != ((bf.r + 63) / 64 + BWD_CHUNK - 1) / BWD_CHUNK) return 7;
                if (bf.part_off != slots) return 8;
                slots += (int64_t)bf.npiv * bf.nchunk;
                nfwd += bf.nslab;
                nbwd += bf.npiv;
                nrect += (int64_t)bf.npiv * bf.nchunk;
                // child_lo[child][slab] = first entry of the child's rel[] that maps to a row >= the slab's first row
                for (int c = 0; c < bf.nchild; c++) {
                    const SolveChild& sc = P.solve_children[bf.child_list + c];
                    const int* rl = P.rel.data() + sc.rel_off;
                    for (int j = 0; j < bf.nslab; j++) {
                        int row0 = j < bf.npiv ? 64 * j : bf.s + 64 * (j - bf.npiv);
                        int lo = P.child_lo[bf.clo_off + (int64_t)c * bf.nslab + j];
                        if (lo < 0 || lo > sc.r) return 9;
                        if (lo < sc.r && rl[lo] < row0) return 10;
                        if (lo > 0 && rl[lo - 1] >= row0) return 11;
                    }
                }
            }
            if (slots > P.bwd_part_slots) return 12;
            if (L.fwd_item_end - L.fwd_item_begin != nfwd || L.bwd_item_end - L.bwd_item_begin != nbwd ||
                L.rect_item_end - L.rect_item_begin != nrect)
                return 13;
            for (int q = L.fwd_item_begin; q < L.fwd_item_end; q++) {
                const BigItem& it = P.big_fwd_items[q];
                if (it.front < L.big_begin || it.front >= L.big_end) return 14;
                const BigFront& bf = P.big_fronts[it.front];
                if (it.idx < 0 || it.idx >= bf.nslab) return 15;
                if (it.idx < bf.npiv) {
                    if (it.idx != last_piv[it.front] + 1) return 16;
                    last_piv[it.front] = it.idx;
                } else if (last_piv[it.front] != bf.npiv - 1) {
                    return 17;   // a boundary slab listed before the last pivot slab of its front
                }
            }
            std::vector<int> next_panel(P.big_fronts.size(), -2);
            for (int q = L.bwd_item_begin; q < L.bwd_item_end; q++) {
                const BigItem& it = P.big_bwd_items[q];
                const BigFront& bf = P.big_fronts[it.front];
                int expect = next_panel[it.front] == -2 ? bf.npiv - 1 : next_panel[it.front];
                if (it.idx != expect) return 18;   // panels of a front from the last one down
                next_panel[it.front] = it.idx - 1;
            }
        }
    }
    return 0;
}

// the fronts of the single-GPU plan: out[4 t .. 4 t + 3] = first pivot, pivots, boundary size, level; returns their number
// (parent_out, may be null: index of the parent front, -1 for roots)
int nkp_sim_fronts2(int n, const int* rowptr, const int* colind, const int* ci, const int* cj, const int* ck, int nb,
                    int leaf, int* out, int cap, int* perm_out, int* parent_out);
int nkp_sim_fronts(int n, const int* rowptr, const int* colind, const int* ci, const int* cj, const int* ck, int nb,
                   int leaf, int* out, int cap, int* perm_out) {
    return nkp_sim_fronts2(n, rowptr, colind, ci, cj, ck, nb, leaf, out, cap, perm_out, nullptr);
}
int nkp_sim_fronts2(int n, const int* rowptr, const int* colind, const int* ci, const int* cj, const int* ck, int nb,
                    int leaf, int* out, int cap, int* perm_out, int* parent_out) {
    Options opt;
    opt.nb = nb;
    opt.leaf = leaf;
    tree_env(opt);
    if (opt.tn > nb) opt.tn = nb;
    const int* coords[3] = {ci, cj, ck};
    Plan P;
    if (analyse(n, rowptr, colind, (ci || cj || ck) ? coords : nullptr, opt, P)) return -1;
    const int nf = (int)P.fronts.size();
    for (int t = 0; t < nf && t < cap; t++) {
        out[4 * t] = P.fronts[t].first;
        out[4 * t + 1] = P.fronts[t].s;
        out[4 * t + 2] = P.fronts[t].r;
        out[4 * t + 3] = P.fronts[t].level;
        if (parent_out) parent_out[t] = P.fronts[t].parent;
    }
    if (perm_out) memcpy(perm_out, P.perm.data(), sizeof(int) * n);
    return nf;
}

// INDEPENDENT symbolic factorisation (shares nothing with csrc/analysis.cpp but the permutation it is given):
// below-diagonal column counts of the Cholesky factor of pattern(A + A^T) in the numbering perm[old] = new --
// elimination tree by path compression (Liu), then one row-subtree traversal per row, which visits every nonzero
// of L exactly once.  O(nnz(L)) time, O(n + nnz(A)) memory.  From the counts c_j: nnz(L + U) = 2 sum c_j + n and
// the LU flops of a structurally symmetric elimination, sum (2 c_j^2 + c_j) -- the minimum any method pays for
// this ordering; the plan's dense fronts pay for padding on top (tests/test_oracle_and_plan.py).
int nkp_true_colcounts(int n, const int* rowptr, const int* colind, const int* perm, long long* counts) {
    std::vector<int64_t> lp((size_t)n + 1, 0);
    for (int i = 0; i < n; i++)
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const int a = perm[i], b = perm[colind[p]];
            if (a != b) lp[std::max(a, b) + 1]++;
        }
    for (int i = 0; i < n; i++) lp[i + 1] += lp[i];
    std::vector<int> lo((size_t)lp[n]);   // for every row (new numbering): its neighbours with a smaller number
    {
        std::vector<int64_t> pos(lp.begin(), lp.end() - 1);
        for (int i = 0; i < n; i++)
            for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
                const int a = perm[i], b = perm[colind[p]];
                if (a != b) lo[pos[std::max(a, b)]++] = std::min(a, b);
            }
    }
    std::vector<int> parent((size_t)n, -1), anc((size_t)n, -1);
    for (int i = 0; i < n; i++)
        for (int64_t p = lp[i]; p < lp[i + 1]; p++) {
            int k = lo[p];
            while (k != -1 && k < i) {
                const int next = anc[k];
                anc[k] = i;
                if (next == -1) parent[k] = i;
                k = next;
            }
        }
    std::vector<int> mark((size_t)n, -1);
    for (int i = 0; i < n; i++) counts[i] = 0;
    for (int i = 0; i < n; i++) {
        mark[i] = i;
        for (int64_t p = lp[i]; p < lp[i + 1]; p++)
            for (int k = lo[p]; k != -1 && mark[k] != i; k = parent[k]) {
                mark[k] = i;
                counts[k]++;
            }
    }
    return 0;
}

int nkp_sim_run(int n, const int* rowptr, const int* colind, const double* val, const int* ci,
                const int* cj, const int* ck, int nb, int leaf, const double* B, int nrhs,
                double* X, double* stats_out, int* perm_out, int analysis_only) {
    int nranks = getenv("NKP_SIM_RANKS") ? atoi(getenv("NKP_SIM_RANKS")) : 1;
    return nkp_sim_run_dist(n, rowptr, colind, val, ci, cj, ck, nb, leaf, B, nrhs, X, stats_out, perm_out,
                            analysis_only, nranks, nullptr);
}
}
