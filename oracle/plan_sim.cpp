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
            int64_t dst = front_entry(rl[a], rl[b], t.sp, t.mp, nb, t.Loff, t.UToff, t.F22off);
            H[dst] += C[a + (int64_t)b * t.rc];
        }
}

}  // namespace

extern "C" {

// stats_out[0..7]: n_fronts, n_levels, max_front, nnz_lu, heap_len, flops, tiny_pivots, analysis seconds
int nkp_sim_run(int n, const int* rowptr, const int* colind, const double* val, const int* ci,
                const int* cj, const int* ck, int nb, int leaf, const double* B, int nrhs,
                double* X, double* stats_out, int* perm_out, int analysis_only) {
    Options opt;
    opt.nb = nb;
    opt.leaf = leaf;
    if (opt.tn > nb) opt.tn = nb;
    if (getenv("NKP_OUTER")) opt.outer = atoi(getenv("NKP_OUTER"));
    if (getenv("NKP_SIM_TM")) opt.tm = atoi(getenv("NKP_SIM_TM"));
    if (getenv("NKP_SIM_TN")) opt.tn = atoi(getenv("NKP_SIM_TN"));
    opt.verbose = getenv("NKP_SIM_VERBOSE") ? 1 : 0;
    const int* coords[3] = {ci, cj, ck};
    Plan P;
    int rc = analyse(n, rowptr, colind, (ci || cj || ck) ? coords : nullptr, opt, P);
    if (rc) return rc;
    if (stats_out) {
        stats_out[0] = (double)P.fronts.size();
        stats_out[1] = P.nlevels;
        stats_out[2] = P.max_front;
        stats_out[3] = (double)P.nnz_lu;
        stats_out[4] = (double)P.heap_len;
        stats_out[5] = P.flops;
        stats_out[7] = P.t_order + P.t_symbolic + P.t_plan;
    }
    if (perm_out) memcpy(perm_out, P.perm.data(), sizeof(int) * n);
    if (analysis_only) return 0;

    std::vector<double> heap((size_t)P.heap_len, 0.0);
    double* H = heap.data();
    double amax = 0;
    for (int64_t p = 0; p < P.nnz; p++) amax = std::max(amax, std::fabs(val[p]));
    double tiny = std::sqrt(2.220446049250313e-16) * amax;
    for (int64_t p = 0; p < P.nnz; p++) H[P.scatter[p]] += val[p];

    int nrepl = 0;
    for (int l = P.nlevels - 1; l >= 0; l--) {
        const LevelPlan& L = P.levels[l];
        memset(H + L.f22_zero_off, 0, sizeof(double) * (size_t)L.f22_zero_len);
        int npass = (int)L.add_tiles.size();
        for (int pass = 0; pass < npass; pass++) {
#pragma omp parallel for schedule(dynamic)
            for (int q = L.add_begin[pass]; q < L.add_begin[pass + 1]; q++) sim_add(H, P.add_tasks[q], P.rel.data(), nb);
        }
        for (int step = 0; step < L.nsteps; step++) {
#pragma omp parallel for schedule(dynamic) reduction(+ : nrepl)
            for (int q = L.diag_begin[step]; q < L.diag_begin[step + 1]; q++) sim_diag(H, P.diag_tasks[q], tiny, &nrepl);
            if (getenv("NKP_SIM_POISON")) {
                for (int q = L.diag_begin[step]; q < L.diag_begin[step + 1]; q++) {
                    const DiagTask& d = P.diag_tasks[q];
                    for (int a = 0; a < d.kb; a++) for (int b = 0; b < d.kb; b++)
                        if (std::isnan(H[d.Doff + a + (int64_t)b * d.ld])) { fprintf(stderr, "NaN in diag block level %d step %d task %d (%d,%d) kb=%d ld=%d\n", l, step, q - L.diag_begin[step], a, b, d.kb, d.ld); a = b = 1 << 20; }
                }
                for (int q = L.trsm_begin[step]; q < L.trsm_begin[step + 1]; q++) {
                    const TrsmTask& d = P.trsm_tasks[q];
                    for (int a = 0; a < d.nrows; a++) for (int b = 0; b < d.kb; b++)
                        if (std::isnan(H[d.Xoff + a + (int64_t)b * d.ld])) { fprintf(stderr, "NaN in trsm X level %d step %d unit %d (%d,%d) nrows=%d kb=%d ld=%d\n", l, step, d.unit, a, b, d.nrows, d.kb, d.ld); a = b = 1 << 20; }
                }
            }
            for (int q = L.trsm_begin[step]; q < L.trsm_begin[step + 1]; q++) sim_trsm(H, P.trsm_tasks[q]);
            for (int q = L.gemm_begin[step]; q < L.gemm_begin[step + 1]; q++) sim_gemm(H, P.gemm_tasks[q], opt);
        }
    }
    if (stats_out) stats_out[6] = nrepl;
    if (nrhs <= 0) return 0;

    // ---- solves: y = permuted rhs; forward deepest -> root, backward root -> deepest --------
    std::vector<double> W((size_t)P.solve_pool_len);
    std::vector<double> y(n);
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
        for (int i = 0; i < n; i++) y[P.perm[i]] = rhs[i];
        for (int l = P.nlevels - 1; l >= 0; l--) {
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
                for (int p = 0; p < t.s; p++) {
                    double yp = w[p];
                    for (int a = p + 1; a < t.m; a++) w[a] -= Lr[a + (int64_t)p * t.m] * yp;
                }
                for (int a = 0; a < t.s; a++) y[t.first + a] = w[a];
            }
        }
        for (int l = 0; l < P.nlevels; l++) {
            const LevelPlan& L = P.levels[l];
#pragma omp parallel for schedule(dynamic)
            for (int q = L.solve_begin; q < L.solve_end; q++) {
                const SolveTask& t = P.solve_tasks[q];
                double* w = W.data() + t.woff;
                const int* bi = P.bidx.data() + t.bidx_off;
                for (int a = 0; a < t.r; a++) w[t.s + a] = y[bi[a]];
                const double* UT = H + t.UToff;
                for (int p = t.s - 1; p >= 0; p--) {
                    double z = y[t.first + p];
                    for (int a = p + 1; a < t.m; a++) z -= UT[a + (int64_t)p * t.m] * w[a];
                    w[p] = z / UT[p + (int64_t)p * t.m];
                    y[t.first + p] = w[p];
                }
            }
        }
        for (int i = 0; i < n; i++) {
            xacc[i] += y[P.perm[i]];
            X[i + (int64_t)c * n] = xacc[i];
        }
      }
    return 0;
}
}
