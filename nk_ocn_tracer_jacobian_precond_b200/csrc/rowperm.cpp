// rowperm.cpp -- static row permutation for a large diagonal ("LargeDiag").
//
// What SuperLU_DIST's pdgssvx* does before the factorisation when options->RowPerm == LargeDiag -- the default that
// set_default_options_dist leaves behind, and the reference never changes it (src/solve_ABglobal.c:332-334,
// src/solve_ABdist.c:493-495): HSL MC64 job 5 (Duff & Koster, SIAM J. Matrix Anal. Appl. 22, 2001), i.e. the row
// permutation that maximises the PRODUCT of the diagonal magnitudes, together with row / column scalings from the
// dual variables under which every diagonal entry has magnitude 1 and no off-diagonal entry exceeds 1.  That is the
// property static pivoting (no row exchanges during the numeric phase) relies on.
//
// MC64 is not part of the reference tree (it sits inside the external SuperLU_DIST 5.1.3, src/Makefile:3); this is an
// independent implementation of the published algorithm on the CRS operand of src/matrix.h:64-68:
//   cost(i,j) = log(max_k |a_ik|) - log|a_ij|  >= 0
//   minimum-cost perfect matching by successive shortest augmenting paths (Dijkstra on reduced costs) after a
//   greedy start on the tight edges; duals u (rows), v (columns) with u_i + v_j <= cost(i,j), equality on the matching
//   R_i = exp(u_i) / max_k |a_ik| ,  C_j = exp(v_j)   =>   R_i |a_ij| C_j = exp(u_i + v_j - cost(i,j)) <= 1.
// Host code: the matching runs once per sparsity pattern, next to the ordering (SamePattern_SameRowPerm afterwards).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <utility>
#include <vector>

#include "../../include/nkprecond.h"

namespace {
const double INF = std::numeric_limits<double>::infinity();
}

extern "C" int nkp_rowperm_largediag(int n, const int* rowptr, const int* colind, const double* nzval, int* rowmap,
                                     double* R, double* C) {
    if (n <= 0 || !rowptr || !colind || !nzval || !rowmap) return NKP_EINVAL;
    const int64_t nnz = rowptr[n];
    std::vector<double> cost((size_t)nnz), rlog((size_t)n);
    bool empty_row = false;
#pragma omp parallel for schedule(static) reduction(|| : empty_row)
    for (int i = 0; i < n; i++) {
        double mx = 0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            if (colind[p] < 0 || colind[p] >= n) {
                empty_row = true;
                continue;
            }
            mx = std::max(mx, std::fabs(nzval[p]));
        }
        if (!(mx > 0) || !std::isfinite(mx)) {
            empty_row = true;
            continue;
        }
        const double lm = std::log(mx);
        rlog[i] = lm;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const double a = std::fabs(nzval[p]);
            cost[p] = a > 0 ? std::max(0.0, lm - std::log(a)) : INF;   // explicit zeros do not take part
        }
    }
    if (empty_row) return NKP_EANALYSIS;   // a row without a nonzero (or a bad index): structurally singular

    // dual start: v_j = min_i cost(i,j), u_i = min_j (cost(i,j) - v_j): feasible, at least one tight edge per row
    std::vector<double> u((size_t)n, 0.0), v((size_t)n, INF);
    for (int64_t p = 0; p < nnz; p++) v[colind[p]] = std::min(v[colind[p]], cost[p]);
    for (int j = 0; j < n; j++)
        if (v[j] == INF) return NKP_EANALYSIS;   // empty column
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        double m = INF;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) m = std::min(m, cost[p] - v[colind[p]]);
        u[i] = m;
    }
    std::vector<int> match_row((size_t)n, -1), match_col((size_t)n, -1);   // row -> column, column -> row
    // greedy start on tight edges, the diagonal first (it is the answer for almost every row of this operator family)
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < n; i++) {
            if (match_row[i] >= 0) continue;
            for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
                const int j = colind[p];
                if (pass == 0 && j != i) continue;
                if (match_col[j] >= 0 || cost[p] == INF) continue;
                if (cost[p] - u[i] - v[j] <= 0.0) {
                    match_row[i] = j;
                    match_col[j] = i;
                    break;
                }
            }
        }

    // shortest augmenting paths from the rows that are still free
    std::vector<double> d((size_t)n, INF);
    std::vector<int> pred((size_t)n, -1);
    std::vector<char> scanned((size_t)n, 0);
    std::vector<int> touched, order;
    typedef std::pair<double, int> Entry;
    for (int i0 = 0; i0 < n; i0++) {
        if (match_row[i0] >= 0) continue;
        std::priority_queue<Entry, std::vector<Entry>, std::greater<Entry>> heap;
        touched.clear();
        order.clear();
        double lower = 0;
        int i = i0, sink = -1;
        for (;;) {
            for (int p = rowptr[i]; p < rowptr[i + 1]; p++) {
                const int j = colind[p];
                if (scanned[j] || cost[p] == INF) continue;
                const double dn = lower + std::max(0.0, cost[p] - u[i] - v[j]);
                if (dn < d[j]) {
                    if (d[j] == INF) touched.push_back(j);
                    d[j] = dn;
                    pred[j] = i;
                    heap.push(Entry(dn, j));
                }
            }
            int j = -1;
            while (!heap.empty()) {
                const Entry e = heap.top();
                heap.pop();
                if (!scanned[e.second] && e.first <= d[e.second]) {
                    j = e.second;
                    break;
                }
            }
            if (j < 0) break;   // no augmenting path: structurally singular
            scanned[j] = 1;
            order.push_back(j);
            lower = d[j];
            if (match_col[j] < 0) {
                sink = j;
                break;
            }
            i = match_col[j];
        }
        if (sink < 0) return NKP_EANALYSIS;
        // duals: every edge of the shortest-path tree becomes tight, feasibility is kept
        u[i0] += lower;
        for (int j : order) {
            const double delta = lower - d[j];
            if (j != sink) u[match_col[j]] += delta;
            v[j] -= delta;
        }
        // augment along the predecessor chain
        for (int j = sink;;) {
            const int ii = pred[j];
            const int jprev = match_row[ii];
            match_row[ii] = j;
            match_col[j] = ii;
            if (ii == i0) break;
            j = jprev;
        }
        for (int j : touched) {
            d[j] = INF;
            scanned[j] = 0;
            pred[j] = -1;
        }
    }
    for (int i = 0; i < n; i++) rowmap[i] = match_row[i];
    if (R)
        for (int i = 0; i < n; i++) R[i] = std::exp(u[i] - rlog[i]);
    if (C)
        for (int j = 0; j < n; j++) C[j] = std::exp(v[j]);
    return NKP_OK;
}
