"""Static row permutation for a large diagonal (csrc/rowperm.cpp, nkp_rowperm_largediag / nkp_create_rowperm): what
pdgssvx* does first under RowPerm = LargeDiag, the default the reference keeps (src/solve_ABglobal.c:332-334).

CPU side: the matching is a host computation (no GPU needed to call it), checked against scipy's independent
minimum-weight bipartite matching; the plan built on the permuted pattern is interpreted on the CPU
(oracle/plan_sim.cpp) and compared with the pivoted oracle."""
import ctypes

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.csgraph import min_weight_full_bipartite_matching

from nk_ocn_tracer_jacobian_precond_b200 import solver
from oracle import oracle_solve

P = ctypes.POINTER


def _ip(a):
    return a.ctypes.data_as(P(ctypes.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(P(ctypes.c_double)) if a is not None else None


def pow2(x):
    """the rounding nkp_create_rowperm applies to the scalings"""
    return np.ldexp(1.0, np.rint(np.log2(x)).astype(np.int32))


def sim_rowperm(lib, n, rp, ci, nz, coords, B, rowmap=None, R=None, Cs=None, nranks=1, nb=64, leaf=96):
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    nz = np.ascontiguousarray(nz, np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    X = np.zeros_like(B, order="F")
    stats = np.zeros(8)
    i, j, k = coords if coords is not None else (None, None, None)
    rc = lib.nkp_sim_run_rowperm(n, _ip(rp), _ip(ci), _dp(nz), _ip(i), _ip(j), _ip(k), nb, leaf, _dp(B), B.shape[1], _dp(X),
                                 _dp(stats), nranks, _ip(rowmap), _dp(R), _dp(Cs))
    return rc, X, stats


def _coords(m):
    return tuple(np.ascontiguousarray(m["tracer_state_ind_to_" + q], np.int32) for q in "ijk")


@pytest.mark.parametrize("seed", range(8))
def test_matching_maximises_the_diagonal_product(seed):
    """Objective equal to scipy's independent optimal assignment; the duals give |R a C| <= 1 with 1 on the matching."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 300))
    A = sp.random(n, n, density=rng.uniform(0.02, 0.3), random_state=rng, format="csr",
                  data_rvs=lambda k: rng.standard_normal(k) * 10.0 ** rng.uniform(-3, 3, k))
    hidden = rng.permutation(n)   # guarantees a perfect matching
    A = (A + sp.csr_matrix((rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n), (np.arange(n), hidden)), shape=(n, n))).tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    rowmap, R, Cs = solver.rowperm_largediag(n, A.indptr, A.indices, A.data)
    assert sorted(rowmap.tolist()) == list(range(n))
    W = A.copy()
    W.data = (np.log(np.abs(A.data)).max() + 1.0) - np.log(np.abs(A.data))   # positive weights, minimised
    r, c = min_weight_full_bipartite_matching(W)
    ref = np.sum(np.log(np.abs(np.asarray(A[r, c]).ravel())))
    got = np.sum(np.log(np.abs(np.asarray(A[np.arange(n), rowmap]).ravel())))
    assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref))
    S = sp.diags(R) @ abs(A) @ sp.diags(Cs)
    assert S.max() <= 1.0 + 1e-12
    assert np.abs(np.asarray(S[np.arange(n), rowmap]).ravel() - 1.0).max() <= 1e-12


def test_matching_edge_cases():
    # already diagonal: identity, scalings 1 / |a_ii|
    rowmap, R, Cs = solver.rowperm_largediag(3, [0, 1, 2, 3], [0, 1, 2], [2.0, -4.0, 0.5])
    assert rowmap.tolist() == [0, 1, 2]
    assert np.allclose(R * np.array([2.0, 4.0, 0.5]) * Cs, 1.0)
    # explicit zeros do not take part: the only nonzero transversal is the anti-diagonal
    rowmap, _, _ = solver.rowperm_largediag(2, [0, 2, 4], [0, 1, 0, 1], [0.0, 5.0, 3.0, 0.0])
    assert rowmap.tolist() == [1, 0]
    # structurally singular: two rows that only reach the same column / an empty row
    with pytest.raises(solver.NkpError):
        solver.rowperm_largediag(2, [0, 1, 2], [0, 0], [1.0, 1.0])
    with pytest.raises(solver.NkpError):
        solver.rowperm_largediag(2, [0, 2, 2], [0, 1], [1.0, 1.0])
    # explicit zeros only on the one transversal that exists structurally
    with pytest.raises(solver.NkpError):
        solver.rowperm_largediag(2, [0, 1, 2], [0, 1], [1.0, 0.0])


def test_tracer_operand_what_largediag_does_to_it(golden_matrix, reftest_matrix):
    """The reference's SuperLU_DIST call permutes rows of this operator family: with centred advection many rows
    carry their largest entry off the diagonal.  Recorded here because it decides the default."""
    for m, lo in ((golden_matrix, 500), (reftest_matrix, 50)):
        n = m["n"]
        rowmap, R, Cs = solver.rowperm_largediag(n, m["rowptr"], m["colind"], m["nzval_row_wise"])
        moved = int((rowmap != np.arange(n)).sum())
        assert lo <= moved < n
        A = oracle_solve.csr(n, m["rowptr"], m["colind"], m["nzval_row_wise"])
        d_old = np.abs(A.diagonal())
        d_new = np.abs(np.asarray(A[np.arange(n), rowmap]).ravel())
        assert np.sum(np.log(d_new)) > np.sum(np.log(d_old))


@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_plan_on_the_permuted_pattern_matches_oracle(sim_lib, golden_matrix, golden_rhs, nranks):
    """Ordering / symbolic phase / scatter map built for Pr A: the interpreted plan solves the ORIGINAL system."""
    m = golden_matrix
    n = m["n"]
    rowmap, R, Cs = solver.rowperm_largediag(n, m["rowptr"], m["colind"], m["nzval_row_wise"])
    rc, X, stats = sim_rowperm(sim_lib, n, m["rowptr"], m["colind"], m["nzval_row_wise"], _coords(m), golden_rhs["B"],
                               rowmap, pow2(R), pow2(Cs), nranks=nranks)
    assert rc == 0
    rel = np.linalg.norm(X - golden_rhs["X"], axis=0) / np.linalg.norm(golden_rhs["X"], axis=0)
    assert rel.max() <= 1e-8, rel
    assert stats[6] == 0


def test_shuffled_rows_need_the_row_permutation(sim_lib, reftest_matrix, reftest_rhs):
    """Rows of the operand in a scrambled order (zero / tiny diagonals): static pivoting alone is lost, with the
    LargeDiag permutation the same plan machinery recovers the pivoted oracle's answer."""
    m = reftest_matrix
    n = m["n"]
    rng = np.random.default_rng(3)
    A = oracle_solve.csr(n, m["rowptr"], m["colind"], m["nzval_row_wise"])
    q = rng.permutation(n)
    As = A[q, :].tocsr()
    As.sort_indices()
    B = np.asfortranarray(reftest_rhs["B"][q, :])          # same equations, scrambled: the solution is unchanged
    rp, ci, nz = As.indptr.astype(np.int32), As.indices.astype(np.int32), As.data.copy()
    coords = _coords(m)
    rc0, X0, st0 = sim_rowperm(sim_lib, n, rp, ci, nz, coords, B)
    bad = rc0 != 0 or st0[6] > 0 or not np.all(np.isfinite(X0)) or \
        (np.linalg.norm(X0 - reftest_rhs["X"], axis=0) / np.linalg.norm(reftest_rhs["X"], axis=0)).max() > 1e-3
    assert bad, "static pivoting was expected to fail on scrambled rows"
    rowmap, R, Cs = solver.rowperm_largediag(n, rp, ci, nz)
    rc, X, st = sim_rowperm(sim_lib, n, rp, ci, nz, coords, B, rowmap, pow2(R), pow2(Cs))
    assert rc == 0 and st[6] == 0
    rel = np.linalg.norm(X - reftest_rhs["X"], axis=0) / np.linalg.norm(reftest_rhs["X"], axis=0)
    assert rel.max() <= 1e-8, rel


def test_power_of_two_scalings_do_not_change_the_arithmetic(sim_lib, golden_matrix, golden_rhs):
    """Without row exchanges LU is invariant under power-of-two scalings bit for bit: the MC64 scalings matter only
    for the tiny-pivot threshold.  (Why nkp_create_rowperm rounds them, and why the default equilibration suffices.)"""
    m = golden_matrix
    n = m["n"]
    rowmap, R, Cs = solver.rowperm_largediag(n, m["rowptr"], m["colind"], m["nzval_row_wise"])
    args = (sim_lib, n, m["rowptr"], m["colind"], m["nzval_row_wise"], _coords(m), golden_rhs["B"][:, :1])
    _, X1, _ = sim_rowperm(*args, rowmap, pow2(R), pow2(Cs))
    _, X2, _ = sim_rowperm(*args, rowmap, np.ones(n), np.ones(n))
    assert np.array_equal(X1, X2)


def test_rowmap_must_be_a_permutation(sim_lib, golden_matrix, golden_rhs):
    m = golden_matrix
    n = m["n"]
    rowmap = np.zeros(n, dtype=np.int32)
    rc, _, _ = sim_rowperm(sim_lib, n, m["rowptr"], m["colind"], m["nzval_row_wise"], _coords(m), golden_rhs["B"][:, :1], rowmap)
    assert rc == -8
