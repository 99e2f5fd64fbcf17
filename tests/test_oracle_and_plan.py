"""CPU side: the oracle against the golden vectors, and the host analysis (ordering,
symbolic structure, memory plan, task lists) validated by interpreting the plan on the
CPU (oracle/plan_sim.cpp) and comparing with the oracle."""
import ctypes
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import synth_case
from oracle import oracle_solve

P = ctypes.POINTER


def _ip(a):
    return a.ctypes.data_as(P(ctypes.c_int)) if a is not None else None


def _dp(a):
    return a.ctypes.data_as(P(ctypes.c_double))


def run_sim(lib, n, rp, ci, nz, coords, B, nb=64, leaf=96, analysis_only=0):
    B = np.asfortranarray(B, dtype=np.float64)
    nrhs = B.shape[1]
    X = np.zeros_like(B, order="F")
    stats = np.zeros(8)
    perm = np.zeros(n, dtype=np.int32)
    i, j, k = coords if coords is not None else (None, None, None)
    rc = lib.nkp_sim_run(n, _ip(rp), _ip(ci), _dp(nz), _ip(i), _ip(j), _ip(k), nb, leaf, _dp(B), nrhs, _dp(X),
                         _dp(stats), _ip(perm), analysis_only)
    assert rc == 0, rc
    return X, stats, perm


def test_oracle_reproduces_golden_solutions(golden_matrix, golden_rhs):
    m = golden_matrix
    X, info = oracle_solve.solve(m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], golden_rhs["B"], return_info=True)
    assert np.allclose(X, golden_rhs["X"], rtol=1e-11, atol=0)
    A = oracle_solve.csr(m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"])
    for c in range(X.shape[1]):
        berr, _ = oracle_solve.berr_of(A, X[:, c], golden_rhs["B"][:, c])
        assert berr <= 4 * oracle_solve.EPS


def test_oracle_manufactured_solution(golden_matrix):
    """KAT-2: b = A x*, recover x* to <= 1e-8 (BASELINE.json tolerance), residual <= 1e-10."""
    m = golden_matrix
    n = m["n"]
    rng = np.random.default_rng(5)
    xs = rng.standard_normal(n)
    b = oracle_solve.spmv(n, m["rowptr"], m["colind"], m["nzval_row_wise"], xs)
    x = oracle_solve.solve(n, m["rowptr"], m["colind"], m["nzval_row_wise"], b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) <= 1e-8
    r = oracle_solve.spmv(n, m["rowptr"], m["colind"], m["nzval_row_wise"], x) - b
    assert np.linalg.norm(r) / np.linalg.norm(b) <= 1e-10


@pytest.mark.parametrize("nb,leaf", [(64, 96), (16, 32), (8, 400)])
def test_plan_interpreter_matches_oracle_on_golden(sim_lib, golden_matrix, golden_rhs, nb, leaf):
    m = golden_matrix
    coords = (m["tracer_state_ind_to_i"], m["tracer_state_ind_to_j"], m["tracer_state_ind_to_k"])
    X, stats, perm = run_sim(sim_lib, m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], coords, golden_rhs["B"],
                             nb=nb, leaf=leaf)
    assert sorted(perm.tolist()) == list(range(m["n"]))
    # static pivoting without refinement: a few digits worse than the pivoted oracle
    rel = np.linalg.norm(X - golden_rhs["X"], axis=0) / np.linalg.norm(golden_rhs["X"], axis=0)
    assert rel.max() <= 1e-8, rel
    assert stats[6] == 0  # no tiny pivots replaced


@pytest.mark.parametrize("shape", [(12, 10, 5), (24, 28, 16), (30, 34, 20)])
def test_plan_interpreter_multistep_fronts(sim_lib, shape):
    """Fronts with several pivot blocks exercise the tile-skip rules and partial blocks."""
    c = synth_case(*shape, seed=2)
    n = c["n"]
    A = sp.csr_matrix((c["nzval"], c["colind"], c["rowptr"]), shape=(n, n))
    A = (A - 50.0 * sp.eye(n)).tocsr()   # well conditioned: isolates plan logic from pivot growth
    A.sort_indices()
    rng = np.random.default_rng(1)
    xs = rng.standard_normal((n, 2))
    B = A @ xs
    X, stats, _ = run_sim(sim_lib, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data,
                          (c["i"], c["j"], c["k"]), B)
    assert np.linalg.norm(X - xs) / np.linalg.norm(xs) <= 1e-12


def test_plan_without_coordinates_uses_graph_dissection(sim_lib):
    c = synth_case(16, 14, 8, seed=4)
    n = c["n"]
    A = sp.csr_matrix((c["nzval"], c["colind"], c["rowptr"]), shape=(n, n))
    A = (A - 50.0 * sp.eye(n)).tocsr()
    A.sort_indices()
    xs = np.random.default_rng(2).standard_normal((n, 1))
    X, stats, perm = run_sim(sim_lib, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data, None, A @ xs,
                             leaf=48)
    assert sorted(perm.tolist()) == list(range(n))
    assert stats[0] > 4  # really dissected
    assert np.linalg.norm(X - xs) / np.linalg.norm(xs) <= 1e-12


def test_analysis_reports_fill_and_flops(sim_lib):
    c = synth_case(40, 46, 24, seed=1)
    _, stats, _ = run_sim(sim_lib, c["n"], c["rowptr"], c["colind"], c["nzval"], (c["i"], c["j"], c["k"]),
                          np.zeros((c["n"], 1)), analysis_only=1)
    nfronts, nlevels, maxfront, nnz_lu, heap, flops = stats[:6]
    assert nnz_lu >= len(c["nzval"])
    assert heap >= nnz_lu
    assert flops > 0 and nlevels >= 5 and maxfront < c["n"]


def test_plan_interpreter_on_reference_test_options(sim_lib, reftest_matrix, reftest_rhs):
    """19-point rows with +-2 reach (upwind3) and diagonal (k+-1,i+-1) couplings (isop): the
    graph-derived separators must still separate, and the answer must match the oracle."""
    m = reftest_matrix
    coords = (m["tracer_state_ind_to_i"], m["tracer_state_ind_to_j"], m["tracer_state_ind_to_k"])
    X, stats, perm = run_sim(sim_lib, m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], coords, reftest_rhs["B"],
                             nb=64, leaf=64)
    rel = np.linalg.norm(X - reftest_rhs["X"], axis=0) / np.linalg.norm(reftest_rhs["X"], axis=0)
    assert rel.max() <= 1e-8, rel
    Xo = oracle_solve.solve(m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], reftest_rhs["B"])
    assert np.allclose(Xo, reftest_rhs["X"], rtol=1e-11, atol=0)


def test_ordering_cache_round_trip(sim_lib, golden_matrix, golden_rhs, tmp_path):
    """the nested-dissection tree is stored on disk under a pattern hash and read
    back by a later analysis; the cached plan is the same plan (identical permutation, statistics and
    solution); a different pattern gets its own file; a damaged file only means 'recompute'."""
    import ctypes
    m = golden_matrix
    coords = (m["tracer_state_ind_to_i"], m["tracer_state_ind_to_j"], m["tracer_state_ind_to_k"])
    args = (m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], coords, golden_rhs["B"])
    sim_lib.nkp_sim_set_cache_dir.argtypes = [ctypes.c_char_p]
    sim_lib.nkp_sim_set_cache_dir(None)
    X0, st0, p0 = run_sim(sim_lib, *args)
    assert sim_lib.nkp_sim_last_order_cached() == 0
    cache = tmp_path / "cache"
    cache.mkdir()
    sim_lib.nkp_sim_set_cache_dir(str(cache).encode())
    try:
        X1, st1, p1 = run_sim(sim_lib, *args)                 # miss: computes and stores
        assert sim_lib.nkp_sim_last_order_cached() == 0
        files = sorted(cache.iterdir())
        assert len(files) == 1 and files[0].name.startswith("nkp_order_")
        X2, st2, p2 = run_sim(sim_lib, *args)                 # hit
        assert sim_lib.nkp_sim_last_order_cached() == 1
        assert np.array_equal(p0, p1) and np.array_equal(p0, p2)
        assert np.array_equal(st0[:6], st2[:6])
        assert np.array_equal(X0, X2)                          # same plan, same arithmetic: bitwise
        # without coordinates the ordering differs, and so does the key
        run_sim(sim_lib, m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], None, golden_rhs["B"])
        assert sim_lib.nkp_sim_last_order_cached() == 0 and len(list(cache.iterdir())) == 2
        # a truncated file is ignored and rewritten
        data = files[0].read_bytes()
        files[0].write_bytes(data[: len(data) // 2])
        X3, _, p3 = run_sim(sim_lib, *args)
        assert sim_lib.nkp_sim_last_order_cached() == 0 and np.array_equal(p0, p3)
        assert files[0].stat().st_size == len(data)
    finally:
        sim_lib.nkp_sim_set_cache_dir(None)


@pytest.mark.parametrize("shape,nranks,leaf", [((20, 24, 10), 1, 96), ((40, 46, 24), 1, 96), ((40, 46, 24), 3, 96),
                                               ((64, 74, 38), 2, 48)])
def test_plan_invariants_the_kernels_rely_on(sim_lib, shape, nranks, leaf):
    """Even leading dimensions and aligned panel offsets (16-byte cp.async), one inversion task per
    diagonal block, complete and dependency-ordered item lists of the dataflow sweeps, the child run
    starts of the forward slabs, the partial-product slots of the backward rectangles -- for one and for
    several ranks (oracle/plan_sim.cpp::nkp_sim_check_plan)."""
    c = synth_case(*shape, seed=2)
    rc = sim_lib.nkp_sim_check_plan(c["n"], _ip(c["rowptr"]), _ip(c["colind"]), _ip(c["i"]), _ip(c["j"]), _ip(c["k"]),
                                    64, leaf, nranks)
    assert rc == 0, f"plan invariant {rc} violated"


def test_analysis_is_independent_of_the_thread_count(tmp_path):
    """The dissection halves are OpenMP tasks; the tree numbering must not depend on scheduling (every
    rank of a multi-GPU run derives the same plan): same permutation with 1, 2 and 5 threads."""
    import subprocess, sys
    script = tmp_path / "perm.py"
    script.write_text(
        "import sys, ctypes, numpy as np\n"
        f"sys.path.insert(0, {repr(os.path.dirname(os.path.abspath(__file__)))})\n"
        f"sys.path.insert(0, {repr(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))})\n"
        "from conftest import synth_case\n"
        "from test_oracle_and_plan import run_sim\n"
        "lib = ctypes.CDLL(sys.argv[1])\n"
        "c = synth_case(64, 74, 38, seed=3)\n"
        "X, st, perm = run_sim(lib, c['n'], c['rowptr'], c['colind'], c['nzval'], (c['i'], c['j'], c['k']),\n"
        "                      np.zeros((c['n'], 1)), analysis_only=1)\n"
        "np.save(sys.argv[2], perm)\n")
    libpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "libnkp_sim.so")
    perms = []
    for nt in (1, 2, 5):
        out = tmp_path / f"perm{nt}.npy"
        subprocess.check_call([sys.executable, str(script), libpath, str(out)], env=dict(os.environ, OMP_NUM_THREADS=str(nt)))
        perms.append(np.load(out))
    assert np.array_equal(perms[0], perms[1]) and np.array_equal(perms[0], perms[2])


@pytest.mark.parametrize("shape", [(40, 46, 24), (64, 74, 38)])
def test_oracle_floor_of_the_manufactured_solution(shape):
    """Where is the floor of ||x - x*|| / ||x*|| for the PIVOTED oracle (scipy SuperLU + pdgsrfs refinement) when
    b = A x* is formed in working precision and when it is formed in extended precision and rounded once
    ?  At these sizes cond(A) * eps is ~1e-11..1e-10, both floors sit there, and the extended b
    is never worse.  The floor grows with the grid (cond ~ 1e8 at gx1v6-shape: ~1e-8, measured on the GPU in
    profiles/r02_refine_probe_gx1v6.log) -- the reason tests/test_gpu_parity.py::test_full_size_gx1v6_properties
    checks the solution DIFFERENCE of two factorisations at 1e-10 and x* at 3e-8."""
    import bench
    from conftest import synth_case
    c = synth_case(*shape, seed=1)
    n = c["n"]
    A = oracle_solve.csr(n, c["rowptr"], c["colind"], c["nzval"])
    xs = np.random.default_rng(0).standard_normal((n, 2))
    b_dbl = np.asfortranarray(A @ xs)
    b_ext = bench.spmv_extended(c["rowptr"], c["colind"], c["nzval"], xs)
    assert 0 < np.abs(b_dbl - b_ext).max() <= 4 * oracle_solve.EPS * np.abs(b_ext).max()
    lu = oracle_solve.factor(n, c["rowptr"], c["colind"], c["nzval"])
    err = {}
    for name, b in (("double", b_dbl), ("extended", b_ext)):
        X = oracle_solve.solve(n, c["rowptr"], c["colind"], c["nzval"], b, lu=lu)
        err[name] = float((np.linalg.norm(X - xs, axis=0) / np.linalg.norm(xs, axis=0)).max())
    print(f"oracle floor {shape}: n={n} double b {err['double']:.2e}, extended b {err['extended']:.2e}")
    assert err["double"] <= 1e-9 and err["extended"] <= 1e-9
    assert err["extended"] <= 1.5 * err["double"]


def test_crs_post_reproduces_gen_A(golden_matrix):
    """oracle/crs_post.py (restatement of sum_dup_vals / strip_matrix_zeros / sort_cols_all_rows, src/matrix.c:3621-3770)
    turns the pre-processing form of the golden operand into exactly the CRS the reference's gen_A wrote."""
    from conftest import synth_case
    from nk_ocn_tracer_jacobian_precond_b200 import synth
    from oracle import crs_post
    c = synth_case(20, 24, 10, seed=1)
    n, rp, ci, nz, _ = synth.assemble_crs(c["grid"], c["circ"], raw=True)
    assert np.any(np.diff(ci)[np.diff(np.repeat(np.arange(n), np.diff(rp))) == 0] < 0)     # really unsorted
    rp2, ci2, nz2, dup = crs_post.finalize(rp, ci, nz)
    assert dup == 0
    assert np.array_equal(rp2, golden_matrix["rowptr"]) and np.array_equal(ci2, golden_matrix["colind"])
    assert np.array_equal(nz2, golden_matrix["nzval_row_wise"])


def test_crs_post_known_answers():
    """Hand-checkable case: duplicates are summed INTO THE FIRST occurrence in order, a sum that cancels is stripped,
    zeros left by merged duplicates are stripped, columns end up ascending."""
    from oracle import crs_post
    rp = np.array([0, 5, 7, 9], dtype=np.int32)
    ci = np.array([2, 0, 2, 1, 2, 1, 1, 2, 0], dtype=np.int32)
    v = np.array([1e16, 3.0, 1.0, 0.0, -1e16, 2.0, -2.0, 5.0, 4.0])
    rp2, ci2, v2, dup = crs_post.finalize(rp, ci, v)
    # row 0: column 2 = (1e16 + 1.0) + -1e16 = 0.0 in double (the reference's order), stripped; explicit 0 stripped
    # the second and third entry of column 2 also match each other: dup_cnt counts that pair too (3 + 1)
    assert dup == 4
    assert rp2.tolist() == [0, 1, 1, 3] and ci2.tolist() == [0, 0, 2] and v2.tolist() == [3.0, 4.0, 5.0]
    rp3, ci3, v3, _ = crs_post.finalize(rp, ci, v, strip_zeros=False)
    assert rp3.tolist() == rp.tolist() and ci3.tolist() == [0, 1, 2, 2, 2, 1, 1, 0, 2]
    assert v3.tolist() == [3.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 4.0, 5.0]


def _true_counts(sim_lib, c, perm):
    cnt = np.zeros(c["n"], dtype=np.int64)
    rc = sim_lib.nkp_true_colcounts(c["n"], _ip(c["rowptr"]), _ip(c["colind"]), _ip(perm),
                                    cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    assert rc == 0
    return cnt


def test_independent_symbolic_factorisation_matches_plain_elimination(sim_lib):
    """oracle/plan_sim.cpp::nkp_true_colcounts (elimination tree + row subtrees) against a set-based symbolic
    elimination written out in Python: struct(j) = adj(j) above j  U  struct(children) minus j."""
    c = synth_case(12, 10, 5, seed=4)
    n = c["n"]
    _, _, perm = run_sim(sim_lib, n, c["rowptr"], c["colind"], c["nzval"], (c["i"], c["j"], c["k"]), np.zeros((n, 1)),
                         analysis_only=1)
    A = sp.csr_matrix((np.ones(len(c["colind"])), c["colind"], c["rowptr"]), shape=(n, n))
    S = (A + A.T).tocsr()
    order = np.argsort(perm)
    S = S[order, :][:, order].tocsr()
    S.sort_indices()
    struct, children, ref = [None] * n, [[] for _ in range(n)], np.zeros(n, dtype=np.int64)
    for j in range(n):
        row = S.indices[S.indptr[j]:S.indptr[j + 1]]
        st = set(int(x) for x in row[row > j])
        for ch in children[j]:
            st |= struct[ch]
            struct[ch] = None
        st.discard(j)
        ref[j] = len(st)
        if st:
            children[min(st)].append(j)
        struct[j] = st
    assert np.array_equal(_true_counts(sim_lib, c, perm), ref)


@pytest.mark.parametrize("shape", [(30, 34, 20), (40, 46, 24)])
def test_plan_flops_against_the_minimum_of_its_ordering(sim_lib, shape, monkeypatch):
    """The plan's flop and storage figures (dense fronts of the ordering actually used) against
    an INDEPENDENT symbolic factorisation of the same ordering: never below the minimum, and -- with the assembly tree
    taken from the elimination tree -- within a small factor of it.  One dense front per dissection node (the round-1
    tree, NKP_SUPERNODES=0) pays 2-3x: ocean subdomains cut by coordinate planes are often disconnected."""
    c = synth_case(*shape, seed=2)
    n = c["n"]
    ratios = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NKP_SUPERNODES", mode)
        _, stats, perm = run_sim(sim_lib, n, c["rowptr"], c["colind"], c["nzval"], (c["i"], c["j"], c["k"]), np.zeros((n, 1)),
                                 leaf=16, analysis_only=1)
        cnt = _true_counts(sim_lib, c, perm).astype(float)
        true_flops = float((2.0 * cnt ** 2 + cnt).sum())
        true_nnz = 2.0 * cnt.sum() + n
        assert stats[3] >= true_nnz and stats[5] >= 0.999 * true_flops
        ratios[mode] = (stats[5] / true_flops, stats[3] / true_nnz)
    assert ratios["1"][0] <= 1.7 and ratios["1"][1] <= 1.6, ratios   # measured 1.52 / 1.42 and 1.30 / 1.30 with leaf = 16; 1.07 / 1.15 at gx3v7-shape with the default leaf
    assert ratios["0"][0] >= 1.5 * ratios["1"][0], ratios


def test_degenerate_trees_through_the_plan_interpreter(sim_lib):
    """Shapes of elimination tree the ocean operand does not produce but a caller's matrix may: forests (diagonal and
    block-diagonal matrices), a star (arrow matrix: every leaf column is its own front), a path (tridiagonal), 1 x 1."""
    rng = np.random.default_rng(0)
    n = 300
    arrow = sp.lil_matrix((n, n))
    arrow.setdiag(4.0)
    arrow[n - 1, :] = 1.0
    arrow[:, n - 1] = 1.0
    arrow[n - 1, n - 1] = 400.0
    cases = {
        "diagonal": sp.diags(rng.uniform(1, 2, 10)),
        "1x1": sp.diags([3.0]),
        "block diagonal": sp.block_diag([sp.random(7, 7, 0.5, random_state=k) + 10 * sp.eye(7) for k in range(5)] + [2 * sp.eye(3)]),
        "arrow": arrow,
        "path": sp.diags([-np.ones(999), 4 * np.ones(1000), -np.ones(999)], [-1, 0, 1]),
    }
    for name, A in cases.items():
        A = sp.csr_matrix(A)
        A.sort_indices()
        m = A.shape[0]
        xs = rng.standard_normal((m, 2))
        X, stats, perm = run_sim(sim_lib, m, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64),
                                 None, A @ xs)
        assert sorted(perm.tolist()) == list(range(m)), name
        assert np.abs(X - xs).max() <= 1e-12, name


@pytest.mark.parametrize("nranks", [1, 4])
def test_dissection_node_fronts_remain_available(sim_lib, golden_matrix, golden_rhs, monkeypatch, nranks):
    """NKP_SUPERNODES=0: one dense front per dissection node (the assembly tree of round 1), kept for A/B runs."""
    monkeypatch.setenv("NKP_SUPERNODES", "0")
    monkeypatch.setenv("NKP_SIM_RANKS", str(nranks))
    m = golden_matrix
    coords = (m["tracer_state_ind_to_i"], m["tracer_state_ind_to_j"], m["tracer_state_ind_to_k"])
    X, stats, _ = run_sim(sim_lib, m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], coords, golden_rhs["B"])
    rel = np.linalg.norm(X - golden_rhs["X"], axis=0) / np.linalg.norm(golden_rhs["X"], axis=0)
    assert rel.max() <= 1e-8 and stats[6] == 0
    monkeypatch.setenv("NKP_SUPERNODES", "1")
    _, stats_new, _ = run_sim(sim_lib, m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"], coords, golden_rhs["B"])
    assert stats_new[5] < stats[5]      # the elimination-tree supernodes need fewer flops


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_patterns_through_analysis_and_interpreter(sim_lib, seed, monkeypatch):
    """Arbitrary sparsity (not the stencil family): random unsymmetric / partly symmetric patterns without coordinates,
    random panel width, leaf size, rank count and supernode relaxation -- ordering, elimination tree, supernodes, plan
    and the lockstep interpreter must solve every one of them (diagonally dominant, so static pivoting is safe)."""
    from test_rowperm import sim_rowperm
    rng = np.random.default_rng(seed)
    for trial in range(20):
        n = int(rng.integers(2, 400))
        A = sp.random(n, n, density=rng.choice([0.002, 0.01, 0.03, 0.1]), random_state=rng, format="csr")
        if rng.random() < 0.5:
            A = A + A.T * rng.random()
        A = sp.csr_matrix(A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0))
        A.sort_indices()
        xs = rng.standard_normal((n, 2))
        monkeypatch.setenv("NKP_RELAX_FRAC", str(rng.choice([0.0, 0.1, 0.3])))
        monkeypatch.setenv("NKP_RELAX_SMALL", str(int(rng.choice([0, 8, 32]))))
        rc, X, _ = sim_rowperm(sim_lib, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64),
                               None, A @ xs, nranks=int(rng.choice([1, 2, 3, 5])), nb=int(rng.choice([8, 16, 64])),
                               leaf=int(rng.choice([4, 16, 48, 96])))
        assert rc == 0, (seed, trial, rc)
        assert np.abs(X - xs).max() <= 1e-9, (seed, trial)
