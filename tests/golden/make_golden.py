"""Regenerates the golden fixtures (run HERE, where /root/reference exists):

  circ_20x24x10.nc   synthetic POP-style circulation file (nk..b200.synth, seed 1)
  A_20x24x10.nc      matrix file written by the reference's UNCHANGED gen_A
                     (oracle/_ref/gen_A, built by oracle/Makefile from /root/reference/src)
                     with the option set of SURVEY.md Appendix B "minimal-input"
  opts_20x24x10.txt  the gen_A option file used
  rhs_x_20x24x10.npz seeded right-hand sides and the oracle's solutions (scipy SuperLU +
                     pdgsrfs refinement, oracle/oracle_solve.py)
  A_reftest_20x24x10.nc / rhs_x_reftest_20x24x10.npz
                     the same for the reference's own test option set (upwind3 + isop_file +
                     vmix file, test/test_gen_A.csh:22-23): 19-point rows
"""
import os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from nk_ocn_tracer_jacobian_precond_b200 import synth
from oracle import oracle_solve

def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])
    g = synth.make_grid(20, 24, 10, seed=1)
    c = synth.make_circulation(g, seed=1)
    os.chdir(HERE)
    synth.write_circ_file("circ_20x24x10.nc", g, c)
    open("opts_20x24x10.txt", "w").write(synth.MINIMAL_OPTS.format(circ="circ_20x24x10.nc"))
    subprocess.check_call([os.path.join(ROOT, "oracle/_ref/gen_A"), "-o", "opts_20x24x10.txt", "A_20x24x10.nc"])
    m = synth.read_matrix_file("A_20x24x10.nc")
    n = len(m["rowptr"]) - 1
    rng = np.random.default_rng(123)
    B = np.asfortranarray(rng.standard_normal((n, 3)))
    X, info = oracle_solve.solve(n, m["rowptr"], m["colind"], m["nzval_row_wise"], B, return_info=True)
    np.savez_compressed("rhs_x_20x24x10.npz", B=B, X=X, berr=np.array([i[0] for i in info]))
    print("golden written: n =", n, "nnz =", len(m["colind"]), "oracle berr", [i[0] for i in info])

    # the reference's own test option set (test/test_gen_A.csh:22-23): upwind3 + isop_file + vmix file.
    # The 2 MB circulation file is not committed; it is regenerated from the seed.
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        full = synth.make_full_fields(g, c, seed=1)
        circ = os.path.join(td, "circ_full.nc")
        synth.write_circ_file(circ, g, c, full)
        opts = os.path.join(td, "opts.txt")
        open(opts, "w").write(synth.REFTEST_OPTS.format(circ=circ))
        subprocess.check_call([os.path.join(ROOT, "oracle/_ref/gen_A"), "-o", opts, os.path.join(HERE, "A_reftest_20x24x10.nc")])
    m = synth.read_matrix_file("A_reftest_20x24x10.nc")
    n = len(m["rowptr"]) - 1
    B = np.asfortranarray(np.random.default_rng(321).standard_normal((n, 2)))
    X, info = oracle_solve.solve(n, m["rowptr"], m["colind"], m["nzval_row_wise"], B, return_info=True)
    np.savez_compressed("rhs_x_reftest_20x24x10.npz", B=B, X=X, berr=np.array([i[0] for i in info]))
    print("golden (reference test options) written: n =", n, "nnz =", len(m["colind"]), "oracle berr", [i[0] for i in info])

if __name__ == "__main__":
    main()
