"""Operand side of the path: the CRS matrix file (src/matrix.c:3845-4031), the NetCDF
provider underneath the reference's file layer, and the bit-exactness of the synthetic
generator against the reference's own gen_A."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, synth_case
from nk_ocn_tracer_jacobian_precond_b200 import synth

GEN_A = os.path.join(ROOT, "oracle", "_ref", "gen_A")


def test_golden_matrix_known_answers(golden_matrix):
    """KAT-1: rows sorted, diagonal present and non-zero, no explicit zeros,
    nnz within the comp_nnz bound of src/matrix.c:612-651 (7-point stencil here)."""
    m = golden_matrix
    n, rp, ci, nz = m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"]
    assert rp[0] == 0 and rp[-1] == len(ci) == len(nz)
    assert int(m["coupled_tracer_cnt"]) == 1
    assert n == int(m["KMT"].sum())
    assert np.all(nz != 0.0)
    for r in range(n):
        cols = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(cols) > 0)
        assert r in cols
        assert len(cols) <= 7
    # index maps: j outer, i middle, k inner (src/matrix.c:239-251)
    i, j, k = m["tracer_state_ind_to_i"], m["tracer_state_ind_to_j"], m["tracer_state_ind_to_k"]
    key = (j.astype(np.int64) * m["imt"] + i) * m["km"] + k
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(m["int3_to_tracer_state_ind"][k, j, i], np.arange(n))


def test_synth_assembler_bit_exact_vs_golden(golden_matrix):
    """numpy restatement of gen_sparse_matrix == reference gen_A output, bit for bit."""
    c = synth_case(20, 24, 10, seed=1)
    assert c["n"] == golden_matrix["n"]
    assert np.array_equal(c["rowptr"], golden_matrix["rowptr"])
    assert np.array_equal(c["colind"], golden_matrix["colind"])
    assert np.array_equal(c["nzval"], golden_matrix["nzval_row_wise"])  # exact, not allclose


@pytest.mark.skipif(not os.path.exists(GEN_A), reason="oracle/_ref/gen_A not built (reference tree absent)")
@pytest.mark.parametrize("shape,seed", [((20, 24, 10), 1), ((12, 10, 5), 7), ((36, 30, 16), 3)])
def test_reference_gen_A_matches_assembler(tmp_path, shape, seed):
    """Run the reference's unchanged gen_A (built in place from /root/reference/src) on a
    freshly written circulation file and compare its CRS with the numpy assembler."""
    c = synth_case(*shape, seed=seed)
    circ = tmp_path / "circ.nc"
    synth.write_circ_file(str(circ), c["grid"], c["circ"])
    (tmp_path / "opts.txt").write_text(synth.MINIMAL_OPTS.format(circ=str(circ)))
    subprocess.check_call([GEN_A, "-o", str(tmp_path / "opts.txt"), str(tmp_path / "A.nc")])
    m = synth.read_matrix_file(str(tmp_path / "A.nc"))
    assert np.array_equal(c["rowptr"], m["rowptr"])
    assert np.array_equal(c["colind"], m["colind"])
    assert np.array_equal(c["nzval"], m["nzval_row_wise"])
    assert np.array_equal(c["i"], m["tracer_state_ind_to_i"])
    assert np.array_equal(c["k"], m["tracer_state_ind_to_k"])


@pytest.mark.skipif(not os.path.exists(GEN_A), reason="oracle/_ref/gen_A not built")
def test_reference_gen_A_reproduces_golden(tmp_path):
    for f in ("circ_20x24x10.nc", "opts_20x24x10.txt"):
        shutil.copy(os.path.join(GOLDEN, f), tmp_path / f)
    subprocess.check_call([GEN_A, "-o", "opts_20x24x10.txt", "A.nc"], cwd=tmp_path)
    a = synth.read_matrix_file(str(tmp_path / "A.nc"))
    b = synth.read_matrix_file(os.path.join(GOLDEN, "A_20x24x10.nc"))
    for k in ("nzval_row_wise", "colind", "rowptr", "KMT", "int3_to_tracer_state_ind"):
        assert np.array_equal(a[k], b[k]), k


# ---- NetCDF-3 provider through its C ABI -------------------------------------------------


def _nc3():
    path = os.path.join(ROOT, "nk_ocn_tracer_jacobian_precond_b200", "libnkp_nc3.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "nk_ocn_tracer_jacobian_precond_b200", "csrc"),
                               path])
    lib = ctypes.CDLL(path)
    lib.nc_strerror.restype = ctypes.c_char_p
    return lib


def test_nc3_create_redef_roundtrip(tmp_path):
    """create -> close -> put -> reopen + redef (header grows, data relocates:
    src/matrix.c:288,3865) -> put; then read everything back with scipy."""
    from scipy.io import netcdf_file
    lib = _nc3()
    fn = str(tmp_path / "t.nc").encode()
    ncid = ctypes.c_int()
    did = ctypes.c_int()
    vid = ctypes.c_int()
    NC_64BIT_OFFSET, NC_WRITE, NC_INT, NC_DOUBLE = 0x0200, 1, 4, 6
    assert lib.nc_create(fn, NC_64BIT_OFFSET, ctypes.byref(ncid)) == 0
    assert lib.nc_def_dim(ncid, b"x", ctypes.c_size_t(5), ctypes.byref(did)) == 0
    dims = (ctypes.c_int * 1)(did.value)
    assert lib.nc_def_var(ncid, b"a", NC_DOUBLE, 1, dims, ctypes.byref(vid)) == 0
    assert lib.nc_put_att_text(ncid, vid, b"units", ctypes.c_size_t(2), b"cm") == 0
    assert lib.nc_close(ncid) == 0
    a = np.arange(5, dtype=np.float64) * 1.5
    assert lib.nc_open(fn, NC_WRITE, ctypes.byref(ncid)) == 0
    assert lib.nc_inq_varid(ncid, b"a", ctypes.byref(vid)) == 0
    assert lib.nc_put_var_double(ncid, vid, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    assert lib.nc_close(ncid) == 0
    # grow the header with a long-named variable + attributes, data of "a" must survive
    assert lib.nc_open(fn, NC_WRITE, ctypes.byref(ncid)) == 0
    assert lib.nc_redef(ncid) == 0
    d2 = ctypes.c_int()
    assert lib.nc_def_dim(ncid, b"a_rather_long_dimension_name_to_grow_the_header", ctypes.c_size_t(3), ctypes.byref(d2)) == 0
    dims = (ctypes.c_int * 1)(d2.value)
    v2 = ctypes.c_int()
    assert lib.nc_def_var(ncid, b"another_quite_long_variable_name", NC_INT, 1, dims, ctypes.byref(v2)) == 0
    m1 = (ctypes.c_int * 1)(-1)
    assert lib.nc_put_att_int(ncid, v2, b"_FillValue", NC_INT, ctypes.c_size_t(1), m1) == 0
    assert lib.nc_close(ncid) == 0
    b = np.array([7, -8, 9], dtype=np.int32)
    assert lib.nc_open(fn, NC_WRITE, ctypes.byref(ncid)) == 0
    assert lib.nc_inq_varid(ncid, b"another_quite_long_variable_name", ctypes.byref(v2)) == 0
    assert lib.nc_put_var_int(ncid, v2, b.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    assert lib.nc_inq_varid(ncid, b"nope", ctypes.byref(v2)) == -49  # NC_ENOTVAR
    assert lib.nc_close(ncid) == 0
    f = netcdf_file(fn.decode(), "r", mmap=False)
    assert f.version_byte == 2
    assert np.array_equal(f.variables["a"].data, a)
    assert f.variables["a"].units == b"cm"
    assert np.array_equal(f.variables["another_quite_long_variable_name"].data, b)
    assert f.variables["another_quite_long_variable_name"]._FillValue == -1
    f.close()


def test_nc3_reads_scipy_written_file(tmp_path):
    from scipy.io import netcdf_file
    lib = _nc3()
    fn = str(tmp_path / "s.nc")
    f = netcdf_file(fn, "w", version=2)
    f.createDimension("z", 4)
    v = f.createVariable("T", "d", ("z",))
    v[:] = [1.0, -2.5, 3.25, 1e300]
    v._FillValue = np.float64(9.96921e36)
    w = f.createVariable("K", "i", ("z",))
    w[:] = [1, 2, 3, -4]
    f.close()
    ncid = ctypes.c_int()
    vid = ctypes.c_int()
    assert lib.nc_open(fn.encode(), 0, ctypes.byref(ncid)) == 0
    out = np.zeros(4)
    assert lib.nc_inq_varid(ncid, b"T", ctypes.byref(vid)) == 0
    assert lib.nc_get_var_double(ncid, vid, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) == 0
    assert np.array_equal(out, [1.0, -2.5, 3.25, 1e300])
    fv = ctypes.c_double()
    assert lib.nc_get_att_double(ncid, vid, b"_FillValue", ctypes.byref(fv)) == 0
    assert fv.value == 9.96921e36
    assert lib.nc_get_att_double(ncid, vid, b"missing", ctypes.byref(fv)) == -43  # NC_ENOTATT
    ki = np.zeros(4, dtype=np.int32)
    assert lib.nc_inq_varid(ncid, b"K", ctypes.byref(vid)) == 0
    assert lib.nc_get_var_int(ncid, vid, ki.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    assert np.array_equal(ki, [1, 2, 3, -4])
    ln = ctypes.c_size_t()
    did = ctypes.c_int()
    assert lib.nc_inq_dimid(ncid, b"z", ctypes.byref(did)) == 0
    assert lib.nc_inq_dimlen(ncid, did, ctypes.byref(ln)) == 0 and ln.value == 4
    assert lib.nc_close(ncid) == 0
    assert lib.nc_open(str(tmp_path / "missing.nc").encode(), 0, ctypes.byref(ncid)) != 0


def _scipy_record_file(fn, nrec):
    from scipy.io import netcdf_file
    f = netcdf_file(fn, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("z", 3)
    fx = f.createVariable("fixed", "d", ("z",))
    fx[:] = [7.0, 8.0, 9.0]
    v = f.createVariable("T", "d", ("time", "z"))
    w = f.createVariable("S", "d", ("time", "z"))
    for r in range(nrec):
        v[r] = [1.0 + 10 * r, 2.0 + 10 * r, 3.0 + 10 * r]
        w[r] = [-1.0 - 10 * r, -2.0 - 10 * r, -3.0 - 10 * r]
    f.close()


def test_nc3_record_variable_one_record_read_write_redef(tmp_path):
    """A POP tracer file whose fields carry an unlimited time dimension of length 1 works with the
    reference on real libnetcdf (get_var_3d_double / put_var_3d_double, src/file_io.c:273-334): the
    provider reads AND writes such a variable, and nc_redef + nc_enddef relocates it intact."""
    from scipy.io import netcdf_file
    lib = _nc3()
    fn = str(tmp_path / "rec1.nc")
    _scipy_record_file(fn, 1)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    ncid, vid, sid = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.nc_open(fn.encode(), 1, ctypes.byref(ncid)) == 0      # NC_WRITE
    assert lib.nc_inq_varid(ncid, b"T", ctypes.byref(vid)) == 0
    out = np.zeros(3)
    assert lib.nc_get_var_double(ncid, vid, dp(out)) == 0 and np.array_equal(out, [1.0, 2.0, 3.0])
    new = np.array([0.5, 0.25, 0.125])
    assert lib.nc_put_var_double(ncid, vid, dp(new)) == 0
    # grow the header: fixed and record data both move, nothing may be lost or overwritten
    assert lib.nc_redef(ncid) == 0
    d2, v2 = ctypes.c_int(), ctypes.c_int()
    assert lib.nc_def_dim(ncid, b"a_long_dimension_name_that_grows_the_header_by_a_lot", ctypes.c_size_t(4), ctypes.byref(d2)) == 0
    dims = (ctypes.c_int * 1)(d2.value)
    assert lib.nc_def_var(ncid, b"extra_fixed_variable_with_a_long_name", 6, 1, dims, ctypes.byref(v2)) == 0
    assert lib.nc_enddef(ncid) == 0
    assert lib.nc_close(ncid) == 0
    f = netcdf_file(fn, "r", mmap=False)
    assert np.array_equal(f.variables["fixed"].data, [7.0, 8.0, 9.0])
    assert np.array_equal(f.variables["T"].data, [[0.5, 0.25, 0.125]])
    assert np.array_equal(f.variables["S"].data, [[-1.0, -2.0, -3.0]])
    f.close()


def test_nc3_refuses_interleaved_records_without_damage(tmp_path):
    """numrecs > 1: whole-variable get/put and the relocation of nc_redef are refused, and the file is
    left byte for byte as it was (no truncation, no partial move)."""
    lib = _nc3()
    fn = str(tmp_path / "rec2.nc")
    _scipy_record_file(fn, 2)
    before = open(fn, "rb").read()
    ncid, vid = ctypes.c_int(), ctypes.c_int()
    assert lib.nc_open(fn.encode(), 1, ctypes.byref(ncid)) == 0
    assert lib.nc_inq_varid(ncid, b"T", ctypes.byref(vid)) == 0
    buf = np.zeros(6)
    assert lib.nc_get_var_double(ncid, vid, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) != 0
    assert lib.nc_put_var_double(ncid, vid, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double))) != 0
    assert lib.nc_redef(ncid) == 0
    d2 = ctypes.c_int()
    assert lib.nc_def_dim(ncid, b"another_long_dimension_name_to_grow_the_header", ctypes.c_size_t(4), ctypes.byref(d2)) == 0
    assert lib.nc_enddef(ncid) != 0
    lib.nc_abort(ncid) if hasattr(lib, "nc_abort") else lib.nc_close(ncid)
    assert open(fn, "rb").read() == before


def test_reftest_golden_known_answers(reftest_matrix):
    """The reference's own test options (test/test_gen_A.csh:22-23): upwind3 + isop_file rows have
    up to 21 entries (src/matrix.c:621-650), sorted, diagonal present, zeros stripped."""
    m = reftest_matrix
    n, rp, ci, nz = m["n"], m["rowptr"], m["colind"], m["nzval_row_wise"]
    assert rp[-1] == len(nz) and np.all(nz != 0.0)
    lens = np.diff(rp)
    assert 7 < lens.max() <= 21
    for r in range(n):
        cols = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(cols) > 0) and r in cols


@pytest.mark.skipif(not os.path.exists(GEN_A), reason="oracle/_ref/gen_A not built")
def test_reference_gen_A_reproduces_reftest_golden(tmp_path, reftest_matrix):
    c = synth_case(20, 24, 10, seed=1)
    full = synth.make_full_fields(c["grid"], c["circ"], seed=1)
    circ = tmp_path / "circ_full.nc"
    synth.write_circ_file(str(circ), c["grid"], c["circ"], full)
    (tmp_path / "opts.txt").write_text(synth.REFTEST_OPTS.format(circ=str(circ)))
    subprocess.check_call([GEN_A, "-o", str(tmp_path / "opts.txt"), str(tmp_path / "A.nc")])
    a = synth.read_matrix_file(str(tmp_path / "A.nc"))
    for k in ("nzval_row_wise", "colind", "rowptr"):
        assert np.array_equal(a[k], reftest_matrix[k]), k


def test_raw_variable_extent_for_device_ingest(golden_matrix):
    """nkp_nc3_inq_var_extent (include/nkp_nc3.h): the bytes at the reported extent are the big-endian
    doubles nc_get_var_double would have swapped on the host (src/matrix.c:3996) -- the input of nkp_factor_be."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "nk_ocn_tracer_jacobian_precond_b200", "libnkp_nc3.so"))
    path = os.path.join(GOLDEN, "A_20x24x10.nc")
    ncid, varid = ctypes.c_int(), ctypes.c_int()
    assert lib.nc_open(path.encode(), 0, ctypes.byref(ncid)) == 0
    assert lib.nc_inq_varid(ncid, b"nzval_row_wise", ctypes.byref(varid)) == 0
    off, nbytes, xtype = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_int()
    assert lib.nkp_nc3_inq_var_extent(ncid, varid, ctypes.byref(off), ctypes.byref(nbytes), ctypes.byref(xtype)) == 0
    nz = golden_matrix["nzval_row_wise"]
    assert xtype.value == 6 and nbytes.value == 8 * nz.size          # NC_DOUBLE
    with open(path, "rb") as f:
        f.seek(off.value)
        raw = f.read(nbytes.value)
    assert np.array_equal(np.frombuffer(raw, dtype=">f8").astype(np.float64), nz)
    assert lib.nkp_nc3_inq_var_extent(ncid, 9999, None, None, None) != 0
    assert lib.nc_close(ncid) == 0
