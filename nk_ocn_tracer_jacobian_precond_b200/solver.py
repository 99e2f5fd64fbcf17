"""ctypes binding of the C ABI in include/nkprecond.h (libnkprecond.so).

This is host-side plumbing only: every numeric operation happens in the CUDA library.
If the library (or a GPU) is missing the calls raise -- there is no CPU fallback.

The class mirrors the two-phase protocol of the reference drivers
(src/solve_ABglobal.c:350-395): one ``factor`` (pdgssvx_ABglobal with nrhs=0) followed by
any number of ``solve`` calls (options.Fact = FACTORED, nrhs >= 1), B overwritten by X.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NKP_LIB", os.path.join(_HERE, "libnkprecond.so"))   # NKP_LIB: developer override


class NkpOptions(C.Structure):
    _fields_ = [("nb", C.c_int), ("leaf", C.c_int), ("equil", C.c_int), ("refine_max", C.c_int),
                ("device", C.c_int), ("verbose", C.c_int), ("refine_rule", C.c_int), ("residual_extra", C.c_int),
                ("reserved", C.c_int * 8)]


class NkpMinFields(C.Structure):
    _fields_ = [("imt", C.c_int), ("jmt", C.c_int), ("km", C.c_int), ("n", C.c_int)] + \
               [(k, C.c_void_p) for k in ("KMT", "ind_i", "ind_j", "ind_k", "int3_to_tracer_state_ind", "dz", "z_t", "TAREA",
                                          "HTE", "HUS", "HTN", "HUW", "DXU", "DYU", "UVEL", "VVEL", "WVEL")] + \
               [("fill_value", C.c_double)]


class NkpStats(C.Structure):
    _fields_ = [("n", C.c_int), ("nnz", C.c_int64), ("n_fronts", C.c_int), ("n_levels", C.c_int),
                ("max_front", C.c_int), ("nnz_lu", C.c_int64), ("factor_flops", C.c_double),
                ("heap_bytes", C.c_double), ("t_analysis", C.c_double), ("t_factor", C.c_double),
                ("t_scatter", C.c_double), ("t_solve", C.c_double), ("refine_steps", C.c_int),
                ("tiny_pivots", C.c_int), ("kernel_launches", C.c_int64), ("solve_bytes", C.c_double),
                ("t_gemm", C.c_double), ("gemm_flops", C.c_double), ("n_gemm", C.c_int64),
                ("t_trsm", C.c_double), ("t_diag", C.c_double), ("t_extend_add", C.c_double),
                ("t_sweeps", C.c_double), ("factor_flops_local", C.c_double), ("nnz_lu_local", C.c_double),
                ("n_xfers", C.c_double), ("order_cached", C.c_double), ("reserved", C.c_double * 4)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib = None


def load_library():
    """Load libnkprecond.so; raises OSError with a build hint if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp = C.c_void_p
    lib.nkp_default_options.argtypes = [P(NkpOptions)]
    lib.nkp_default_options.restype = None
    lib.nkp_create.argtypes = [P(vp), C.c_int, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(NkpOptions)]
    lib.nkp_create_dist.argtypes = [P(vp), C.c_int, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int),
                                    P(NkpOptions), C.c_int, C.c_int, C.c_char_p]
    lib.nkp_comm_unique_id.argtypes = [C.c_char_p]
    lib.nkp_rowperm_largediag.argtypes = [C.c_int, P(C.c_int), P(C.c_int), P(C.c_double), P(C.c_int), P(C.c_double),
                                          P(C.c_double)]
    lib.nkp_create_rowperm.argtypes = [P(vp), C.c_int, P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int), P(C.c_int),
                                       P(NkpOptions), P(C.c_int), P(C.c_double), P(C.c_double), C.c_int, C.c_int, C.c_char_p]
    lib.nkp_create_be.argtypes = [P(vp), C.c_int, C.c_longlong, vp, vp, P(C.c_int), P(C.c_int), P(C.c_int), P(NkpOptions)]
    lib.nkp_crs_finalize_device.argtypes = [C.c_int, vp, vp, vp, C.c_int, P(C.c_longlong), P(C.c_int)]
    lib.nkp_bswap32_device.argtypes = [vp, C.c_longlong]
    lib.nkp_assemble_min_device.argtypes = [P(NkpMinFields), C.c_double, C.c_double, C.c_double, vp, vp, vp, C.c_longlong,
                                            P(C.c_longlong)]
    lib.nkp_factor.argtypes = [vp, P(C.c_double)]
    lib.nkp_factor_be.argtypes = [vp, vp]
    lib.nkp_factor_device.argtypes = [vp, vp]
    lib.nkp_solve.argtypes = [vp, P(C.c_double), C.c_int, C.c_int, P(C.c_double)]
    lib.nkp_solve_device.argtypes = [vp, vp, C.c_int, C.c_int, P(C.c_double)]
    lib.nkp_solve_dist.argtypes = [vp, P(C.c_double), C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_double)]
    lib.nkp_set_tracer_maps.argtypes = [vp, C.c_int, C.c_int, P(C.c_int), P(C.c_int), P(C.c_int), C.c_int, C.c_int, C.c_int]
    lib.nkp_solve_fields.argtypes = [vp, P(P(C.c_double)), C.c_int, P(C.c_double)]
    lib.nkp_residual_device.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.nkp_sweeps_device.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.nkp_get_perm.argtypes = [vp, P(C.c_int)]
    lib.nkp_get_stats.argtypes = [vp, P(NkpStats)]
    lib.nkp_sync.argtypes = [vp]
    lib.nkp_set_analysis_cache.argtypes = [C.c_char_p]
    lib.nkp_set_profile.argtypes = [vp, C.c_int]
    lib.nkp_set_refine_rule.argtypes = [vp, C.c_int]
    lib.nkp_set_residual_extra.argtypes = [vp, C.c_int]
    lib.nkp_set_verbose.argtypes = [vp, C.c_int]
    lib.nkp_destroy.argtypes = [vp]
    lib.nkp_destroy.restype = None
    lib.nkp_last_error.restype = C.c_char_p
    lib.nkp_version.restype = C.c_char_p
    _lib = lib
    return lib


class NkpError(RuntimeError):
    pass


UNIQUE_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """NCCL unique id for nkp_create_dist (call on rank 0, ship to the other ranks)."""
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(load_library().nkp_comm_unique_id(buf), "nkp_comm_unique_id")
    return buf.raw


def set_analysis_cache(path):
    """Directory of the on-disk ordering cache (None disables); see nkp_set_analysis_cache."""
    _check(load_library().nkp_set_analysis_cache(None if path is None else os.fsencode(path)), "nkp_set_analysis_cache")


def _check(rc, what):
    if rc != 0:
        msg = load_library().nkp_last_error().decode()
        raise NkpError(f"{what} failed with code {rc}: {msg}")


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def rowperm_largediag(n, rowptr, colind, nzval):
    """Static row permutation for a large diagonal (nkp_rowperm_largediag: what pdgssvx does under
    RowPerm = LargeDiag, MC64 job 5).  Host computation.  Returns (rowmap, row_scale, col_scale)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    nzval = np.ascontiguousarray(nzval, dtype=np.float64)
    rowmap = np.zeros(int(n), dtype=np.int32)
    R = np.zeros(int(n))
    Cs = np.zeros(int(n))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = load_library().nkp_rowperm_largediag(int(n), _iptr(rowptr), _iptr(colind), dp(nzval), _iptr(rowmap), dp(R), dp(Cs))
    if rc != 0:
        raise NkpError(f"nkp_rowperm_largediag failed with code {rc} (structurally singular matrix?)")
    return rowmap, R, Cs


def crs_finalize_device(n, d_rowptr, d_colind, d_val, strip_zeros=True):
    """sum_dup_vals + (strip_matrix_zeros) + sort_cols_all_rows (src/matrix.c:3621-3770) on device arrays, in place;
    arguments are device addresses.  Returns (nnz, dup_cnt)."""
    nnz, dup = C.c_longlong(), C.c_int()
    _check(load_library().nkp_crs_finalize_device(int(n), C.c_void_p(d_rowptr), C.c_void_p(d_colind), C.c_void_p(d_val),
                                                  int(bool(strip_zeros)), C.byref(nnz), C.byref(dup)), "nkp_crs_finalize_device")
    return nnz.value, dup.value


def assemble_min_device(shape, n, dev_ptrs, fill_value, d_rowptr, d_colind, d_val, capacity, day_cnt=365.0,
                        sink_rate=365.0, sink_depth=10.0e2):
    """nkp_assemble_min_device: stencil values of the centered / const / const / const_shallow option set on the device.
    dev_ptrs maps the field names of nkp_min_fields to device addresses.  Returns the number of entries written."""
    f = NkpMinFields()
    f.imt, f.jmt, f.km = (int(v) for v in shape)
    f.n = int(n)
    for k, v in dev_ptrs.items():
        setattr(f, k, int(v))
    f.fill_value = float(fill_value)
    nnz = C.c_longlong()
    _check(load_library().nkp_assemble_min_device(C.byref(f), day_cnt, sink_rate, sink_depth, C.c_void_p(d_rowptr),
                                                  C.c_void_p(d_colind), C.c_void_p(d_val), int(capacity), C.byref(nnz)),
           "nkp_assemble_min_device")
    return nnz.value


class TracerJacobianSolver:
    """Analysis handle + numeric factors for one sparsity pattern (CRS, 0-based).

    coords: optional (i, j, k) int arrays per unknown (tracer_state_ind_to_{i,j,k},
    src/matrix.c:322-329) enabling the geometric nested dissection.
    """

    def __init__(self, n, rowptr, colind, coords=None, comm=None, file_byte_order=False, rowperm=None, **opts):
        """comm: None (one GPU) or (rank, nranks, unique_id_bytes) for one-process-per-GPU runs.
        rowperm: None, or (rowmap, row_scale, col_scale) as returned by rowperm_largediag (the scalings may be None):
        static row permutation of the factored matrix (nkp_create_rowperm).
        file_byte_order=True: rowptr / colind are the big-endian NC_INT bytes of the matrix file (bytes-like, n + 1 and
        nnz values); they are converted on the device (nkp_create_be)."""
        lib = load_library()
        self._lib = lib
        self.n = int(n)
        if file_byte_order:
            rp_raw = np.frombuffer(rowptr, dtype=np.uint8)
            ci_raw = np.frombuffer(colind, dtype=np.uint8)
            assert rp_raw.size == 4 * (self.n + 1) and comm is None
            self.nnz = ci_raw.size // 4
        else:
            rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
            colind = np.ascontiguousarray(colind, dtype=np.int32)
            self.nnz = int(rowptr[-1])
        o = NkpOptions()
        lib.nkp_default_options(C.byref(o))
        for k, v in opts.items():
            setattr(o, k, int(v))
        ci = cj = ck = None
        if coords is not None:
            ci, cj, ck = (np.ascontiguousarray(c, dtype=np.int32) if c is not None else None for c in coords)
        self._h = C.c_void_p()
        if rowperm is not None:
            assert not file_byte_order
            rowmap, rs, cs = rowperm
            rowmap = np.ascontiguousarray(rowmap, dtype=np.int32)
            dp = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
            rs_keep = None if rs is None else np.ascontiguousarray(rs, dtype=np.float64)
            cs_keep = None if cs is None else np.ascontiguousarray(cs, dtype=np.float64)
            rank, nranks, uid = comm if comm is not None else (0, 1, None)
            _check(lib.nkp_create_rowperm(C.byref(self._h), self.n, _iptr(rowptr), _iptr(colind), _iptr(ci), _iptr(cj), _iptr(ck),
                                          C.byref(o), _iptr(rowmap), dp(rs_keep), dp(cs_keep), int(rank), int(nranks),
                                          None if uid is None else bytes(uid)), "nkp_create_rowperm")
        elif file_byte_order:
            _check(lib.nkp_create_be(C.byref(self._h), self.n, self.nnz, C.c_void_p(rp_raw.ctypes.data), C.c_void_p(ci_raw.ctypes.data),
                                     _iptr(ci), _iptr(cj), _iptr(ck), C.byref(o)), "nkp_create_be")
        elif comm is None:
            _check(lib.nkp_create(C.byref(self._h), self.n, _iptr(rowptr), _iptr(colind), _iptr(ci), _iptr(cj),
                                  _iptr(ck), C.byref(o)), "nkp_create")
        else:
            rank, nranks, uid = comm
            assert len(uid) == UNIQUE_ID_BYTES
            _check(lib.nkp_create_dist(C.byref(self._h), self.n, _iptr(rowptr), _iptr(colind), _iptr(ci), _iptr(cj),
                                       _iptr(ck), C.byref(o), int(rank), int(nranks), bytes(uid)), "nkp_create_dist")

    # -- numeric phase -------------------------------------------------------------------
    def factor(self, nzval):
        """Numeric LU from host values (includes the host->device copy)."""
        nzval = np.ascontiguousarray(nzval, dtype=np.float64)
        assert nzval.size == self.nnz
        _check(self._lib.nkp_factor(self._h, nzval.ctypes.data_as(C.POINTER(C.c_double))), "nkp_factor")

    def factor_be(self, raw):
        """Numeric LU from nnz big-endian doubles (bytes / buffer as they lie in the matrix file)."""
        buf = np.frombuffer(raw, dtype=np.uint8)
        assert buf.size == 8 * self.nnz
        _check(self._lib.nkp_factor_be(self._h, C.c_void_p(buf.ctypes.data)), "nkp_factor_be")

    def factor_device(self, d_ptr):
        """Numeric LU with values already on the device (d_ptr: int device address)."""
        _check(self._lib.nkp_factor_device(self._h, C.c_void_p(d_ptr)), "nkp_factor_device")

    def solve(self, B):
        """Solve A X = B in place (host memory, Fortran order n x nrhs); returns berr[nrhs]."""
        assert B.dtype == np.float64
        if B.ndim == 1:
            ldb, nrhs = B.shape[0], 1
            assert B.flags.c_contiguous
        else:
            assert B.flags.f_contiguous
            ldb, nrhs = B.shape
        berr = np.zeros(max(nrhs, 1))
        _check(self._lib.nkp_solve(self._h, B.ctypes.data_as(C.POINTER(C.c_double)), ldb, nrhs,
                                   berr.ctypes.data_as(C.POINTER(C.c_double))), "nkp_solve")
        return berr

    def solve_dist(self, B_loc, fst_row):
        """Multi-GPU solve with a row-distributed right-hand side (solve_ABdist, src/solve_ABdist.c:141-144): B_loc is
        this rank's slab of rows [fst_row, fst_row + m_loc), Fortran order m_loc x nrhs, overwritten by the same rows
        of X.  Returns berr[nrhs]."""
        assert B_loc.dtype == np.float64
        if B_loc.ndim == 1:
            m_loc, nrhs = B_loc.shape[0], 1
            assert B_loc.flags.c_contiguous
        else:
            assert B_loc.flags.f_contiguous
            m_loc, nrhs = B_loc.shape
        berr = np.zeros(max(nrhs, 1))
        _check(self._lib.nkp_solve_dist(self._h, B_loc.ctypes.data_as(C.POINTER(C.c_double)), max(m_loc, 1), nrhs,
                                        int(fst_row), m_loc, berr.ctypes.data_as(C.POINTER(C.c_double))), "nkp_solve_dist")
        return berr

    def solve_device(self, d_ptr, ldb, nrhs):
        berr = np.zeros(max(nrhs, 1))
        _check(self._lib.nkp_solve_device(self._h, C.c_void_p(d_ptr), ldb, nrhs,
                                          berr.ctypes.data_as(C.POINTER(C.c_double))), "nkp_solve_device")
        return berr

    def set_tracer_maps(self, ind_i, ind_j, ind_k, shape, coupled_tracer_cnt=1):
        """Register tracer_state_ind_to_{i,j,k} (src/matrix.c:322-329) and the grid shape (imt, jmt, km)."""
        ii, jj, kk = (np.ascontiguousarray(a, dtype=np.int32) for a in (ind_i, ind_j, ind_k))
        imt, jmt, km = (int(v) for v in shape)
        self._field_shape = (km, jmt, imt)
        self._ct = int(coupled_tracer_cnt)
        _check(self._lib.nkp_set_tracer_maps(self._h, int(ii.size), self._ct, _iptr(ii), _iptr(jj), _iptr(kk),
                                             imt, jmt, km), "nkp_set_tracer_maps")

    def solve_fields(self, fields):
        """get_B + solve + put_B (src/solve_ABglobal.c:154-267) for a list of 3-D float64 fields
        [k][j][i], in place; every coupled_tracer_cnt consecutive fields form one system and all systems
        are solved as one batch.  Returns berr per system."""
        shape = getattr(self, "_field_shape", None)
        for f in fields:
            assert f.dtype == np.float64 and f.flags.c_contiguous and (shape is None or f.shape == shape)
        ptrs = (C.POINTER(C.c_double) * max(len(fields), 1))(*[f.ctypes.data_as(C.POINTER(C.c_double)) for f in fields])
        berr = np.zeros(max(len(fields) // getattr(self, "_ct", 1), 1))
        _check(self._lib.nkp_solve_fields(self._h, ptrs, len(fields), berr.ctypes.data_as(C.POINTER(C.c_double))),
               "nkp_solve_fields")
        return berr

    def sweeps_device(self, d_ptr, ldb, nrhs):
        _check(self._lib.nkp_sweeps_device(self._h, C.c_void_p(d_ptr), ldb, nrhs), "nkp_sweeps_device")

    def residual_device(self, d_x, d_b, d_r, nrhs):
        _check(self._lib.nkp_residual_device(self._h, C.c_void_p(d_x), C.c_void_p(d_b), C.c_void_p(d_r), nrhs),
               "nkp_residual_device")

    def set_profile(self, on=True):
        _check(self._lib.nkp_set_profile(self._h, int(bool(on))), "nkp_set_profile")

    def set_refine_rule(self, rule):
        """0: SuperLU's componentwise berr rule (default); 1: normwise ||r|| <= 1e-14 ||b||."""
        _check(self._lib.nkp_set_refine_rule(self._h, int(rule)), "nkp_set_refine_rule")

    def set_residual_extra(self, on=True):
        """Refinement residual accumulated in twice the working precision (nkp_options.residual_extra)."""
        _check(self._lib.nkp_set_residual_extra(self._h, int(bool(on))), "nkp_set_residual_extra")

    def sync(self):
        _check(self._lib.nkp_sync(self._h), "nkp_sync")

    # -- introspection -------------------------------------------------------------------
    def perm(self):
        p = np.zeros(self.n, dtype=np.int32)
        _check(self._lib.nkp_get_perm(self._h, _iptr(p)), "nkp_get_perm")
        return p

    def stats(self):
        st = NkpStats()
        _check(self._lib.nkp_get_stats(self._h, C.byref(st)), "nkp_get_stats")
        return st.as_dict()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.nkp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
