"""Synthetic POP-style inputs for the tracer-Jacobian preconditioner.

BASELINE.json: "Synthetic POP-style KMT grids of the named shapes give the inputs".

Three things live here:

* ``make_grid`` / ``make_circulation`` -- a deterministic (seeded) lat-lon ocean grid
  with continents (KMT), POP-like layer thicknesses and a smooth circulation with
  realistic advection/diffusion ratios.
* ``write_circ_file`` / ``write_tracer_file`` -- the NetCDF (CDF-2) files the
  reference's ``gen_A`` and solver drivers read: variable names and attributes as
  consumed by src/grid.c:91-214 and src/matrix.c:984-1217,2575-2640.
* ``assemble_crs`` -- a numpy restatement of ``gen_sparse_matrix``
  (src/matrix.c:3775-3840) for the option set
  ``adv_type centered / hmix_type const / vmix_type const / sink const_shallow``.
  Floating-point expressions keep the reference's evaluation order so the CRS is
  bit-identical to the unchanged ``gen_A`` (proved in tests/test_synth_vs_gen_A.py
  against oracle/_ref/gen_A and the committed golden file).  bench.py uses it to
  build gx3v7 / gx1v6-shape operands on the GPU box, where the reference tree and
  its generator do not exist.
"""
from __future__ import annotations

import numpy as np

FILL = 9.969209968386869e36  # POP / netCDF default double fill value

R_EARTH_CM = 6.37122e8


# --------------------------------------------------------------------------- grid


def make_grid(imt: int, jmt: int, km: int, seed: int = 0) -> dict:
    """Lat-lon sphere grid with analytic continents; KMT == 0 on j=0 and j=jmt-1
    (required by src/grid.c:162-180)."""
    rng = np.random.default_rng(seed)
    # layer thicknesses: 10 m at the surface growing to 250 m (cm), POP-like
    s = np.linspace(0.0, 1.0, km)
    dz = 1000.0 + (25000.0 - 1000.0) * s**2
    dz = np.round(dz, 3)
    z_w = np.concatenate([[0.0], np.cumsum(dz)])
    z_t = 0.5 * (z_w[:-1] + z_w[1:])

    lat_s, lat_n = -78.0, 88.0
    dlat = (lat_n - lat_s) / jmt
    dlon = 360.0 / imt
    tlat1 = lat_s + dlat * (np.arange(jmt) + 0.5)
    tlon1 = dlon * (np.arange(imt) + 0.5)
    ulat1 = tlat1 + 0.5 * dlat  # U points sit at the NE corner of T cells
    TLAT = np.repeat(tlat1[:, None], imt, axis=1)
    TLONG = np.repeat(tlon1[None, :], jmt, axis=0)

    dlat_r = np.deg2rad(dlat)
    dlon_r = np.deg2rad(dlon)
    cos_t = np.cos(np.deg2rad(tlat1))[:, None] * np.ones((1, imt))
    cos_u = np.cos(np.deg2rad(np.clip(ulat1, -89.5, 89.5)))[:, None] * np.ones((1, imt))
    DXT = R_EARTH_CM * cos_t * dlon_r
    DYT = R_EARTH_CM * dlat_r * np.ones((jmt, imt))
    DXU = R_EARTH_CM * cos_u * dlon_r
    DYU = R_EARTH_CM * dlat_r * np.ones((jmt, imt))
    TAREA = DXT * DYT
    HTN = DXU.copy()  # length of the north face of a T cell
    HTE = DYT.copy()  # length of the east face of a T cell
    HUS = DXT.copy()  # zonal distance used by the const-hmix east/west weights
    HUW = DYT.copy()  # meridional distance used by the const-hmix north/south weights

    # bathymetry: smooth function of (lon, lat) with a few "continents"
    lon_r = np.deg2rad(TLONG)
    lat_r = np.deg2rad(TLAT)
    ph = rng.uniform(0, 2 * np.pi, size=6)
    h = (
        0.55
        + 0.50 * np.sin(2 * lon_r + ph[0]) * np.cos(1.5 * lat_r + ph[1])
        + 0.35 * np.sin(3 * lon_r + ph[2]) * np.sin(2.0 * lat_r + ph[3])
        + 0.20 * np.cos(5 * lon_r + ph[4]) * np.cos(4.0 * lat_r + ph[5])
    )
    # ~60 % ocean: shift so that the 40th percentile is the coastline
    h = h - np.quantile(h, 0.40)
    depth_frac = np.clip(h / max(np.quantile(h[h > 0], 0.30), 1e-6), 0.0, 1.0)  # ~70 % of the ocean at full depth
    KMT = np.where(h > 0, np.maximum(3, np.ceil(depth_frac * km)), 0).astype(np.int32)
    KMT = np.minimum(KMT, km)
    KMT[0, :] = 0
    KMT[-1, :] = 0

    return dict(
        imt=imt, jmt=jmt, km=km, z_t=z_t, dz=dz, TLONG=TLONG, TLAT=TLAT, KMT=KMT,
        TAREA=TAREA, DXU=DXU, DYU=DYU, HTE=HTE, HTN=HTN, HUS=HUS, HUW=HUW, seed=seed,
    )


def kmu_from_kmt(KMT: np.ndarray) -> np.ndarray:
    """src/grid.c:187-203."""
    jmt, imt = KMT.shape
    KMU = np.zeros_like(KMT)
    ip1 = np.r_[1:imt, 0]
    a = KMT[:-1, :]
    b = KMT[1:, :]
    KMU[:-1, :] = np.minimum(np.minimum(a, b), np.minimum(a[:, ip1], b[:, ip1]))
    return KMU


def _face_transports(grid: dict, UVEL, VVEL):
    """UTE, VTN as load_UTE / load_VTN build them (src/matrix.c:1024-1031,1103-1111)."""
    KMT = grid["KMT"]
    km, jmt, imt = grid["km"], grid["jmt"], grid["imt"]
    KMU = kmu_from_kmt(KMT)
    kk = np.arange(km)[:, None, None]
    mU = kk < KMU[None, :, :]
    DY = grid["DYU"]
    DX = grid["DXU"]
    UTE = np.zeros((km, jmt, imt))
    t1 = np.where(mU, 0.5 * UVEL * DY[None], 0.0)
    # UTE[k][j][i] += 0.5*U[k][j][i]*DY[j][i]; then += 0.5*U[k][j-1][i]*DY[j-1][i]   (j = 1..jmt-2)
    UTE[:, 1:-1, :] = (0.0 + t1[:, 1:-1, :]) + t1[:, 0:-2, :]
    im1 = np.r_[imt - 1, 0:imt - 1]
    VTN = np.zeros((km, jmt, imt))
    t2 = np.where(mU, 0.5 * VVEL * DX[None], 0.0)
    VTN[:, 1:-1, :] = (0.0 + t2[:, 1:-1, :]) + t2[:, 1:-1, :][:, :, im1]
    return UTE, VTN


def make_circulation(grid: dict, seed: int = 0, u_scale: float = 5.0) -> dict:
    """Smooth gyre-like flow (|u| ~ cm/s) plus seeded noise; W from continuity."""
    rng = np.random.default_rng(seed + 1000)
    km, jmt, imt = grid["km"], grid["jmt"], grid["imt"]
    lon = np.deg2rad(grid["TLONG"])
    lat = np.deg2rad(grid["TLAT"])
    prof = np.exp(-grid["z_t"] / 1.0e5)[:, None, None]  # e-folding 1000 m
    u2 = u_scale * (np.cos(3 * lat) * np.sin(lon + 0.3) + 0.5 * np.sin(2 * lat))
    v2 = u_scale * 0.6 * (np.sin(2 * lon + 1.1) * np.cos(lat))
    UVEL = prof * u2[None] + 0.2 * u_scale * prof * rng.standard_normal((km, jmt, imt))
    VVEL = prof * v2[None] + 0.2 * u_scale * prof * rng.standard_normal((km, jmt, imt))
    KMU = kmu_from_kmt(grid["KMT"])
    kk = np.arange(km)[:, None, None]
    land_u = kk >= KMU[None]
    UVEL = np.where(land_u, FILL, UVEL)
    VVEL = np.where(land_u, FILL, VVEL)

    # W at the top of each T cell from continuity, zero at the column bottom
    UTE, VTN = _face_transports(grid, np.where(land_u, 0.0, UVEL), np.where(land_u, 0.0, VVEL))
    im1 = np.r_[imt - 1, 0:imt - 1]
    div = np.zeros((km, jmt, imt))
    div[:, 1:-1, :] = (
        UTE[:, 1:-1, :] - UTE[:, 1:-1, :][:, :, im1] + VTN[:, 1:-1, :] - VTN[:, 0:-2, :]
    ) / grid["TAREA"][None, 1:-1, :]
    ocean = kk < grid["KMT"][None]
    div = np.where(ocean, div, 0.0)
    W = np.zeros((km + 1, jmt, imt))
    for k in range(km - 1, -1, -1):
        W[k] = W[k + 1] + grid["dz"][k] * div[k]
    WVEL = np.where(ocean, W[:km], FILL)
    return dict(UVEL=UVEL, VVEL=VVEL, WVEL=WVEL)


# --------------------------------------------------------------------------- files


def _nc():
    from scipy.io import netcdf_file  # imported lazily: scipy is harness-side only
    return netcdf_file


def make_full_fields(grid: dict, circ: dict, seed: int = 0) -> dict:
    """Extra circulation-file fields for the reference's own test option set
    (test/test_gen_A.csh:22-23: adv upwind3, hmix isop_file, vmix file):

    * UTE/VTN/WTK_{POS,NEG}: positive / negative parts of the face transports
      (src/matrix.c:1471-1558),
    * 36 HDIF_EXPLICIT_3D_IRF_{1..4}_{1..3}_{1..3}: impulse responses of an isopycnal-style
      diffusion operator; entry [k][j][i] of class (i',j',k') is the coefficient (1/s) of the
      unique stencil neighbour whose indices are congruent to (i',j',k') mod (4,3,3)
      (src/matrix.c:2233-2376) -- needs imt % 4 == 0,
    * VDC_S, VDC_GM: vertical diffusivities (src/matrix.c:2869-2885).
    """
    rng = np.random.default_rng(seed + 2000)
    km, jmt, imt = grid["km"], grid["jmt"], grid["imt"]
    assert imt % 4 == 0, "IRF residue classes need imt % 4 == 0 (periodic seam)"
    KMT = grid["KMT"]
    kk = np.arange(km)[:, None, None]
    ocean = kk < KMT[None]
    U = np.where(circ["UVEL"] == FILL, 0.0, circ["UVEL"])
    V = np.where(circ["VVEL"] == FILL, 0.0, circ["VVEL"])
    W = np.where(circ["WVEL"] == FILL, 0.0, circ["WVEL"])
    UTE, VTN = _face_transports(grid, U, V)
    out = {}
    landT = ~ocean
    for name, f in (("UTE", UTE), ("VTN", VTN), ("WTK", W)):
        out[name + "_POS"] = np.where(landT, FILL, np.maximum(f, 0.0))
        out[name + "_NEG"] = np.where(landT, FILL, np.minimum(f, 0.0))
    # vertical diffusivity 0.1 .. 10 cm^2/s
    vdc = 0.1 + 9.9 * np.exp(-grid["z_t"] / 5.0e3)[:, None, None] * (0.5 + 0.5 * rng.random((km, jmt, imt)))
    out["VDC_S"] = np.where(landT, FILL, vdc)
    out["VDC_GM"] = np.where(landT, FILL, 0.05 * vdc)

    # isopycnal-style diffusion stencil: 15 neighbours, conservative (self = -sum of the others)
    lon = np.deg2rad(grid["TLONG"])
    lat = np.deg2rad(grid["TLAT"])
    kappa = 4.0e6 + 2.0e6 * (0.5 + 0.5 * np.sin(2 * lon + 0.7) * np.cos(lat))     # cm^2/s
    sx = 1.0e-3 * np.sin(3 * lon) * np.cos(2 * lat)                                 # isopycnal slopes
    sy = 1.0e-3 * np.cos(2 * lon + 0.4) * np.sin(3 * lat)
    dz = grid["dz"][:, None, None]
    dx = grid["HUS"][None]
    dy = grid["HUW"][None]
    ta = grid["TAREA"][None]
    ii = np.arange(imt)
    ip1 = np.r_[1:imt, 0]
    im1 = np.r_[imt - 1, 0:imt - 1]

    def shifted(mask, dk, dj, di):
        """ocean mask of the neighbour (k+dk, j+dj, i+di), False outside the grid."""
        m = np.zeros_like(mask)
        ks = slice(max(0, -dk), km - max(0, dk))
        kd = slice(max(0, dk), km - max(0, -dk))
        js = slice(max(0, -dj), jmt - max(0, dj))
        jd = slice(max(0, dj), jmt - max(0, -dj))
        idx = ii if di == 0 else (ip1 if di == 1 else im1)
        m[ks, js, :] = mask[kd, jd, :][:, :, idx]
        return m

    coef = {}
    k2 = kappa[None]
    coef[(0, 0, 1)] = k2 * grid["HTE"][None] / dx / ta
    coef[(0, 0, -1)] = (k2 * grid["HTE"][None] / dx)[:, :, im1] / ta
    coef[(0, 1, 0)] = k2 * grid["HTN"][None] / dy / ta
    cs = np.zeros((1, jmt, imt))
    cs[:, 1:, :] = (k2 * grid["HTN"][None] / dy)[:, :-1, :]
    coef[(0, -1, 0)] = cs / ta
    slope2 = (sx**2 + sy**2)[None]
    dzu = np.empty((km, 1, 1)); dzu[1:] = 0.5 * (dz[1:] + dz[:-1]); dzu[0] = dz[0]
    dzd = np.empty((km, 1, 1)); dzd[:-1] = 0.5 * (dz[1:] + dz[:-1]); dzd[-1] = dz[-1]
    coef[(-1, 0, 0)] = k2 * slope2 / dzu / dz * np.ones((km, jmt, imt))
    coef[(1, 0, 0)] = k2 * slope2 / dzd / dz * np.ones((km, jmt, imt))
    for dk in (-1, 1):
        for di in (-1, 1):
            coef[(dk, 0, di)] = -dk * di * k2 * sx[None] / (4.0 * dx * dz) * np.ones((km, jmt, imt))
        for dj in (-1, 1):
            coef[(dk, dj, 0)] = -dk * dj * k2 * sy[None] / (4.0 * dy * dz) * np.ones((km, jmt, imt))
    self_c = np.zeros((km, jmt, imt))
    for off, c in list(coef.items()):
        c = np.broadcast_to(c, (km, jmt, imt)) * (ocean & shifted(ocean, *off))
        coef[off] = c
        self_c -= c
    coef[(0, 0, 0)] = np.where(ocean, self_c, 0.0)
    # every (cell, offset) pair belongs to exactly one residue class, and the 15 offsets of one cell fall
    # into 15 different classes: one scatter per offset fills all 36 fields
    ka, ja, ia = np.arange(km)[:, None, None], np.arange(jmt)[None, :, None], np.arange(imt)[None, None, :]
    irf = np.zeros((36, km, jmt, imt))
    for (dk, dj, di), c in coef.items():
        cls = ((((ia + di) % imt) % 4) * 3 + (ja + dj) % 3) * 3 + (ka + dk) % 3
        np.put_along_axis(irf, np.broadcast_to(cls, (km, jmt, imt))[None], (c + 0.0)[None], axis=0)   # + 0.0: no negative zeros
    for ip in range(4):
        for jp in range(3):
            for kp in range(3):
                out[f"HDIF_EXPLICIT_3D_IRF_{ip + 1}_{jp + 1}_{kp + 1}"] = irf[(ip * 3 + jp) * 3 + kp]
    return out


REFTEST_OPTS = (
    "circ_fname {circ}\n"
    "day_cnt 365.0\n"
    "adv_type upwind3\n"
    "hmix_type isop_file\n"
    "vmix_type file\n"
    "sink_type const_shallow 365.0 10.0e2\n"
)


def write_circ_file(path: str, grid: dict, circ: dict, full: dict | None = None) -> None:
    """POP-history-like file for gen_A's ``circ_fname``; ``full`` (make_full_fields) adds the
    fields of the reference's own test option set."""
    f = _nc()(path, "w", version=2)
    km, jmt, imt = grid["km"], grid["jmt"], grid["imt"]
    f.createDimension("nlon", imt)
    f.createDimension("nlat", jmt)
    f.createDimension("z_t", km)
    for name in ("z_t", "dz"):
        v = f.createVariable(name, "d", ("z_t",))
        v[:] = grid[name]
    for name in ("TLONG", "TLAT", "TAREA"):
        v = f.createVariable(name, "d", ("nlat", "nlon"))
        v[:] = grid[name]
    v = f.createVariable("KMT", "i", ("nlat", "nlon"))
    v[:] = grid["KMT"]
    for name in ("DXU", "DYU", "HTE", "HTN", "HUS", "HUW"):
        v = f.createVariable(name, "d", ("nlat", "nlon"))
        v[:] = grid[name]
        v._FillValue = np.float64(FILL)
    for name in ("UVEL", "VVEL", "WVEL"):
        v = f.createVariable(name, "d", ("z_t", "nlat", "nlon"))
        v[:] = circ[name]
        v._FillValue = np.float64(FILL)
    for name, arr in (full or {}).items():
        v = f.createVariable(name, "d", ("z_t", "nlat", "nlon"))
        v[:] = arr
        if not name.startswith("HDIF"):   # IRF variables get no fill masking (src/matrix.c:2259)
            v._FillValue = np.float64(FILL)
    f.close()


def write_tracer_file(path: str, grid: dict, fields: dict) -> None:
    """Tracer (RHS / solution) file: 3-D doubles (z_t, nlat, nlon), src/file_io.c:273-295."""
    f = _nc()(path, "w", version=2)
    km, jmt, imt = grid["km"], grid["jmt"], grid["imt"]
    f.createDimension("nlon", imt)
    f.createDimension("nlat", jmt)
    f.createDimension("z_t", km)
    for name, arr in fields.items():
        v = f.createVariable(name, "d", ("z_t", "nlat", "nlon"))
        v[:] = np.asarray(arr, dtype=np.float64).reshape(km, jmt, imt)
    f.close()


def read_matrix_file(path: str) -> dict:
    """Read the reference's matrix file into numpy arrays."""
    f = _nc()(path, "r", mmap=False)
    out = {}
    for name in ("nzval_row_wise", "colind", "rowptr", "coupled_tracer_cnt", "KMT",
                 "tracer_state_ind_to_i", "tracer_state_ind_to_j", "tracer_state_ind_to_k",
                 "int3_to_tracer_state_ind"):
        if name in f.variables:
            a = np.array(f.variables[name].data)
            # the file is big-endian; hand out native-endian arrays (they go through ctypes)
            out[name] = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("=")))
    out["imt"] = f.dimensions["nlon"]
    out["jmt"] = f.dimensions["nlat"]
    out["km"] = f.dimensions["z_t"]
    f.close()
    return out


def read_tracer(path: str, name: str) -> np.ndarray:
    f = _nc()(path, "r", mmap=False)
    a = np.array(f.variables[name].data, dtype=np.float64).copy()
    f.close()
    return a


MINIMAL_OPTS = (
    "circ_fname {circ}\n"
    "adv_type centered\n"
    "hmix_type const\n"
    "vmix_type const\n"
    "sink_type const_shallow 365.0 10.0e2\n"
)


# --------------------------------------------------------------------------- index maps


def index_maps(KMT: np.ndarray, km: int):
    """gen_ind_maps (src/matrix.c:239-251): j outer, i middle, k inner."""
    jmt, imt = KMT.shape
    cnt = KMT.astype(np.int64).ravel()  # (j, i) row-major
    n = int(cnt.sum())
    start = np.concatenate([[0], np.cumsum(cnt)])
    col_of = np.repeat(np.arange(jmt * imt), cnt)
    k = np.arange(n) - start[col_of]
    j = col_of // imt
    i = col_of % imt
    int3 = -np.ones((km, jmt, imt), dtype=np.int32)
    int3[k, j, i] = np.arange(n, dtype=np.int32)
    return n, i.astype(np.int32), j.astype(np.int32), k.astype(np.int32), int3


# --------------------------------------------------------------------------- CRS assembly


def assemble_crs(grid: dict, circ: dict, day_cnt: float = 365.0,
                 sink_rate: float = 365.0, sink_depth: float = 10.0e2, raw: bool = False):
    """numpy restatement of gen_sparse_matrix for centered / const / const / const_shallow.

    Returns (n, rowptr[int32 n+1], colind[int32 nnz], nzval[float64 nnz], maps) where
    maps = (i, j, k, int3_to_tracer_state_ind).
    raw=True returns the matrix as it stands BEFORE the post-processing of gen_sparse_matrix (sum_dup_vals,
    strip_matrix_zeros, sort_cols_all_rows, src/matrix.c:3829-3835): every stencil slot of a row in slot order
    (self, k-1, k+1, east, west, north, south), exact zeros kept, columns unsorted -- the input of
    nkp_crs_finalize_device.
    Operation order per slot follows src/matrix.c:
      add_UTE_coeffs :1239-1273, add_VTN_coeffs :1320-1360, add_WVEL_coeffs :1401-1430,
      adv_enforce_divfree :2094-2206, add_hmix_const :2656-2710, add_vmix_const :2978-3004,
      add_sink_pure_diag :3084-3091, strip_matrix_zeros :3657-3688, sort :3753-3765.
    """
    KMT = grid["KMT"]
    km, jmt, imt = grid["km"], grid["jmt"], grid["imt"]
    dz = grid["dz"]
    z_t = grid["z_t"]
    TAREA = grid["TAREA"]
    delta_t = 60.0 * 60.0 * 24.0 * day_cnt
    year_cnt = day_cnt / 365.0

    n, ii, jj, kk, int3 = index_maps(KMT, km)
    ip1 = np.where(ii < imt - 1, ii + 1, 0)
    im1 = np.where(ii > 0, ii - 1, imt - 1)

    def fv0(a):
        return np.where(a == FILL, 0.0, a)

    UVEL = fv0(circ["UVEL"])
    VVEL = fv0(circ["VVEL"])
    UTE, VTN = _face_transports(grid, UVEL, VVEL)
    # load_WVEL :1168-1198
    ocean3 = np.arange(km)[:, None, None] < KMT[None]
    W = np.zeros((km + 1, jmt, imt))
    W[:km] = np.where(ocean3, 0.0 + fv0(circ["WVEL"]), 0.0)
    W[:km, 0, :] = 0.0
    W[:km, -1, :] = 0.0
    W[0, 1:-1, :] = 0.0

    ta = TAREA[jj, ii]
    has_up = kk - 1 >= 0
    has_dn = kk + 1 < KMT[jj, ii]
    has_e = kk < KMT[jj, ip1]
    has_w = kk < KMT[jj, im1]
    has_n = kk < KMT[jj + 1, ii]
    has_s = kk < KMT[jj - 1, ii]

    w = 0.5
    ute_e = UTE[kk, jj, ii]
    ute_w = UTE[kk, jj, im1]
    vtn_n = VTN[kk, jj, ii]
    vtn_s = VTN[kk, jj - 1, ii]
    w_top = W[kk, jj, ii]
    w_bot = W[kk + 1, jj, ii]
    dzk = dz[kk]

    # advection, non-self slots (each starts from 0.0)
    a_e = 0.0 - (1.0 - w) * ute_e / ta * delta_t
    a_w = 0.0 + (1.0 - w) * ute_w / ta * delta_t
    a_n = 0.0 - (1.0 - w) * vtn_n / ta * delta_t
    a_s = 0.0 + (1.0 - w) * vtn_s / ta * delta_t
    a_up = 0.0 - (1.0 - w) * w_top / dzk * delta_t
    a_dn = 0.0 + (1.0 - w) * w_bot / dzk * delta_t

    # adv_enforce_divfree: self = -(sum of non-self in slot order k-1,k+1,e,w,n,s)
    ssum = np.zeros(n)
    for present, val in ((has_up, a_up), (has_dn, a_dn), (has_e, a_e), (has_w, a_w), (has_n, a_n), (has_s, a_s)):
        ssum = np.where(present, ssum + val, ssum)
    v_self = -ssum

    # hmix const
    ah = 4.0e6
    HTE, HUS, HTN, HUW = (fv0(grid[x]) for x in ("HTE", "HUS", "HTN", "HUW"))
    with np.errstate(divide="ignore", invalid="ignore"):
        ce = np.where(has_e, ah * HTE[jj, ii] / HUS[jj, ii] / ta * delta_t, 0.0)
        cw = np.where(has_w, ah * HTE[jj, im1] / HUS[jj, im1] / ta * delta_t, 0.0)
        cn = np.where(has_n, ah * HTN[jj, ii] / HUW[jj, ii] / ta * delta_t, 0.0)
        cs = np.where(has_s, ah * HTN[jj - 1, ii] / HUW[jj - 1, ii] / ta * delta_t, 0.0)
    v_self = v_self - (ce + cw + cn + cs)
    v_e = a_e + ce
    v_w = a_w + cw
    v_n = a_n + cn
    v_s = a_s + cs

    # vmix const
    vdc = 0.1
    dz_up = dz[np.maximum(kk - 1, 0)]
    dz_dn = dz[np.minimum(kk + 1, km - 1)]
    ct = np.where(has_up, vdc / (0.5 * (dz_up + dzk)) / dzk * delta_t, 0.0)
    cb = np.where(has_dn, vdc / (0.5 * (dzk + dz_dn)) / dzk * delta_t, 0.0)
    v_self = v_self - (ct + cb)
    v_up = a_up + ct
    v_dn = a_dn + cb

    # sink const_shallow
    v_self = np.where(z_t[kk] < sink_depth, v_self + (-year_cnt * sink_rate), v_self)

    # column indices of the neighbours
    me = np.arange(n, dtype=np.int64)
    c_up = me - 1
    c_dn = me + 1
    km1 = km - 1

    def nbr(j2, i2):
        return int3[np.minimum(kk, km1), j2, i2].astype(np.int64)

    c_e = nbr(jj, ip1)
    c_w = nbr(jj, im1)
    c_n = nbr(jj + 1, ii)
    c_s = nbr(jj - 1, ii)

    rows = []
    cols = []
    vals = []
    slots = []
    for slot, (present, c, v) in enumerate(((np.ones(n, bool), me, v_self), (has_up, c_up, v_up), (has_dn, c_dn, v_dn),
                                            (has_e, c_e, v_e), (has_w, c_w, v_w), (has_n, c_n, v_n), (has_s, c_s, v_s))):
        keep = present if raw else present & (v != 0.0)  # strip_matrix_zeros drops exact zeros
        rows.append(me[keep])
        cols.append(c[keep])
        vals.append(v[keep])
        slots.append(np.full(int(keep.sum()), slot))
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    vals = np.concatenate(vals)
    if raw:
        order = np.lexsort((np.concatenate(slots), rows))
        rows, cols, vals = rows[order], cols[order], vals[order]
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(rowptr, rows + 1, 1)
        return n, np.cumsum(rowptr).astype(np.int32), cols.astype(np.int32), vals, (ii, jj, kk, int3)
    # east and west may name the same column on a 2-wide periodic grid; the reference
    # would merge them in sum_dup_vals -- not reproduced: require imt >= 3
    assert imt >= 3
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    return n, rowptr.astype(np.int32), cols.astype(np.int32), vals, (ii, jj, kk, int3)


def make_rhs(n: int, nrhs: int = 1, seed: int = 0) -> np.ndarray:
    """Seeded N(0,1) right-hand sides on ocean points, column-major (n, nrhs)."""
    rng = np.random.default_rng(seed + 7)
    return np.asfortranarray(rng.standard_normal((n, nrhs)))
