/*
 * nkprecond.h -- C ABI of the B200 sparse direct solver for the Newton-Krylov ocean
 * tracer-Jacobian preconditioner.
 *
 * This is the drop-in boundary for the ONE hot path of the reference
 * (klindsay28/NK_ocn_tracer_jacobian_precond): the external SuperLU_DIST calls
 *
 *     pdgssvx_ABglobal(..., nrhs=0, ...)   src/solve_ABglobal.c:353   (factor)
 *     pdgssvx_ABglobal(..., nrhs=1, ...)   src/solve_ABglobal.c:395   (solve + refine)
 *     pdgssvx(..., nrhs=0, ...)            src/solve_ABdist.c:518     (factor)
 *     pdgssvx(..., nrhs=1, ...)            src/solve_ABdist.c:571     (solve + refine)
 *
 * operating on the CRS operand of src/matrix.h:64-68 (flat_len, nnz, nzval_row_wise,
 * colind, rowptr) and the flattened tracer right-hand side of src/solve_ABglobal.c:184-191.
 * include/compat/superlu_ddefs.h maps those SuperLU names onto the calls below so the
 * reference's main()s build unchanged; new callers use this header directly.
 *
 * Plain C: pointers and sizes only.  All functions return 0 on success, a negative
 * NKP_E* code on failure; nothing here ever falls back to a CPU solver -- if CUDA or a
 * GPU is unavailable the call fails with NKP_ECUDA.
 */
#ifndef NKPRECOND_H
#define NKPRECOND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NKP_OK 0
#define NKP_EINVAL (-1)
#define NKP_EANALYSIS (-2)
#define NKP_ECUDA (-3)
#define NKP_ENOMEM (-4)
#define NKP_ESTATE (-5)

typedef struct nkp_solver nkp_solver;

/* Options; zero-initialise then call nkp_default_options. */
typedef struct nkp_options {
    int nb;              /* pivot block width (<= 64)                                    */
    int leaf;            /* nested dissection stops below this many unknowns; subtrees of */
                         /* the elimination tree up to this size stay one front          */
    int equil;           /* 1: row/column equilibration (SuperLU Equil=YES)              */
    int refine_max;      /* max refinement steps (SuperLU IterRefine=SLU_DOUBLE, ITMAX)  */
    int device;          /* CUDA device ordinal                                          */
    int verbose;         /* 0 silent, 1 phase summary on stderr                          */
    int refine_rule;     /* 0: SuperLU's pdgsrfs rule -- componentwise berr <= eps or no     */
                         /*    halving (default, what the reference gets);                    */
                         /* 1: normwise -- stop once ||b - A x||_2 <= 1e-14 ||b||_2 for every */
                         /*    right-hand side (or no halving); typically 2-3 steps fewer     */
    int residual_extra;  /* 1 (default): the refinement residual b - A x is accumulated in twice the      */
                         /*    working precision (LAPACK xGERFSX-style extra-precise refinement): the     */
                         /*    iteration converges to the solution of the double-precision system instead */
                         /*    of wandering at cond(A) * eps, and SuperLU's berr rule is met 2-3 steps    */
                         /*    earlier (gx1v6-shape: 3 steps instead of 4-7, profiles/r02_refine_probe_   */
                         /*    gx1v6.log);  0: working precision, exactly what pdgsrfs does               */
    int reserved[8];
} nkp_options;

/* Statistics in the spirit of PStatPrint (src/solve_ABglobal.c:351-360). */
typedef struct nkp_stats {
    int n;
    int64_t nnz;
    int n_fronts;
    int n_levels;
    int max_front;
    int64_t nnz_lu;          /* stored entries of L + U (incl. diagonal)          */
    double factor_flops;     /* algorithmic flops of the numeric factorisation    */
    double heap_bytes;       /* device bytes: factors + update-matrix pools       */
    double t_analysis;       /* host seconds: ordering + symbolic + plan          */
    double t_factor;         /* device seconds of the last numeric factorisation  */
    double t_scatter;        /* ... of which zero-fill + CRS->front scatter       */
    double t_solve;          /* device seconds of the last solve (all rhs, incl. refinement) */
    int refine_steps;        /* refinement steps taken by the last solve          */
    int tiny_pivots;         /* pivots replaced in the last factorisation         */
    int64_t kernel_launches; /* kernels launched by this handle so far            */
    double solve_bytes;      /* algorithmic bytes of one forward+backward sweep   */
    /* per-kernel device times of the last factorisation (only with nkp_set_profile(s,1)) */
    double t_gemm;           /* Schur-update GEMM launches, seconds                */
    double gemm_flops;       /* algorithmic flops of those launches                */
    int64_t n_gemm;          /* number of GEMM launches                            */
    double t_trsm;
    double t_diag;
    double t_extend_add;
    double t_sweeps;         /* device seconds of the last raw sweep pair(s) (nkp_sweeps_device) */
    double factor_flops_local; /* multi-GPU: flops of the fronts owned by this rank       */
    double nnz_lu_local;
    double n_xfers;          /* parent/child pairs whose data crosses GPUs               */
    double order_cached;     /* 1: the ordering came from the on-disk analysis cache      */
    double reserved[4];
} nkp_stats;

void nkp_default_options(nkp_options* opt);

/* On-disk cache of the pattern-dependent ordering (the expensive part of the analysis), keyed by a hash
 * of (rowptr, colind, coordinates, ordering options).  The reference repeats the whole analysis in every
 * process (src/solve_ABglobal.c:350-353); with a cache directory set, nkp_create stores the nested-
 * dissection ordering and its assembly tree there and later processes -- e.g. the next Newton iteration's solver run -- read it
 * back.  Process-wide; dir = NULL or "" disables it.  Without this call the environment variable
 * NKP_ANALYSIS_CACHE is consulted.  A missing, stale or damaged file only means "recompute". */
int nkp_set_analysis_cache(const char* dir);

/* Analysis (ordering, symbolic factorisation, memory plan, task lists) for the pattern
 * (n, rowptr, colind), 0-based CRS as in src/matrix.c:84-88.  coord_i/j/k are optional
 * per-unknown grid coordinates (tracer_state_ind_to_{i,j,k}, src/matrix.c:322-329; repeat
 * them per coupled tracer); pass NULL to order from the graph alone.  Replaces the
 * ordering/symbolic half of the first pdgssvx* call.  The handle is reusable for any
 * number of nkp_factor calls with new values on the same pattern. */
int nkp_create(nkp_solver** out, int n, const int* rowptr, const int* colind,
               const int* coord_i, const int* coord_j, const int* coord_k,
               const nkp_options* opt);

/* Same as nkp_create with the pattern in FILE byte order: the big-endian NC_INT bytes of the `rowptr` (n + 1 values)
 * and `colind` (nnz values) variables exactly as they lie in the NetCDF-3 matrix file (src/matrix.c:3884,3888; locate
 * them with nkp_nc3_inq_var_extent).  The bytes are converted on the device; replaces the host loops behind
 * nc_get_var_int in get_sparse_matrix (src/matrix.c:3944-4031).  Coordinates stay host-order ints. */
int nkp_create_be(nkp_solver** out, int n, long long nnz, const void* rowptr_be, const void* colind_be,
                  const int* coord_i, const int* coord_j, const int* coord_k, const nkp_options* opt);

/* Multi-GPU (one process per GPU): every rank calls nkp_create_dist with the
 * same pattern; rank 0 obtains `unique_id` (NKP_UNIQUE_ID_BYTES bytes) from
 * nkp_comm_unique_id and ships it to the other ranks by any means (the bench uses
 * torch.distributed).  Independent nested-dissection subtrees are factored and swept on their
 * owner GPU; only the update matrices / vectors of the top separator fronts and the
 * separator solutions cross NVLink (NCCL send/recv/broadcast).  A, B and X are replicated
 * on every rank, like solve_ABglobal (src/solve_ABglobal.c:132-139,194). */
#define NKP_UNIQUE_ID_BYTES 128
int nkp_comm_unique_id(void* unique_id);
int nkp_create_dist(nkp_solver** out, int n, const int* rowptr, const int* colind,
                    const int* coord_i, const int* coord_j, const int* coord_k,
                    const nkp_options* opt, int rank, int nranks, const void* unique_id);

/* Static row permutation for a large diagonal -- what pdgssvx* does first under options->RowPerm = LargeDiag, the
 * default that set_default_options_dist leaves and the reference keeps (src/solve_ABglobal.c:332-334,
 * src/solve_ABdist.c:493-495): HSL MC64 job 5 (Duff & Koster 2001), independently implemented (csrc/rowperm.cpp).
 * HOST function, no GPU involved.  rowmap[i] = the column whose diagonal position row i takes (the permutation that
 * maximises the product of the diagonal magnitudes); row_scale / col_scale (may be NULL) receive the scalings from
 * the dual variables: row_scale[i] * |a_ij| * col_scale[j] <= 1 everywhere, = 1 on the new diagonal.
 * NKP_EANALYSIS: the matrix is structurally singular (empty row / column, no perfect matching). */
int nkp_rowperm_largediag(int n, const int* rowptr, const int* colind, const double* nzval, int* rowmap,
                          double* row_scale, double* col_scale);

/* nkp_create / nkp_create_dist with a static row permutation: row i of the operand becomes row rowmap[i] of the
 * matrix that is ordered and factored (from nkp_rowperm_largediag, or the caller's own -- SuperLU's MY_PERMR).
 * row_scale / col_scale (both or neither) replace the solver's own equilibration in every factorisation of the
 * handle -- SuperLU's SamePattern_SameRowPerm reuse of perm_r, R, C; they are rounded to powers of two.  Nothing
 * else changes for the caller: values keep their CRS order, B, X, residuals and berr are in terms of the original
 * rows.  nranks = 1: single GPU (unique_id may be NULL); otherwise as nkp_create_dist.
 * The ordering then works on the pattern of the permuted matrix, which for this operator family is wider than the
 * stencil (more fill); the default path (no row permutation, equilibration + tiny-pivot replacement + refinement)
 * meets the reference's accuracy on every operand measured so far, see DESIGN.md section 2. */
int nkp_create_rowperm(nkp_solver** out, int n, const int* rowptr, const int* colind,
                       const int* coord_i, const int* coord_j, const int* coord_k, const nkp_options* opt,
                       const int* rowmap, const double* row_scale, const double* col_scale,
                       int rank, int nranks, const void* unique_id);

/* Numeric factorisation from HOST values (nzval_row_wise, src/matrix.c:84), includes the
 * host->device copy.  Replaces pdgssvx*(nrhs = 0). */
int nkp_factor(nkp_solver* s, const double* nzval);
/* Same from values in FILE byte order: nnz big-endian IEEE doubles exactly as they lie in the
 * nzval_row_wise variable of the NetCDF-3 matrix file (src/matrix.c:3880; locate them with
 * nkp_nc3_inq_var_extent, include/nkp_nc3.h).  The bytes travel to the device unchanged and are swapped
 * there, replacing the host loop behind nc_get_var_double in get_sparse_matrix (src/matrix.c:3996). */
int nkp_factor_be(nkp_solver* s, const void* nzval_be);
/* Same with the values already resident in device memory. */
int nkp_factor_device(nkp_solver* s, const double* d_nzval);

/* Solve A X = B in place, B column-major n x nrhs with leading dimension ldb, HOST
 * memory; iterative refinement included; berr[nrhs] (may be NULL) receives the
 * componentwise backward errors.  Replaces pdgssvx*(Fact = FACTORED, nrhs >= 1). */
int nkp_solve(nkp_solver* s, double* B, int ldb, int nrhs, double* berr);
/* Same with B in device memory (berr stays a host pointer). */
int nkp_solve_device(nkp_solver* s, double* d_B, int ldb, int nrhs, double* berr);

/* Multi-GPU with a DISTRIBUTED right-hand side -- what pdgssvx does for solve_ABdist (src/solve_ABdist.c:141-144,
 * :571): rank r passes rows [fst_row, fst_row + m_loc) of B (HOST memory, column-major m_loc x nrhs, leading
 * dimension ldb >= m_loc) and receives the same rows of X in place; the slabs of all ranks must tile [0, n) (any
 * sizes; the reference uses n / P with the remainder on the last rank).  Every rank moves only its slab across
 * PCIe; the slabs are exchanged between the GPUs over NVLink (NCCL).  berr[nrhs] is the same on every rank.
 * With one rank this is nkp_solve (fst_row = 0, m_loc = n). */
int nkp_solve_dist(nkp_solver* s, double* B_loc, int ldb, int nrhs, int fst_row, int m_loc, double* berr);

/* Tracer fields in, tracer fields out -- get_B_global + pdgssvx*(FACTORED) + put_B_global of the
 * reference in one call (src/solve_ABglobal.c:154-208, :395, :213-267).
 * nkp_set_tracer_maps registers the index maps of the matrix file (tracer_state_ind_to_{i,j,k},
 * src/matrix.c:322-329; grid dimensions as read by get_grid_dims, src/grid.c:34); n of the handle must
 * equal coupled_tracer_cnt * tracer_state_len (src/matrix.c:471).
 * nkp_solve_fields takes `nfields` HOST pointers to 3-D double fields of km*jmt*imt values, [k][j][i]
 * (what get_var_3d_double returns, src/file_io.c:273).  Every `coupled_tracer_cnt` consecutive fields
 * form one right-hand side (src/solve_ABglobal.c:373-388); ALL systems are solved as one batched
 * multi-RHS solve (the reference loops with nrhs = 1).  The ocean points of the fields are gathered
 * and scattered by device kernels; land values are left untouched (src/solve_ABglobal.c:236-248).
 * berr (may be NULL) receives nfields / coupled_tracer_cnt backward errors.  A field count that is not
 * a multiple of coupled_tracer_cnt is NKP_EINVAL (fatal in the reference, src/solve_ABglobal.c:376-379). */
int nkp_set_tracer_maps(nkp_solver* s, int tracer_state_len, int coupled_tracer_cnt, const int* ind_i,
                        const int* ind_j, const int* ind_k, int imt, int jmt, int km);
int nkp_solve_fields(nkp_solver* s, double* const* fields, int nfields, double* berr);

/* The post-processing that ends every assembly of the reference's generator (gen_sparse_matrix, src/matrix.c:3775-3840)
 * on a CRS held in DEVICE memory, in place: sum_dup_vals (src/matrix.c:3621-3650; same order of additions: bit-exact),
 * then -- if strip_zeros != 0 -- strip_matrix_zeros (:3657-3688; rowptr is rebuilt, *nnz_out receives the new count),
 * then sort_cols_all_rows (:3753-3765).  strip_zeros = 0 keeps explicit zeros, i.e. the slot pattern, which is what a
 * refactorisation with new values on the same analysis needs.  dup_cnt_out (may be NULL) receives the reference's
 * dup_cnt.  Uses the current CUDA device. */
int nkp_crs_finalize_device(int n, int* d_rowptr, int* d_colind, double* d_val, int strip_zeros, long long* nnz_out,
                            int* dup_cnt_out);
/* Stencil values on the device for the generator's option set  adv_type centered / hmix_type const / vmix_type const /
 * sink_type const_shallow <sink_rate> <sink_depth>  (what gen_sparse_matrix computes before its post-processing for
 * that option set, src/matrix.c:3790-3827; bit-identical to gen_A after nkp_crs_finalize_device).  All pointers of
 * nkp_min_fields are DEVICE pointers: KMT[jmt][imt]; the index maps of the matrix file (src/matrix.c:309-329);
 * dz, z_t [km]; 2-D fields [jmt][imt]; velocities [km][jmt][imt] with fill_value marking land (src/matrix.c:984-1217).
 * Writes rowptr (n + 1), colind and val in slot order (self, k-1, k+1, east, west, north, south), exact zeros kept;
 * capacity = number of entries colind / val can hold (7 n always suffices); *nnz_out = entries written. */
typedef struct nkp_min_fields {
    int imt, jmt, km, n;
    const int *KMT, *ind_i, *ind_j, *ind_k, *int3_to_tracer_state_ind;
    const double *dz, *z_t, *TAREA, *HTE, *HUS, *HTN, *HUW, *DXU, *DYU, *UVEL, *VVEL, *WVEL;
    double fill_value;
} nkp_min_fields;
int nkp_assemble_min_device(const nkp_min_fields* fields, double day_cnt, double sink_rate, double sink_depth,
                            int* d_rowptr, int* d_colind, double* d_val, long long capacity, long long* nnz_out);
/* count big-endian 32-bit integers in device memory -> host byte order, in place (NC_INT arrays of the matrix file). */
int nkp_bswap32_device(void* d_data, long long count);

/* Residual r = b - A x for the currently loaded values (device pointers, column-major
 * with leading dimension n); the refinement SpMV exposed for testing and measurement. */
int nkp_residual_device(nkp_solver* s, const double* d_x, const double* d_b, double* d_r, int nrhs);

/* One forward+backward sweep pair without refinement (device pointers): d_B is
 * overwritten by the solution of the factored (equilibrated, permuted) system. */
int nkp_sweeps_device(nkp_solver* s, double* d_B, int ldb, int nrhs);

/* The fill-reducing permutation: perm[old] = new (n entries). */
int nkp_get_perm(const nkp_solver* s, int* perm);
int nkp_get_stats(const nkp_solver* s, nkp_stats* st);
/* on != 0: bracket every kernel of nkp_factor* with CUDA events on the solver's stream and
 * report per-kernel-class times in nkp_stats (small overhead; off by default). */
int nkp_set_profile(nkp_solver* s, int on);
/* Change the refinement stopping rule of an existing handle (see nkp_options.refine_rule). */
int nkp_set_refine_rule(nkp_solver* s, int rule);
/* Switch the extra-precise residual of an existing handle on or off (see nkp_options.residual_extra). */
int nkp_set_residual_extra(nkp_solver* s, int on);
/* Diagnostics on stderr: 0 silent, 1 phase summary, 2 timeline of the factorisation, 3 + per-launch sweep trace. */
int nkp_set_verbose(nkp_solver* s, int level);
/* Block until all device work of this handle has finished. */
int nkp_sync(nkp_solver* s);
void nkp_destroy(nkp_solver* s);

const char* nkp_last_error(void);
const char* nkp_version(void);

#ifdef __cplusplus
}
#endif
#endif
