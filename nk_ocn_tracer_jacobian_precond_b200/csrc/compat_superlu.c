/*
 * compat_superlu.c -- SuperLU_DIST-named entry points (include/compat/superlu_ddefs.h)
 * implemented on the B200 solver's native C ABI (include/nkprecond.h), so that the
 * reference's solve_ABglobal.c and solve_ABdist.c compile and run UNCHANGED:
 *
 *   pdgssvx_ABglobal(nrhs=0) / pdgssvx(nrhs=0)   -> nkp_create + nkp_factor
 *   pdgssvx_ABglobal(nrhs>0) / pdgssvx(nrhs>0)   -> nkp_solve
 *   Destroy_LU                                   -> nkp_destroy
 *
 * (call sites: src/solve_ABglobal.c:353,395,413 ; src/solve_ABdist.c:518,571,593).
 *
 * Process model: one host process drives the GPU(s); the drivers' MPI calls are served by
 * compat_mpi.c (rank 0 only, mailbox for the virtual ranks).  pdgssvx() reassembles the
 * block-row slabs that get_sparse_matrix_dist / get_B_dist "sent" (tags 0-3) and posts the
 * solution slabs for put_B_dist (tag 4).
 *
 * Ordering: the geometric nested dissection wants the (i,j,k) coordinates of the unknowns.
 * The SuperLU API has no slot for them, but both drivers keep the matrix file name in the
 * global `matrix_fname`; when that symbol is visible (executables linked with -rdynamic)
 * the index maps of the matrix file (src/matrix.c:322-329) are read through whatever
 * NetCDF provider the program is linked with.  Otherwise the graph-based dissection is used.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nkprecond.h"
#include "superlu_ddefs.h"

extern char *matrix_fname __attribute__ ((weak));
extern int dbg_lvl __attribute__ ((weak));

typedef struct {
   nkp_solver *h;
   int n;
   int nnz;
   int *rowptr;
   int *colind;
   double *val;
   int *csc_to_crs;             /* ABglobal: CRS position of every CSC entry */
   int nparts;                  /* ABdist: virtual ranks */
} compat_state;

/* ---- memory / abort ---------------------------------------------------------------------- */

void *
superlu_malloc_dist (size_t size)
{
   return malloc (size ? size : 1);
}

void
superlu_free_dist (void *addr)
{
   free (addr);
}

void
superlu_abort_and_exit_dist (char *msg)
{
   fprintf (stderr, "%s", msg);
   exit (-1);
}

double *
doubleMalloc_dist (int_t n)
{
   return (double *) superlu_malloc_dist ((size_t) (n > 0 ? n : 0) * sizeof (double));
}

int_t *
intMalloc_dist (int_t n)
{
   return (int_t *) superlu_malloc_dist ((size_t) (n > 0 ? n : 0) * sizeof (int_t));
}

/* ---- grid ------------------------------------------------------------------------------------ */

void
superlu_gridinit (MPI_Comm Bcomm, int_t nprow, int_t npcol, gridinfo_t * grid)
{
   memset (grid, 0, sizeof (*grid));
   grid->comm = Bcomm;
   grid->rscp.comm = grid->cscp.comm = Bcomm;
   grid->rscp.Np = (int) npcol;
   grid->cscp.Np = (int) nprow;
   grid->iam = 0;
   grid->nprow = nprow;
   grid->npcol = npcol;
}

void
superlu_gridexit (gridinfo_t * grid)
{
   (void) grid;
}

/* ---- matrices -------------------------------------------------------------------------------- */

void
dallocateA_dist (int_t n, int_t nnz, double **a, int_t ** asub, int_t ** xa)
{
   *a = doubleMalloc_dist (nnz);
   *asub = intMalloc_dist (nnz);
   *xa = intMalloc_dist (n + 1);
}

void
dCompRow_to_CompCol_dist (int_t m, int_t n, int_t nnz, double *a, int_t * colind, int_t * rowptr,
                          double **at, int_t ** rowind, int_t ** colptr)
{
   int_t i, j, p;
   int_t *marker;
   dallocateA_dist (n, nnz, at, rowind, colptr);
   marker = intMalloc_dist (n + 1);
   for (j = 0; j <= n; j++)
      marker[j] = 0;
   for (p = 0; p < nnz; p++)
      marker[colind[p] + 1]++;
   (*colptr)[0] = 0;
   for (j = 0; j < n; j++)
      (*colptr)[j + 1] = (*colptr)[j] + marker[j + 1];
   for (j = 0; j < n; j++)
      marker[j] = (*colptr)[j];
   for (i = 0; i < m; i++)
      for (p = rowptr[i]; p < rowptr[i + 1]; p++) {
         j = colind[p];
         (*rowind)[marker[j]] = i;
         (*at)[marker[j]] = a[p];
         marker[j]++;
      }
   superlu_free_dist (marker);
}

void
dCreate_CompCol_Matrix_dist (SuperMatrix * A, int_t m, int_t n, int_t nnz, double *nzval, int_t * rowind,
                             int_t * colptr, Stype_t stype, Dtype_t dtype, Mtype_t mtype)
{
   NCformat *st = (NCformat *) superlu_malloc_dist (sizeof (NCformat));
   A->Stype = stype;
   A->Dtype = dtype;
   A->Mtype = mtype;
   A->nrow = m;
   A->ncol = n;
   A->Store = st;
   st->nnz = nnz;
   st->nzval = nzval;
   st->rowind = rowind;
   st->colptr = colptr;
}

void
dCreate_CompRowLoc_Matrix_dist (SuperMatrix * A, int_t m, int_t n, int_t nnz_loc, int_t m_loc, int_t fst_row,
                                double *nzval, int_t * colind, int_t * rowptr, Stype_t stype, Dtype_t dtype,
                                Mtype_t mtype)
{
   NRformat_loc *st = (NRformat_loc *) superlu_malloc_dist (sizeof (NRformat_loc));
   A->Stype = stype;
   A->Dtype = dtype;
   A->Mtype = mtype;
   A->nrow = m;
   A->ncol = n;
   A->Store = st;
   st->nnz_loc = nnz_loc;
   st->m_loc = m_loc;
   st->fst_row = fst_row;
   st->nzval = nzval;
   st->rowptr = rowptr;
   st->colind = colind;
}

void
Destroy_CompCol_Matrix_dist (SuperMatrix * A)
{
   NCformat *st = (NCformat *) A->Store;
   if (st == NULL)
      return;
   superlu_free_dist (st->rowind);
   superlu_free_dist (st->colptr);
   superlu_free_dist (st->nzval);
   superlu_free_dist (st);
   A->Store = NULL;
}

void
Destroy_CompRowLoc_Matrix_dist (SuperMatrix * A)
{
   NRformat_loc *st = (NRformat_loc *) A->Store;
   if (st == NULL)
      return;
   superlu_free_dist (st->rowptr);
   superlu_free_dist (st->colind);
   superlu_free_dist (st->nzval);
   superlu_free_dist (st);
   A->Store = NULL;
}

/* ---- options / structs --------------------------------------------------------------------------- */

void
set_default_options_dist (superlu_dist_options_t * o)
{
   memset (o, 0, sizeof (*o));
   o->Fact = DOFACT;
   o->Equil = YES;
   o->ParSymbFact = NO;
   o->ColPerm = METIS_AT_PLUS_A;
   o->RowPerm = LargeDiag;
   o->ReplaceTinyPivot = YES;
   o->IterRefine = SLU_DOUBLE;
   o->Trans = NOTRANS;
   o->SolveInitialized = NO;
   o->RefineInitialized = NO;
   o->PrintStat = YES;
   o->num_lookaheads = 10;
   o->lookahead_etree = NO;
   o->SymPattern = NO;
}

void
print_options_dist (superlu_dist_options_t * o)
{
   printf ("**************************************************\n");
   printf (".. options (B200 solver behind the SuperLU_DIST API):\n");
   printf ("**    Fact             : %4d\n", o->Fact);
   printf ("**    Equil            : %4d\n", o->Equil);
   printf ("**    ParSymbFact      : %4d  (analysis runs on the host)\n", o->ParSymbFact);
   printf ("**    ColPerm          : %4d  (nested dissection, built in)\n", o->ColPerm);
   printf ("**    RowPerm          : %4d  (static pivoting; LargeDiag row permutation only with NKP_ROWPERM=1)\n", o->RowPerm);
   printf ("**    ReplaceTinyPivot : %4d\n", o->ReplaceTinyPivot);
   printf ("**    IterRefine       : %4d\n", o->IterRefine);
   printf ("**************************************************\n");
}

void
ScalePermstructInit (const int_t m, const int_t n, ScalePermstruct_t * sp)
{
   int_t i;
   sp->DiagScale = NOEQUIL;
   sp->R = NULL;
   sp->C = NULL;
   sp->perm_r = intMalloc_dist (m);
   sp->perm_c = intMalloc_dist (n);
   for (i = 0; i < m; i++)
      sp->perm_r[i] = i;
   for (i = 0; i < n; i++)
      sp->perm_c[i] = i;
}

void
ScalePermstructFree (ScalePermstruct_t * sp)
{
   superlu_free_dist (sp->perm_r);
   superlu_free_dist (sp->perm_c);
   superlu_free_dist (sp->R);
   superlu_free_dist (sp->C);
   sp->perm_r = sp->perm_c = NULL;
   sp->R = sp->C = NULL;
}

void
LUstructInit (const int_t n, LUstruct_t * lu)
{
   (void) n;
   memset (lu, 0, sizeof (*lu));
}

static void
free_state (compat_state * cs)
{
   if (cs == NULL)
      return;
   if (cs->h)
      nkp_destroy (cs->h);
   free (cs->rowptr);
   free (cs->colind);
   free (cs->val);
   free (cs->csc_to_crs);
   free (cs);
}

void
Destroy_LU (int_t n, gridinfo_t * grid, LUstruct_t * lu)
{
   (void) n;
   (void) grid;
   free_state ((compat_state *) lu->nkp);
   lu->nkp = NULL;
}

void
LUstructFree (LUstruct_t * lu)
{
   (void) lu;
}

void
dSolveFinalize (superlu_dist_options_t * options, SOLVEstruct_t * SOLVEstruct)
{
   (void) SOLVEstruct;
   options->SolveInitialized = NO;
}

/* ---- statistics ---------------------------------------------------------------------------------- */

void
PStatInit (SuperLUStat_t * stat)
{
   memset (stat, 0, sizeof (*stat));
}

void
PStatFree (SuperLUStat_t * stat)
{
   (void) stat;
}

void
PStatPrint (superlu_dist_options_t * options, SuperLUStat_t * stat, gridinfo_t * grid)
{
   (void) options;
   if (grid->iam != 0 || !stat->valid)
      return;
   printf ("**************************************************\n");
   printf ("**** Time (seconds) ****\n");
   if (stat->t_analysis > 0)
      printf ("\tANALYSIS time   %8.3f\n", stat->t_analysis);
   if (stat->t_factor > 0) {
      printf ("\tSCATTER time    %8.3f\n", stat->t_scatter);
      printf ("\tFACTOR time     %8.3f\n", stat->t_factor);
      printf ("\tFactor flops\t%e\tGflops \t%8.2f\n", stat->flops_factor, stat->flops_factor * 1e-9 / stat->t_factor);
      printf ("\tnnz(L+U)        %lld\n", stat->nnz_lu);
   }
   if (stat->t_solve > 0) {
      printf ("\tSOLVE+REFINE time %8.3f\n", stat->t_solve);
      printf ("\tRefinement steps  %d\n", stat->refine_steps);
   }
   printf ("**************************************************\n");
}

static void
fill_stat (compat_state * cs, SuperLUStat_t * stat, int factored_now, int solved_now)
{
   nkp_stats st;
   if (stat == NULL || nkp_get_stats (cs->h, &st))
      return;
   stat->valid = 1;
   if (factored_now) {
      stat->t_analysis = st.t_analysis;
      stat->t_scatter = st.t_scatter;
      stat->t_factor = st.t_factor;
      stat->flops_factor = st.factor_flops;
      stat->nnz_lu = (long long) st.nnz_lu;
   }
   if (solved_now) {
      stat->t_solve = st.t_solve;
      stat->refine_steps = st.refine_steps;
   }
}

/* ---- coordinates side channel ------------------------------------------------------------------ */

typedef int (*nc_open_t) (const char *, int, int *);
typedef int (*nc_close_t) (int);
typedef int (*nc_inq_varid_t) (int, const char *, int *);
typedef int (*nc_inq_dimid_t) (int, const char *, int *);
typedef int (*nc_inq_dimlen_t) (int, int, size_t *);
typedef int (*nc_get_var_int_t) (int, int, int *);

/* returns 0 and three malloc'ed arrays of length n on success */
static int
load_coords (int n, int **ci, int **cj, int **ck)
{
   nc_open_t p_open;
   nc_close_t p_close;
   nc_inq_varid_t p_varid;
   nc_inq_dimid_t p_dimid;
   nc_inq_dimlen_t p_dimlen;
   nc_get_var_int_t p_get;
   const char *names[3] = { "tracer_state_ind_to_i", "tracer_state_ind_to_j", "tracer_state_ind_to_k" };
   int *out[3] = { NULL, NULL, NULL };
   int ncid, dimid, varid, d, rep;
   size_t len = 0;

   if (&matrix_fname == NULL || matrix_fname == NULL)
      return 1;
   p_open = (nc_open_t) dlsym (RTLD_DEFAULT, "nc_open");
   p_close = (nc_close_t) dlsym (RTLD_DEFAULT, "nc_close");
   p_varid = (nc_inq_varid_t) dlsym (RTLD_DEFAULT, "nc_inq_varid");
   p_dimid = (nc_inq_dimid_t) dlsym (RTLD_DEFAULT, "nc_inq_dimid");
   p_dimlen = (nc_inq_dimlen_t) dlsym (RTLD_DEFAULT, "nc_inq_dimlen");
   p_get = (nc_get_var_int_t) dlsym (RTLD_DEFAULT, "nc_get_var_int");
   if (!p_open || !p_close || !p_varid || !p_dimid || !p_dimlen || !p_get)
      return 1;
   if (p_open (matrix_fname, 0, &ncid))
      return 1;
   if (p_dimid (ncid, "tracer_state_len", &dimid) || p_dimlen (ncid, dimid, &len) || len == 0 || n % (int) len != 0) {
      p_close (ncid);
      return 1;
   }
   rep = n / (int) len;
   for (d = 0; d < 3; d++) {
      int t;
      out[d] = (int *) malloc ((size_t) n * sizeof (int));
      if (out[d] == NULL || p_varid (ncid, names[d], &varid) || p_get (ncid, varid, out[d])) {
         p_close (ncid);
         free (out[0]);
         free (out[1]);
         free (out[2]);
         return 1;
      }
      for (t = 1; t < rep; t++)
         memcpy (out[d] + (size_t) t * len, out[d], len * sizeof (int));
   }
   p_close (ncid);
   *ci = out[0];
   *cj = out[1];
   *ck = out[2];
   return 0;
}

static int
create_and_factor (compat_state * cs, superlu_dist_options_t * options, ScalePermstruct_t * sp, int *info)
{
   int *ci = NULL, *cj = NULL, *ck = NULL;
   nkp_options o;
   int rc;
   int have = (load_coords (cs->n, &ci, &cj, &ck) == 0);
   nkp_default_options (&o);
   if (&dbg_lvl != NULL && dbg_lvl > 0)
      o.verbose = 1;
   if (&dbg_lvl != NULL && dbg_lvl)
      printf ("(0) nkp: ordering = %s nested dissection\n", have ? "geometric (index maps of the matrix file)" : "graph");
   /* options->RowPerm = LargeDiag is what set_default_options_dist leaves and the reference keeps
    * (src/solve_ABglobal.c:332-334).  The B200 solver meets the reference's accuracy on this operator family without
    * a row permutation (it only adds fill here), so the permutation is computed only on request:
    * NKP_ROWPERM=1 in the environment.  perm_r then reports it, as SuperLU's ScalePermstruct does. */
   if (options && options->RowPerm == LargeDiag && getenv ("NKP_ROWPERM") && atoi (getenv ("NKP_ROWPERM")) > 0) {
      int *rowmap = (int *) malloc ((size_t) cs->n * sizeof (int));
      double *rs = (double *) malloc ((size_t) cs->n * sizeof (double));
      double *csc = (double *) malloc ((size_t) cs->n * sizeof (double));
      rc = (rowmap && rs && csc) ? nkp_rowperm_largediag (cs->n, cs->rowptr, cs->colind, cs->val, rowmap, rs, csc) : NKP_ENOMEM;
      if (rc == 0) {
         int i, moved = 0;
         for (i = 0; i < cs->n; i++)
            moved += rowmap[i] != i;
         if (&dbg_lvl != NULL && dbg_lvl)
            printf ("(0) nkp: RowPerm = LargeDiag moves %d of %d rows\n", moved, cs->n);
         rc = nkp_create_rowperm (&cs->h, cs->n, cs->rowptr, cs->colind, have ? ci : NULL, have ? cj : NULL,
                                  have ? ck : NULL, &o, rowmap, rs, csc, 0, 1, NULL);
         if (rc == 0 && sp && sp->perm_r)
            for (i = 0; i < cs->n; i++)
               sp->perm_r[i] = rowmap[i];
      }
      else
         fprintf (stderr, "(0) nkp_rowperm_largediag failed with code %d\n", rc);
      free (rowmap);
      free (rs);
      free (csc);
   }
   else
      rc = nkp_create (&cs->h, cs->n, cs->rowptr, cs->colind, have ? ci : NULL, have ? cj : NULL, have ? ck : NULL, &o);
   free (ci);
   free (cj);
   free (ck);
   if (rc) {
      fprintf (stderr, "(0) nkp_create failed: %s\n", nkp_last_error ());
      *info = -1;
      return rc;
   }
   rc = nkp_factor (cs->h, cs->val);
   if (rc) {
      fprintf (stderr, "(0) nkp_factor failed: %s\n", nkp_last_error ());
      *info = -1;
      return rc;
   }
   if (sp && sp->perm_c)
      nkp_get_perm (cs->h, sp->perm_c);
   if (sp)
      sp->DiagScale = BOTH;
   *info = 0;
   return 0;
}

/* ---- the drivers ------------------------------------------------------------------------------------ */

void
pdgssvx_ABglobal (superlu_dist_options_t * options, SuperMatrix * A, ScalePermstruct_t * ScalePermstruct,
                  double B[], int ldb, int nrhs, gridinfo_t * grid, LUstruct_t * LUstruct, double *berr,
                  SuperLUStat_t * stat, int *info)
{
   compat_state *cs = (compat_state *) LUstruct->nkp;
   int factored_now = 0;
   (void) grid;
   *info = 0;
   if (options->Fact != FACTORED) {
      NCformat *st = (NCformat *) A->Store;
      int n = A->nrow, nnz = st->nnz, i, j, p;
      const double *a = (const double *) st->nzval;
      int *next;
      free_state (cs);
      cs = (compat_state *) calloc (1, sizeof (compat_state));
      LUstruct->nkp = cs;
      cs->n = n;
      cs->nnz = nnz;
      cs->rowptr = (int *) calloc ((size_t) n + 1, sizeof (int));
      cs->colind = (int *) malloc ((size_t) nnz * sizeof (int));
      cs->val = (double *) malloc ((size_t) nnz * sizeof (double));
      cs->csc_to_crs = (int *) malloc ((size_t) nnz * sizeof (int));
      next = (int *) malloc ((size_t) n * sizeof (int));
      if (!cs->rowptr || !cs->colind || !cs->val || !cs->csc_to_crs || !next)
         ABORT ("Malloc fails in pdgssvx_ABglobal.");
      for (p = 0; p < nnz; p++)
         cs->rowptr[st->rowind[p] + 1]++;
      for (i = 0; i < n; i++)
         cs->rowptr[i + 1] += cs->rowptr[i];
      for (i = 0; i < n; i++)
         next[i] = cs->rowptr[i];
      for (j = 0; j < n; j++)
         for (p = st->colptr[j]; p < st->colptr[j + 1]; p++) {
            int q = next[st->rowind[p]]++;
            cs->colind[q] = j;
            cs->val[q] = a[p];
            cs->csc_to_crs[p] = q;
         }
      free (next);
      if (create_and_factor (cs, options, ScalePermstruct, info))
         return;
      factored_now = 1;
   }
   if (cs == NULL || cs->h == NULL) {
      *info = -1;
      return;
   }
   if (nrhs > 0) {
      if (nkp_solve (cs->h, B, ldb, nrhs, berr)) {
         fprintf (stderr, "(0) nkp_solve failed: %s\n", nkp_last_error ());
         *info = -1;
         return;
      }
   }
   fill_stat (cs, stat, factored_now, nrhs > 0);
}

static int
slab_count (int n, int nparts, int part)
{
   int count = n / nparts;
   if (part == nparts - 1)
      count = n - part * count;
   return count;
}

void
pdgssvx (superlu_dist_options_t * options, SuperMatrix * A, ScalePermstruct_t * ScalePermstruct,
         double B[], int ldb, int nrhs, gridinfo_t * grid, LUstruct_t * LUstruct,
         SOLVEstruct_t * SOLVEstruct, double *berr, SuperLUStat_t * stat, int *info)
{
   compat_state *cs = (compat_state *) LUstruct->nkp;
   int factored_now = 0;
   (void) SOLVEstruct;
   *info = 0;
   if (options->Fact != FACTORED) {
      NRformat_loc *st = (NRformat_loc *) A->Store;
      int n = A->nrow, nparts = (int) (grid->nprow * grid->npcol), d, i;
      int **rp = (int **) calloc ((size_t) nparts, sizeof (int *));
      int **cid = (int **) calloc ((size_t) nparts, sizeof (int *));
      double **vl = (double **) calloc ((size_t) nparts, sizeof (double *));
      long long nnz = st->nnz_loc;
      int pos;
      if (st->m_loc == n)
         nparts = 1;            /* a single real rank already holds everything */
      for (d = 1; d < nparts; d++) {
         size_t nb;
         int count = slab_count (n, nparts, d);
         rp[d] = (int *) nkp_mpi_mailbox_take (d, 0, &nb);
         cid[d] = (int *) nkp_mpi_mailbox_take (d, 1, &nb);
         vl[d] = (double *) nkp_mpi_mailbox_take (d, 2, &nb);
         if (!rp[d] || !cid[d] || !vl[d]) {
            fprintf (stderr, "(0) pdgssvx: slab of virtual rank %d not found; run with -n 1 or let rank 0 distribute\n", d);
            *info = -1;
            return;
         }
         nnz += rp[d][count] - rp[d][0];
      }
      free_state (cs);
      cs = (compat_state *) calloc (1, sizeof (compat_state));
      LUstruct->nkp = cs;
      cs->n = n;
      cs->nnz = (int) nnz;
      cs->nparts = nparts;
      cs->rowptr = (int *) malloc (((size_t) n + 1) * sizeof (int));
      cs->colind = (int *) malloc ((size_t) nnz * sizeof (int));
      cs->val = (double *) malloc ((size_t) nnz * sizeof (double));
      if (!cs->rowptr || !cs->colind || !cs->val)
         ABORT ("Malloc fails in pdgssvx.");
      for (i = 0; i <= st->m_loc; i++)
         cs->rowptr[i] = st->rowptr[i];
      memcpy (cs->colind, st->colind, (size_t) st->nnz_loc * sizeof (int));
      memcpy (cs->val, st->nzval, (size_t) st->nnz_loc * sizeof (double));
      pos = st->nnz_loc;
      for (d = 1; d < nparts; d++) {
         int count = slab_count (n, nparts, d), fst = d * (n / nparts);
         int cnt = rp[d][count] - rp[d][0];
         for (i = 0; i <= count; i++)
            cs->rowptr[fst + i] = pos + (rp[d][i] - rp[d][0]);
         memcpy (cs->colind + pos, cid[d], (size_t) cnt * sizeof (int));
         memcpy (cs->val + pos, vl[d], (size_t) cnt * sizeof (double));
         pos += cnt;
         free (rp[d]);
         free (cid[d]);
         free (vl[d]);
      }
      free (rp);
      free (cid);
      free (vl);
      if (create_and_factor (cs, options, ScalePermstruct, info))
         return;
      factored_now = 1;
      options->SolveInitialized = YES;
   }
   if (cs == NULL || cs->h == NULL) {
      *info = -1;
      return;
   }
   if (nrhs > 0) {
      int n = cs->n, nparts = cs->nparts, d, c;
      int m_loc = slab_count (n, nparts, 0);
      double *X = (double *) malloc ((size_t) n * (size_t) nrhs * sizeof (double));
      if (X == NULL)
         ABORT ("Malloc fails in pdgssvx.");
      for (c = 0; c < nrhs; c++)
         memcpy (X + (size_t) c * n, B + (size_t) c * ldb, (size_t) m_loc * sizeof (double));
      for (d = 1; d < nparts; d++) {
         size_t nb;
         int count = slab_count (n, nparts, d), fst = d * (n / nparts);
         double *slab = (double *) nkp_mpi_mailbox_take (d, 3, &nb);
         if (slab == NULL || nb < (size_t) count * (size_t) nrhs * sizeof (double)) {
            fprintf (stderr, "(0) pdgssvx: right-hand-side slab of virtual rank %d not found\n", d);
            *info = -1;
            free (X);
            return;
         }
         for (c = 0; c < nrhs; c++)
            memcpy (X + (size_t) c * n + fst, slab + (size_t) c * count, (size_t) count * sizeof (double));
         free (slab);
      }
      if (nkp_solve (cs->h, X, n, nrhs, berr)) {
         fprintf (stderr, "(0) nkp_solve failed: %s\n", nkp_last_error ());
         *info = -1;
         free (X);
         return;
      }
      for (c = 0; c < nrhs; c++)
         memcpy (B + (size_t) c * ldb, X + (size_t) c * n, (size_t) m_loc * sizeof (double));
      for (d = 1; d < nparts; d++) {
         int count = slab_count (n, nparts, d), fst = d * (n / nparts);
         double *slab = (double *) malloc ((size_t) count * (size_t) nrhs * sizeof (double));
         for (c = 0; c < nrhs; c++)
            memcpy (slab + (size_t) c * count, X + (size_t) c * n + fst, (size_t) count * sizeof (double));
         nkp_mpi_mailbox_post (d, 4, slab, (size_t) count * (size_t) nrhs * sizeof (double));
         free (slab);
      }
      free (X);
   }
   fill_stat (cs, stat, factored_now, nrhs > 0);
}
