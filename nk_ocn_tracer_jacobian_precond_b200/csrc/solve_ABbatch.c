/*
 * solve_ABbatch.c -- native C driver of the B200 tracer-Jacobian solver.
 *
 * Same program surface as the reference's solve_ABglobal / solve_ABdist
 * (src/solve_ABglobal.c:41, src/solve_ABdist.c:46):
 *
 *     solve_ABbatch [-D dbg_lvl] [-n nprow[,npcol]] [-v vars] matrix_fname inout_fname
 *
 * same CRS matrix file (src/matrix.c:3845-3939), same in-place tracer read/write on
 * inout_fname (src/solve_ABglobal.c:154-267), same "(%d) " output prefix and exit codes
 * (EXIT_FAILURE on parse / I/O / allocation errors, EXIT_SUCCESS otherwise -- also when the
 * solver reports info != 0, src/solve_ABglobal.c:354-357,430).
 *
 * What differs from the reference drivers: every tracer group named by -v is read first and
 * ALL groups are solved in ONE batched multi-RHS call (nkp_solve_fields) instead of one
 * nrhs = 1 call per group (src/solve_ABglobal.c:370-409), and the gather of ocean points /
 * scatter of the solution (src/solve_ABglobal.c:184-191, :242-248) runs on the GPU.
 * -n is accepted for command-line compatibility; this program drives one GPU.
 *
 * Host code stays in C and reaches CUDA only through include/nkprecond.h; files are read
 * with the nc_* subset of include/compat/netcdf.h (libnkp_nc3.so, or a real libnetcdf).
 */
#define _FILE_OFFSET_BITS 64
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <errno.h>
#include <limits.h>

#include "netcdf.h"
#include "nkprecond.h"
#ifndef NKP_NO_NC3_EXTENT
#include "nkp_nc3.h"            /* raw variable extents: only libnkp_nc3.so has them */
#endif

static int iam = 0;
static int dbg_lvl = 0;
static long nprow = 4, npcol = 4;   /* reference default grid, src/solve_ABglobal.c:296 */
static char *vars = NULL;
static char *matrix_fname = NULL;
static char *inout_fname = NULL;

static int
to_long (const char *str, long *out)
{
   char *end;
   long v;

   if (str == NULL || *str == '\0')
      return 1;
   errno = 0;
   v = strtol (str, &end, 10);
   if (errno != 0 || *end != '\0')
      return 1;
   *out = v;
   return 0;
}

static int
parse_cmd_line (int argc, char **argv)
{
   const char *usage_msg = "usage: jacobian_precond [-D dbg_lvl] [-n nprow[,npcol]] [-v vars] matrix_fname inout_fname";
   int opt;
   long lval;
   char *cp;

   while ((opt = getopt (argc, argv, "D:n:v:h")) != -1) {
      switch (opt) {
      case 'D':
         if (to_long (optarg, &lval) || lval < INT_MIN || lval > INT_MAX) {
            fprintf (stderr, "(%d) error parsing argument '%s' for option '%c'\n", iam, optarg, opt);
            return 1;
         }
         dbg_lvl = (int) lval;
         break;
      case 'n':
         cp = strtok (optarg, ",");
         if (to_long (cp, &lval)) {
            fprintf (stderr, "(%d) error parsing argument '%s' for option '%c'\n", iam, cp ? cp : "", opt);
            return 1;
         }
         nprow = npcol = lval;
         if ((cp = strtok (NULL, ",")) != NULL) {
            if (to_long (cp, &lval)) {
               fprintf (stderr, "(%d) error parsing argument '%s' for option '%c'\n", iam, cp, opt);
               return 1;
            }
            npcol = lval;
         }
         break;
      case 'v':
         if ((vars = malloc (strlen (optarg) + 1)) == NULL) {
            fprintf (stderr, "(%d) malloc failed in parse_cmd_line for vars\n", iam);
            return 1;
         }
         strcpy (vars, optarg);
         break;
      default:                 /* '?', 'h' */
         fprintf (stderr, "(%d) %s\n", iam, usage_msg);
         return 1;
      }
   }
   if (optind != argc - 2) {
      fprintf (stderr, "(%d) unexpected number of arguments\n%s\n", iam, usage_msg);
      return 1;
   }
   matrix_fname = argv[optind++];
   inout_fname = argv[optind++];
   return 0;
}

static int
nc_fail (int status, const char *what, const char *name)
{
   fprintf (stderr, "(%d) netCDF error in %s (%s): %s\n", iam, what, name ? name : "", nc_strerror (status));
   return 1;
}

static int
dim_len (int ncid, const char *name, size_t *len)
{
   int status, dimid;

   if ((status = nc_inq_dimid (ncid, name, &dimid)) != NC_NOERR)
      return nc_fail (status, "nc_inq_dimid", name);
   if ((status = nc_inq_dimlen (ncid, dimid, len)) != NC_NOERR)
      return nc_fail (status, "nc_inq_dimlen", name);
   return 0;
}

static int
read_int_var (int ncid, const char *name, int *dst)
{
   int status, varid;

   if ((status = nc_inq_varid (ncid, name, &varid)) != NC_NOERR)
      return nc_fail (status, "nc_inq_varid", name);
   if ((status = nc_get_var_int (ncid, varid, dst)) != NC_NOERR)
      return nc_fail (status, "nc_get_var_int", name);
   return 0;
}

/* the operand and the index maps of the matrix file */
typedef struct {
   int imt, jmt, km;
   int tracer_state_len, coupled_tracer_cnt;
   int flat_len, nnz;
   int *ind_i, *ind_j, *ind_k;   /* tracer_state_ind_to_{i,j,k} */
   int *rowptr, *colind;
   double *nzval;               /* host byte order, or ... */
   int nzval_is_file_order;     /* ... 1: the raw big-endian bytes of the file (swapped on the GPU) */
   int index_is_file_order;     /* 1: rowptr / colind hold the raw big-endian NC_INT bytes of the file, too */
} matrix_file_t;

#ifndef NKP_NO_NC3_EXTENT
/* Matrix-file ingest without the host conversion loops of nc_get_var_* (src/matrix.c:3978-4000): the bytes of a
 * fixed-size variable exactly as they lie in the file; the GPU converts them (nkp_create_be / nkp_factor_be). */
static int
read_raw_var (const char *fname, int ncid, const char *name, int want_type, size_t want_bytes, void *dst)
{
   long long off = 0, nbytes = 0;
   int xtype = 0, varid, ok = 0;
   FILE *fp;

   if (nc_inq_varid (ncid, name, &varid) != NC_NOERR)
      return 0;
   if (nkp_nc3_inq_var_extent (ncid, varid, &off, &nbytes, &xtype) != NC_NOERR || xtype != want_type
       || nbytes != (long long) want_bytes || (fp = fopen (fname, "rb")) == NULL)
      return 0;
   if (fseeko (fp, (off_t) off, SEEK_SET) == 0 && fread (dst, 1, want_bytes, fp) == want_bytes)
      ok = 1;
   fclose (fp);
   return ok;
}
#endif

static int
read_matrix_file (const char *fname, matrix_file_t * mf)
{
   int status, ncid, varid;
   size_t len;

   memset (mf, 0, sizeof (*mf));
   if ((status = nc_open (fname, NC_NOWRITE, &ncid)) != NC_NOERR)
      return nc_fail (status, "nc_open", fname);
   if (dim_len (ncid, "nlon", &len))
      return 1;
   mf->imt = (int) len;
   if (dim_len (ncid, "nlat", &len))
      return 1;
   mf->jmt = (int) len;
   if (dim_len (ncid, "z_t", &len))
      return 1;
   mf->km = (int) len;
   if (dim_len (ncid, "tracer_state_len", &len))
      return 1;
   mf->tracer_state_len = (int) len;
   if (dim_len (ncid, "nnz", &len))
      return 1;
   mf->nnz = (int) len;
   if (dim_len (ncid, "flat_len_p1", &len))
      return 1;
   mf->flat_len = (int) len - 1;
   if (read_int_var (ncid, "coupled_tracer_cnt", &mf->coupled_tracer_cnt))
      return 1;
   if (mf->coupled_tracer_cnt < 1 || (long) mf->coupled_tracer_cnt * mf->tracer_state_len != mf->flat_len) {
      fprintf (stderr, "(%d) inconsistent matrix file: flat_len = %d, coupled_tracer_cnt = %d, tracer_state_len = %d\n",
               iam, mf->flat_len, mf->coupled_tracer_cnt, mf->tracer_state_len);
      return 1;
   }
   mf->ind_i = malloc (sizeof (int) * (size_t) mf->tracer_state_len);
   mf->ind_j = malloc (sizeof (int) * (size_t) mf->tracer_state_len);
   mf->ind_k = malloc (sizeof (int) * (size_t) mf->tracer_state_len);
   mf->rowptr = malloc (sizeof (int) * ((size_t) mf->flat_len + 1));
   mf->colind = malloc (sizeof (int) * (size_t) mf->nnz);
   mf->nzval = malloc (sizeof (double) * (size_t) mf->nnz);
   if (!mf->ind_i || !mf->ind_j || !mf->ind_k || !mf->rowptr || !mf->colind || !mf->nzval) {
      fprintf (stderr, "(%d) malloc failed in read_matrix_file\n", iam);
      return 1;
   }
   if (read_int_var (ncid, "tracer_state_ind_to_i", mf->ind_i) || read_int_var (ncid, "tracer_state_ind_to_j", mf->ind_j)
       || read_int_var (ncid, "tracer_state_ind_to_k", mf->ind_k))
      return 1;
#ifndef NKP_NO_NC3_EXTENT
   if (read_raw_var (fname, ncid, "rowptr", NC_INT, sizeof (int) * ((size_t) mf->flat_len + 1), mf->rowptr)
       && read_raw_var (fname, ncid, "colind", NC_INT, sizeof (int) * (size_t) mf->nnz, mf->colind))
      mf->index_is_file_order = 1;
#endif
   if (!mf->index_is_file_order && (read_int_var (ncid, "rowptr", mf->rowptr) || read_int_var (ncid, "colind", mf->colind)))
      return 1;
   if ((status = nc_inq_varid (ncid, "nzval_row_wise", &varid)) != NC_NOERR)
      return nc_fail (status, "nc_inq_varid", "nzval_row_wise");
#ifndef NKP_NO_NC3_EXTENT
   if (read_raw_var (fname, ncid, "nzval_row_wise", NC_DOUBLE, sizeof (double) * (size_t) mf->nnz, mf->nzval))
      mf->nzval_is_file_order = 1;
#endif
   if (!mf->nzval_is_file_order && (status = nc_get_var_double (ncid, varid, mf->nzval)) != NC_NOERR)
      return nc_fail (status, "nc_get_var_double", "nzval_row_wise");
   if ((status = nc_close (ncid)) != NC_NOERR)
      return nc_fail (status, "nc_close", fname);
   return 0;
}

int
main (int argc, char **argv)
{
   matrix_file_t mf;
   nkp_solver *solver = NULL;
   nkp_options opt;
   nkp_stats st;
   int rc, q, t, nvars = 0, ngroups, ncid, status;
   int *ci, *cj, *ck;
   char **names = NULL, *cp;
   double **fields = NULL, *berr = NULL;
   size_t ncell;

   if (parse_cmd_line (argc, argv))
      exit (EXIT_FAILURE);
   if (dbg_lvl) {
      printf ("(%d) dbg_lvl            = %d\n", iam, dbg_lvl);
      printf ("(%d) nprow,npcol        = %ld,%ld (accepted for compatibility; one GPU is used)\n", iam, nprow, npcol);
      printf ("(%d) vars               = %s\n", iam, vars ? vars : "(none)");
      printf ("(%d) matrix_fname       = %s\n", iam, matrix_fname);
      printf ("(%d) inout_fname        = %s\n\n", iam, inout_fname);
   }

   if (read_matrix_file (matrix_fname, &mf))
      exit (EXIT_FAILURE);
   if (dbg_lvl)
      printf ("(%d) flat_len = %d, nnz = %d, coupled_tracer_cnt = %d\n", iam, mf.flat_len, mf.nnz, mf.coupled_tracer_cnt);

   /* grid coordinates of every unknown (repeated per coupled tracer) for the geometric ordering */
   ci = malloc (sizeof (int) * (size_t) mf.flat_len);
   cj = malloc (sizeof (int) * (size_t) mf.flat_len);
   ck = malloc (sizeof (int) * (size_t) mf.flat_len);
   if (!ci || !cj || !ck) {
      fprintf (stderr, "(%d) malloc failed for coordinates\n", iam);
      exit (EXIT_FAILURE);
   }
   for (t = 0; t < mf.coupled_tracer_cnt; t++)
      for (q = 0; q < mf.tracer_state_len; q++) {
         ci[t * mf.tracer_state_len + q] = mf.ind_i[q];
         cj[t * mf.tracer_state_len + q] = mf.ind_j[q];
         ck[t * mf.tracer_state_len + q] = mf.ind_k[q];
      }

   /* analysis + numeric factorisation: the nrhs = 0 call of the reference (src/solve_ABglobal.c:353) */
   nkp_default_options (&opt);
   opt.verbose = dbg_lvl;
   if (dbg_lvl)
      printf ("(%d) calling %s\n", iam, mf.index_is_file_order ? "nkp_create_be (rowptr, colind in file byte order)" : "nkp_create");
   rc = mf.index_is_file_order ? nkp_create_be (&solver, mf.flat_len, mf.nnz, mf.rowptr, mf.colind, ci, cj, ck, &opt)
      : nkp_create (&solver, mf.flat_len, mf.rowptr, mf.colind, ci, cj, ck, &opt);
   if (rc != NKP_OK) {
      fprintf (stderr, "(%d) nkp_create failed: %s\n", iam, nkp_last_error ());
      exit (EXIT_FAILURE);
   }
   free (ci);
   free (cj);
   free (ck);
   if (dbg_lvl)
      printf ("(%d) calling %s\n", iam, mf.nzval_is_file_order ? "nkp_factor_be (values in file byte order)" : "nkp_factor");
   rc = mf.nzval_is_file_order ? nkp_factor_be (solver, mf.nzval) : nkp_factor (solver, mf.nzval);
   if (dbg_lvl)
      printf ("(%d) factor info = %d\n", iam, rc);
   if (rc != NKP_OK) {
      fprintf (stderr, "(%d) nkp_factor failed: %s\n", iam, nkp_last_error ());
      exit (EXIT_FAILURE);
   }
   rc = nkp_set_tracer_maps (solver, mf.tracer_state_len, mf.coupled_tracer_cnt, mf.ind_i, mf.ind_j, mf.ind_k, mf.imt,
                             mf.jmt, mf.km);
   if (rc != NKP_OK) {
      fprintf (stderr, "(%d) nkp_set_tracer_maps failed: %s\n", iam, nkp_last_error ());
      exit (EXIT_FAILURE);
   }

   /* -v list: coupled_tracer_cnt names per solve; running out inside a group is fatal
    * (src/solve_ABglobal.c:370-388) */
   if (vars != NULL)
      for (cp = vars; *cp;) {
         nvars++;
         if ((cp = strchr (cp, ',')) == NULL)
            break;
         cp++;
      }
   if (nvars % mf.coupled_tracer_cnt != 0) {
      fprintf (stderr, "(%d) not enough vars for coupled_tracer_cnt = %d\n", iam, mf.coupled_tracer_cnt);
      exit (EXIT_FAILURE);
   }
   ngroups = nvars / mf.coupled_tracer_cnt;
   ncell = (size_t) mf.imt * mf.jmt * mf.km;
   names = malloc (sizeof (char *) * (size_t) (nvars + 1));
   fields = malloc (sizeof (double *) * (size_t) (nvars + 1));
   berr = malloc (sizeof (double) * (size_t) (ngroups + 1));
   if (!names || !fields || !berr) {
      fprintf (stderr, "(%d) malloc failed for the tracer list\n", iam);
      exit (EXIT_FAILURE);
   }
   for (q = 0, cp = vars ? strtok (vars, ",") : NULL; cp != NULL; cp = strtok (NULL, ","))
      names[q++] = cp;
   if (q != nvars) {
      fprintf (stderr, "(%d) empty name in the -v list\n", iam);
      exit (EXIT_FAILURE);
   }

   /* read every field (get_B_global, src/solve_ABglobal.c:176-183) */
   if (nvars > 0) {
      if ((status = nc_open (inout_fname, NC_WRITE, &ncid)) != NC_NOERR) {
         nc_fail (status, "nc_open", inout_fname);
         exit (EXIT_FAILURE);
      }
      for (q = 0; q < nvars; q++) {
         int varid;

         if ((fields[q] = malloc (sizeof (double) * ncell)) == NULL) {
            fprintf (stderr, "(%d) malloc failed for field %s\n", iam, names[q]);
            exit (EXIT_FAILURE);
         }
         if (dbg_lvl)
            printf ("(%d) reading %s\n", iam, names[q]);
         if ((status = nc_inq_varid (ncid, names[q], &varid)) != NC_NOERR) {
            nc_fail (status, "nc_inq_varid", names[q]);
            exit (EXIT_FAILURE);
         }
         if ((status = nc_get_var_double (ncid, varid, fields[q])) != NC_NOERR) {
            nc_fail (status, "nc_get_var_double", names[q]);
            exit (EXIT_FAILURE);
         }
      }

      /* one batched solve for all groups (the reference: one pdgssvx* call per group, :395) */
      if (dbg_lvl)
         printf ("(%d) calling nkp_solve_fields for %d right-hand side(s)\n", iam, ngroups);
      rc = nkp_solve_fields (solver, fields, nvars, berr);
      if (dbg_lvl) {
         printf ("(%d) solve info = %d\n", iam, rc);
         if (rc == NKP_OK)
            for (q = 0; q < ngroups; q++)
               printf ("(%d) berr[%d] = %e\n", iam, q, berr[q]);
      }
      if (rc != NKP_OK)
         fprintf (stderr, "(%d) nkp_solve_fields failed: %s\n", iam, nkp_last_error ());

      /* write the solutions over the inputs; land values were never touched (put_B_global, :236-253) */
      for (q = 0; rc == NKP_OK && q < nvars; q++) {
         int varid;

         if (dbg_lvl)
            printf ("(%d) writing %s\n", iam, names[q]);
         if ((status = nc_inq_varid (ncid, names[q], &varid)) != NC_NOERR
             || (status = nc_put_var_double (ncid, varid, fields[q])) != NC_NOERR) {
            nc_fail (status, "nc_put_var_double", names[q]);
            exit (EXIT_FAILURE);
         }
      }
      if ((status = nc_close (ncid)) != NC_NOERR) {
         nc_fail (status, "nc_close", inout_fname);
         exit (EXIT_FAILURE);
      }
   }

   if (dbg_lvl && nkp_get_stats (solver, &st) == NKP_OK) {
      /* the PStatPrint block of the reference (src/solve_ABglobal.c:351-360) */
      printf ("(%d) analysis %.3f s, factor %.3f s (%.2f TFLOP/s), solve %.3f s, refinement steps %d, tiny pivots %d\n",
              iam, st.t_analysis, st.t_factor, st.t_factor > 0 ? st.factor_flops / st.t_factor * 1e-12 : 0.0, st.t_solve,
              st.refine_steps, st.tiny_pivots);
      printf ("(%d) fronts %d, levels %d, largest front %d, nnz(L+U) %lld\n", iam, st.n_fronts, st.n_levels, st.max_front,
              (long long) st.nnz_lu);
   }
   for (q = 0; q < nvars; q++)
      free (fields[q]);
   free (fields);
   free (names);
   free (berr);
   nkp_destroy (solver);
   free (mf.ind_i);
   free (mf.ind_j);
   free (mf.ind_k);
   free (mf.rowptr);
   free (mf.colind);
   free (mf.nzval);
   free (vars);
   exit (EXIT_SUCCESS);
}
