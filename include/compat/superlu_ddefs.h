/*
 * superlu_ddefs.h -- the SuperLU_DIST 5.1.3 C API subset that the reference's
 * programs are written against (pin: /root/reference/src/Makefile:3), re-declared
 * so that src/gen_A.c, src/matrix.c, src/solve_ABglobal.c and src/solve_ABdist.c
 * compile UNCHANGED and link against libnkprecond (the B200 solver) instead of
 * SuperLU_DIST + ParMETIS + MPI + BLAS.
 *
 * gen_A / matrix.c need only `int_t` (README:13-16).  The solver drivers use the
 * calls listed in SURVEY.md section 8(b); each declaration below cites the
 * reference call site it serves.  Names, argument order and argument meaning
 * follow the published SuperLU_DIST 5.x interface; the structs keep the public
 * fields the drivers touch and carry an opaque handle to the nkp_solver that
 * does the work (include/nkprecond.h).
 *
 * Implementation: nk_ocn_tracer_jacobian_precond_b200/csrc/compat_superlu.c
 */
#ifndef NKP_COMPAT_SUPERLU_DDEFS_H
#define NKP_COMPAT_SUPERLU_DDEFS_H

#include <stdio.h>
#include <stdlib.h>
#include "mpi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* README:13-16 -- the generator's only dependency on the solver library */
typedef int int_t;
#define mpi_int_t MPI_INT
#define IFMT "%8d"

typedef enum { NO, YES } yes_no_t;
typedef enum { DOFACT, SamePattern, SamePattern_SameRowPerm, FACTORED } fact_t;
typedef enum { NOROWPERM, LargeDiag, MY_PERMR } rowperm_t;
typedef enum { NATURAL, MMD_ATA, MMD_AT_PLUS_A, COLAMD, METIS_AT_PLUS_A, PARMETIS, ZOLTAN, MY_PERMC } colperm_t;
typedef enum { NOTRANS, TRANS, CONJ } trans_t;
typedef enum { NOEQUIL, ROW, COL, BOTH } DiagScale_t;
typedef enum { NOREFINE, SLU_SINGLE = 1, SLU_DOUBLE, SLU_EXTRA } IterRefine_t;

typedef enum { SLU_NC, SLU_NCP, SLU_NR, SLU_SC, SLU_SCP, SLU_SR, SLU_DN, SLU_NR_loc } Stype_t;
typedef enum { SLU_S, SLU_D, SLU_C, SLU_Z } Dtype_t;
typedef enum { SLU_GE, SLU_TRLU, SLU_TRUU, SLU_TRL, SLU_TRU, SLU_SYL, SLU_SYU, SLU_HEL, SLU_HEU } Mtype_t;

typedef struct {
   Stype_t Stype;
   Dtype_t Dtype;
   Mtype_t Mtype;
   int_t nrow;
   int_t ncol;
   void *Store;
} SuperMatrix;

/* SLU_NC store (src/solve_ABglobal.c:327) */
typedef struct {
   int_t nnz;
   void *nzval;
   int_t *rowind;
   int_t *colptr;
} NCformat;

/* SLU_NR_loc store (src/solve_ABdist.c:482-483) */
typedef struct {
   int_t nnz_loc;
   int_t m_loc;
   int_t fst_row;
   void *nzval;
   int_t *rowptr;
   int_t *colind;
} NRformat_loc;

/* fields set/read at src/solve_ABglobal.c:332-334,358,363; src/solve_ABdist.c:488-495,523,539,595 */
typedef struct {
   fact_t Fact;
   yes_no_t Equil;
   colperm_t ColPerm;
   trans_t Trans;
   IterRefine_t IterRefine;
   double DiagPivotThresh;
   yes_no_t SymmetricMode;
   yes_no_t PivotGrowth;
   yes_no_t ConditionNumber;
   rowperm_t RowPerm;
   yes_no_t ParSymbFact;
   yes_no_t ReplaceTinyPivot;
   yes_no_t SolveInitialized;
   yes_no_t RefineInitialized;
   yes_no_t PrintStat;
   int nnzL, nnzU;
   int num_lookaheads;
   yes_no_t lookahead_etree;
   yes_no_t SymPattern;
} superlu_dist_options_t;

typedef struct {
   MPI_Comm comm;
   int Np;
   int Iam;
} superlu_scope_t;

/* field `comm` used at src/solve_ABglobal.c:132-139 */
typedef struct {
   MPI_Comm comm;
   superlu_scope_t rscp;
   superlu_scope_t cscp;
   int iam;
   int_t nprow;
   int_t npcol;
} gridinfo_t;

/* perm_c referenced (dead code) at src/solve_ABdist.c:527-533 */
typedef struct {
   DiagScale_t DiagScale;
   double *R;
   double *C;
   int_t *perm_r;
   int_t *perm_c;
} ScalePermstruct_t;

/* the factors live on the GPUs; `nkp` is the opaque nkp_solver handle */
typedef struct {
   int_t *etree;
   void *Glu_persist;
   void *Llu;
   void *nkp;
} LUstruct_t;

typedef struct {
   void *reserved;
} SOLVEstruct_t;

/* phase timers in the spirit of PStatPrint (src/solve_ABglobal.c:351-360) */
typedef struct {
   double t_analysis;
   double t_scatter;
   double t_factor;
   double t_solve;
   double t_refine;
   double flops_factor;
   int refine_steps;
   long long nnz_lu;
   int valid;
} SuperLUStat_t;

#define SUPERLU_MALLOC(size) superlu_malloc_dist(size)
#define SUPERLU_FREE(addr) superlu_free_dist(addr)
#define ABORT(err_msg) \
 { char msg[256]; \
   sprintf(msg, "%s at line %d in file %s\n", err_msg, __LINE__, __FILE__); \
   superlu_abort_and_exit_dist(msg); }

void *superlu_malloc_dist (size_t size);
void superlu_free_dist (void *addr);
void superlu_abort_and_exit_dist (char *msg);

/* src/solve_ABglobal.c:307,425 ; src/solve_ABdist.c:461,604 */
void superlu_gridinit (MPI_Comm Bcomm, int_t nprow, int_t npcol, gridinfo_t * grid);
void superlu_gridexit (gridinfo_t * grid);

/* src/solve_ABglobal.c:126 */
void dCompRow_to_CompCol_dist (int_t m, int_t n, int_t nnz, double *a, int_t * colind, int_t * rowptr,
                               double **at, int_t ** rowind, int_t ** colptr);
/* src/solve_ABglobal.c:136 */
void dallocateA_dist (int_t n, int_t nnz, double **a, int_t ** asub, int_t ** xa);
/* src/solve_ABglobal.c:327 */
void dCreate_CompCol_Matrix_dist (SuperMatrix * A, int_t m, int_t n, int_t nnz, double *nzval, int_t * rowind,
                                  int_t * colptr, Stype_t stype, Dtype_t dtype, Mtype_t mtype);
/* src/solve_ABdist.c:482 */
void dCreate_CompRowLoc_Matrix_dist (SuperMatrix * A, int_t m, int_t n, int_t nnz_loc, int_t m_loc, int_t fst_row,
                                     double *nzval, int_t * colind, int_t * rowptr, Stype_t stype, Dtype_t dtype,
                                     Mtype_t mtype);
/* src/solve_ABglobal.c:412 ; src/solve_ABdist.c:591 */
void Destroy_CompCol_Matrix_dist (SuperMatrix * A);
void Destroy_CompRowLoc_Matrix_dist (SuperMatrix * A);

/* src/solve_ABglobal.c:332,336 */
void set_default_options_dist (superlu_dist_options_t * options);
void print_options_dist (superlu_dist_options_t * options);

/* src/solve_ABglobal.c:339-340,413-415 */
void ScalePermstructInit (const int_t m, const int_t n, ScalePermstruct_t * ScalePermstruct);
void ScalePermstructFree (ScalePermstruct_t * ScalePermstruct);
void LUstructInit (const int_t n, LUstruct_t * LUstruct);
void LUstructFree (LUstruct_t * LUstruct);
void Destroy_LU (int_t n, gridinfo_t * grid, LUstruct_t * LUstruct);

/* src/solve_ABglobal.c:344,346 ; src/solve_ABdist.c:146,183,185 */
double *doubleMalloc_dist (int_t n);
int_t *intMalloc_dist (int_t n);

/* src/solve_ABglobal.c:351,359,361 */
void PStatInit (SuperLUStat_t * stat);
void PStatPrint (superlu_dist_options_t * options, SuperLUStat_t * stat, gridinfo_t * grid);
void PStatFree (SuperLUStat_t * stat);

/* src/solve_ABglobal.c:353,395 -- A and B replicated */
void pdgssvx_ABglobal (superlu_dist_options_t * options, SuperMatrix * A, ScalePermstruct_t * ScalePermstruct,
                       double B[], int ldb, int nrhs, gridinfo_t * grid, LUstruct_t * LUstruct, double *berr,
                       SuperLUStat_t * stat, int *info);
/* src/solve_ABdist.c:518,571 -- A and B block-row distributed */
void pdgssvx (superlu_dist_options_t * options, SuperMatrix * A, ScalePermstruct_t * ScalePermstruct,
              double B[], int ldb, int nrhs, gridinfo_t * grid, LUstruct_t * LUstruct,
              SOLVEstruct_t * SOLVEstruct, double *berr, SuperLUStat_t * stat, int *info);
/* src/solve_ABdist.c:596 */
void dSolveFinalize (superlu_dist_options_t * options, SOLVEstruct_t * SOLVEstruct);

#ifdef __cplusplus
}
#endif
#endif
