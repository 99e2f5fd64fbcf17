/*
 * mpi.h -- single-process stand-in for the MPI subset the reference's solver
 * drivers use (src/solve_ABglobal.c:132-139,194,199,292-298,427;
 * src/solve_ABdist.c:117,163,170,211-212,223-224,235,308,320,377,406).
 *
 * The B200 build runs ONE host process that drives all GPUs, so the world has
 * exactly one real rank (rank 0).  The reference's drivers are rank-0-centric:
 * rank 0 reads files and MPI_Send()s row slabs to ranks 1..P-1.  To let those
 * mains compile and run UNCHANGED with "-n nprow,npcol", sends to the (virtual)
 * ranks >= 1 are retained in a mailbox keyed by (peer, tag); the solver shim
 * (pdgssvx in superlu_ddefs.h) reassembles the full operand / right-hand side
 * from the mailbox and deposits each virtual rank's slab of the solution where
 * the driver's MPI_Recv(src, tag) will find it.
 * Implementation: nk_ocn_tracer_jacobian_precond_b200/csrc/compat_mpi.c
 */
#ifndef NKP_COMPAT_MPI_H
#define NKP_COMPAT_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct {
   int MPI_SOURCE;
   int MPI_TAG;
   int MPI_ERROR;
} MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_LONG_LONG_INT 3
#define MPI_CHAR 4

int MPI_Init (int *argc, char ***argv);
int MPI_Finalize (void);
int MPI_Comm_rank (MPI_Comm comm, int *rank);
int MPI_Comm_size (MPI_Comm comm, int *size);
int MPI_Bcast (void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Send (const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm);
int MPI_Recv (void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status * status);
int MPI_Barrier (MPI_Comm comm);
int MPI_Abort (MPI_Comm comm, int code);

/* mailbox access for the solver shim (not part of MPI) */
/* take (and remove) the oldest retained message sent to virtual rank `peer` with `tag`;
 * returns a malloc'ed buffer the caller frees, or NULL; *nbytes receives its size */
void *nkp_mpi_mailbox_take (int peer, int tag, size_t *nbytes);
/* deposit a message as if virtual rank `peer` had sent it to rank 0 with `tag` */
int nkp_mpi_mailbox_post (int peer, int tag, const void *buf, size_t nbytes);

#ifdef __cplusplus
}
#endif
#endif
