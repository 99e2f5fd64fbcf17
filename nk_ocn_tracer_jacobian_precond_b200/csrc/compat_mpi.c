/*
 * compat_mpi.c -- single-process stand-in for the MPI subset of the reference's solver
 * drivers (include/compat/mpi.h).  One real rank (0); point-to-point messages addressed to
 * the virtual ranks >= 1 are retained in a mailbox so that the solver shim can reassemble
 * the block-row distributed operand / right-hand side that src/solve_ABdist.c:116-244,
 * 249-330 scatter from rank 0, and can hand the solution slabs back to
 * put_B_dist's MPI_Recv (src/solve_ABdist.c:377).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpi.h"

typedef struct msg {
   int peer, tag;
   size_t nbytes;
   void *data;
   struct msg *next;
} msg;

static msg *head = NULL, *tail = NULL;

static size_t
type_bytes (MPI_Datatype t)
{
   switch (t) {
   case MPI_INT:
      return sizeof (int);
   case MPI_DOUBLE:
      return sizeof (double);
   case MPI_LONG_LONG_INT:
      return sizeof (long long);
   default:
      return 1;
   }
}

int
nkp_mpi_mailbox_post (int peer, int tag, const void *buf, size_t nbytes)
{
   msg *m = malloc (sizeof (msg));
   if (m == NULL)
      return 1;
   m->peer = peer;
   m->tag = tag;
   m->nbytes = nbytes;
   m->data = malloc (nbytes ? nbytes : 1);
   if (m->data == NULL) {
      free (m);
      return 1;
   }
   memcpy (m->data, buf, nbytes);
   m->next = NULL;
   if (tail)
      tail->next = m;
   else
      head = m;
   tail = m;
   return 0;
}

void *
nkp_mpi_mailbox_take (int peer, int tag, size_t *nbytes)
{
   msg *m = head, *prev = NULL;
   while (m) {
      if (m->peer == peer && m->tag == tag) {
         void *d = m->data;
         if (nbytes)
            *nbytes = m->nbytes;
         if (prev)
            prev->next = m->next;
         else
            head = m->next;
         if (tail == m)
            tail = prev;
         free (m);
         return d;
      }
      prev = m;
      m = m->next;
   }
   return NULL;
}

int
MPI_Init (int *argc, char ***argv)
{
   (void) argc;
   (void) argv;
   return MPI_SUCCESS;
}

int
MPI_Finalize (void)
{
   size_t nb;
   void *d;
   while (head) {
      d = nkp_mpi_mailbox_take (head->peer, head->tag, &nb);
      free (d);
   }
   return MPI_SUCCESS;
}

int
MPI_Comm_rank (MPI_Comm comm, int *rank)
{
   (void) comm;
   *rank = 0;
   return MPI_SUCCESS;
}

int
MPI_Comm_size (MPI_Comm comm, int *size)
{
   (void) comm;
   *size = 1;
   return MPI_SUCCESS;
}

int
MPI_Bcast (void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
   (void) buf;
   (void) count;
   (void) type;
   (void) root;
   (void) comm;
   return MPI_SUCCESS;
}

int
MPI_Barrier (MPI_Comm comm)
{
   (void) comm;
   return MPI_SUCCESS;
}

int
MPI_Send (const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm)
{
   (void) comm;
   if (count < 0)
      return 1;
   return nkp_mpi_mailbox_post (dest, tag, buf, (size_t) count * type_bytes (type));
}

int
MPI_Recv (void *buf, int count, MPI_Datatype type, int src, int tag, MPI_Comm comm, MPI_Status * status)
{
   size_t nb = 0, want = (size_t) count * type_bytes (type);
   void *d = nkp_mpi_mailbox_take (src, tag, &nb);
   (void) comm;
   if (d == NULL) {
      fprintf (stderr, "(0) MPI_Recv: no message from virtual rank %d with tag %d\n", src, tag);
      abort ();
   }
   memcpy (buf, d, nb < want ? nb : want);
   free (d);
   if (status) {
      status->MPI_SOURCE = src;
      status->MPI_TAG = tag;
      status->MPI_ERROR = MPI_SUCCESS;
   }
   return MPI_SUCCESS;
}

int
MPI_Abort (MPI_Comm comm, int code)
{
   (void) comm;
   exit (code ? code : 1);
}
