/*
 * Serial MPI stand-in (NTask = 1).  TEST INFRASTRUCTURE ONLY -- see include/mpi.h.
 * With one task LeftTask == RightTask == 0 (reference 2LPT.c:66-81), so the
 * halo MPI_Sendrecv calls are self-sends and must really copy the bytes.
 */
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int code) {
  (void) comm;
  fprintf(stderr, "[mpi shim] MPI_Abort(%d)\n", code);
  fflush(NULL);
  exit(code ? code : 1);
}
int MPI_Barrier(MPI_Comm comm) { (void) comm; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void) comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void) comm; *size = 1; return MPI_SUCCESS; }

/* Test tap: the reference only writes its P(k) bins to text files with %10.5f (compute_pofk.c:251-258).
 * To pin them as in-memory doubles the stand-in remembers the payload of the most recent
 * MPI_Allreduce calls on arrays of >= 4 doubles (the three per-bin sums of compute_pofk.c:225-227,
 * the five of 693-697).  Nothing in the reference is changed by this. */
#define TAP_SLOTS 16
#define TAP_MAX 8192
static double tap_data[TAP_SLOTS][TAP_MAX];
static int tap_count[TAP_SLOTS];
static long tap_total = 0;
void mgp_shim_tap_reset(void) { tap_total = 0; }
long mgp_shim_tap_total(void) { return tap_total; }
/* i = 0 is the oldest retained call; returns the element count (0 if out of range) */
int mgp_shim_tap_get(long i, double *out) {
  long first = tap_total > TAP_SLOTS ? tap_total - TAP_SLOTS : 0;
  if (i < first || i >= tap_total) return 0;
  int s = (int) (i % TAP_SLOTS);
  memcpy(out, tap_data[s], (size_t) tap_count[s] * sizeof(double));
  return tap_count[s];
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm) {
  (void) op; (void) comm;
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memmove(recvbuf, sendbuf, (size_t) count * (size_t) t);
  if (t == MPI_DOUBLE && count >= 4 && count <= TAP_MAX) {
    int s = (int) (tap_total % TAP_SLOTS);
    memcpy(tap_data[s], recvbuf, (size_t) count * sizeof(double));
    tap_count[s] = count;
    tap_total++;
  }
  return MPI_SUCCESS;
}
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm) {
  (void) root;
  return MPI_Allreduce(sendbuf, recvbuf, count, t, op, comm);
}
int MPI_Allgather(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, int rcount, MPI_Datatype rt, MPI_Comm comm) {
  (void) rcount; (void) rt; (void) comm;
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memmove(recvbuf, sendbuf, (size_t) scount * (size_t) st);
  return MPI_SUCCESS;
}
int MPI_Gather(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, int rcount, MPI_Datatype rt, int root, MPI_Comm comm) {
  (void) root;
  return MPI_Allgather(sendbuf, scount, st, recvbuf, rcount, rt, comm);
}
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm) {
  (void) buf; (void) count; (void) t; (void) root; (void) comm;
  return MPI_SUCCESS;
}
int MPI_Sendrecv(const void *sendbuf, int scount, MPI_Datatype st, int dest, int stag,
                 void *recvbuf, int rcount, MPI_Datatype rt, int source, int rtag,
                 MPI_Comm comm, MPI_Status *status) {
  (void) stag; (void) rtag; (void) comm;
  if (dest == MPI_PROC_NULL || source == MPI_PROC_NULL) return MPI_SUCCESS;
  size_t sb = (size_t) scount * (size_t) st, rb = (size_t) rcount * (size_t) rt;
  size_t n = sb < rb ? sb : rb;
  if (n && sendbuf != recvbuf) memmove(recvbuf, sendbuf, n);
  if (status) { status->MPI_SOURCE = 0; status->MPI_TAG = rtag; status->MPI_ERROR = MPI_SUCCESS; }
  return MPI_SUCCESS;
}
int MPI_Type_match_size(int typeclass, int size, MPI_Datatype *t) {
  (void) typeclass; *t = (MPI_Datatype) size; return MPI_SUCCESS;
}
int MPI_Type_create_struct(int count, const int *blocklengths, const MPI_Aint *offsets,
                           const MPI_Datatype *types, MPI_Datatype *newtype) {
  /* extent = furthest member end, rounded up to the widest member alignment */
  long end = 0, align = 1;
  for (int i = 0; i < count; i++) {
    long e = (long) offsets[i] + (long) blocklengths[i] * (long) types[i];
    if (e > end) end = e;
    if ((long) types[i] > align) align = (long) types[i];
  }
  end = (end + align - 1) / align * align;
  *newtype = (MPI_Datatype) end;
  return MPI_SUCCESS;
}
int MPI_Type_commit(MPI_Datatype *t) { (void) t; return MPI_SUCCESS; }

int MPI_Type_struct(int count, int *blocklengths, MPI_Aint *offsets, MPI_Datatype *types, MPI_Datatype *newtype) {
  return MPI_Type_create_struct(count, blocklengths, offsets, types, newtype);
}
int MPI_Gatherv(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, const int *rcounts, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm) {
  (void) rcounts; (void) rt; (void) root; (void) comm;
  memcpy((char *) recvbuf + (size_t) displs[0] * (size_t) st, sendbuf, (size_t) scount * (size_t) st);
  return MPI_SUCCESS;
}
