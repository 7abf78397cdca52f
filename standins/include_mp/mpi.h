/*
 * Multi-process stand-in for <mpi.h>: the ranks are processes started by oracle/mprun.py that share one
 * memory segment (standins/shim_mpi_mp.c).  TEST INFRASTRUCTURE ONLY.  Differs from the serial header
 * (../include/mpi.h) in one thing: a datatype carries its KIND as well as its size, because reductions over
 * several ranks must know whether 4 bytes are an int or a float.
 *
 * Serial header's description follows.
 *
 * Purpose: lets the UNMODIFIED reference sources under /root/reference/src be
 * compiled in a container that has no MPI installation, so
 * that the reference's own arithmetic can be executed as the parity oracle
 * (oracle/_ref).  Only the ~25 symbols the reference uses are provided
 * (SURVEY.md section 8(c)).  A datatype is represented by its size in bytes.
 */
#ifndef MGP_SHIM_MPI_H
#define MGP_SHIM_MPI_H
#define MGP_SHIM_MPI_MP 1

#include <stddef.h>

typedef int  MPI_Comm;
typedef int  MPI_Datatype;   /* (kind << 16) | element size in bytes; kind 0 opaque bytes, 1 signed, 2 unsigned, 3 real */
#define MGP_DT(kind, size) ((MPI_Datatype) (((kind) << 16) | (int) (size)))
#define MGP_DT_SIZE(t) ((size_t) ((t) & 0xffff))
#define MGP_DT_KIND(t) ((int) ((t) >> 16))
typedef int  MPI_Op;
typedef long MPI_Aint;
typedef int  MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_PROC_NULL  (-1)
#define MPI_IN_PLACE   ((void *) 1)
#define MPI_SUCCESS    0

#define MPI_BYTE               MGP_DT(0, 1)
#define MPI_CHAR               MGP_DT(0, 1)
#define MPI_INT                MGP_DT(1, sizeof(int))
#define MPI_UNSIGNED           MGP_DT(2, sizeof(unsigned))
#define MPI_LONG               MGP_DT(1, sizeof(long))
#define MPI_LONG_LONG          MGP_DT(1, sizeof(long long))
#define MPI_UNSIGNED_LONG_LONG MGP_DT(2, sizeof(unsigned long long))
#define MPI_FLOAT              MGP_DT(3, sizeof(float))
#define MPI_DOUBLE             MGP_DT(3, sizeof(double))

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

#define MPI_TYPECLASS_REAL 1

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allgather(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, int rcount, MPI_Datatype rt, MPI_Comm comm);
int MPI_Gather(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, int rcount, MPI_Datatype rt, int root, MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Sendrecv(const void *sendbuf, int scount, MPI_Datatype st, int dest, int stag,
                 void *recvbuf, int rcount, MPI_Datatype rt, int source, int rtag,
                 MPI_Comm comm, MPI_Status *status);
int MPI_Type_match_size(int typeclass, int size, MPI_Datatype *t);
int MPI_Type_create_struct(int count, const int *blocklengths, const MPI_Aint *offsets,
                           const MPI_Datatype *types, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *t);
/* MatchMaker (mm_main.c:72, 124; mm_snap_io.c:335): the MPI-1 name of MPI_Type_create_struct, and a gather of unequal pieces */
int MPI_Type_struct(int count, int *blocklengths, MPI_Aint *offsets, MPI_Datatype *types, MPI_Datatype *newtype);
int MPI_Gatherv(const void *sendbuf, int scount, MPI_Datatype st, void *recvbuf, const int *rcounts, const int *displs,
                MPI_Datatype rt, int root, MPI_Comm comm);

#endif
