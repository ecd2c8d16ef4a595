/* Single-rank stand-in for <mpi.h>, used ONLY to compile the unmodified
 * reference sources into oracle/_ref (test infrastructure, see oracle/README.md).
 * Semantics: one rank; collectives copy send->recv; point-to-point are no-ops
 * (the reference's move_data returns before any MPI call when numprocs == 1,
 * /root/reference/src/multiproc/move_data.cc:53-54). */
#ifndef PSPH_ORACLE_MPI_STUB_H
#define PSPH_ORACLE_MPI_STUB_H
#include <string.h>
#include <stdlib.h>
#include <time.h>

typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Comm;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_INFO_NULL 0
#define MPI_SUCCESS 0
enum { MPI_INT = 4, MPI_UNSIGNED = 5, MPI_FLOAT = 6, MPI_DOUBLE = 8, MPI_CHAR = 1, MPI_BYTE = 2 };
enum { MPI_SUM = 1, MPI_MIN = 2, MPI_MAX = 3 };

static inline size_t psph_mpi_size_(MPI_Datatype t)
{
  switch (t) { case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
               case MPI_DOUBLE: return 8; case MPI_CHAR: case MPI_BYTE: return 1; default: return 0; }
}
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Abort(MPI_Comm, int e) { exit(e ? e : 1); return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline double MPI_Wtime(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{ memcpy(r, s, n * psph_mpi_size_(t)); return 0; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{ memcpy(r, s, n * psph_mpi_size_(t)); return 0; }
static inline int MPI_Allgather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm)
{ memcpy(r, s, n * psph_mpi_size_(t)); return 0; }
static inline int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm)
{ memcpy(r, s, n * psph_mpi_size_(t)); return 0; }
static inline int MPI_Alltoall(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, MPI_Comm)
{ memcpy(r, s, n * psph_mpi_size_(t)); return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *q) { *q = 0; return 0; }
static inline int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *q) { *q = 0; return 0; }
static inline int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
static inline int MPI_Address(void *p, MPI_Aint *a) { *a = (MPI_Aint) p; return 0; }
static inline int MPI_Type_struct(int, int *, MPI_Aint *, MPI_Datatype *, MPI_Datatype *n) { *n = 100; return 0; }
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
#endif
