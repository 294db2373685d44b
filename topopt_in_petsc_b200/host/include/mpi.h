/* mpi.h -- minimal stand-in for the MPI calls the untouched main.cc / MMA.cc make.
 * One process per GPU; ranks are started by host/tobrun (env TOB_RANK / TOB_NRANKS / TOB_JOB) and
 * exchange small scalars through a POSIX shared-memory segment.  Everything heavy (halos, dots of
 * the solver) goes over NCCL inside libtopopt_b200.so, not through this shim. */
#ifndef TOB_MPI_SHIM_H
#define TOB_MPI_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 1
#define MPI_INT 2
#define MPI_UNSIGNED_LONG 3
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_SUCCESS 0
int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int scount, MPI_Datatype stype, void* recvbuf, int rcount, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
#ifdef __cplusplus
}
#endif
#endif
