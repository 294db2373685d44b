/* mpi.h (petsc_lite) -- single-process stand-in for the MPI calls the reference makes.  TEST INFRASTRUCTURE:
 * part of oracle/petsc_lite, used only to run the reference's own sources serially on the CPU (oracle/Makefile ->
 * oracle/_ref/topopt_ref); never linked into the product. */
#ifndef PETSC_LITE_MPI_H
#define PETSC_LITE_MPI_H
#include <stddef.h>

typedef int MPI_Comm;
typedef int MPI_Op;
typedef int MPI_Info;
typedef long long MPI_Offset;
typedef struct { int dummy; } MPI_Status;
struct plite_datatype;
typedef struct plite_datatype* MPI_Datatype;
struct plite_file;
typedef struct plite_file* MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_SUCCESS 0
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_INFO_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_RDONLY 4

extern MPI_Datatype MPI_DOUBLE, MPI_INT, MPI_UNSIGNED_LONG, MPI_FLOAT, MPI_CHAR;

int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void* sendbuf, int scount, MPI_Datatype stype, void* recvbuf, int rcount, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);
int MPI_Type_size(MPI_Datatype type, int* size);
int MPI_Type_vector(int count, int blocklength, int stride, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* type);
int MPI_Type_free(MPI_Datatype* type);
int MPI_File_open(MPI_Comm comm, const char* filename, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_close(MPI_File* fh);
int MPI_File_delete(const char* filename, MPI_Info info);
int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype etype, MPI_Datatype filetype, const char* datarep, MPI_Info info);
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status);
int MPI_File_write_all(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status);
#endif
