// oracle/shim/mpi.h -- TEST INFRASTRUCTURE. Stand-in for <mpi.h>: the "ranks" of a run are threads of one
// process (oracle/shim/ref_capi.cpp); only what the reference's hot-path sources reference is declared.
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
#include <cstddef>
#define MPI_VERSION 3
#define MPI_SUBVERSION 1
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_UNSIGNED_LONG 3
#define MPI_SUM 0
#define MPI_BYTE 4
// MPI-IO names used by the reference's dump path (frontend/io/atom_dump.cpp); the file is a memory sink here
typedef void *MPI_File;
typedef int MPI_Info;
#define MPI_INFO_NULL 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 4
#define MPI_SUCCESS 0
#define MPI_COMM_TYPE_SHARED 1
extern "C" {
// the accelerator hook shim (arch_cuda/cuda_hooks.cpp) asks for its node-local rank; the stand-in says "MPI not initialised"
// (ranks are threads of one process here), which sends the shim to the launcher's environment variables
int MPI_Initialized(int *flag);
int MPI_Comm_split_type(MPI_Comm comm, int split_type, int key, MPI_Info info, MPI_Comm *newcomm);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_free(MPI_Comm *comm);
double MPI_Wtime();
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Get_address(const void *location, MPI_Aint *address);
int MPI_Type_create_struct(int count, const int *blocklengths, const MPI_Aint *displacements, const MPI_Datatype *types,
                           MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Type_free(MPI_Datatype *type);
int MPI_File_open(MPI_Comm comm, const char *filename, int amode, MPI_Info info, MPI_File *fh);
int MPI_File_close(MPI_File *fh);
}
#endif
