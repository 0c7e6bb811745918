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
extern "C" {
double MPI_Wtime();
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Get_address(const void *location, MPI_Aint *address);
int MPI_Type_create_struct(int count, const int *blocklengths, const MPI_Aint *displacements, const MPI_Datatype *types,
                           MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);
int MPI_Type_free(MPI_Datatype *type);
}
#endif
