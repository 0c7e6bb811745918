// oracle/shim/comm/types_define.h -- TEST INFRASTRUCTURE. libcomm v0.3.3 (absent) restated: basic types.
#ifndef ORACLE_SHIM_COMM_TYPES_H
#define ORACLE_SHIM_COMM_TYPES_H
#include <mpi.h>
#ifndef DIMENSION_SIZE
#define DIMENSION_SIZE 3
#endif
namespace comm {
    typedef int _type_lattice_size;
    typedef int _type_lattice_coord;
    typedef int _MPI_Rank;
    const int DIR_LOWER = 0;
    const int DIR_HIGHER = 1;
    struct mpi_process {
        int own_rank, all_ranks;
        MPI_Comm comm;
    };
}
#endif
