// oracle/shim/utils/mpi_utils.h -- TEST INFRASTRUCTURE. kiwi::mpi_process stand-in.
#ifndef ORACLE_SHIM_KIWI_MPI_UTILS_H
#define ORACLE_SHIM_KIWI_MPI_UTILS_H
#include <mpi.h>
#include "data_def.h"
namespace kiwi {
    struct mpi_process {
        int own_rank, all_ranks;
        MPI_Comm comm;
    };
}
#endif
