// oracle/shim/io/local_storage.h -- TEST INFRASTRUCTURE. kiwi::LocalStorage stand-in for the reference's dump
// path (frontend/io/atom_dump.cpp, frontend/io/buffered_io.cpp): instead of an MPI-IO shared file, the byte stream
// the reference writes is appended to a per-thread memory sink that oracle/shim/ref_dump.cpp hands back.
#ifndef ORACLE_SHIM_KIWI_LOCAL_STORAGE_H
#define ORACLE_SHIM_KIWI_LOCAL_STORAGE_H
#include <cstddef>
#include <vector>
#include <mpi.h>
#include <utils/mpi_utils.h>
namespace shim { extern thread_local std::vector<unsigned char> *tl_dump_sink; }
namespace kiwi {
    typedef unsigned char byte;
    struct SinkWriter {
        void write(const void *data, size_t bytes) {
            const unsigned char *p = static_cast<const unsigned char *>(data);
            if (shim::tl_dump_sink) shim::tl_dump_sink->insert(shim::tl_dump_sink->end(), p, p + bytes);
        }
    };
    class LocalStorage {
    public:
        LocalStorage(MPI_File, size_t, size_t, size_t) {}
        void make(MPI_Datatype, mpi_process) {}
        void writeHeader(byte *, size_t, mpi_process) {}
        SinkWriter writer;
    };
}
#endif
