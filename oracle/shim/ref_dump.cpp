// oracle/shim/ref_dump.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the reference's OWN dump code -- AtomDump::dump (frontend/io/atom_dump.cpp:39-75) and
// BufferedFileWriter::write/flush (frontend/io/buffered_io.cpp:18-45), compiled in place -- over one sub-box and
// returns the byte stream it hands to kiwi::LocalStorage (72-byte atom_dump::AtomInfoDump records,
// frontend/io/atom_info_dump.h:14-22). A separate translation unit because ref_capi.cpp declares its own
// `AtomDump` friend stand-in to reach AtomList's private members.
#include <cstring>
#include <string>
#include <vector>

#include <mpi.h>
#include <comm/domain/bcc_domain.h>

#include "io/atom_dump.h"

namespace shim { thread_local std::vector<unsigned char> *tl_dump_sink = nullptr; }

extern "C" {
int MPI_File_open(MPI_Comm, const char *, int, MPI_Info, MPI_File *fh) {
    static int token;
    *fh = &token;
    return 0;
}
int MPI_File_close(MPI_File *fh) {
    *fh = NULL;
    return 0;
}
}

// region = OutputBaseInterface's constructor (frontend/io/output_base_interface.h:26-31): the owned sub-box in
// ghost-inclusive doubled-x coordinates
size_t ref_dump_rank(AtomList *atom_list, InterAtomList *inter_list, const comm::BccDomain *d, size_t time_step,
                     unsigned char *out, size_t cap) {
    _type_lattice_coord begin[DIMENSION], end[DIMENSION];
    begin[0] = d->dbx_sub_box_lattice_region.x_low - d->dbx_ghost_ext_lattice_region.x_low;
    begin[1] = d->dbx_sub_box_lattice_region.y_low - d->dbx_ghost_ext_lattice_region.y_low;
    begin[2] = d->dbx_sub_box_lattice_region.z_low - d->dbx_ghost_ext_lattice_region.z_low;
    for (int k = 0; k < DIMENSION; k++) end[k] = begin[k] + d->dbx_sub_box_lattice_size[k];
    const _type_lattice_size atoms_size =
        d->dbx_sub_box_lattice_size[0] * d->dbx_sub_box_lattice_size[1] * d->dbx_sub_box_lattice_size[2];
    std::vector<unsigned char> sink;
    shim::tl_dump_sink = &sink;
    {
        AtomDump dump("memory", atoms_size, begin, end);
        dump.dump(atom_list, inter_list, time_step);
    }
    shim::tl_dump_sink = nullptr;
    if (out && sink.size() <= cap) memcpy(out, sink.data(), sink.size());
    return sink.size();
}
