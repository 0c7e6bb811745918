// oracle/shim/comm/comm.hpp -- TEST INFRASTRUCTURE. libcomm v0.3.3 comm::neiSendReceive<T, F> restated
// (source absent; semantics recovered from the packers): per dimension stage (x,y,z; reversed
// when F), pack LOWER and HIGHER, send to neighbours[dim][dir], receive what neighbour [dim][(dir+1)%2] sent
// with the same dir, and unpack with onReceive(.., dim, dir) -- "mirror with send"
// (reference src/pack/lat_particle_packer.cpp:66).
#ifndef ORACLE_SHIM_COMM_HPP
#define ORACLE_SHIM_COMM_HPP
#include "packer.h"
#include "thread_world.h"
#include "types_define.h"

namespace comm {
    template<typename T, bool F = false>
    void neiSendReceive(Packer<T> *packer, const mpi_process, const MPI_Datatype, const _MPI_Rank (*neighbours)[2]) {
        shim::ThreadWorld *w = shim::tl_world;
        const int me = shim::tl_rank;
        for (int s = 0; s < DIMENSION_SIZE; s++) {
            const int d = F ? DIMENSION_SIZE - 1 - s : s;
            T *send[2];
            for (int dir = DIR_LOWER; dir <= DIR_HIGHER; dir++) {
                const unsigned long n = packer->sendLength(d, dir);
                send[dir] = new T[n];
                packer->onSend(send[dir], n, d, dir);
                w->box[2 * me + dir].ptr = send[dir];
                w->box[2 * me + dir].count = n;
            }
            w->barrier.wait();
            for (int dir = DIR_LOWER; dir <= DIR_HIGHER; dir++) {
                const int src = neighbours[d][(dir + 1) % 2];
                const shim::Slot &m = w->box[2 * src + dir];
                packer->onReceive(static_cast<T *>(m.ptr), m.count, d, dir);
            }
            w->barrier.wait();
            delete[] send[0];
            delete[] send[1];
        }
    }
}
#endif
