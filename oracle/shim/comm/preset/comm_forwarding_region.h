// oracle/shim/comm/preset/comm_forwarding_region.h -- TEST INFRASTRUCTURE. libcomm comm::fwCommLocalRegion
// restated from its two call sites (reference src/atom/atom_list.cpp:33-40, src/pack/inter_border_packer.cpp:24)
// and the receive slabs that must mirror it (src/pack/lat_particle_packer.cpp:65-76,97-108,128-139): the region
// of LOCAL sites, in ghost-extended doubled-x indices, forwarded to neighbour [dim][dir]. The message of a later
// dimension spans the full ghost-extended range of the earlier ones (that is how edges and corners propagate).
#ifndef ORACLE_SHIM_COMM_FW_REGION_H
#define ORACLE_SHIM_COMM_FW_REGION_H
#include "../domain/bcc_domain.h"
#include "../domain/region.hpp"

namespace comm {
    inline Region<_type_lattice_size> fwCommLocalRegion(const BccDomain *d, const int dim, const int dir) {
        const _type_lattice_size *g = d->dbx_lattice_size_ghost, *b = d->dbx_sub_box_lattice_size,
                                 *e = d->dbx_ghost_extended_lattice_size;
        Region<_type_lattice_size> r;
        for (int k = 0; k < 3; k++) {
            if (k == dim) {
                if (dir == DIR_LOWER) { r.low[k] = g[k]; r.high[k] = 2 * g[k]; }
                else { r.low[k] = b[k]; r.high[k] = b[k] + g[k]; }
            } else if (k < dim) { r.low[k] = 0; r.high[k] = e[k]; }
            else { r.low[k] = g[k]; r.high[k] = g[k] + b[k]; }
        }
        return r;
    }
}
#endif
