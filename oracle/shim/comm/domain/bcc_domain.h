// oracle/shim/comm/domain/bcc_domain.h -- TEST INFRASTRUCTURE. libcomm comm::BccDomain restated: Domain plus the
// doubled-x ("dbx") twins of the lattice sizes/regions (x index counts corner AND body-centre sites).
#ifndef ORACLE_SHIM_COMM_BCC_DOMAIN_H
#define ORACLE_SHIM_COMM_BCC_DOMAIN_H
#include "domain.h"

namespace comm {
    class BccDomain : public Domain {
    public:
        _type_lattice_size dbx_sub_box_lattice_size[DIMENSION_SIZE];
        _type_lattice_size dbx_lattice_size_ghost[DIMENSION_SIZE];
        _type_lattice_size dbx_ghost_extended_lattice_size[DIMENSION_SIZE];
        Region<_type_lattice_coord> dbx_sub_box_lattice_region, dbx_ghost_ext_lattice_region;

        class Builder : public Domain::Builder {
        public:
            Builder &setComm(mpi_process p, MPI_Comm *c) { Domain::Builder::setComm(p, c); return *this; }
            Builder &setPhaseSpace(const int64_t ps[DIMENSION_SIZE]) { Domain::Builder::setPhaseSpace(ps); return *this; }
            Builder &setLatticeConst(const double a) { Domain::Builder::setLatticeConst(a); return *this; }
            Builder &setCutoffRadius(const double crf) { Domain::Builder::setCutoffRadius(crf); return *this; }
            Builder &setGhostSize(const _type_lattice_size g) { Domain::Builder::setGhostSize(g); return *this; }
            BccDomain *localBuild(const int grid_size[DIMENSION_SIZE], const int grid_coord[DIMENSION_SIZE]) {
                BccDomain *d = new BccDomain();
                fill(d, grid_size, grid_coord);
                for (int k = 0; k < 3; k++) {
                    const int m = k == 0 ? 2 : 1;
                    d->dbx_sub_box_lattice_size[k] = m * d->sub_box_lattice_size[k];
                    d->dbx_lattice_size_ghost[k] = m * d->lattice_size_ghost[k];
                    d->dbx_ghost_extended_lattice_size[k] = m * d->ghost_extended_lattice_size[k];
                    d->dbx_sub_box_lattice_region.low[k] = m * d->sub_box_lattice_region.low[k];
                    d->dbx_sub_box_lattice_region.high[k] = m * d->sub_box_lattice_region.high[k];
                    d->dbx_ghost_ext_lattice_region.low[k] = m * d->ghost_ext_lattice_region.low[k];
                    d->dbx_ghost_ext_lattice_region.high[k] = m * d->ghost_ext_lattice_region.high[k];
                }
                return d;
            }
            // single-process build (the reference calls build() after setComm; here: a 1x1x1 grid)
            BccDomain *build() {
                const int g[3] = {1, 1, 1}, c[3] = {0, 0, 0};
                return localBuild(g, c);
            }
        };
    };
}
#endif
