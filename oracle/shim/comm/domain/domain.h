// oracle/shim/comm/domain/domain.h -- TEST INFRASTRUCTURE. libcomm v0.3.3 comm::Domain (+ Builder) restated:
// exactly the fields the reference reads. localBuild() = the sub-box a rank WOULD own in
// a virtual process grid, the same call the reference's own unit tests use (tests/unit/inter_atom_test.cpp:10-24).
#ifndef ORACLE_SHIM_COMM_DOMAIN_H
#define ORACLE_SHIM_COMM_DOMAIN_H
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include "../types_define.h"
#include "region.hpp"

namespace comm {
    class Domain {
    public:
        double lattice_const;
        double cutoff_radius_factor;
        _type_lattice_size cut_lattice;
        int64_t phase_space[DIMENSION_SIZE];
        int grid_size[DIMENSION_SIZE];
        int grid_coord[DIMENSION_SIZE];
        _MPI_Rank rank_id_neighbours[DIMENSION_SIZE][2];
        double meas_global_length[DIMENSION_SIZE];
        Region<double> meas_global_region, meas_sub_box_region, meas_ghost_ext_region;
        _type_lattice_size sub_box_lattice_size[DIMENSION_SIZE];
        _type_lattice_size lattice_size_ghost[DIMENSION_SIZE];
        _type_lattice_size ghost_extended_lattice_size[DIMENSION_SIZE];
        Region<_type_lattice_coord> sub_box_lattice_region, ghost_ext_lattice_region;
        int rank; // rank id in MPI_Cart order (x slowest)

        class Builder {
        public:
            Builder() : _a(0), _crf(0), _ghost(-1) { _ps[0] = _ps[1] = _ps[2] = 0; }
            Builder &setComm(mpi_process, MPI_Comm *) { return *this; }
            Builder &setPhaseSpace(const int64_t ps[DIMENSION_SIZE]) { for (int i = 0; i < 3; i++) _ps[i] = ps[i]; return *this; }
            Builder &setLatticeConst(const double a) { _a = a; return *this; }
            Builder &setCutoffRadius(const double crf) { _crf = crf; return *this; }
            Builder &setGhostSize(const _type_lattice_size g) { _ghost = g; return *this; }
            Domain *localBuild(const int grid_size[DIMENSION_SIZE], const int grid_coord[DIMENSION_SIZE]) {
                Domain *d = new Domain();
                fill(d, grid_size, grid_coord);
                return d;
            }
        protected:
            int64_t _ps[3];
            double _a, _crf;
            int _ghost;
            static int rankOf(const int c[3], const int g[3]) { return (c[0] * g[1] + c[1]) * g[2] + c[2]; }
            void fill(Domain *d, const int grid_size[3], const int grid_coord[3]) const {
                d->lattice_const = _a;
                d->cutoff_radius_factor = _crf;
                d->cut_lattice = static_cast<_type_lattice_size>(std::ceil(_crf));
                const int ghost = _ghost < 0 ? d->cut_lattice : _ghost;
                d->rank = rankOf(grid_coord, grid_size);
                for (int k = 0; k < 3; k++) {
                    if (_ps[k] % grid_size[k] != 0) throw std::invalid_argument("phase space must divide by the grid");
                    d->phase_space[k] = _ps[k];
                    d->grid_size[k] = grid_size[k];
                    d->grid_coord[k] = grid_coord[k];
                    const int n = static_cast<int>(_ps[k] / grid_size[k]);
                    d->sub_box_lattice_size[k] = n;
                    d->lattice_size_ghost[k] = ghost;
                    d->ghost_extended_lattice_size[k] = n + 2 * ghost;
                    d->sub_box_lattice_region.low[k] = grid_coord[k] * n;
                    d->sub_box_lattice_region.high[k] = grid_coord[k] * n + n;
                    d->ghost_ext_lattice_region.low[k] = grid_coord[k] * n - ghost;
                    d->ghost_ext_lattice_region.high[k] = grid_coord[k] * n + n + ghost;
                    d->meas_global_length[k] = _ps[k] * _a;
                    d->meas_global_region.low[k] = 0.0;
                    d->meas_global_region.high[k] = d->meas_global_length[k];
                    d->meas_sub_box_region.low[k] = d->sub_box_lattice_region.low[k] * _a;
                    d->meas_sub_box_region.high[k] = d->sub_box_lattice_region.high[k] * _a;
                    d->meas_ghost_ext_region.low[k] = d->ghost_ext_lattice_region.low[k] * _a;
                    d->meas_ghost_ext_region.high[k] = d->ghost_ext_lattice_region.high[k] * _a;
                    int c[3] = {grid_coord[0], grid_coord[1], grid_coord[2]};
                    c[k] = (grid_coord[k] - 1 + grid_size[k]) % grid_size[k];
                    d->rank_id_neighbours[k][DIR_LOWER] = rankOf(c, grid_size);
                    c[k] = (grid_coord[k] + 1) % grid_size[k];
                    d->rank_id_neighbours[k][DIR_HIGHER] = rankOf(c, grid_size);
                }
            }
        };
    };
}
#endif
