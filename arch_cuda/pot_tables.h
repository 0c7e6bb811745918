// arch_cuda/pot_tables.h -- how the hooks read the spline tables out of the reference's `eam` object.
//
// The C ABI wants, per tabulated function, what libpot holds after eam::interpolateFile() (call site: reference
// src/simulation.cpp:131): n, 1/dx and the (n+1) x 7 coefficient rows. libpot v0.1.0 (reference pkg.yaml:16) is
// not vendored in the reference tree, so two adapters exist:
//   * ORACLE_SHIM_EAM_H (the test build of this repository: oracle/shim/eam.h wraps oracle/pot.c), and
//   * the libpot build, written against libpot's public members as published upstream (github.com/misa-md/potential:
//     eam::electron_density / eam::embedded of type EamBaseList with getEamItemByType(key) -> InterpolationObject*,
//     eam::eam_phi of type OneWayEamList with getPhiByEamPhiByType(from, to); InterpolationObject {n, invDx, spline}).
//     UNVERIFIED here (no libpot source in this environment) -- see INTEGRATION.md section 3.
#ifndef ARCH_CUDA_POT_TABLES_H
#define ARCH_CUDA_POT_TABLES_H
#include <eam.h>
#include "misa_b200.h"

namespace arch_cuda {
#ifdef ORACLE_SHIM_EAM_H
    inline bool from_pot_table(const pot_table &t, misa_b200_table *out) {
        if (!t.spline) return false;
        out->n = t.n;
        out->inv_dx = t.inv_dx;
        out->spline = t.spline;
        return true;
    }
    inline bool electron_density_table(eam *e, unsigned short key, misa_b200_table *out) {
        const int i = pot_index_of_key(e->p, key);
        return i >= 0 && from_pot_table(e->p->elec[i], out);
    }
    inline bool embedded_table(eam *e, unsigned short key, misa_b200_table *out) {
        const int i = pot_index_of_key(e->p, key);
        return i >= 0 && from_pot_table(e->p->embed[i], out);
    }
    inline bool pair_table(eam *e, unsigned short key_from, unsigned short key_to, misa_b200_table *out) {
        const int i = pot_index_of_key(e->p, key_from), j = pot_index_of_key(e->p, key_to);
        return i >= 0 && j >= 0 && from_pot_table(e->p->phi[i][j], out);
    }
#else
    template<class Interp>
    inline bool from_interpolation_object(const Interp *o, misa_b200_table *out) {
        if (o == nullptr || o->spline == nullptr) return false;
        out->n = o->n;
        out->inv_dx = o->invDx;
        out->spline = &o->spline[0][0];
        return true;
    }
    inline bool electron_density_table(eam *e, unsigned short key, misa_b200_table *out) {
        return from_interpolation_object(e->electron_density.getEamItemByType(key), out);
    }
    inline bool embedded_table(eam *e, unsigned short key, misa_b200_table *out) {
        return from_interpolation_object(e->embedded.getEamItemByType(key), out);
    }
    inline bool pair_table(eam *e, unsigned short key_from, unsigned short key_to, misa_b200_table *out) {
        return from_interpolation_object(e->eam_phi.getPhiByEamPhiByType(key_from, key_to), out);
    }
#endif
}
#endif
