/*
 * oracle/md_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the MISA-MD EAM hot path: lattice-indexed neighbour
 * stencil, Wigner-Seitz site mapping, the half-list rho / df / force loops incl. off-lattice "inter"
 * atoms, velocity-Verlet, run-away detection, and the six staged ghost exchanges -- for any number of
 * sub-boxes ("ranks") held in ONE process (the reference uses one MPI rank per sub-box; here the staged
 * x->y->z exchange of libcomm v0.3.3 -- absent from /root/reference, pkg.yaml:17 -- is restated as an
 * in-process mailbox swap with the same packer semantics).
 *
 * Every function cites the reference file:line it follows. The EAM arithmetic (libpot) is in pot.[ch]
 * and is "parity unpinned" (see pot.h). The integer/lattice parts are pinned by the reference's own unit
 * tests re-expressed in tests/test_oracle_kat.py, and the whole file is cross-checked against the
 * reference's own sources compiled in place (oracle/_ref, see oracle/Makefile) in tests/test_oracle_vs_ref.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 */
#ifndef MISA_MD_ORACLE_H
#define MISA_MD_ORACLE_H

#include <stddef.h>
#include "pot.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/atom/atom_element.h:18-41 -- 104-byte AoS record (id@0 type@8 x@16 v@40 f@64 rho@88 df@96) */
typedef struct ora_atom {
    unsigned long id;
    int type; /* reference src/types/atom_types.h:23-25: INVALID=-1, Fe=0, Cu=1, Ni=2 */
    int _pad;
    double x[3];
    double v[3];
    double f[3];
    double rho;
    double df;
} ora_atom;

#define ORA_INVALID (-1)
#define ORA_DIR_LOWER 0
#define ORA_DIR_HIGHER 1

/* reference src/lattice/box.h:10-20 */
#define ORA_IN_BOX 0u
#define ORA_OUT_X_LITTER 1u
#define ORA_OUT_X_BIG 2u
#define ORA_OUT_Y_LITTER 4u
#define ORA_OUT_Y_BIG 8u
#define ORA_OUT_Z_LITTER 16u
#define ORA_OUT_Z_BIG 32u
#define ORA_INDEX_NOT_EXISTS (-1L)

typedef struct ora_iregion { int x_low, y_low, z_low, x_high, y_high, z_high; } ora_iregion;

/* libcomm comm::Domain / comm::BccDomain restated: the fields the reference reads. */
typedef struct ora_domain {
    long phase_space[3];
    int grid_size[3], grid_coord[3];
    int rank, n_ranks;
    int rank_id_neighbours[3][2];
    double lattice_const, cutoff_radius_factor;
    int cut_lattice;
    double meas_global_length[3];
    double meas_global_low[3], meas_global_high[3];
    double meas_sub_box_low[3], meas_sub_box_high[3];
    int sub_box_lattice_size[3];
    int lattice_size_ghost[3];
    int ghost_extended_lattice_size[3];
    ora_iregion sub_box_lattice_region;
    ora_iregion ghost_ext_lattice_region;
    /* doubled-x twins (comm::BccDomain) */
    int dbx_sub_box_lattice_size[3];
    int dbx_lattice_size_ghost[3];
    int dbx_ghost_extended_lattice_size[3];
    ora_iregion dbx_sub_box_lattice_region;
    ora_iregion dbx_ghost_ext_lattice_region;
} ora_domain;

typedef struct ora_ivec { long *v; size_t n, cap; } ora_ivec;

/* one sub-box: reference class `atom` (AtomSet + AtomList + InterAtomList + NeighbourIndex) */
typedef struct ora_rank {
    ora_domain dom;
    /* BccLattice, reference src/lattice/lattice.h:48-66 */
    long size_x, size_y, size_z, size; /* size_x is doubled */
    ora_atom *atoms;                   /* AtomList::_atoms */
    ora_ivec sendlist[6], recvlist[6];  /* reference src/atom/atom_list.h:196-197 */
    /* NeighbourIndex, reference src/atom/neighbour_index.h:66-69 */
    ora_ivec nei_even, nei_odd, nei_half_even, nei_half_odd;
    /* InterAtomList, reference src/atom/inter_atom_list.h:21-40 */
    ora_atom *inter; size_t n_inter, cap_inter;
    ora_atom *ghost; size_t n_ghost, cap_ghost;
    ora_ivec intersend[6], interrecv[6]; /* refs: >=0 local inter index, <0 => ~ghost index */
    /* inter_map (site -> refs), rebuilt by makeIndex */
    long *map_site; long *map_ref; size_t n_map, cap_map;
    double cutoff_radius;
} ora_rank;

typedef struct ora_world {
    int n_ranks;
    ora_rank *ranks;
    const pot_eam *pot;
    double dt;
    double dt_inv_m[3];
    int n_threads; /* ranks are processed by up to this many OpenMP threads */
} ora_world;

/* ---- construction -------------------------------------------------------------------------- */
/* libcomm Domain::Builder::localBuild restated (phase space must divide evenly by the grid). */
int ora_domain_build(ora_domain *d, const long phase_space[3], const int grid_size[3], const int grid_coord[3],
                     double lattice_const, double cutoff_radius_factor, int ghost_size /* <0: default = cut_lattice */);

ora_world *ora_world_create(const long phase_space[3], const int grid_size[3], double lattice_const,
                            double cutoff_radius_factor, const pot_eam *pot, double dt, int n_threads);
void ora_world_free(ora_world *w);
ora_rank *ora_world_rank(ora_world *w, int r);
ora_atom *ora_rank_atoms(ora_rank *rk);
long ora_rank_size(const ora_rank *rk);

/* perfect bcc lattice + ids for the owned sites, reference src/world_builder.cpp:105-131 (positions);
 * ids are global-lattice ids 1 + ((gz*PY + gy)*2PX + gx) so they do not depend on the decomposition
 * (identical to the reference for a 1x1x1 grid). types/velocities are set by the caller. */
void ora_world_fill_lattice(ora_world *w);

/* ---- pieces (reference function in brackets) ---------------------------------------------- */
void ora_nei_make(ora_rank *rk, int cut_lattice, double cutoff_radius_factor); /* NeighbourIndex::make */
int ora_is_positive_index(double x, double y, double z);                        /* ::isPositiveIndex */
void ora_voronoy(double X, double Y, double Z, double LC, long out[3]);         /* VORONOY macro */
unsigned ora_is_out_box(const ora_atom *a, const ora_domain *d);                /* ws::isOutBox */
void ora_near_lat_coord(const ora_atom *a, const ora_domain *d, long c[3]);     /* ws::getNearLatCoord */
void ora_near_lat_sub_box_coord(const ora_atom *a, const ora_domain *d, long c[3]); /* ws::getNearLatSubBoxCoord */
long ora_find_near_lat_index_in_sub_box(const ora_rank *rk, const ora_atom *a);  /* ws::findNearLatIndexInSubBox */
ora_iregion ora_fw_comm_local_region(const ora_domain *d, int dim, int dir);    /* comm::fwCommLocalRegion */

void ora_set_dt(ora_world *w, double dt);           /* NewtonMotion::setTimestepLength */
void ora_exchange_atom_first(ora_world *w);         /* AtomList::exchangeAtomFirst */
void ora_exchange_atom(ora_world *w);               /* AtomList::exchangeAtom */
void ora_exchange_inter(ora_world *w);              /* InterAtomList::exchangeInter */
void ora_border_inter(ora_world *w);                /* InterAtomList::borderInter */
void ora_clear_force(ora_world *w);                 /* atom::clearForce */
void ora_compute_eam(ora_world *w);                 /* atom::computeEam */
void ora_first_step(ora_world *w);                  /* NewtonMotion::firststep */
void ora_second_step(ora_world *w);                 /* NewtonMotion::secondstep */
int ora_decide(ora_world *w);                       /* atom::decide (returns OR of nflag) */
void ora_setv(ora_world *w, const int lat[4], const double direction[3], double energy); /* atom::setv */
void ora_collision_step(ora_world *w, const int lat[4], const double direction[3], double energy);
void ora_prepare(ora_world *w);  /* simulation::prepareForStart tail: exchangeAtomFirst, clearForce, computeEam */
void ora_step(ora_world *w);     /* one iteration of simulation::simulate */

/* individual lattice passes on ONE rank, exposed for kernel-level parity tests */
void ora_lat_rho(ora_rank *rk, const pot_eam *pot);
void ora_lat_df(ora_rank *rk, const pot_eam *pot);
void ora_lat_force(ora_rank *rk, const pot_eam *pot);

/* diagnostics */
double ora_mvv(const ora_world *w);                 /* configuration::mvv summed over ranks */
double ora_kinetic_energy(const ora_world *w);      /* 0.5*mvv*mvv2e [eV] */
double ora_temperature(const ora_world *w);         /* configuration::temperature */
void ora_rescale(ora_world *w, double T);           /* configuration::rescale */
double ora_potential_energy(ora_world *w);          /* ours: sum F(rho) + pair sum (needs rho current) */
size_t ora_total_inter(const ora_world *w);
void ora_test_set_inter(ora_world *w, int r, const ora_atom *atoms, size_t n); /* test helper */
size_t ora_rank_n_inter(const ora_rank *rk);
ora_atom *ora_rank_inter(ora_rank *rk);
size_t ora_rank_n_ghost_inter(const ora_rank *rk);

/* ---- dump record stream ------------------------------------------------------------------ */
/* atom_dump::AtomInfoDump, reference frontend/io/atom_info_dump.h:14-22 -- 72 bytes:
 * id@0 step@8 type@16 inter_type@20 (short; 2 pad bytes, zero here) atom_location@24 atom_velocity@48 */
typedef struct ora_dump_record {
    unsigned long id;
    size_t step;
    int type;
    short inter_type;
    short _pad;
    double atom_location[3];
    double atom_velocity[3];
} ora_dump_record;
/* AtomDump::dump of one sub-box (reference frontend/io/atom_dump.cpp:39-75): inter atoms in list order, then the
 * valid sites of the owned region (frontend/io/output_base_interface.h:26-31) in z,y,x order. Returns the record
 * count; records beyond `cap` are not written. */
size_t ora_dump(const ora_rank *rk, size_t time_step, ora_dump_record *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
