/*
 * oracle/md_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE. See md_oracle.h.
 * Compile with -ffp-contract=off: the reference semantics are taken to be IEEE double without FMA
 * contraction (what `g++ -std=c++11` gives for the reference's C++11 build, CMakeLists.txt:4).
 */
#include "md_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* reference src/types/pre_define.h:11-19 */
#define BOLTZ 8.617343e-5
#define mvv2e 1.0364269e-4
#define ftm2v (1.0 / mvv2e)

/* reference src/types/atom_types.h:17-19,35-46,61-73 */
static inline double atom_mass(int type) {
    switch (type) {
        case 0: return 55.845;
        case 1: return 63.546;
        case 2: return 58.6934;
        default: return 0;
    }
}
static inline int type_key(int type) {
    switch (type) {
        case 0: return 26;
        case 1: return 29;
        case 2: return 28;
        default: return 0;
    }
}

/* ---- small containers ----------------------------------------------------------------------- */
static void ivec_push(ora_ivec *v, long x) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 64;
        v->v = (long *)realloc(v->v, v->cap * sizeof(long));
    }
    v->v[v->n++] = x;
}
static void ivec_clear(ora_ivec *v) { v->n = 0; }
static void ivec_free(ora_ivec *v) { free(v->v); v->v = NULL; v->n = v->cap = 0; }

static void inter_push(ora_atom **arr, size_t *n, size_t *cap, const ora_atom *a) {
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 64;
        *arr = (ora_atom *)realloc(*arr, *cap * sizeof(ora_atom));
    }
    (*arr)[(*n)++] = *a;
}
/* std::list::erase keeps the order of the remaining elements */
static void inter_erase(ora_rank *rk, size_t i) {
    memmove(&rk->inter[i], &rk->inter[i + 1], (rk->n_inter - i - 1) * sizeof(ora_atom));
    rk->n_inter--;
}
static inline ora_atom *inter_ref(ora_rank *rk, long ref) { return ref >= 0 ? &rk->inter[ref] : &rk->ghost[~ref]; }

/* ---- domain (libcomm Domain::Builder restated; fields per SURVEY.md section 8c) ---------------- */
int ora_domain_build(ora_domain *d, const long phase_space[3], const int grid_size[3], const int grid_coord[3],
                     double lattice_const, double cutoff_radius_factor, int ghost_size) {
    memset(d, 0, sizeof *d);
    d->lattice_const = lattice_const;
    d->cutoff_radius_factor = cutoff_radius_factor;
    d->cut_lattice = (int)ceil(cutoff_radius_factor);
    if (ghost_size < 0) ghost_size = d->cut_lattice;
    d->n_ranks = grid_size[0] * grid_size[1] * grid_size[2];
    d->rank = (grid_coord[0] * grid_size[1] + grid_coord[1]) * grid_size[2] + grid_coord[2]; /* MPI_Cart order */
    int lo[3], hi[3], glo[3], ghi[3];
    for (int k = 0; k < 3; k++) {
        if (phase_space[k] % grid_size[k] != 0) return -1;
        d->phase_space[k] = phase_space[k];
        d->grid_size[k] = grid_size[k];
        d->grid_coord[k] = grid_coord[k];
        const int n = (int)(phase_space[k] / grid_size[k]);
        d->sub_box_lattice_size[k] = n;
        d->lattice_size_ghost[k] = ghost_size;
        d->ghost_extended_lattice_size[k] = n + 2 * ghost_size;
        lo[k] = grid_coord[k] * n;
        hi[k] = lo[k] + n;
        glo[k] = lo[k] - ghost_size;
        ghi[k] = hi[k] + ghost_size;
        d->meas_global_length[k] = phase_space[k] * lattice_const;
        d->meas_global_low[k] = 0.0;
        d->meas_global_high[k] = d->meas_global_length[k];
        d->meas_sub_box_low[k] = lo[k] * lattice_const;
        d->meas_sub_box_high[k] = hi[k] * lattice_const;
        int c[3] = {grid_coord[0], grid_coord[1], grid_coord[2]};
        c[k] = (grid_coord[k] - 1 + grid_size[k]) % grid_size[k];
        d->rank_id_neighbours[k][ORA_DIR_LOWER] = (c[0] * grid_size[1] + c[1]) * grid_size[2] + c[2];
        c[k] = (grid_coord[k] + 1) % grid_size[k];
        d->rank_id_neighbours[k][ORA_DIR_HIGHER] = (c[0] * grid_size[1] + c[1]) * grid_size[2] + c[2];
        d->dbx_sub_box_lattice_size[k] = (k == 0 ? 2 : 1) * n;
        d->dbx_lattice_size_ghost[k] = (k == 0 ? 2 : 1) * ghost_size;
        d->dbx_ghost_extended_lattice_size[k] = (k == 0 ? 2 : 1) * (n + 2 * ghost_size);
    }
    d->sub_box_lattice_region = (ora_iregion){lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
    d->ghost_ext_lattice_region = (ora_iregion){glo[0], glo[1], glo[2], ghi[0], ghi[1], ghi[2]};
    d->dbx_sub_box_lattice_region = (ora_iregion){2 * lo[0], lo[1], lo[2], 2 * hi[0], hi[1], hi[2]};
    d->dbx_ghost_ext_lattice_region = (ora_iregion){2 * glo[0], glo[1], glo[2], 2 * ghi[0], ghi[1], ghi[2]};
    return 0;
}

/* comm::fwCommLocalRegion restated from the receive slabs of LatPackerFirst::onReceive
 * (reference src/pack/lat_particle_packer.cpp:65-76,97-108,128-139) and its use at
 * reference src/atom/atom_list.cpp:33-40: region of LOCAL sites (ghost-extended, doubled-x indices)
 * whose data is forwarded to neighbour [dim][dir]. */
ora_iregion ora_fw_comm_local_region(const ora_domain *d, int dim, int dir) {
    const int *g = d->dbx_lattice_size_ghost, *b = d->dbx_sub_box_lattice_size, *e = d->dbx_ghost_extended_lattice_size;
    int lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        if (k == dim) {
            if (dir == ORA_DIR_LOWER) { lo[k] = g[k]; hi[k] = 2 * g[k]; }
            else { lo[k] = b[k]; hi[k] = b[k] + g[k]; }
        } else if (k < dim) { lo[k] = 0; hi[k] = e[k]; }
        else { lo[k] = g[k]; hi[k] = g[k] + b[k]; }
    }
    return (ora_iregion){lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]};
}
static inline int region_is_in(const ora_iregion *r, long x, long y, long z) {
    return x >= r->x_low && x < r->x_high && y >= r->y_low && y < r->y_high && z >= r->z_low && z < r->z_high;
}

/* ---- lattice index math: reference src/lattice/lattice.h:48-66 ------------------------------- */
static inline long idx3(const ora_rank *rk, long x, long y, long z) { return (z * rk->size_y + y) * rk->size_x + x; }
static inline void idx_to_3d(const ora_rank *rk, long index, long *x, long *y, long *z) {
    *x = index % rk->size_x;
    index = index / rk->size_x;
    *y = index % rk->size_y;
    *z = index / rk->size_y;
}

/* ---- NeighbourIndex: reference src/atom/neighbour_index.inl:13-92 ---------------------------- */
int ora_is_positive_index(double x, double y, double z) {
    if (z > 0) return 1;
    else if (z == 0) {
        if (y > 0) return 1;
        else if (y == 0) { if (x > 0) return 1; }
    }
    return 0;
}

void ora_nei_make(ora_rank *rk, int cut_lattice, double cutoff_radius_factor) {
    ivec_clear(&rk->nei_even); ivec_clear(&rk->nei_odd); ivec_clear(&rk->nei_half_even); ivec_clear(&rk->nei_half_odd);
    const double cutoff_lat_factor = cutoff_radius_factor + 2 * 0.5; /* config::nei_lat_cutoff, md_building_config.h.in:24-29 */
    for (long zI = -cut_lattice - 1; zI <= cut_lattice + 1; zI++)
        for (long yI = -cut_lattice - 1; yI <= cut_lattice + 1; yI++)
            for (long xI = -2 * cut_lattice - 2; xI <= 2 * cut_lattice + 2; xI++) {
                double z = (double)zI + (((double)(xI % 2)) / 2);
                double y = (double)yI + (((double)(xI % 2)) / 2);
                double x = ((double)xI) / 2;
                const double r = x * x + y * y + z * z;
                if (r < cutoff_lat_factor * cutoff_lat_factor && r > 0) {
                    const long ix = xI;
                    const long iy = (xI < 0 && xI % 2 != 0) ? yI - 1 : yI;
                    const long iz = (xI < 0 && xI % 2 != 0) ? zI - 1 : zI;
                    const long off = idx3(rk, ix, iy, iz);
                    ivec_push(&rk->nei_even, off);
                    if (ora_is_positive_index(x, y, z)) ivec_push(&rk->nei_half_even, off);
                }
            }
    for (long zI = -cut_lattice - 1; zI <= cut_lattice + 1; zI++)
        for (long yI = -cut_lattice - 1; yI <= cut_lattice + 1; yI++)
            for (long xI = -2 * cut_lattice - 2; xI <= 2 * cut_lattice + 2; xI++) {
                double z = (double)zI - (((double)(xI % 2)) / 2);
                double y = (double)yI - (((double)(xI % 2)) / 2);
                double x = (double)xI / 2;
                const double r = x * x + y * y + z * z;
                if (r < cutoff_lat_factor * cutoff_lat_factor && r > 0) {
                    const long ix = xI;
                    const long iy = (xI < 0 && xI % 2 != 0) ? yI + 1 : yI;
                    const long iz = (xI < 0 && xI % 2 != 0) ? zI + 1 : zI;
                    const long off = idx3(rk, ix, iy, iz);
                    ivec_push(&rk->nei_odd, off);
                    if (ora_is_positive_index(x, y, z)) ivec_push(&rk->nei_half_odd, off);
                }
            }
}

/* NeighbourIndex::begin/end flag selection: reference src/atom/neighbour_index.inl:96-132 */
static inline const ora_ivec *nei_list(const ora_rank *rk, int half, long x) {
    const int flag = (half ? 2 : 0) | (x % 2 == 0 ? 1 : 0);
    switch (flag) {
        case 0: return &rk->nei_odd;
        case 1: return &rk->nei_even;
        case 2: return &rk->nei_half_odd;
        default: return &rk->nei_half_even;
    }
}

/* ---- Wigner-Seitz mapping: reference src/lattice/ws_utils.cpp:15-163 -------------------------- */
static const long ws_offset[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 0, -1}, {1, 0, -1},
                                     {-1, -1, 0},  {1, -1, 0},  {-1, 0, 0},  {1, 0, 0}};
static const double ws_normal[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {1, 1, -1},
                                       {-1, -1, 1},  {1, -1, 1},  {-1, 1, 1},  {1, 1, 1}};
static const double ws_d = -3.0 / 4.0;

void ora_voronoy(double X, double Y, double Z, double LC, long out[3]) { /* ws_utils.cpp:66-89 */
    long cx = (long)lround(X / LC), cy = (long)lround(Y / LC), cz = (long)lround(Z / LC);
    const double dx = X / LC - cx, dy = Y / LC - cy, dz = Z / LC - cz;
    const unsigned flag = (dz > 0 ? 4u : 0u) | (dy > 0 ? 2u : 0u) | (dx > 0 ? 1u : 0u);
    cx = 2 * cx;
    if (ws_normal[flag][0] * dx + ws_normal[flag][1] * dy + ws_normal[flag][2] * dz + ws_d >= 0.0) {
        cx += ws_offset[flag][0];
        cy += ws_offset[flag][1];
        cz += ws_offset[flag][2];
    }
    out[0] = cx; out[1] = cy; out[2] = cz;
}

unsigned ora_is_out_box(const ora_atom *a, const ora_domain *d) { /* ws_utils.cpp:91-115 */
    long c[3];
    ora_voronoy(a->x[0], a->x[1], a->x[2], d->lattice_const, c);
    c[0] -= 2 * d->sub_box_lattice_region.x_low;
    c[1] -= d->sub_box_lattice_region.y_low;
    c[2] -= d->sub_box_lattice_region.z_low;
    unsigned flag = ORA_IN_BOX;
    if (c[0] < 0) flag |= ORA_OUT_X_LITTER;
    else if (c[0] >= 2 * d->sub_box_lattice_size[0]) flag |= ORA_OUT_X_BIG;
    if (c[1] < 0) flag |= ORA_OUT_Y_LITTER;
    else if (c[1] >= d->sub_box_lattice_size[1]) flag |= ORA_OUT_Y_BIG;
    if (c[2] < 0) flag |= ORA_OUT_Z_LITTER;
    else if (c[2] >= d->sub_box_lattice_size[2]) flag |= ORA_OUT_Z_BIG;
    return flag;
}

void ora_near_lat_coord(const ora_atom *a, const ora_domain *d, long c[3]) { /* ws_utils.cpp:156-163 */
    ora_voronoy(a->x[0], a->x[1], a->x[2], d->lattice_const, c);
    c[0] -= 2 * d->ghost_ext_lattice_region.x_low;
    c[1] -= d->ghost_ext_lattice_region.y_low;
    c[2] -= d->ghost_ext_lattice_region.z_low;
}

void ora_near_lat_sub_box_coord(const ora_atom *a, const ora_domain *d, long c[3]) { /* ws_utils.cpp:181-189 */
    ora_voronoy(a->x[0], a->x[1], a->x[2], d->lattice_const, c);
    c[0] -= 2 * d->sub_box_lattice_region.x_low;
    c[1] -= d->sub_box_lattice_region.y_low;
    c[2] -= d->sub_box_lattice_region.z_low;
}

long ora_find_near_lat_index_in_sub_box(const ora_rank *rk, const ora_atom *a) { /* ws_utils.cpp:133-154 */
    const ora_domain *d = &rk->dom;
    long c[3];
    ora_near_lat_sub_box_coord(a, d, c);
    long j = c[0], k = c[1], l = c[2];
    if (j < 0 || k < 0 || l < 0 || j >= 2 * d->sub_box_lattice_size[0] || k >= d->sub_box_lattice_size[1] ||
        l >= d->sub_box_lattice_size[2])
        return ORA_INDEX_NOT_EXISTS;
    j += 2 * d->lattice_size_ghost[0];
    k += d->lattice_size_ghost[1];
    l += d->lattice_size_ghost[2];
    return idx3(rk, j, k, l);
}

/* ---- world ---------------------------------------------------------------------------------- */
ora_world *ora_world_create(const long phase_space[3], const int grid_size[3], double lattice_const,
                            double cutoff_radius_factor, const pot_eam *pot, double dt, int n_threads) {
    ora_world *w = (ora_world *)calloc(1, sizeof *w);
    w->n_ranks = grid_size[0] * grid_size[1] * grid_size[2];
    w->ranks = (ora_rank *)calloc((size_t)w->n_ranks, sizeof(ora_rank));
    w->pot = pot;
    w->n_threads = n_threads > 0 ? n_threads : 1;
    for (int cx = 0; cx < grid_size[0]; cx++)
        for (int cy = 0; cy < grid_size[1]; cy++)
            for (int cz = 0; cz < grid_size[2]; cz++) {
                const int coord[3] = {cx, cy, cz};
                ora_domain d;
                /* ghost = ceil(crf)+1: reference src/simulation.cpp:40-46 */
                if (ora_domain_build(&d, phase_space, grid_size, coord, lattice_const, cutoff_radius_factor,
                                     (int)ceil(cutoff_radius_factor) + 1)) {
                    ora_world_free(w);
                    return NULL;
                }
                ora_rank *rk = &w->ranks[d.rank];
                rk->dom = d;
                /* AtomSet ctor, reference src/atom/atom_set.cpp:12-31: x doubled */
                rk->size_x = d.dbx_ghost_extended_lattice_size[0];
                rk->size_y = d.dbx_ghost_extended_lattice_size[1];
                rk->size_z = d.dbx_ghost_extended_lattice_size[2];
                rk->size = rk->size_x * rk->size_y * rk->size_z;
                rk->atoms = (ora_atom *)calloc((size_t)rk->size, sizeof(ora_atom));
                rk->cutoff_radius = d.lattice_const * d.cutoff_radius_factor; /* reference src/atom.cpp:15 */
                ora_nei_make(rk, d.cut_lattice, d.cutoff_radius_factor);       /* reference src/simulation.cpp:64 */
            }
    ora_set_dt(w, dt);
    return w;
}

void ora_world_free(ora_world *w) {
    if (!w) return;
    for (int r = 0; r < w->n_ranks; r++) {
        ora_rank *rk = &w->ranks[r];
        free(rk->atoms); free(rk->inter); free(rk->ghost); free(rk->map_site); free(rk->map_ref);
        ivec_free(&rk->nei_even); ivec_free(&rk->nei_odd); ivec_free(&rk->nei_half_even); ivec_free(&rk->nei_half_odd);
        for (int i = 0; i < 6; i++) {
            ivec_free(&rk->sendlist[i]); ivec_free(&rk->recvlist[i]);
            ivec_free(&rk->intersend[i]); ivec_free(&rk->interrecv[i]);
        }
    }
    free(w->ranks);
    free(w);
}

ora_rank *ora_world_rank(ora_world *w, int r) { return &w->ranks[r]; }
ora_atom *ora_rank_atoms(ora_rank *rk) { return rk->atoms; }
long ora_rank_size(const ora_rank *rk) { return rk->size; }
size_t ora_rank_n_inter(const ora_rank *rk) { return rk->n_inter; }
ora_atom *ora_rank_inter(ora_rank *rk) { return rk->inter; }
size_t ora_rank_n_ghost_inter(const ora_rank *rk) { return rk->n_ghost; }
/* test helper: replace rank r's local inter-atom list (InterAtomList::addInterAtom per element) */
void ora_test_set_inter(ora_world *w, int r, const ora_atom *atoms, size_t n) {
    ora_rank *rk = &w->ranks[r];
    rk->n_inter = 0;
    for (size_t i = 0; i < n; i++) inter_push(&rk->inter, &rk->n_inter, &rk->cap_inter, &atoms[i]);
}
size_t ora_total_inter(const ora_world *w) {
    size_t n = 0;
    for (int r = 0; r < w->n_ranks; r++) n += w->ranks[r].n_inter;
    return n;
}

void ora_world_fill_lattice(ora_world *w) { /* positions: reference src/world_builder.cpp:120-124 */
    for (int r = 0; r < w->n_ranks; r++) {
        ora_rank *rk = &w->ranks[r];
        const ora_domain *d = &rk->dom;
        const double a = d->lattice_const;
        for (long i = 0; i < rk->size; i++) rk->atoms[i].type = ORA_INVALID; /* ghosts are filled by the first exchange */
        for (int k = 0; k < d->dbx_sub_box_lattice_size[2]; k++)
            for (int j = 0; j < d->dbx_sub_box_lattice_size[1]; j++)
                for (int i = 0; i < d->dbx_sub_box_lattice_size[0]; i++) {
                    ora_atom *at = &rk->atoms[idx3(rk, i + d->dbx_lattice_size_ghost[0], j + d->dbx_lattice_size_ghost[1],
                                                   k + d->dbx_lattice_size_ghost[2])];
                    const long gx = d->dbx_sub_box_lattice_region.x_low + i, gy = d->dbx_sub_box_lattice_region.y_low + j,
                               gz = d->dbx_sub_box_lattice_region.z_low + k;
                    at->id = 1ul + (unsigned long)((gz * d->phase_space[1] + gy) * (2 * d->phase_space[0]) + gx);
                    at->type = 0;
                    at->x[0] = (d->dbx_sub_box_lattice_region.x_low + i) * 0.5 * (a);
                    at->x[1] = (d->dbx_sub_box_lattice_region.y_low + j) * a + (i % 2) * (a / 2);
                    at->x[2] = (d->dbx_sub_box_lattice_region.z_low + k) * a + (i % 2) * (a / 2);
                    at->v[0] = at->v[1] = at->v[2] = 0.0;
                }
    }
}

/* ---- NewtonMotion: reference src/newton_motion.cpp:13-74 -------------------------------------- */
void ora_set_dt(ora_world *w, double dt) {
    w->dt = dt;
    for (int i = 0; i < 3; i++) {
        const double dt_halve = 0.5 * dt * ftm2v;
        w->dt_inv_m[i] = dt_halve / atom_mass(i);
    }
}

#define FOR_RANKS(w, rk) \
    _Pragma("omp parallel for schedule(static) num_threads(w->n_threads)") \
    for (int _r = 0; _r < (w)->n_ranks; _r++) { ora_rank *rk = &(w)->ranks[_r];
#define END_RANKS }

#define FOR_SUBBOX(rk, at)                                                                                        \
    for (long _z = (rk)->dom.dbx_lattice_size_ghost[2]; _z < (rk)->dom.dbx_sub_box_lattice_size[2] + (rk)->dom.dbx_lattice_size_ghost[2]; _z++) \
        for (long _y = (rk)->dom.dbx_lattice_size_ghost[1]; _y < (rk)->dom.dbx_sub_box_lattice_size[1] + (rk)->dom.dbx_lattice_size_ghost[1]; _y++) \
            for (long _x = (rk)->dom.dbx_lattice_size_ghost[0]; _x < (rk)->dom.dbx_sub_box_lattice_size[0] + (rk)->dom.dbx_lattice_size_ghost[0]; _x++) { \
                ora_atom *at = &(rk)->atoms[idx3(rk, _x, _y, _z)];
#define END_SUBBOX }

void ora_first_step(ora_world *w) { /* newton_motion.cpp:30-55 */
    const double dt = w->dt;
    const double *dt_inv_m = w->dt_inv_m;
    FOR_RANKS(w, rk)
        FOR_SUBBOX(rk, a)
            if (a->type != ORA_INVALID) {
                a->v[0] = a->v[0] + dt_inv_m[a->type] * a->f[0];
                a->v[1] = a->v[1] + dt_inv_m[a->type] * a->f[1];
                a->v[2] = a->v[2] + dt_inv_m[a->type] * a->f[2];
                a->x[0] += dt * a->v[0];
                a->x[1] += dt * a->v[1];
                a->x[2] += dt * a->v[2];
            }
        END_SUBBOX
        for (size_t i = 0; i < rk->n_inter; i++) {
            ora_atom *a = &rk->inter[i];
            for (int d = 0; d < 3; ++d) {
                a->v[d] = a->v[d] + dt_inv_m[a->type] * a->f[d];
                a->x[d] += dt * a->v[d];
            }
        }
    END_RANKS
}

void ora_second_step(ora_world *w) { /* newton_motion.cpp:57-74 */
    const double *dt_inv_m = w->dt_inv_m;
    FOR_RANKS(w, rk)
        FOR_SUBBOX(rk, a)
            if (a->type != ORA_INVALID) {
                a->v[0] += dt_inv_m[a->type] * a->f[0];
                a->v[1] += dt_inv_m[a->type] * a->f[1];
                a->v[2] += dt_inv_m[a->type] * a->f[2];
            }
        END_SUBBOX
        for (size_t i = 0; i < rk->n_inter; i++) {
            ora_atom *a = &rk->inter[i];
            a->v[0] += dt_inv_m[a->type] * a->f[0];
            a->v[1] += dt_inv_m[a->type] * a->f[1];
            a->v[2] += dt_inv_m[a->type] * a->f[2];
        }
    END_RANKS
}

/* ---- atom::decide: reference src/atom.cpp:21-84 ---------------------------------------------- */
static int decide_rank(ora_rank *rk) {
    const ora_domain *d = &rk->dom;
    rk->n_ghost = 0; /* inter_atom_list->clearGhost() */
    int nflag = 0;
    for (long k = 0; k < d->dbx_sub_box_lattice_size[2]; k++)
        for (long j = 0; j < d->dbx_sub_box_lattice_size[1]; j++)
            for (long i = 0; i < d->dbx_sub_box_lattice_size[0]; i++) {
                ora_atom *a = &rk->atoms[idx3(rk, d->dbx_lattice_size_ghost[0] + i, d->dbx_lattice_size_ghost[1] + j,
                                              d->dbx_lattice_size_ghost[2] + k)];
                if (a->type != ORA_INVALID) {
                    const double xt = (i + d->dbx_sub_box_lattice_region.x_low) * 0.5 * d->lattice_const;
                    const double yt = (j + d->dbx_sub_box_lattice_region.y_low + (i % 2) * 0.5) * d->lattice_const;
                    const double zt = (k + d->dbx_sub_box_lattice_region.z_low + (i % 2) * 0.5) * d->lattice_const;
                    double dist = (a->x[0] - xt) * (a->x[0] - xt);
                    dist += (a->x[1] - yt) * (a->x[1] - yt);
                    dist += (a->x[2] - zt) * (a->x[2] - zt);
                    if (dist > (pow(0.2 * d->lattice_const, 2.0))) {
                        inter_push(&rk->inter, &rk->n_inter, &rk->cap_inter, a);
                        a->type = ORA_INVALID;
                        a->v[0] = 0; a->v[1] = 0; a->v[2] = 0;
                        nflag = 1;
                    }
                }
            }
    for (size_t it = 0; it < rk->n_inter;) {
        ora_atom *in = &rk->inter[it];
        const long near_idx = ora_find_near_lat_index_in_sub_box(rk, in);
        ora_atom *near = near_idx == ORA_INDEX_NOT_EXISTS ? NULL : &rk->atoms[near_idx];
        if (near != NULL && near->type == ORA_INVALID && ora_is_out_box(near, d) == ORA_IN_BOX) {
            near->id = in->id;
            near->type = in->type;
            near->x[0] = in->x[0]; near->x[1] = in->x[1]; near->x[2] = in->x[2];
            near->v[0] = in->v[0]; near->v[1] = in->v[1]; near->v[2] = in->v[2];
            inter_erase(rk, it);
        } else {
            it++;
        }
    }
    return nflag;
}

int ora_decide(ora_world *w) {
    int flag = 0;
    for (int r = 0; r < w->n_ranks; r++) flag |= decide_rank(&w->ranks[r]);
    return flag;
}

/* ---- atom::clearForce: reference src/atom.cpp:86-100 ------------------------------------------ */
void ora_clear_force(ora_world *w) {
    FOR_RANKS(w, rk)
        for (long i = 0; i < rk->size; i++) {
            ora_atom *a = &rk->atoms[i];
            a->f[0] = 0; a->f[1] = 0; a->f[2] = 0; a->rho = 0;
        }
        for (size_t i = 0; i < rk->n_inter; i++) {
            ora_atom *a = &rk->inter[i];
            a->f[0] = 0; a->f[1] = 0; a->f[2] = 0; a->rho = 0;
        }
    END_RANKS
}

/* ---- exchange engine: libcomm comm::neiSendReceive<T, reverse> restated ----
 * per dimension stage (x,y,z or reversed): every rank packs LOWER and HIGHER, messages travel to
 * rank_id_neighbours[dim][dir], each rank then unpacks the message that arrived from neighbour
 * [dim][(dir+1)%2] with onReceive(.., dim, dir) -- "mirror with send", lat_particle_packer.cpp:66. */
typedef struct packer {
    size_t elem;
    int reverse;
    size_t (*send_len)(ora_world *, ora_rank *, int, int);
    void (*on_send)(ora_world *, ora_rank *, void *, size_t, int, int);
    void (*on_recv)(ora_world *, ora_rank *, const void *, size_t, int, int);
} packer;

static void nei_send_receive(ora_world *w, const packer *pk) {
    const int nr = w->n_ranks;
    void **buf = (void **)calloc((size_t)nr * 2, sizeof(void *));
    size_t *cnt = (size_t *)calloc((size_t)nr * 2, sizeof(size_t));
    for (int s = 0; s < 3; s++) {
        const int dim = pk->reverse ? 2 - s : s;
#pragma omp parallel for schedule(static) num_threads(w->n_threads)
        for (int r = 0; r < nr; r++)
            for (int dir = 0; dir < 2; dir++) {
                const size_t n = pk->send_len(w, &w->ranks[r], dim, dir);
                cnt[2 * r + dir] = n;
                buf[2 * r + dir] = malloc(n * pk->elem + 8);
                pk->on_send(w, &w->ranks[r], buf[2 * r + dir], n, dim, dir);
            }
#pragma omp parallel for schedule(static) num_threads(w->n_threads)
        for (int r = 0; r < nr; r++)
            for (int dir = 0; dir < 2; dir++) {
                const int src = w->ranks[r].dom.rank_id_neighbours[dim][(dir + 1) % 2];
                pk->on_recv(w, &w->ranks[r], buf[2 * src + dir], cnt[2 * src + dir], dim, dir);
            }
        for (int i = 0; i < 2 * nr; i++) { free(buf[i]); buf[i] = NULL; }
    }
    free(buf);
    free(cnt);
}

/* reference src/pack/lat_particle_data.h:11-20 (32 B) and src/pack/particledata.h:11-22 (64 B) */
typedef struct { int type; int _pad; double r[3]; } lat_particle_data;
typedef struct { unsigned long id; int type; int _pad; double r[3]; double v[3]; } particle_data;

static void periodic_offset(const ora_domain *d, double off[3], int dim, int dir) { /* lat_particle_packer.cpp:22-32 */
    if (d->grid_coord[dim] == 0 && dir == ORA_DIR_LOWER) off[dim] = d->meas_global_length[dim];
    if (d->grid_coord[dim] == d->grid_size[dim] - 1 && dir == ORA_DIR_HIGHER) off[dim] = -((d->meas_global_length[dim]));
}

/* LatParticlePacker / LatPackerFirst / LatPacker: reference src/pack/lat_particle_packer.cpp:17-193 */
static size_t lat_send_len(ora_world *w, ora_rank *rk, int dim, int dir) { (void)w; return rk->sendlist[2 * dim + dir].n; }
static void lat_on_send(ora_world *w, ora_rank *rk, void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    lat_particle_data *buf = (lat_particle_data *)vbuf;
    double off[3] = {0.0, 0.0, 0.0};
    periodic_offset(&rk->dom, off, dim, dir);
    const ora_ivec *sl = &rk->sendlist[2 * dim + dir];
    for (size_t i = 0; i < n; i++) {
        const ora_atom *a = &rk->atoms[sl->v[i]];
        buf[i].type = a->type;
        buf[i]._pad = 0;
        buf[i].r[0] = a->x[0] + off[0];
        buf[i].r[1] = a->x[1] + off[1];
        buf[i].r[2] = a->x[2] + off[2];
    }
}
static void lat_first_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    const lat_particle_data *buf = (const lat_particle_data *)vbuf;
    const int *ghost = rk->dom.dbx_lattice_size_ghost, *box = rk->dom.dbx_sub_box_lattice_size, *ext = rk->dom.dbx_ghost_extended_lattice_size;
    int lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        if (k == dim) {
            if (dir == 0) { lo[k] = ghost[k] + box[k]; hi[k] = ext[k]; } /* mirror with send */
            else { lo[k] = 0; hi[k] = ghost[k]; }
        } else if (k < dim) { lo[k] = 0; hi[k] = ext[k]; }
        else { lo[k] = ghost[k]; hi[k] = ghost[k] + box[k]; }
    }
    ora_ivec *rl = &rk->recvlist[2 * dim + dir];
    size_t m = 0;
    for (int k = lo[2]; k < hi[2]; k++)
        for (int j = lo[1]; j < hi[1]; j++)
            for (int i = lo[0]; i < hi[0]; i++) {
                ora_atom *a = &rk->atoms[idx3(rk, i, j, k)];
                a->type = buf[m].type;
                a->x[0] = buf[m].r[0];
                a->x[1] = buf[m].r[1];
                a->x[2] = buf[m++].r[2];
                ivec_push(rl, idx3(rk, i, j, k));
            }
    if (n != rl->n) fprintf(stderr, "unpack_recvfirst: received data size does not match, expected %zu, but got %zu.\n", n, rl->n);
}
static void lat_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    const lat_particle_data *buf = (const lat_particle_data *)vbuf;
    const ora_ivec *rl = &rk->recvlist[2 * dim + dir];
    for (size_t i = 0; i < n; i++) {
        ora_atom *a = &rk->atoms[rl->v[i]];
        a->type = buf[i].type;
        a->x[0] = buf[i].r[0];
        a->x[1] = buf[i].r[1];
        a->x[2] = buf[i].r[2];
    }
}

void ora_exchange_atom_first(ora_world *w) { /* reference src/atom/atom_list.cpp:25-49 */
    for (int r = 0; r < w->n_ranks; r++) {
        ora_rank *rk = &w->ranks[r];
        for (int d = 0; d < 3; d++)
            for (int dir = ORA_DIR_LOWER; dir <= ORA_DIR_HIGHER; dir++) {
                ora_ivec *sl = &rk->sendlist[2 * d + dir];
                ivec_clear(sl);
                ivec_clear(&rk->recvlist[2 * d + dir]);
                const ora_iregion rg = ora_fw_comm_local_region(&rk->dom, d, dir);
                for (int iz = rg.z_low; iz < rg.z_high; iz++)
                    for (int iy = rg.y_low; iy < rg.y_high; iy++)
                        for (int ix = rg.x_low; ix < rg.x_high; ix++) ivec_push(sl, idx3(rk, ix, iy, iz));
            }
    }
    const packer pk = {sizeof(lat_particle_data), 0, lat_send_len, lat_on_send, lat_first_on_recv};
    nei_send_receive(w, &pk);
}

void ora_exchange_atom(ora_world *w) { /* reference src/atom/atom_list.cpp:51-57 */
    const packer pk = {sizeof(lat_particle_data), 0, lat_send_len, lat_on_send, lat_on_recv};
    nei_send_receive(w, &pk);
}

/* RhoPacker: reference src/pack/rho_packer.cpp:13-46 (reverse: ghosts -> owners, ADD) */
static size_t rho_send_len(ora_world *w, ora_rank *rk, int dim, int dir) {
    (void)w;
    return rk->recvlist[2 * dim + (dir == ORA_DIR_LOWER ? 1 : 0)].n;
}
static void rho_on_send(ora_world *w, ora_rank *rk, void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    double *buf = (double *)vbuf;
    const ora_ivec *rl = &rk->recvlist[2 * dim + (dir == ORA_DIR_LOWER ? 1 : 0)];
    for (size_t i = 0; i < n; i++) buf[i] = rk->atoms[rl->v[i]].rho;
}
static void rho_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w; (void)n;
    const double *buf = (const double *)vbuf;
    const ora_ivec *sl = &rk->sendlist[2 * dim + (dir == ORA_DIR_LOWER ? ORA_DIR_HIGHER : ORA_DIR_LOWER)];
    for (size_t i = 0; i < sl->n; i++) rk->atoms[sl->v[i]].rho += buf[i];
}

/* ForcePacker: reference src/pack/force_packer.cpp:11-40 (reverse, 3 doubles/site, ADD) */
static size_t force_send_len(ora_world *w, ora_rank *rk, int dim, int dir) {
    (void)w;
    return rk->recvlist[2 * dim + (dir == ORA_DIR_LOWER ? 1 : 0)].n * 3;
}
static void force_on_send(ora_world *w, ora_rank *rk, void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    double *buf = (double *)vbuf;
    const ora_ivec *rl = &rk->recvlist[2 * dim + (dir == ORA_DIR_LOWER ? 1 : 0)];
    size_t m = 0;
    for (size_t i = 0; i < n / 3; i++) {
        const ora_atom *a = &rk->atoms[rl->v[i]];
        buf[m++] = a->f[0]; buf[m++] = a->f[1]; buf[m++] = a->f[2];
    }
}
static void force_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w; (void)n;
    const double *buf = (const double *)vbuf;
    const ora_ivec *sl = &rk->sendlist[2 * dim + (dir == ORA_DIR_LOWER ? ORA_DIR_HIGHER : ORA_DIR_LOWER)];
    size_t m = 0;
    for (size_t i = 0; i < sl->n; i++) {
        ora_atom *a = &rk->atoms[sl->v[i]];
        a->f[0] += buf[m++]; a->f[1] += buf[m++]; a->f[2] += buf[m++];
    }
}

/* DfEmbedPacker: reference src/pack/df_embed_packer.cpp:17-68 (forward, ASSIGN, lattice then inter) */
static size_t df_send_len(ora_world *w, ora_rank *rk, int dim, int dir) {
    (void)w;
    return rk->sendlist[2 * dim + dir].n + rk->intersend[2 * dim + dir].n;
}
static void df_on_send(ora_world *w, ora_rank *rk, void *vbuf, size_t n, int dim, int dir) {
    (void)w; (void)n;
    double *buf = (double *)vbuf;
    const ora_ivec *sl = &rk->sendlist[2 * dim + dir], *isl = &rk->intersend[2 * dim + dir];
    size_t m = 0;
    for (size_t i = 0; i < sl->n; i++) buf[m++] = rk->atoms[sl->v[i]].df;
    for (size_t i = 0; i < isl->n; i++) buf[m++] = inter_ref(rk, isl->v[i])->df;
}
static void df_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    const double *buf = (const double *)vbuf;
    const ora_ivec *rl = &rk->recvlist[2 * dim + dir], *irl = &rk->interrecv[2 * dim + dir];
    if (n != rl->n + irl->n) {
        fprintf(stderr, "wrong number of dfembed recv!!!\n");
        abort(); /* MPI_Abort(MPI_COMM_WORLD, 2) */
    }
    size_t m = 0;
    for (size_t i = 0; i < rl->n; i++) rk->atoms[rl->v[i]].df = buf[m++];
    for (size_t i = 0; i < irl->n; i++) inter_ref(rk, irl->v[i])->df = buf[m++];
}

/* InterParticlePacker: reference src/pack/inter_particle_packer.cpp:17-121 (migration of out-of-box inter atoms) */
static const unsigned out_box_flags[3][2] = {{ORA_OUT_X_LITTER, ORA_OUT_X_BIG}, {ORA_OUT_Y_LITTER, ORA_OUT_Y_BIG}, {ORA_OUT_Z_LITTER, ORA_OUT_Z_BIG}};
static size_t interp_send_len(ora_world *w, ora_rank *rk, int dim, int dir) {
    (void)w;
    size_t n = 0;
    for (size_t i = 0; i < rk->n_inter; i++)
        if (ora_is_out_box(&rk->inter[i], &rk->dom) & out_box_flags[dim][dir]) n++;
    return n;
}
static void interp_on_send(ora_world *w, ora_rank *rk, void *vbuf, size_t n, int dim, int dir) {
    (void)w; (void)n;
    particle_data *buf = (particle_data *)vbuf;
    double off[3] = {0.0, 0.0, 0.0};
    periodic_offset(&rk->dom, off, dim, dir);
    size_t i = 0;
    for (size_t it = 0; it < rk->n_inter;) {
        const ora_atom *a = &rk->inter[it];
        if (ora_is_out_box(a, &rk->dom) & out_box_flags[dim][dir]) {
            buf[i].id = a->id;
            buf[i].type = a->type;
            buf[i]._pad = 0;
            for (int k = 0; k < 3; k++) { buf[i].r[k] = a->x[k] + off[k]; buf[i].v[k] = a->v[k]; }
            inter_erase(rk, it);
            i++;
        } else {
            it++;
        }
    }
}
static void interp_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w; (void)dim; (void)dir;
    const particle_data *buf = (const particle_data *)vbuf;
    ora_atom a;
    memset(&a, 0, sizeof a);
    for (size_t i = 0; i < n; i++) {
        a.id = buf[i].id;
        a.type = buf[i].type;
        for (int k = 0; k < 3; k++) { a.x[k] = buf[i].r[k]; a.v[k] = buf[i].v[k]; }
        inter_push(&rk->inter, &rk->n_inter, &rk->cap_inter, &a);
    }
}
void ora_exchange_inter(ora_world *w) { /* reference src/atom/inter_atom_list.cpp:19-25 */
    const packer pk = {sizeof(particle_data), 0, interp_send_len, interp_on_send, interp_on_recv};
    nei_send_receive(w, &pk);
}

/* InterBorderPacker: reference src/pack/inter_border_packer.cpp:12-106 (ghost copies of border inter atoms) */
static size_t interb_send_len(ora_world *w, ora_rank *rk, int dim, int dir) {
    (void)w;
    ora_ivec *sl = &rk->intersend[2 * dim + dir];
    const ora_iregion rg = ora_fw_comm_local_region(&rk->dom, dim, dir);
    long c[3] = {0, 0, 0};
    for (size_t i = 0; i < rk->n_inter; i++) {
        ora_near_lat_coord(&rk->inter[i], &rk->dom, c);
        if (region_is_in(&rg, c[0], c[1], c[2])) ivec_push(sl, (long)i);
    }
    for (size_t i = 0; i < rk->n_ghost; i++) {
        ora_near_lat_coord(&rk->ghost[i], &rk->dom, c);
        if (region_is_in(&rg, c[0], c[1], c[2])) ivec_push(sl, ~(long)i);
    }
    return sl->n;
}
static void interb_on_send(ora_world *w, ora_rank *rk, void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    lat_particle_data *buf = (lat_particle_data *)vbuf;
    double shift = 0.0;
    const ora_domain *d = &rk->dom;
    if (d->grid_coord[dim] == 0 && dir == ORA_DIR_LOWER) shift = d->meas_global_length[dim];
    if (d->grid_coord[dim] == d->grid_size[dim] - 1 && dir == ORA_DIR_HIGHER) shift = -((d->meas_global_length[dim]));
    const ora_ivec *sl = &rk->intersend[2 * dim + dir];
    for (size_t i = 0; i < n; i++) {
        const ora_atom *a = inter_ref(rk, sl->v[i]);
        buf[i].type = a->type;
        buf[i]._pad = 0;
        for (int k = 0; k < 3; k++) buf[i].r[k] = a->x[k] + (k == dim ? shift : 0.0);
    }
}
static void interb_on_recv(ora_world *w, ora_rank *rk, const void *vbuf, size_t n, int dim, int dir) {
    (void)w;
    const lat_particle_data *buf = (const lat_particle_data *)vbuf;
    ora_ivec *rl = &rk->interrecv[2 * dim + dir];
    ivec_clear(rl);
    ora_atom e;
    memset(&e, 0, sizeof e);
    for (size_t i = 0; i < n; i++) {
        e.type = buf[i].type;
        e.x[0] = buf[i].r[0]; e.x[1] = buf[i].r[1]; e.x[2] = buf[i].r[2];
        inter_push(&rk->ghost, &rk->n_ghost, &rk->cap_ghost, &e);
        ivec_push(rl, ~(long)(rk->n_ghost - 1));
    }
}
void ora_border_inter(ora_world *w) { /* reference src/atom/inter_atom_list.cpp:47-53 */
    for (int r = 0; r < w->n_ranks; r++)
        for (int i = 0; i < 6; i++) { ivec_clear(&w->ranks[r].intersend[i]); ivec_clear(&w->ranks[r].interrecv[i]); }
    const packer pk = {sizeof(lat_particle_data), 0, interb_send_len, interb_on_send, interb_on_recv};
    nei_send_receive(w, &pk);
}

/* ---- InterAtomList::makeIndex: reference src/atom/inter_atom_list.cpp:27-45 -------------------
 * unordered_multimap<site, AtomElement*> restated as arrays sorted by (site, insertion order). */
static int map_cmp(const void *a, const void *b) {
    const long *pa = (const long *)a, *pb = (const long *)b;
    if (pa[0] != pb[0]) return pa[0] < pb[0] ? -1 : 1;
    return pa[1] < pb[1] ? -1 : (pa[1] > pb[1] ? 1 : 0);
}
static void make_index(ora_rank *rk) {
    const size_t n = rk->n_inter + rk->n_ghost;
    if (n > rk->cap_map) {
        rk->cap_map = n * 2;
        rk->map_site = (long *)realloc(rk->map_site, rk->cap_map * sizeof(long));
        rk->map_ref = (long *)realloc(rk->map_ref, rk->cap_map * sizeof(long));
    }
    rk->n_map = n;
    if (n == 0) return;
    long *tmp = (long *)malloc(n * 3 * sizeof(long));
    long c[3];
    for (size_t i = 0; i < n; i++) {
        const ora_atom *a = i < rk->n_inter ? &rk->inter[i] : &rk->ghost[i - rk->n_inter];
        ora_near_lat_coord(a, &rk->dom, c);
        tmp[3 * i] = idx3(rk, c[0], c[1], c[2]);
        tmp[3 * i + 1] = (long)i;
        tmp[3 * i + 2] = i < rk->n_inter ? (long)i : ~(long)(i - rk->n_inter);
    }
    qsort(tmp, n, 3 * sizeof(long), map_cmp);
    for (size_t i = 0; i < n; i++) { rk->map_site[i] = tmp[3 * i]; rk->map_ref[i] = tmp[3 * i + 2]; }
    free(tmp);
}
static void map_equal_range(const ora_rank *rk, long site, size_t *first, size_t *last) {
    size_t lo = 0, hi = rk->n_map;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (rk->map_site[mid] < site) lo = mid + 1; else hi = mid; }
    *first = lo;
    while (lo < rk->n_map && rk->map_site[lo] == site) lo++;
    *last = lo;
}

/* ---- atom::latRho: reference src/atom.cpp:151-192 --------------------------------------------- */
void ora_lat_rho(ora_rank *rk, const pot_eam *pot) {
    const ora_domain *d = &rk->dom;
    const int xs = d->dbx_lattice_size_ghost[0], ys = d->dbx_lattice_size_ghost[1], zs = d->dbx_lattice_size_ghost[2];
    const double rc2 = rk->cutoff_radius * rk->cutoff_radius;
    for (int k = zs; k < d->dbx_sub_box_lattice_size[2] + zs; k++)
        for (int j = ys; j < d->dbx_sub_box_lattice_size[1] + ys; j++)
            for (int i = xs; i < d->dbx_sub_box_lattice_size[0] + xs; i++) {
                const long ci = idx3(rk, i, j, k);
                ora_atom *c = &rk->atoms[ci];
                if (c->type == ORA_INVALID) continue;
                const ora_ivec *nl = nei_list(rk, 1, i);
                for (size_t q = 0; q < nl->n; q++) {
                    ora_atom *n = &rk->atoms[ci + nl->v[q]];
                    if (n->type == ORA_INVALID) continue;
                    const double delx = c->x[0] - n->x[0], dely = c->x[1] - n->x[1], delz = c->x[2] - n->x[2];
                    const double dist2 = delx * delx + dely * dely + delz * delz;
                    if (dist2 < rc2) {
                        c->rho += pot_charge_density(pot, type_key(n->type), dist2);
                        n->rho += pot_charge_density(pot, type_key(c->type), dist2);
                    }
                }
            }
}

/* ---- atom::interRho: reference src/atom.cpp:194-284 ------------------------------------------- */
static void inter_rho(ora_rank *rk, const pot_eam *pot) {
    const double rc2 = rk->cutoff_radius * rk->cutoff_radius;
    for (size_t it = 0; it < rk->n_inter; it++) {
        ora_atom *in = &rk->inter[it];
        const long near_idx = ora_find_near_lat_index_in_sub_box(rk, in);
        if (near_idx == ORA_INDEX_NOT_EXISTS) continue;
        ora_atom *near = &rk->atoms[near_idx];
        double delx = in->x[0] - near->x[0], dely = in->x[1] - near->x[1], delz = in->x[2] - near->x[2];
        double dist2 = delx * delx + dely * dely + delz * delz;
        if (near->type != ORA_INVALID && dist2 < rc2) {
            in->rho += pot_charge_density(pot, type_key(near->type), dist2);
            near->rho += pot_charge_density(pot, type_key(in->type), dist2);
        }
        long x, y, z;
        idx_to_3d(rk, near_idx, &x, &y, &z);
        const ora_ivec *nl = nei_list(rk, 0, x);
        for (size_t q = 0; q < nl->n; q++) {
            ora_atom *ln = &rk->atoms[near_idx + nl->v[q]];
            if (ln->type != ORA_INVALID) {
                delx = in->x[0] - ln->x[0]; dely = in->x[1] - ln->x[1]; delz = in->x[2] - ln->x[2];
                dist2 = delx * delx + dely * dely + delz * delz;
                if (dist2 < rc2) {
                    in->rho += pot_charge_density(pot, type_key(ln->type), dist2);
                    ln->rho += pot_charge_density(pot, type_key(in->type), dist2);
                }
            }
        }
        size_t b0, b1;
        map_equal_range(rk, near_idx, &b0, &b1);
        for (size_t b = b0; b < b1; b++) {
            const ora_atom *o = inter_ref(rk, rk->map_ref[b]);
            if (o->id == in->id) continue;
            delx = in->x[0] - o->x[0]; dely = in->x[1] - o->x[1]; delz = in->x[2] - o->x[2];
            dist2 = delx * delx + dely * dely + delz * delz;
            if (dist2 < rc2) in->rho += pot_charge_density(pot, type_key(o->type), dist2);
        }
        for (size_t q = 0; q < nl->n; q++) {
            map_equal_range(rk, near_idx + nl->v[q], &b0, &b1);
            for (size_t b = b0; b < b1; b++) {
                const ora_atom *o = inter_ref(rk, rk->map_ref[b]);
                delx = in->x[0] - o->x[0]; dely = in->x[1] - o->x[1]; delz = in->x[2] - o->x[2];
                dist2 = delx * delx + dely * dely + delz * delz;
                if (dist2 < rc2) in->rho += pot_charge_density(pot, type_key(o->type), dist2);
            }
        }
        in->df = pot_d_embed_energy(pot, type_key(in->type), in->rho);
    }
}

/* ---- atom::latDf: reference src/atom.cpp:286-309 ---------------------------------------------- */
void ora_lat_df(ora_rank *rk, const pot_eam *pot) {
    FOR_SUBBOX(rk, a)
        if (a->type == ORA_INVALID) continue;
        a->df = pot_d_embed_energy(pot, type_key(a->type), a->rho);
    END_SUBBOX
}

/* ---- atom::latForce: reference src/atom.cpp:311-358 ------------------------------------------- */
void ora_lat_force(ora_rank *rk, const pot_eam *pot) {
    const ora_domain *d = &rk->dom;
    const int xs = d->dbx_lattice_size_ghost[0], ys = d->dbx_lattice_size_ghost[1], zs = d->dbx_lattice_size_ghost[2];
    const double rc2 = rk->cutoff_radius * rk->cutoff_radius;
    for (int k = zs; k < d->dbx_sub_box_lattice_size[2] + zs; k++)
        for (int j = ys; j < d->dbx_sub_box_lattice_size[1] + ys; j++)
            for (int i = xs; i < d->dbx_sub_box_lattice_size[0] + xs; i++) {
                const long ci = idx3(rk, i, j, k);
                ora_atom *c = &rk->atoms[ci];
                if (c->type == ORA_INVALID) continue;
                const ora_ivec *nl = nei_list(rk, 1, i);
                for (size_t q = 0; q < nl->n; q++) {
                    ora_atom *n = &rk->atoms[ci + nl->v[q]];
                    const double delx = c->x[0] - n->x[0], dely = c->x[1] - n->x[1], delz = c->x[2] - n->x[2];
                    const double dist2 = delx * delx + dely * dely + delz * delz;
                    if (dist2 < rc2 && n->type != ORA_INVALID) {
                        const double fpair = pot_to_force(pot, type_key(c->type), type_key(n->type), dist2, c->df, n->df);
                        c->f[0] += delx * fpair; c->f[1] += dely * fpair; c->f[2] += delz * fpair;
                        n->f[0] -= delx * fpair; n->f[1] -= dely * fpair; n->f[2] -= delz * fpair;
                    }
                }
            }
}

/* ---- atom::interForce: reference src/atom.cpp:360-473 ----------------------------------------- */
static void inter_force(ora_rank *rk, const pot_eam *pot) {
    const double rc2 = rk->cutoff_radius * rk->cutoff_radius;
    for (size_t it = 0; it < rk->n_inter; it++) {
        ora_atom *in = &rk->inter[it];
        const long near_idx = ora_find_near_lat_index_in_sub_box(rk, in);
        if (near_idx == ORA_INDEX_NOT_EXISTS) { assert(0); continue; }
        ora_atom *c = &rk->atoms[near_idx];
        double delx = in->x[0] - c->x[0], dely = in->x[1] - c->x[1], delz = in->x[2] - c->x[2];
        double dist2 = delx * delx + dely * dely + delz * delz;
        double fpair;
        if (dist2 < rc2 && c->type != ORA_INVALID) {
            fpair = pot_to_force(pot, type_key(in->type), type_key(c->type), dist2, in->df, c->df);
            in->f[0] += delx * fpair; in->f[1] += dely * fpair; in->f[2] += delz * fpair;
            c->f[0] -= delx * fpair; c->f[1] -= dely * fpair; c->f[2] -= delz * fpair;
        }
        long x, y, z;
        idx_to_3d(rk, near_idx, &x, &y, &z);
        const ora_ivec *nl = nei_list(rk, 0, x);
        for (size_t q = 0; q < nl->n; q++) {
            ora_atom *ln = &rk->atoms[near_idx + nl->v[q]];
            delx = in->x[0] - ln->x[0]; dely = in->x[1] - ln->x[1]; delz = in->x[2] - ln->x[2];
            dist2 = delx * delx + dely * dely + delz * delz;
            if (dist2 < rc2 && ln->type != ORA_INVALID) {
                fpair = pot_to_force(pot, type_key(in->type), type_key(ln->type), dist2, in->df, ln->df);
                in->f[0] += delx * fpair; in->f[1] += dely * fpair; in->f[2] += delz * fpair;
                ln->f[0] -= delx * fpair; ln->f[1] -= dely * fpair; ln->f[2] -= delz * fpair;
            }
        }
        size_t b0, b1;
        map_equal_range(rk, near_idx, &b0, &b1);
        for (size_t b = b0; b < b1; b++) {
            const ora_atom *o = inter_ref(rk, rk->map_ref[b]);
            if (o->id == in->id) continue;
            delx = in->x[0] - o->x[0]; dely = in->x[1] - o->x[1]; delz = in->x[2] - o->x[2];
            dist2 = delx * delx + dely * dely + delz * delz;
            if (dist2 < rc2) {
                fpair = pot_to_force(pot, type_key(in->type), type_key(o->type), dist2, in->df, o->df);
                in->f[0] += delx * fpair; in->f[1] += dely * fpair; in->f[2] += delz * fpair;
            }
        }
        for (size_t q = 0; q < nl->n; q++) {
            map_equal_range(rk, near_idx + nl->v[q], &b0, &b1);
            for (size_t b = b0; b < b1; b++) {
                const ora_atom *o = inter_ref(rk, rk->map_ref[b]);
                delx = in->x[0] - o->x[0]; dely = in->x[1] - o->x[1]; delz = in->x[2] - o->x[2];
                dist2 = delx * delx + dely * dely + delz * delz;
                if (dist2 < rc2) {
                    fpair = pot_to_force(pot, type_key(in->type), type_key(o->type), dist2, in->df, o->df);
                    in->f[0] += delx * fpair; in->f[1] += dely * fpair; in->f[2] += delz * fpair;
                }
            }
        }
    }
}

/* ---- atom::computeEam: reference src/atom.cpp:102-149 ----------------------------------------- */
void ora_compute_eam(ora_world *w) {
    const pot_eam *pot = w->pot;
    FOR_RANKS(w, rk)
        make_index(rk);
        ora_lat_rho(rk, pot);
        inter_rho(rk, pot);
    END_RANKS
    { const packer pk = {sizeof(double), 1, rho_send_len, rho_on_send, rho_on_recv}; nei_send_receive(w, &pk); }
    FOR_RANKS(w, rk)
        ora_lat_df(rk, pot);
    END_RANKS
    { const packer pk = {sizeof(double), 0, df_send_len, df_on_send, df_on_recv}; nei_send_receive(w, &pk); }
    FOR_RANKS(w, rk)
        ora_lat_force(rk, pot);
        inter_force(rk, pot);
    END_RANKS
    { const packer pk = {sizeof(double), 1, force_send_len, force_on_send, force_on_recv}; nei_send_receive(w, &pk); }
}

/* ---- atom::setv: reference src/atom.cpp:475-494 ----------------------------------------------- */
void ora_setv(ora_world *w, const int lat[4], const double direction[3], double energy) {
    for (int r = 0; r < w->n_ranks; r++) {
        ora_rank *rk = &w->ranks[r];
        const ora_domain *d = &rk->dom;
        if ((lat[0] * 2) >= d->dbx_sub_box_lattice_region.x_low &&
            (lat[0] * 2) < (d->dbx_sub_box_lattice_region.x_low + d->dbx_sub_box_lattice_size[0]) &&
            lat[1] >= d->dbx_sub_box_lattice_region.y_low &&
            lat[1] < (d->dbx_sub_box_lattice_region.y_low + d->dbx_sub_box_lattice_size[1]) &&
            lat[2] >= d->dbx_sub_box_lattice_region.z_low &&
            lat[2] < (d->dbx_sub_box_lattice_region.z_low + d->dbx_sub_box_lattice_size[2])) {
            const long kk = idx3(rk, lat[0] * 2 - d->dbx_ghost_ext_lattice_region.x_low, lat[1] - d->dbx_ghost_ext_lattice_region.y_low,
                                 lat[2] - d->dbx_ghost_ext_lattice_region.z_low) + lat[3];
            ora_atom *a = &rk->atoms[kk];
            const double v_ = sqrt(2 * energy / atom_mass(a->type) / mvv2e);
            const double d_ = sqrt(direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2]);
            a->v[0] += v_ * direction[0] / d_;
            a->v[1] += v_ * direction[1] / d_;
            a->v[2] += v_ * direction[2] / d_;
        }
    }
}

/* ---- driver pieces: reference src/simulation.cpp:137-145,164-194,208-217 ---------------------- */
void ora_prepare(ora_world *w) {
    ora_exchange_atom_first(w);
    ora_clear_force(w);
    ora_compute_eam(w);
}

void ora_step(ora_world *w) {
    ora_first_step(w);
    ora_decide(w);
    ora_exchange_inter(w);
    ora_border_inter(w);
    ora_exchange_atom(w);
    ora_clear_force(w);
    ora_compute_eam(w);
    ora_second_step(w);
}

void ora_collision_step(ora_world *w, const int lat[4], const double direction[3], double energy) {
    ora_setv(w, lat, direction, energy);
    ora_exchange_inter(w);
    ora_border_inter(w);
    ora_exchange_atom(w);
    ora_clear_force(w);
    ora_compute_eam(w);
}

/* ---- diagnostics: reference src/system_configuration.cpp:45-111 ------------------------------- */
double ora_mvv(const ora_world *w) {
    double e = 0.0;
    for (int r = 0; r < w->n_ranks; r++) {
        const ora_rank *rk = &w->ranks[r];
        double er = 0.0;
        FOR_SUBBOX(rk, a)
            if (a->type != ORA_INVALID) er += (a->v[0] * a->v[0] + a->v[1] * a->v[1] + a->v[2] * a->v[2]) * atom_mass(a->type);
        END_SUBBOX
        for (size_t i = 0; i < rk->n_inter; i++) {
            const ora_atom *a = &rk->inter[i];
            er += (a->v[0] * a->v[0] + a->v[1] * a->v[1] + a->v[2] * a->v[2]) * atom_mass(a->type);
        }
        e += er;
    }
    return e;
}
double ora_kinetic_energy(const ora_world *w) { return 0.5 * ora_mvv(w) * mvv2e; }
static double n_atoms_global(const ora_world *w) {
    const ora_domain *d = &w->ranks[0].dom;
    return 2.0 * (double)d->phase_space[0] * (double)d->phase_space[1] * (double)d->phase_space[2];
}
double ora_temperature(const ora_world *w) {
    const double dof = 3 * n_atoms_global(w) - 3;
    return ora_mvv(w) * mvv2e / (dof * BOLTZ);
}
void ora_rescale(ora_world *w, double T) {
    const double scalar = ora_temperature(w);
    const double rescale_factor = sqrt(T / scalar);
    for (int r = 0; r < w->n_ranks; r++) {
        ora_rank *rk = &w->ranks[r];
        FOR_SUBBOX(rk, a)
            a->v[0] *= rescale_factor; a->v[1] *= rescale_factor; a->v[2] *= rescale_factor;
        END_SUBBOX
        for (size_t i = 0; i < rk->n_inter; i++) {
            ora_atom *a = &rk->inter[i];
            a->v[0] *= rescale_factor; a->v[1] *= rescale_factor; a->v[2] *= rescale_factor;
        }
    }
}

/* NOT in the reference (it never evaluates potential energy): E_pot = sum_i F(rho_i) + sum_pairs phi.
 * Every lattice pair is visited once by the half list (ghost partners included), inter-lattice pairs once
 * from the inter atom, inter-inter pairs twice (one-sided loops) hence the 1/2. rho must be current. */
double ora_potential_energy(ora_world *w) {
    const pot_eam *pot = w->pot;
    double total = 0.0;
    for (int r = 0; r < w->n_ranks; r++) {
        ora_rank *rk = &w->ranks[r];
        const ora_domain *d = &rk->dom;
        const double rc2 = rk->cutoff_radius * rk->cutoff_radius;
        const int xs = d->dbx_lattice_size_ghost[0], ys = d->dbx_lattice_size_ghost[1], zs = d->dbx_lattice_size_ghost[2];
        double e = 0.0;
        make_index(rk);
        for (int k = zs; k < d->dbx_sub_box_lattice_size[2] + zs; k++)
            for (int j = ys; j < d->dbx_sub_box_lattice_size[1] + ys; j++)
                for (int i = xs; i < d->dbx_sub_box_lattice_size[0] + xs; i++) {
                    const long ci = idx3(rk, i, j, k);
                    const ora_atom *c = &rk->atoms[ci];
                    if (c->type == ORA_INVALID) continue;
                    e += pot_embed_energy(pot, type_key(c->type), c->rho);
                    const ora_ivec *nl = nei_list(rk, 1, i);
                    for (size_t q = 0; q < nl->n; q++) {
                        const ora_atom *n = &rk->atoms[ci + nl->v[q]];
                        if (n->type == ORA_INVALID) continue;
                        const double delx = c->x[0] - n->x[0], dely = c->x[1] - n->x[1], delz = c->x[2] - n->x[2];
                        const double dist2 = delx * delx + dely * dely + delz * delz;
                        if (dist2 < rc2) e += pot_pair_energy(pot, type_key(c->type), type_key(n->type), dist2);
                    }
                }
        for (size_t it = 0; it < rk->n_inter; it++) {
            const ora_atom *in = &rk->inter[it];
            e += pot_embed_energy(pot, type_key(in->type), in->rho);
            const long near_idx = ora_find_near_lat_index_in_sub_box(rk, in);
            if (near_idx == ORA_INDEX_NOT_EXISTS) continue;
            long x, y, z;
            idx_to_3d(rk, near_idx, &x, &y, &z);
            const ora_ivec *nl = nei_list(rk, 0, x);
            for (size_t q = 0; q <= nl->n; q++) {
                const long site = q < nl->n ? near_idx + nl->v[q] : near_idx;
                const ora_atom *ln = &rk->atoms[site];
                if (ln->type != ORA_INVALID) {
                    const double delx = in->x[0] - ln->x[0], dely = in->x[1] - ln->x[1], delz = in->x[2] - ln->x[2];
                    const double dist2 = delx * delx + dely * dely + delz * delz;
                    if (dist2 < rc2) e += pot_pair_energy(pot, type_key(in->type), type_key(ln->type), dist2);
                }
                size_t b0, b1;
                map_equal_range(rk, site, &b0, &b1);
                for (size_t b = b0; b < b1; b++) {
                    const ora_atom *o = inter_ref(rk, rk->map_ref[b]);
                    if (q == nl->n && o->id == in->id) continue;
                    const double delx = in->x[0] - o->x[0], dely = in->x[1] - o->x[1], delz = in->x[2] - o->x[2];
                    const double dist2 = delx * delx + dely * dely + delz * delz;
                    if (dist2 < rc2) e += 0.5 * pot_pair_energy(pot, type_key(in->type), type_key(o->type), dist2);
                }
            }
        }
        total += e;
    }
    return total;
}

/* ---- dump record stream ---------------------------------------------------------------------------- */
/* BufferedFileWriter::write, reference frontend/io/buffered_io.cpp:18-36 (the buffering itself does not change
 * the byte stream: full buffers and the final flush are written back to back) */
static void dump_one(const ora_atom *a, size_t time_step, ora_dump_record *out, size_t cap, size_t *n) {
    if (*n < cap) {
        ora_dump_record *r = out + *n;
        memset(r, 0, sizeof *r);
        r->id = a->id;
        r->step = time_step;
        r->type = a->type;
        r->inter_type = 0; /* normal */
        for (int d = 0; d < 3; d++) {
            r->atom_location[d] = a->x[d];
            r->atom_velocity[d] = a->v[d];
        }
    }
    (*n)++;
}

/* AtomDump::dump, reference frontend/io/atom_dump.cpp:39-75 */
size_t ora_dump(const ora_rank *rk, size_t time_step, ora_dump_record *out, size_t cap) {
    const ora_domain *d = &rk->dom;
    size_t n = 0;
    /* region of OutputBaseInterface, reference frontend/io/output_base_interface.h:26-31 */
    const int b0 = d->dbx_sub_box_lattice_region.x_low - d->dbx_ghost_ext_lattice_region.x_low;
    const int b1 = d->dbx_sub_box_lattice_region.y_low - d->dbx_ghost_ext_lattice_region.y_low;
    const int b2 = d->dbx_sub_box_lattice_region.z_low - d->dbx_ghost_ext_lattice_region.z_low;
    const int e0 = b0 + d->dbx_sub_box_lattice_size[0], e1 = b1 + d->dbx_sub_box_lattice_size[1], e2 = b2 + d->dbx_sub_box_lattice_size[2];
    for (size_t i = 0; i < rk->n_inter; i++) /* :59-61 inter atoms first, list order */
        dump_one(&rk->inter[i], time_step, out, cap, &n);
    for (int k = b2; k < e2; k++)           /* :63-72 */
        for (int j = b1; j < e1; j++)
            for (int i = b0; i < e0; i++) {
                const ora_atom *a = &rk->atoms[((long)k * rk->size_y + j) * rk->size_x + i]; /* getAtomEleByGhostIndex */
                if (a->type == ORA_INVALID) continue;
                dump_one(a, time_step, out, cap, &n);
            }
    return n;
}
