// arch_cuda/cuda_hooks.cpp -- the reference-side binding: MISA-MD's eight accelerator hooks for ARCH_NAME=cuda
// (reference src/arch/arch_imp.h:17-31, name-mangled by src/arch/arch_macros.h:10-11, dispatched from
// src/arch/hardware_accelerate.hpp:17-69 and src/arch/arch_env.hpp:15-25), implemented as a thin C++ shim that
// flattens the reference's C++ types to plain pointers and sizes and calls the C ABI of libmisa_b200.so
// (include/misa_b200.h). Build contract: target `md_arch_cuda` (reference src/arch/arch_libs.cmake:24-37), see
// arch_cuda/CMakeLists.txt and INTEGRATION.md.
//
// Error convention: the hooks return void; like the reference (src/simulation.cpp:100-102) a fatal error is
// logged and the run aborted through MPI_Abort.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <mpi.h>
#include <comm/domain/bcc_domain.h>
#include <eam.h>

#include "arch/arch_imp.h"          // declares the cuda_* hooks (needs ACCELERATE_ENABLED, ARCH_NAME=cuda)
#include "atom/atom_element.h"
#include "atom/neighbour_index.h"
#include "types/atom_types.h"

#include "misa_b200.h"
#include "pot_tables.h"

static misa_b200_ctx *g_ctx = nullptr;
static void *g_registered = nullptr;

static void check(int rc, const char *what) {
    if (rc != MISA_B200_OK) {
        fprintf(stderr, "[arch_cuda] %s failed (%d): %s\n", what, rc, misa_b200_last_error());
        MPI_Abort(MPI_COMM_WORLD, 1);
    }
}

// One MPI rank <-> one GPU: the device is this rank's index among the ranks of its NODE (shared-memory split of
// MPI_COMM_WORLD), modulo the visible devices -- works under mpirun / srun / torchrun alike, no launcher variable needed.
void cuda_env_init() {
    int local = -1, inited = 0;
    MPI_Initialized(&inited);
    if (inited) {
        MPI_Comm node;
        if (MPI_Comm_split_type(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, &node) == MPI_SUCCESS) {
            MPI_Comm_rank(node, &local);
            MPI_Comm_free(&node);
        }
    }
    const int n = misa_b200_device_count();
    check(misa_b200_env_init(local >= 0 && n > 0 ? local % n : -1), "cuda_env_init");
}

void cuda_env_clean() {
    if (g_registered) { misa_b200_host_unregister(g_registered); g_registered = nullptr; }
    if (g_ctx) { misa_b200_destroy(g_ctx); g_ctx = nullptr; }
    misa_b200_env_clean();
}

void cuda_domain_init(const comm::BccDomain *domain) {
    misa_b200_domain d;
    for (int k = 0; k < 3; k++) {
        d.phase_space[k] = domain->phase_space[k];
        d.grid_size[k] = domain->grid_size[k];
        d.grid_coord[k] = domain->grid_coord[k];
        d.sub_box_lattice_size[k] = domain->sub_box_lattice_size[k];
        d.lattice_size_ghost[k] = domain->lattice_size_ghost[k];
        d.sub_box_lattice_low[k] = domain->sub_box_lattice_region.low[k];
        d.rank_id_neighbours[k][0] = domain->rank_id_neighbours[k][0];
        d.rank_id_neighbours[k][1] = domain->rank_id_neighbours[k][1];
        d.meas_global_length[k] = domain->meas_global_length[k];
    }
    d.rank = 0; // only used by the resident multi-GPU mode; the compat hooks never exchange
    d.lattice_const = domain->lattice_const;
    d.cutoff_radius_factor = domain->cutoff_radius_factor;
    check(misa_b200_create(&d, &g_ctx), "cuda_domain_init");
}

// friend of NeighbourIndex<T> (reference src/atom/neighbour_index.h:22-25): reads the four protected vectors
void cuda_nei_offset_init(const NeighbourIndex<AtomElement> *nei) {
    std::vector<int64_t> v[4];
    const std::vector<NeiOffset> *src[4] = {&nei->nei_even_offsets, &nei->nei_odd_offsets, &nei->nei_half_even_offsets,
                                            &nei->nei_half_odd_offsets};
    for (int i = 0; i < 4; i++) v[i].assign(src[i]->begin(), src[i]->end());
    check(misa_b200_set_neighbour_offsets(g_ctx, v[0].data(), v[0].size(), v[1].data(), v[1].size(), v[2].data(), v[2].size(),
                                          v[3].data(), v[3].size()),
          "cuda_nei_offset_init");
}

void cuda_pot_init(eam *pot) {
    // species in atom_type enum order (Fe, Cu, Ni) -> the potential's keys (reference src/types/atom_types.h:61-73)
    const int n_types = atom_type::num_atom_types;
    std::vector<misa_b200_table> elec(n_types), embed(n_types), phi(n_types * n_types);
    for (int i = 0; i < n_types; i++) {
        const unsigned short ki = atom_type::getTypeIdByType(atom_type::getAtomTypeByNum(i));
        if (!arch_cuda::electron_density_table(pot, ki, &elec[i]) || !arch_cuda::embedded_table(pot, ki, &embed[i])) {
            fprintf(stderr, "[arch_cuda] potential has no tables for element key %d\n", (int)ki);
            MPI_Abort(MPI_COMM_WORLD, 1);
        }
        for (int j = 0; j < n_types; j++) {
            const unsigned short kj = atom_type::getTypeIdByType(atom_type::getAtomTypeByNum(j));
            if (!arch_cuda::pair_table(pot, ki, kj, &phi[i * n_types + j])) {
                fprintf(stderr, "[arch_cuda] potential has no pair table for keys %d-%d\n", (int)ki, (int)kj);
                MPI_Abort(MPI_COMM_WORLD, 1);
            }
        }
    }
    check(misa_b200_set_potential(g_ctx, n_types, elec.data(), embed.data(), phi.data()), "cuda_pot_init");
}

// The AoS array lives for the whole run at a fixed address (reference src/atom/atom_list.cpp:12-19):
// page-lock it on first sight so the per-call transfers run at full PCIe speed.
static void pin_once(AtomElement *atoms) {
    if (g_registered == atoms) return;
    if (g_registered) misa_b200_host_unregister(g_registered);
    size_t n_sites = 0;
    misa_b200_site_count(g_ctx, &n_sites);
    g_registered = misa_b200_host_register(atoms, n_sites * sizeof(AtomElement)) == MISA_B200_OK ? atoms : nullptr;
}

void cuda_eam_rho_calc(eam *, AtomElement *atoms, const double cutoff_radius) {
    pin_once(atoms);
    check(misa_b200_eam_rho_calc(g_ctx, atoms, cutoff_radius), "cuda_eam_rho_calc");
}

void cuda_eam_df_calc(eam *, AtomElement *atoms, const double cutoff_radius) {
    check(misa_b200_eam_df_calc(g_ctx, atoms, cutoff_radius), "cuda_eam_df_calc");
}

void cuda_eam_force_calc(eam *, AtomElement *atoms, const double cutoff_radius) {
    check(misa_b200_eam_force_calc(g_ctx, atoms, cutoff_radius), "cuda_eam_force_calc");
}
