/*
 * misa_b200.h -- C ABI of the B200-native EAM hot path for MISA-MD.
 *
 * This is the drop-in boundary. The reference defines its accelerator plug-in as
 * eight C++ hook functions `${ARCH_NAME}_*` (reference src/arch/arch_imp.h:17-31, name-mangled through
 * src/arch/arch_macros.h:10-11, dispatched from src/arch/hardware_accelerate.hpp:17-69 and
 * src/arch/arch_env.hpp:15-25). Those hooks take C++ types (comm::BccDomain*, NeighbourIndex<AtomElement>*,
 * eam*, AtomElement*); the thin C++ shim in arch_cuda/cuda_hooks.cpp flattens them to the plain pointers and
 * sizes below and calls this ABI. Each entry point cites the hook / reference function it replaces.
 *
 * Conventions: every function returns 0 on success, a negative MISA_B200_E* code on failure (the reference's
 * hooks return void and abort through MPI_Abort, src/simulation.cpp:100-102 -- the shim does the abort).
 * All reals are fp64, lattice indices are the reference's linear index idx = (z*Sy + y)*Sx + x over the
 * ghost-extended lattice with doubled x (reference src/atom/atom_list.h:117-119, src/atom/atom_set.cpp:18-26).
 * `atoms` pointers are HOST pointers to the reference's 104-byte AtomElement AoS (src/atom/atom_element.h:18-41)
 * unless stated otherwise. There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef MISA_B200_H
#define MISA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MISA_B200_OK 0
#define MISA_B200_ENODEV (-1)   /* no CUDA device / CUDA runtime error */
#define MISA_B200_EINVAL (-2)   /* bad argument */
#define MISA_B200_ESTATE (-3)   /* call order violated (e.g. compute before potential is set) */
#define MISA_B200_ENCCL (-4)    /* NCCL unavailable or NCCL error */
#define MISA_B200_EOVERFLOW (-5) /* inter-atom capacity exceeded */

typedef struct misa_b200_ctx misa_b200_ctx;

/* The fields of comm::BccDomain the path reads; the doubled-x twins and the
 * ghost-extended sizes are derived inside. */
typedef struct misa_b200_domain {
    int64_t phase_space[3];        /* global box in cells */
    int32_t grid_size[3];          /* process grid */
    int32_t grid_coord[3];         /* this sub-box's coordinate in the grid */
    int32_t sub_box_lattice_size[3]; /* owned cells per dimension (x NOT doubled) */
    int32_t lattice_size_ghost[3];   /* ghost cells per side (x NOT doubled) */
    int32_t sub_box_lattice_low[3];  /* sub_box_lattice_region.{x,y,z}_low (x NOT doubled) */
    int32_t rank_id_neighbours[3][2]; /* [dim][LOWER=0/HIGHER=1] */
    int32_t rank;                   /* own rank id */
    double lattice_const;
    double cutoff_radius_factor;
    double meas_global_length[3];
} misa_b200_domain;

/* One tabulated function as libpot holds it after eam::interpolateFile(): n rows of 7 spline coefficients
 * (row m = 1..n, row 0 unused), uniform grid with spacing 1/inv_dx starting at 0. */
typedef struct misa_b200_table {
    int32_t n;
    double inv_dx;
    const double *spline; /* (n+1)*7 doubles */
} misa_b200_table;

/* per-kernel timing slots for misa_b200_profile_read */
enum {
    MISA_B200_K_VERLET1 = 0, /* firststep + run-away test */
    MISA_B200_K_HALO_X,      /* ghost fill: positions + types */
    MISA_B200_K_RHO,         /* rho (+df when fused) */
    MISA_B200_K_DF,
    MISA_B200_K_HALO_DF,
    MISA_B200_K_FORCE,
    MISA_B200_K_VERLET2,
    MISA_B200_K_INTER,       /* everything on the off-lattice path */
    MISA_B200_K_XFER,        /* AoS<->SoA conversion kernels */
    MISA_B200_K_COUNT
};

/* ---- environment: replaces cuda_env_init / cuda_env_clean (arch_imp.h:17-19; called from
 *      frontend/misa_md.cpp:92,128) ------------------------------------------------------------ */
int misa_b200_env_init(int device /* <0: LOCAL_RANK env or 0 */);
int misa_b200_env_clean(void);
int misa_b200_device_count(void);
const char *misa_b200_last_error(void);

/* ---- setup: replaces cuda_domain_init (arch_imp.h:21, call site simulation.cpp:52-54),
 *      cuda_nei_offset_init (arch_imp.h:23, simulation.cpp:67-69), cuda_pot_init (arch_imp.h:25,
 *      simulation.cpp:133) ------------------------------------------------------------------------ */
int misa_b200_create(const misa_b200_domain *dom, misa_b200_ctx **out);
int misa_b200_destroy(misa_b200_ctx *ctx);
/* the four protected vectors of NeighbourIndex<AtomElement> (src/atom/neighbour_index.h:66-69) */
int misa_b200_set_neighbour_offsets(misa_b200_ctx *ctx, const int64_t *even, size_t n_even, const int64_t *odd,
                                    size_t n_odd, const int64_t *half_even, size_t n_half_even,
                                    const int64_t *half_odd, size_t n_half_odd);
/* Build the same four vectors inside the library (NeighbourIndex::make, neighbour_index.inl:13-76) --
 * for hosts that do not run the reference's AtomSet. Returned copies are optional (may be NULL). */
int misa_b200_make_neighbour_offsets(misa_b200_ctx *ctx, int cut_lattice, double cutoff_radius_factor);
int misa_b200_get_neighbour_offsets(misa_b200_ctx *ctx, int which /*0 even,1 odd,2 half_even,3 half_odd*/,
                                    int64_t *out, size_t cap, size_t *n);
/* Host-only planning (no CUDA device needed): the integer side of the path in the reference's index space.
 * plan_offsets = NeighbourIndex::make for this sub-box's ghost-extended lattice; plan_halo = sendlist /
 * recvlist of AtomList::exchangeAtomFirst (src/atom/atom_list.cpp:25-49, src/pack/lat_particle_packer.cpp:65-139)
 * for message (dim, dir) plus the periodic image shift LatParticlePacker::setOffset adds on send. */
int misa_b200_plan_offsets(const misa_b200_domain *dom, int cut_lattice, double cutoff_radius_factor, int which,
                           int64_t *out, size_t cap, size_t *n);
int misa_b200_plan_halo(const misa_b200_domain *dom, int dim, int dir, int64_t *send, int64_t *recv, size_t cap,
                        size_t *n, double shift[3]);
/* Host-only: how the stencil kernels loop the offsets of plan_offsets (csrc/misa_b200.cu:plan_stencil). `sorted` = the
 * full list of central parity `parity` (reference index space) in device order: the near group first (sites closer than
 * crf + 0.05 lattice constants; its first n_half entries are one of each +v / -v pair, the next n_half the mirrors in the
 * order lower_slot[] names), then the rest by ascending site separation; site_r2 = squared site separation / a^2 of each
 * entry; prefix[L], L = 0..40 = entries a warp loops when (largest displacement level among its atoms) + (bound on the
 * partner's) = L, one level = 0.01 a: every offset whose sites are closer than (crf + 0.01 L) a is inside that prefix,
 * which is what makes the pruning exact. No reference counterpart (the CPU loops all 114 half offsets, atom.cpp:173). */
int misa_b200_plan_stencil(const misa_b200_domain *dom, int cut_lattice, double cutoff_radius_factor, int parity,
                           int64_t *sorted, double *site_r2, size_t cap, size_t *n, int32_t *n_near, int32_t *n_half,
                           int32_t prefix[41], int32_t *lower_slot);
/* Host-only: the boxes of owned cells one stencil launch covers (x0,y0,z0,nx,ny,nz). which = 0 the whole sub-box, 1 its
 * interior (no ghost within the stencil reach), 2 the six slabs around that interior (the interior / boundary split of the
 * overlapped NCCL exchange), 3 interior FIRST then the slabs in one launch, the interior kept 15 more cells from the x faces
 * (no 128-byte line with a ghost in it is touched before the in-kernel wait for the neighbours' push). */
int misa_b200_plan_regions(const misa_b200_domain *dom, int which, int32_t boxes[7][6], int32_t *n_boxes, int64_t *units,
                           int64_t *split);
/* ... and the order in which such a launch visits its 2 * units warp units: u -> (sub-lattice, unit within plan_regions'
 * numbering). which = 3: the interior units of BOTH sub-lattices come before any boundary unit. */
int misa_b200_plan_unit_order(const misa_b200_domain *dom, int which, int64_t u, int32_t *parity, int64_t *unit);
/* Host-only: the three staged exchanges above composed into ONE ghost <- owned map (what the direct NVLink push of
 * the multi-GPU path applies, csrc/p2p.cuh). Entry i: site dst[i] of this sub-box receives site src[i] of the sub-box
 * at offset (sx, sy, sz), code[i] = (sx+1) + 3 (sy+1) + 9 (sz+1); all sub-boxes have the same shape, so read backwards
 * it is what this sub-box pushes to the sub-box at the opposite offset, with image shift shift[code] added. Indices
 * in the reference's index space. Replaces nothing in the reference: it is the closed form of
 * comm::neiSendReceive x 3 with LatPacker (src/atom/atom_list.cpp:51-57, src/pack/lat_particle_packer.cpp:97-139). */
int misa_b200_plan_push(const misa_b200_domain *dom, int64_t *dst, int64_t *src, int8_t *code, size_t cap, size_t *n,
                        double shift[27][3]);
/* n_types species in atom_type enum order (Fe, Cu, Ni); phi is n_types*n_types, symmetric. */
int misa_b200_set_potential(misa_b200_ctx *ctx, int n_types, const misa_b200_table *electron_density,
                            const misa_b200_table *embedded, const misa_b200_table *phi);

/* ---- compat mode: the three per-step hooks on the HOST AoS array -- replace cuda_eam_rho_calc /
 *      cuda_eam_df_calc / cuda_eam_force_calc (arch_imp.h:27-31; call sites atom.cpp:160,294,321).
 *      Post-conditions equal the CPU branches latRho / latDf / latForce (atom.cpp:151-192,286-309,311-358)
 *      as seen after the host's reverse halos: owned sites receive their COMPLETE sums, ghost sites are
 *      left untouched (SURVEY.md section 8b "required post-conditions"). ------------------------- */
int misa_b200_eam_rho_calc(misa_b200_ctx *ctx, void *atoms, double cutoff_radius);
int misa_b200_eam_df_calc(misa_b200_ctx *ctx, void *atoms, double cutoff_radius);
int misa_b200_eam_force_calc(misa_b200_ctx *ctx, void *atoms, double cutoff_radius);
/* number of AtomElement records in the ghost-extended array (AtomList::cap(), src/atom/atom_list.h:170-172) */
int misa_b200_site_count(misa_b200_ctx *ctx, size_t *n_sites);
/* page-lock the host AoS array once so the hook transfers run at PCIe speed (optional) */
int misa_b200_host_register(void *ptr, size_t bytes);
int misa_b200_host_unregister(void *ptr);

/* ---- resident mode: state lives in HBM as SoA across steps; the host loop body of
 *      simulation::simulate (simulation.cpp:164-194) runs on the device ------------------------- */
int misa_b200_upload_atoms(misa_b200_ctx *ctx, const void *atoms);      /* whole ghost-extended array */
int misa_b200_download_atoms(misa_b200_ctx *ctx, void *atoms);
int misa_b200_upload_inter(misa_b200_ctx *ctx, const void *inter_atoms, size_t n); /* InterAtomList::inter_list */
int misa_b200_download_inter(misa_b200_ctx *ctx, void *inter_atoms, size_t cap, size_t *n);
int misa_b200_set_timestep(misa_b200_ctx *ctx, double dt);              /* NewtonMotion::setTimestepLength */
int misa_b200_prepare(misa_b200_ctx *ctx);   /* exchangeAtomFirst + clearForce + computeEam (simulation.cpp:137-145) */
int misa_b200_step(misa_b200_ctx *ctx, int n_steps); /* firststep .. secondstep, n times */
/* the same loop body on a HOST AoS array: upload, n steps on the device, download. On return the OWNED lattice records are
 * current (ghost records are neither read nor written after the first call). Atoms that ran away are not in the array: they
 * live in the library's inter-atom list (misa_b200_download_inter / misa_b200_dump_records) and their sites come back vacant.
 * The host must not change the occupancy (type) of records between calls -- use misa_b200_upload_atoms for that. */
int misa_b200_step_host(misa_b200_ctx *ctx, void *atoms, int n_steps);
int misa_b200_setv(misa_b200_ctx *ctx, const int32_t lat[4], const double direction[3], double energy); /* atom::setv */
int misa_b200_collision_step(misa_b200_ctx *ctx, const int32_t lat[4], const double direction[3], double energy);
/* configuration::rescale's velocity scaling (reference src/system_configuration.cpp:97-110) with the CURRENT global
 * temperature supplied by the caller: v *= sqrt(t_set / t_now) on this sub-box (lattice + inter list) */
int misa_b200_rescale(misa_b200_ctx *ctx, double t_set, double t_now);
/* thermo[0]=sum m v^2 (configuration::mvv), [1]=E_pot local [eV] (ours), [2]=owned valid atoms,
 * [3]=local inter atoms, [4]=ghost inter atoms, [5]=run-aways detected in the last step */
int misa_b200_thermo(misa_b200_ctx *ctx, double out[6]);
int misa_b200_sync(misa_b200_ctx *ctx);

/* ---- the callers / data formats either side of the path, resident mode ------------------ */
/* WorldBuilder::build (reference src/world_builder.cpp:64-199) on the device, BY GLOBAL ATOM ID: ids, species,
 * perfect bcc positions, velocities (std::mt19937(seed)() / 0xFFFFFFFF - 0.5) / mass with draw 3(id-1)+k for
 * component k of atom id (what the reference produces on one rank; independent of the process grid), vcm +
 * zeroMomentum over the global box, configuration::rescale to t_set (0: no rescale). A single non-zero ratio
 * gives a pure lattice; otherwise species = cumulative-ratio rule on a counter-based hash of (alloy_seed, id)
 * (the reference's unseeded rand() has no reproducible stream). Collective only in the sense that every rank
 * must call it with the same arguments; no communication. Replaces misa_b200_upload_atoms for synthetic starts. */
int misa_b200_build_world(misa_b200_ctx *ctx, uint32_t seed, double t_set, const int32_t ratio[3], uint64_t alloy_seed);
/* configuration::temperature / kineticEnergy (reference src/system_configuration.cpp:26-64) over ALL sub-boxes
 * (one NCCL all-reduce when a communicator is set): out[0] = sum m v^2, [1] = T [K] with dof = 3 n - 3,
 * [2] = kinetic energy [eV], [3] = atoms counted (lattice + inter). Every rank must call it. */
int misa_b200_temperature(misa_b200_ctx *ctx, uint64_t n_atoms_global, double out[4]);
/* configuration::rescale (reference src/system_configuration.cpp:86-111), global: temperature as above, then
 * v *= sqrt(t_set / T) on the device. The stage machine's `rescale` (frontend/md_simulation.cpp:62-66). */
int misa_b200_rescale_to(misa_b200_ctx *ctx, double t_set, uint64_t n_atoms_global);
/* AtomDump::dump + BufferedFileWriter::write (reference frontend/io/atom_dump.cpp:39-75, frontend/io/
 * buffered_io.cpp:18-36): the 72-byte atom_dump::AtomInfoDump record stream (frontend/io/atom_info_dump.h:14-22;
 * id, step, type, inter_type = 0, location, velocity; padding bytes zero) of this sub-box -- inter atoms in list
 * order, then the valid sites of [begin, end) (ghost-inclusive doubled-x coordinates; NULL = the owned sub-box of
 * frontend/io/output_base_interface.h:26-31) in z,y,x order -- compacted on the device; only the records cross
 * PCIe. records == NULL: size query. n_records always receives the record count. */
int misa_b200_dump_records(misa_b200_ctx *ctx, const int32_t begin[3], const int32_t end[3], uint64_t time_step,
                           void *records, size_t cap_records, size_t *n_records);

/* single passes on the resident state (kernel-level parity tests and profiling) */
int misa_b200_pass_halo_x(misa_b200_ctx *ctx);  /* AtomList::exchangeAtom */
int misa_b200_pass_clear(misa_b200_ctx *ctx);   /* atom::clearForce */
int misa_b200_pass_rho(misa_b200_ctx *ctx);     /* latRho (+interRho), complete sums on owned sites */
int misa_b200_pass_df(misa_b200_ctx *ctx);      /* latDf */
int misa_b200_pass_halo_df(misa_b200_ctx *ctx); /* DfEmbedPacker exchange */
int misa_b200_pass_force(misa_b200_ctx *ctx);   /* latForce (+interForce) */
int misa_b200_pass_verlet1(misa_b200_ctx *ctx); /* NewtonMotion::firststep + atom::decide */
int misa_b200_pass_verlet2(misa_b200_ctx *ctx); /* NewtonMotion::secondstep */

/* options (A/B switches; the defaults are the production path; also settable as MISA_B200_OPTS="name=value,..."):
 *   "prune" 1  stencil lists pruned by the measured displacements whenever the 0.2a invariant of atom::decide (atom.cpp:42)
 *              holds on the device: per warp, a prefix of the distance-sorted full list
 *   "mark" 1   partner bound of that pruning from the cells marked by far-displaced atoms instead of the global maximum
 *   "fuse" 1   rho + df in one kernel when no inter atoms; "fuse_verlet" 1: second half-kick applied by the next firststep
 *   "smem" 1 / "fast" 1 / "tex" 1 / "novac" 1 / "dilute" 1   kernel generations and their fast paths (csrc/eam_*.cuh)
 *   "sym" 0    pair-symmetric passes (each near pair evaluated once): measured slower
 *   "pipe" 1   step without a host round trip on its critical path; "overlap" -1, "reserve" 8: interior / boundary split
 *              of the stencil launches around a staged NCCL exchange
 *   "p2p" -1   ghost exchange by direct stores into the neighbours' HBM when all of them are peer-mapped (0: NCCL);
 *   "late" 1   wait for the neighbours' push inside the stencil kernels (interior units first)
 *   "push_fused" 1  sync-free step: k_verlet1 / k_rho_f store band sites' positions / df into the neighbours' ghosts themselves
 *              (0: push kernels); "dmax_flags" 1: displacement maxima travel on the push flags (0: all-reduce in front of rho);
 *              "p2p_fence" 2: ARRIVE from a follow-up kernel (push kernels only); "p2p_timeout_s" 30; "p2p_debug" 0
 *   "inter_dev" 1   off-lattice atoms resident on the device (csrc/inter_dev.cuh; 0: the host-list form, before any exists)
 *   "vac_sentinel" 1  vacant sites invisible to the stencil kernels (their record's position kept aside); "low_list" 1
 *   "host_slabs" 1 / "slab_planes" 3   misa_b200_step_host on one sub-box as a pipeline of z-slabs */
int misa_b200_set_option(misa_b200_ctx *ctx, const char *name, int value);

/* read-only introspection (tests, benches): "n_off", "n_full", "n_half", "dmax", "single", "novac", "dilute", "n_minor", "sym",
 * "smem_bytes", "pipe_steps", "pipe_redo", "host_slab_steps", "host_slab_redo", "mark_level", "mark_count", "p2p", "p2p_error",
 * "p2p_dbg_<x|df>_<head|body|tail|all>" */
int misa_b200_query(misa_b200_ctx *ctx, const char *name, double *value);

/* ---- multi-GPU: one sub-box per GPU (replaces libcomm's comm::neiSendReceive over MPI; call sites atom.cpp:114,131,145,
 *      atom_list.cpp:45,53). comm_init builds the NCCL communicator and, when every surrounding sub-box is on a
 *      peer-accessible GPU of this node, maps their arrays (CUDA IPC) for the direct ghost push of csrc/p2p.cuh;
 *      otherwise ghosts travel as staged NCCL send/recv between face neighbours ----- */
int misa_b200_comm_unique_id(void *out128);
int misa_b200_comm_init(misa_b200_ctx *ctx, const void *unique_id128, int rank, int n_ranks);
int misa_b200_comm_destroy(misa_b200_ctx *ctx);

/* ---- measurement: CUDA-event durations per kernel slot on the launching stream ---------------- */
int misa_b200_profile_enable(misa_b200_ctx *ctx, int on);
int misa_b200_profile_read(misa_b200_ctx *ctx, double ms_sum[MISA_B200_K_COUNT], int64_t launches[MISA_B200_K_COUNT]);
int misa_b200_launch_count(misa_b200_ctx *ctx, int64_t *n); /* kernels launched since create */
/* device time of n_steps through CUDA events on the context's stream (ms) */
int misa_b200_timed_steps(misa_b200_ctx *ctx, int n_steps, double *ms);
/* what the stencil kernels loop on the CURRENT resident state (diagnostic launch, not on the step path): out[0] owned sites,
 * [1] offsets looped (per-warp prefixes of the pruned list), [2] pair evaluations executed (branch-free near group + voted
 * far offsets), [3] pairs inside the cutoff -- totals over the sub-box; divide by out[0] for per-atom figures. Feeds the
 * fp64 view of bench.py's roofline ("report both N_off and N_pair actually used"). */
int misa_b200_stencil_stats(misa_b200_ctx *ctx, double out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MISA_B200_H */
