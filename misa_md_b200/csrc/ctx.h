// misa_md_b200/csrc/ctx.h -- internal state of one sub-box on one B200 (not part of the C ABI).
//
// HBM layout: structure-of-arrays over the ghost-extended lattice, PARITY-SPLIT.
// The reference's linear index idx = (z*Sy + y)*Sx + x (x doubled, even x = cube corner, odd x = body
// centre; reference src/atom/atom_list.h:117-119) maps to the device index
//     d(idx) = (idx >> 1) + (idx & 1) * H,        H = n_ext / 2
// i.e. the first half of every array holds the corner sub-lattice as a plain [z][y][cx] cube and the second
// half the body-centre sub-lattice. A warp then works on 32 consecutive cells of ONE sub-lattice: all lanes
// share one neighbour-offset list and every neighbour load is a contiguous 256-byte segment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <string>
#include "../../include/misa_b200.h"

#define MISA_MAX_TYPES 3
#define MISA_ROW 8  // spline rows padded from 7 to 8 doubles (64-byte rows)

struct Geo {
    int nx, ny, nz;     // owned cells
    int gx, gy, gz;     // ghost cells per side
    int sxc, sy, sz;    // ghost-extended cells: sxc = nx + 2gx (cells; the reference's Sx is 2*sxc)
    int lo[3];          // sub_box_lattice_region low (cells)
    long long H;        // n_ext / 2
    long long n_ext;    // 2*sxc*sy*sz
    long long n_cells_owned; // nx*ny*nz
    double a;           // lattice constant
    double rc2;         // (a*crf)^2 computed as cutoff_radius*cutoff_radius (reference src/atom.cpp:177)
    double runaway2;    // pow(0.2*a, 2.0) (reference src/atom.cpp:42)
};

struct DevTables {
    const double *elec;   // [n_types][n_r+1][8]
    const double *embed;  // [n_types][n_rho+1][8]
    const double *phi;    // [n_types*n_types][n_r+1][8]
    int n_types, n_r, n_rho;
    double inv_dr, inv_drho;
};

struct Soa {
    double *x[3], *v[3], *f[3];
    double *rho, *df;
    int8_t *type;
    unsigned long long *id;
    unsigned char *ulev;   // ceil(|x - site| / 0.01a) of the valid atom on the site (k_verlet1 / k_max_displacement): per-warp stencil pruning
    unsigned char *hot;    // [H] per CELL: an atom displaced by more than the marking level sits within reach, or a ghost may (k_verlet1)
    // Vacant sites are INVISIBLE to the stencil kernels: x holds MISA_VACANT_X (no pair is ever in range, no per-neighbour species
    // load is needed) and the position the departed atom left behind -- which the reference keeps in the record and reads in
    // atom::decide (ws::isOutBox(*near_atom), src/atom.cpp:69) -- lives here. Null: feature off (option "vac_sentinel" 0).
    double *sx[3];
};

struct InterSoa {  // off-lattice atoms: [0, n_local) local, [cap/2, cap/2 + n_ghost) ghost copies
    double *x[3], *v[3], *f[3];
    double *rho, *df;
    int8_t *type;
    unsigned long long *id;
    int *site;   // device index of the Wigner-Seitz site (or -1)
    int *next;   // linked list through a site bucket
};

struct HaloList {          // one (dim, dir) message
    int n = 0;
    int *d_send = nullptr; // device indices packed into the message, reference sendlist order
    int *d_recv = nullptr; // device indices the mirrored message is unpacked into, reference recvlist order
    double shift[3] = {0, 0, 0};
};

// direct ghost push over NVLink peer memory (p2p.cuh): per direction code the destination sub-box's arrays and flags
struct P2pPeers {
    double *xyzd[27];              // the destination's x|y|z|df block
    int8_t *type[27];
    unsigned long long *flags[27];
    double shift[27][3];           // periodic image shift of the group
    long long stride;              // field stride of the block (doubles)
    unsigned int mask;             // direction codes in use
    long long spin_limit;          // SM clocks a flag wait may take before it gives up with an error
    unsigned int *fault;           // device word: a READY gate timed out -- producers that push from inside store nothing
    int fence_mode;                // p2p.cuh:p2p_push_tail
    unsigned long long *dbg;       // phase timing of the push kernels (option "p2p_debug"), else null
};

struct misa_b200_ctx {
    misa_b200_domain dom;
    Geo geo;
    int device = 0;
    cudaStream_t stream = nullptr;
    Soa s{};
    unsigned char *d_aos = nullptr;       // staging for AoS <-> SoA (n_ext * 104 B)
    // neighbour stencil, device-index space, per central parity
    std::vector<int64_t> ref_off[4];      // even, odd, half_even, half_odd in REFERENCE index space
    int n_full = 0;
    int *d_off_full = nullptr;            // [2][n_full]
    int *d_off_full_addr = nullptr, *d_off_levels_addr = nullptr;   // the same lists, every (level, parity) segment sorted by address
    // pruned stencils by displacement level L: every valid atom within L*0.01a of its site => two lattice atoms
    // can only be within r_c if their SITES are closer than (crf + 0.02 L) a. L = 20 is atom::decide's own bound.
    static const int kLevels = 21;
    static const int kPairLevels = 41;    // prefix lengths of the distance-sorted full list: sites closer than (crf + 0.01 L) a
    int prefix_n[kPairLevels] = {0};
    // partner bound of the per-warp pruning: atoms above level mark_T mark the cells within reach in k_verlet1; warps that see no
    // mark bound their partners by mark_T instead of the global maximum (eam_smem.cuh:list_len)
    unsigned char *d_hot = nullptr, *d_hot_init = nullptr;
    int opt_mark = 1, mark_T_next = 0, mark_T_used = 0, mark_epoch = 0;   // d_hot: epoch bytes; d_hot_init: static edge map
    bool mark_valid = false;
    unsigned char *d_pmax = nullptr, *d_ptmp = nullptr;   // per-cell partner bound of the serial path (kernels.cuh:k_pmax_*)
    bool pmax_valid = false;
    int *d_low_list = nullptr, *d_low_count = nullptr;   // atoms with a pair below the staged table range, cascade form (eam_fast.cuh:k_low_fix)
    int opt_low_list = 1;
    int level_n[kLevels] = {0};
    int level_near[kLevels] = {0};        // leading entries (lists are sorted by site distance) that are almost surely in range
    int near_full = 0;
    size_t level_ofs[kLevels] = {0};
    int *d_off_levels = nullptr;          // concatenated [L][2][level_n[L]]
    double dmax2 = 0.0;                   // largest squared displacement from the ideal site (valid atoms, ghosts incl.)
    bool dmax_valid = false;
    unsigned long long *d_stepinfo = nullptr, *h_stepinfo = nullptr; // [0] off-lattice activity, [1] dmax2 bits
    // potential
    double *d_elec = nullptr, *d_embed = nullptr, *d_phi = nullptr;
    DevTables tab{};
    bool have_pot = false, have_off = false, have_atoms = false;
    // Hermite (value, knot slope) copies of the r tables + the shared-memory staging plan (eam_smem.cuh)
    double2 *d_herm = nullptr;            // [n_types + n_types^2][n_r + 1]
    double2 *d_ea = nullptr;              // [n_types + n_types^2][n_r + 1]: (knot slope s_m, v_{m+1} - v_m) -- with d_es what a slope-only lookup needs
    double *d_es = nullptr;               // [n_types + n_types^2][n_r + 1] (+ 2 doubles of padding): knot slopes, dense
    double *d_mono = nullptr;             // [n_types + n_types^2][n_r + 1][4]: (c3, c4, c5, c6), one 256-bit load per interval (eam_fast.cuh)
    double2 *d_c34 = nullptr, *d_c56 = nullptr;   // [n_types + n_types^2][n_r + 1]: columns (3,4) and (5,6) of the caller's 7-coefficient rows
    bool hermite_ok = false;              // caller's 7-coefficient rows are Hermite-consistent (checked on the host)
    unsigned long long *d_census = nullptr, *h_census = nullptr; // valid sites per species (ghost-extended array)
    unsigned long long census[MISA_MAX_TYPES] = {0, 0, 0};      // global (all sub-boxes) once prepare() ran
    bool census_valid = false;
    int census_boxes = 1;                 // sub-boxes summed into d_census by the fetch in flight
    int sm_count = 0, smem_optin = 0;
    cudaTextureObject_t tex_x[3] = {0, 0, 0}, tex_df = 0; // int2 views of x,y,z,df (TEX-pipe neighbour loads)
    double *d_xyzd = nullptr;             // ONE allocation behind s.x[0..2] and s.df, field stride xyzd_stride doubles
    long long xyzd_stride = 0;
    cudaTextureObject_t tex_all = 0;      // int2 view of the whole block: one handle for all four fields (eam_fast.cuh)
    int opt_tex = 1, opt_novac = 1;
    int opt_vac_sentinel = 1;             // see Soa::sx
    int opt_fast = 1;                     // third-generation kernels (eam_fast.cuh)
    long long n_valid_sites = -1;         // valid sites at the last census, scaled so that "== geo.n_ext" means none vacant
    bool seen_offlattice = false;         // a run-away / inter atom was reported by any sub-box since the last census
    int opt_fuse_verlet = 1;              // inside a multi-step call: second half-kick of step k folded into k_verlet1 of step k+1
    int opt_dilute = 1;                   // dilute-alloy kernels (eam_fast.cuh, DILUTE variants) when one species holds >= 90 % of the sites
    // static minority-neighbour lists (eam_fast.cuh, built by prepare(); valid until atoms are replaced or anything runs away)
    unsigned char *d_mcount = nullptr, *d_mentry = nullptr;
    int *d_minor = nullptr, *d_minor_count = nullptr; // device indices of the owned minority-species atoms
    int n_minor = 0, minor_maj = 0;
    bool minor_valid = false;
    int opt_smem = 1;                     // use the shared-memory table kernels when possible
    // pair-symmetric stencil passes (eam_sym.cuh): the leading n_half entries of every offset list are the "upper"
    // (reference half-list, src/atom/neighbour_index.inl:79-92) near offsets; their per-pair scalars go through d_pair
    int opt_sym = 0;                      // OFF by default: measured slower than the full-list kernels
    int n_half = 0;                       // 0: lists not symmetric-capable
    int sym_lo[3] = {0, 0, 0}, sym_hi[3] = {0, 0, 0}; // cells below / above the owned box whose sites own a pair with an owned atom
    int2 *d_lo_tab = nullptr;             // [2][n_half] per central parity: (device offset to the lower neighbour, its slot)
    double *d_pair = nullptr;             // [n_half][n_ext]
    size_t pair_elems = 0;
    double stage_r_lo = 2.0;              // tables are staged for r >= stage_r_lo (Angstrom)
    // halo
    HaloList halo[3][2];
    bool all_self = true;                 // 1x1x1 grid: every neighbour is this sub-box
    int n_ghost_map = 0;                  // fused periodic ghost fill (all_self only)
    int *d_ghost_dst = nullptr, *d_ghost_src = nullptr;
    int8_t *d_ghost_shift = nullptr;      // per ghost site: shift code (3 x {-1,0,1}) packed
    std::vector<int> h_ghost_dst, h_ghost_src;   // host copies of the fused fill map (all_self), for the per-slab sub-maps below
    std::vector<int8_t> h_ghost_code;
    // slab-pipelined misa_b200_step_host (misa_b200.cu:step_host_slabs): z-slabs of the owned box travel H2D / D2H on two copy
    // streams while earlier / later slabs are being computed; the fill map re-sorted by the SOURCE site's slab
    int opt_host_slabs = 1, slab_T = 3;
    int n_slabs = 0;
    std::vector<int> slab_z0, slab_map_ofs;      // [n_slabs + 1]
    int *d_slab_dst = nullptr, *d_slab_src = nullptr;
    int8_t *d_slab_code = nullptr;
    unsigned char *d_aos_out = nullptr;          // second staging array: the input records stay intact until the step is known to be good
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_out;
    int64_t host_slab_steps = 0, host_slab_redo = 0;
    double *d_sendbuf[2] = {nullptr, nullptr}, *d_recvbuf[2] = {nullptr, nullptr};
    size_t halo_buf_elems = 0;
    // direct push of the composed ghost <- owned map into the neighbours' HBM (p2p.cuh)
    int opt_p2p_debug = 0;
    int opt_p2p_fence = 2;                // 0: system-scope fence in every CTA of a push kernel; 1: device-scope, the last CTA releases at system scope; 2: ARRIVE from a follow-up kernel
    int opt_dmax_flags = 1;               // sync-free step over peer memory: displacement maxima travel on the push flags (no all-reduce in front of rho)
    int opt_late = 1;                     // wait for the neighbours' push inside the stencil kernels (interior units first)
    int opt_p2p = -1;                     // -1 / 1: whenever every surrounding sub-box is peer-mapped on this node; 0: NCCL send/recv
    bool p2p_active = false;
    int opt_p2p_timeout_s = 30;           // flag waits give up after this many seconds (error, no stores; p2p.cuh)
    unsigned int p2p_last_error = 0;      // flag code of the last timeout reported
    bool p2p_fault = false;               // a wait timed out and the caller has been told: the push stays off until comm_init re-synchronises
    int n_push = 0;
    int *d_push_dst = nullptr, *d_push_src = nullptr;
    int8_t *d_push_code = nullptr;
    bool push_code_used[27] = {false};
    P2pPeers p2p{};
    unsigned long long *d_flags = nullptr, p2p_epoch = 0, p2p_ready_sent = 0;   // flags: [0,27) ready, [32,59) arrive, [63] CTA counter
    unsigned int *h_p2p_err = nullptr, *d_p2p_err = nullptr;
    unsigned long long *d_p2p_dbg = nullptr;   // [2][8]: position / df push
    P2pPeers *d_p2p_dev = nullptr;             // device copy of `p2p` for the producers that push from inside (kernels.cuh:push_site)
    unsigned int *d_p2p_fault = nullptr;
    int opt_push_fused = 1;                    // sync-free step: k_verlet1 pushes positions, k_rho_f's epilogue pushes df, the stencil kernels post ARRIVE
    std::vector<void *> p2p_opened;
    // integrator
    double dt = 0.001;
    double dt_inv_m[MISA_MAX_TYPES] = {0, 0, 0};
    // run-away / inter atoms
    InterSoa inter{};
    int inter_cap = 0;
    int opt_inter_dev = 1;                // the inter-atom list lives on the device (inter_dev.cuh); 0: host list (inter.cuh)
    int *d_counters = nullptr;            // [0] run-aways this step, [1] n_local inter, [2] n_ghost inter, [3] overflow, [4] invariant violations
    int *h_counters = nullptr;            // pinned, mapped mirror (hd_*: the device's view; k_activity stores into it)
    int *hd_counters = nullptr;
    unsigned long long *hd_stepinfo = nullptr;
    int n_inter_local = 0, n_inter_ghost = 0;
    int *d_site_head = nullptr;           // n_ext ints, -1 = empty bucket
    int *d_runaway = nullptr;             // device indices of this step's run-away sites
    bool inter_active = false;            // any sub-box has off-lattice atoms this step (globally agreed)
    double *d_reduce = nullptr;           // scratch for reductions (8 doubles)
    double *h_reduce = nullptr;           // pinned
    int last_runaways = 0;
    // options
    int opt_prune = 1, opt_fuse = 1;
    // pipelined step (misa_b200.cu:step_pipelined): ghost exchange on stream2 while interior cells are computed
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_v1 = nullptr, ev_act = nullptr, ev_hx = nullptr, ev_rho = nullptr, ev_hdf = nullptr;
    unsigned long long *d_stepinfo_g = nullptr; // [0] activity, [1] dmax2 bits: MAX over all sub-boxes (all-reduce result)
    unsigned long long *d_stepinfo_n = nullptr; // [1] dmax2 bits: MAX over this sub-box and the 26 around it (folded from the push flags)
    bool dmax_by_flags = false;                 // this step's pushes carry the displacement maxima (sync-free step over peer memory)
    int opt_pipe = 1;
    int opt_overlap = -1;                       // multi-GPU: interior/boundary split with the exchange on stream2 (-1 auto)
    int opt_reserve = 8;                        // SMs the interior stencil launches leave to the exchange kernels
    int64_t pipe_steps = 0, pipe_redo = 0;      // steps taken by the pipelined path / of those re-done serially (off-lattice activity)
    // dump compaction (dump.cuh)
    unsigned long long *d_dump = nullptr, *d_dump_base = nullptr, *d_dump_total = nullptr, *h_dump_total = nullptr;
    unsigned *d_dump_count = nullptr;
    size_t dump_cap = 0, dump_tiles_cap = 0;
    // NCCL
    void *nccl_comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    // profiling
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev[MISA_B200_K_COUNT];
    double prof_ms[MISA_B200_K_COUNT] = {0};
    int64_t prof_n[MISA_B200_K_COUNT] = {0};
    int64_t launches = 0;
};
