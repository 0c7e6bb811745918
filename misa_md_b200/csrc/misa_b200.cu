// misa_md_b200/csrc/misa_b200.cu -- host side of the C ABI declared in include/misa_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see build.py).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "ctx.h"
#include "util.cuh"
#include "kernels.cuh"
#include "nccl_dl.cuh"
#include "inter.cuh"
#include "inter_dev.cuh"
#include "eam_smem.cuh"
#include "eam_fast.cuh"
#include "eam_sym.cuh"
#include "dump.cuh"
#include "world.cuh"

extern "C" const char *misa_b200_last_error(void) { return g_err.c_str(); }

// reference src/types/pre_define.h:11-19, src/types/atom_types.h:17-19
static const double kMvv2e = 1.0364269e-4;
static const double kFtm2v = 1.0 / 1.0364269e-4;
static const double kBoltz = 8.617343e-5;
static const double kMass[MISA_MAX_TYPES] = {55.845, 63.546, 58.6934};

static int g_device = -1;

extern "C" int misa_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int misa_b200_env_init(int device) {
    int n = misa_b200_device_count();
    REQ(n > 0, MISA_B200_ENODEV, "misa_b200_env_init: no CUDA device visible (there is no CPU fallback)");
    if (device < 0) {
        // node-local rank as the launcher exports it: torchrun, Open MPI, MVAPICH2, Slurm, PMI (Intel MPI / MPICH hydra);
        // the hook shim passes MPI_Comm_split_type(MPI_COMM_TYPE_SHARED)'s rank explicitly when none of them is set
        static const char *const names[] = {"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID",
                                            "MPI_LOCALRANKID", "PMI_LOCAL_RANK"};
        device = 0;
        for (const char *nm : names)
            if (const char *lr = getenv(nm)) { device = atoi(lr) % n; break; }
    }
    REQ(device < n, MISA_B200_EINVAL, "misa_b200_env_init: device index out of range");
    CU(cudaSetDevice(device));
    CU(cudaFree(0));
    g_device = device;
    return MISA_B200_OK;
}

extern "C" int misa_b200_env_clean(void) {
    if (g_device >= 0) cudaDeviceSynchronize();
    g_device = -1;
    return MISA_B200_OK;
}

// -------------------------------------------------------------------------------------------------
// profiling helpers: CUDA events on the launching stream around each kernel slot
// -------------------------------------------------------------------------------------------------
struct Slot {
    misa_b200_ctx *c;
    int k;
    Slot(misa_b200_ctx *ctx, int slot) : c(ctx), k(slot) {
        if (c->prof_on) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, c->stream);
            c->prof_ev[k].push_back(e);
        }
    }
    ~Slot() {
        if (c->prof_on) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, c->stream);
            c->prof_ev[k].push_back(e);
        }
    }
};

static inline int nblk(long long n) { return (int)((n + MISA_BLOCK - 1) / MISA_BLOCK); }

// -------------------------------------------------------------------------------------------------
// shared-memory table kernels (eam_smem.cuh): opt in to the large dynamic shared memory once
// -------------------------------------------------------------------------------------------------
static int census_local(misa_b200_ctx *c);
static int census_fetch(misa_b200_ctx *c);
static void pick_list(const misa_b200_ctx *c, const int *&offs, int &n_off, int *n_near = nullptr);
static bool make_plan(const misa_b200_ctx *c, StagePlan &sp, size_t &smem_bytes);
static inline bool no_vacancy(const misa_b200_ctx *c);
static bool dilute_ok(const misa_b200_ctx *c, const StagePlan &sp, bool accum);
static bool sym_active(const misa_b200_ctx *c, const StagePlan &sp);
static int smem_kernels_init(int optin) {
    static bool done = false;
    if (done) return 0;
#define OPTIN(k) CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 1024)) /* static smem: mbarrier + table directory */
    OPTIN((k_rho_s<true, true, false>)); OPTIN((k_rho_s<true, false, false>)); OPTIN((k_rho_s<true, false, true>));
    OPTIN((k_rho_s<true, true, false, true, false>)); OPTIN((k_rho_s<true, true, false, false, true>)); OPTIN((k_rho_s<true, true, false, true, true>));
    OPTIN((k_force_s<true, false, true, false>)); OPTIN((k_force_s<true, false, false, true>)); OPTIN((k_force_s<true, false, true, true>));
    OPTIN((k_rho_s<false, true, false>)); OPTIN((k_rho_s<false, false, false>)); OPTIN((k_rho_s<false, false, true>));
    OPTIN((k_rho_s<false, true, false, true, false>)); OPTIN((k_force_s<false, false, true, false>));
    OPTIN((k_force_s<true, false>)); OPTIN((k_force_s<true, true>));
    OPTIN((k_force_s<false, false>)); OPTIN((k_force_s<false, true>));
    OPTIN((k_rho_f<true, true, true, false>)); OPTIN((k_rho_f<true, true, false, false>)); OPTIN((k_rho_f<true, true, false, true>));
    OPTIN((k_rho_f<true, false, true, false>)); OPTIN((k_rho_f<true, false, false, false>)); OPTIN((k_rho_f<true, false, false, true>));
    OPTIN((k_rho_f<false, false, true, false>)); OPTIN((k_rho_f<false, false, false, false>)); OPTIN((k_rho_f<false, false, false, true>));
    OPTIN((k_force_f<true, true, false>)); OPTIN((k_force_f<true, true, true>));
    OPTIN((k_force_f<true, false, false>)); OPTIN((k_force_f<true, false, true>));
    OPTIN((k_force_f<false, false, false>)); OPTIN((k_force_f<false, false, true>));
    OPTIN((k_rho_f<true, true, true, false, true>)); OPTIN((k_rho_f<true, true, false, false, true>));
    OPTIN((k_rho_f<true, false, true, false, true>)); OPTIN((k_rho_f<true, false, false, false, true>));
    OPTIN((k_force_f<true, true, false, true>)); OPTIN((k_force_f<true, false, false, true>));
    OPTIN((k_rho_f<true, true, false, false, false, true>)); OPTIN((k_rho_f<true, false, false, false, false, true>));
    OPTIN((k_force_f<true, true, false, false, true>)); OPTIN((k_force_f<true, false, false, false, true>));
    OPTIN((k_rho_a<true, false>)); OPTIN((k_rho_a<false, false>)); OPTIN((k_rho_a<true, true>)); OPTIN((k_rho_a<false, true>));
    OPTIN((k_force_a<true, false>)); OPTIN((k_force_a<false, false>)); OPTIN((k_force_a<true, true>)); OPTIN((k_force_a<false, true>));
#undef OPTIN
    done = true;
    return 0;
}

// -------------------------------------------------------------------------------------------------
// create / destroy
// -------------------------------------------------------------------------------------------------
template <typename T>
static int dmalloc(T **p, size_t n) {
    CU(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}
#include "p2p.cuh"   // needs dmalloc

static int geo_from_domain(const misa_b200_domain *dom, Geo &g) {
    g.nx = dom->sub_box_lattice_size[0]; g.ny = dom->sub_box_lattice_size[1]; g.nz = dom->sub_box_lattice_size[2];
    g.gx = dom->lattice_size_ghost[0]; g.gy = dom->lattice_size_ghost[1]; g.gz = dom->lattice_size_ghost[2];
    g.sxc = g.nx + 2 * g.gx; g.sy = g.ny + 2 * g.gy; g.sz = g.nz + 2 * g.gz;
    for (int k = 0; k < 3; k++) g.lo[k] = dom->sub_box_lattice_low[k];
    g.n_ext = 2LL * g.sxc * g.sy * g.sz;
    g.H = g.n_ext / 2;
    g.n_cells_owned = (long long)g.nx * g.ny * g.nz;
    g.a = dom->lattice_const;
    const double cutoff_radius = dom->lattice_const * dom->cutoff_radius_factor; // reference src/atom.cpp:15
    g.rc2 = cutoff_radius * cutoff_radius;
    g.runaway2 = pow(0.2 * dom->lattice_const, 2.0);                             // reference src/atom.cpp:42
    if (g.nx <= 0 || g.ny <= 0 || g.nz <= 0 || g.gx < 1 || g.gy < 1 || g.gz < 1 || g.n_ext >= (1LL << 31)) return -1;
    // the staged exchange forwards a ghost-wide slab of OWNED sites: a sub-box thinner than its ghost shell
    // would need second-neighbour messages, which neither the reference nor this library sends
    if (g.nx < g.gx || g.ny < g.gy || g.nz < g.gz) return -1;
    return 0;
}

static void build_halo_lists_host(const Geo &g, std::vector<int> send[3][2], std::vector<int> recv[3][2]) {
    // sendlist: comm::fwCommLocalRegion as used at reference src/atom/atom_list.cpp:33-40;
    // recvlist: slabs of LatPackerFirst::onReceive, reference src/pack/lat_particle_packer.cpp:65-76,97-108,128-139.
    const int gh[3] = {2 * g.gx, g.gy, g.gz}, bx[3] = {2 * g.nx, g.ny, g.nz}, ex[3] = {2 * g.sxc, g.sy, g.sz};
    for (int d = 0; d < 3; d++)
        for (int dir = 0; dir < 2; dir++) {
            int slo[3], shi[3], rlo[3], rhi[3];
            for (int k = 0; k < 3; k++) {
                if (k == d) {
                    if (dir == 0) { slo[k] = gh[k]; shi[k] = 2 * gh[k]; rlo[k] = gh[k] + bx[k]; rhi[k] = ex[k]; }
                    else { slo[k] = bx[k]; shi[k] = bx[k] + gh[k]; rlo[k] = 0; rhi[k] = gh[k]; }
                } else if (k < d) { slo[k] = rlo[k] = 0; shi[k] = rhi[k] = ex[k]; }
                else { slo[k] = rlo[k] = gh[k]; shi[k] = rhi[k] = gh[k] + bx[k]; }
            }
            auto fill = [&](std::vector<int> &v, const int *lo, const int *hi) {
                v.clear();
                for (int z = lo[2]; z < hi[2]; z++)
                    for (int y = lo[1]; y < hi[1]; y++)
                        for (int x = lo[0]; x < hi[0]; x++) {
                            const long long idx = ((long long)z * ex[1] + y) * ex[0] + x;
                            v.push_back((int)ref_to_dev(idx, g.H));
                        }
            };
            fill(send[d][dir], slo, shi);
            fill(recv[d][dir], rlo, rhi);
        }
}

// Compose the three staged exchanges into one ghost <- owned map. Every sub-box has this shape and runs these stages, so
// "site b receives what the sender held at site a" chains symbolically: code = direction of the sub-box the value
// ORIGINATES from, (sx+1) + 3 (sy+1) + 9 (sz+1), src = its site there (device index space). Entries come out grouped by
// code, destination ascending inside a group.
static void compose_push_map(const std::vector<int> send[3][2], const std::vector<int> recv[3][2], size_t n_ext, std::vector<int> &dst,
                             std::vector<int> &src, std::vector<int8_t> &codes) {
    std::vector<int> src_of(n_ext), code(n_ext, 13); // 13 = (0+1) + 3*(0+1) + 9*(0+1)
    for (size_t i = 0; i < n_ext; i++) src_of[i] = (int)i;
    static const int mul[3] = {1, 3, 9};
    dst.clear();
    for (int d = 0; d < 3; d++)
        for (int dir = 0; dir < 2; dir++)
            for (size_t i = 0; i < send[d][dir].size(); i++) {
                const int a = send[d][dir][i], b = recv[d][dir][i];
                src_of[b] = src_of[a];
                code[b] = code[a] + mul[d] * (dir == 0 ? 1 : -1);
                dst.push_back(b);
            }
    std::sort(dst.begin(), dst.end(), [&](int a, int b) { return code[a] != code[b] ? code[a] < code[b] : a < b; });
    dst.erase(std::unique(dst.begin(), dst.end()), dst.end());
    src.resize(dst.size());
    codes.resize(dst.size());
    for (size_t i = 0; i < dst.size(); i++) { src[i] = src_of[dst[i]]; codes[i] = (int8_t)code[dst[i]]; }
}
// image shift of group `code` pushed by the sub-box at grid_coord: what the staged path adds hop by hop
// (reference src/pack/lat_particle_packer.cpp:22-32)
static void push_shift(const misa_b200_domain *dom, int code, double shift[3]) {
    const int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
    for (int d = 0; d < 3; d++) {
        shift[d] = 0.0;
        if (s[d] == 1 && dom->grid_coord[d] == 0) shift[d] = dom->meas_global_length[d];
        if (s[d] == -1 && dom->grid_coord[d] == dom->grid_size[d] - 1) shift[d] = -dom->meas_global_length[d];
    }
}
extern "C" int misa_b200_plan_push(const misa_b200_domain *dom, int64_t *dst, int64_t *src, int8_t *code, size_t cap, size_t *n,
                                   double shift[27][3]) {
    REQ(dom && n, MISA_B200_EINVAL, "misa_b200_plan_push: bad argument");
    Geo g;
    REQ(geo_from_domain(dom, g) == 0, MISA_B200_EINVAL, "misa_b200_plan_push: bad sub-box sizes");
    std::vector<int> s[3][2], r[3][2], d_, s_;
    std::vector<int8_t> c_;
    build_halo_lists_host(g, s, r);
    compose_push_map(s, r, (size_t)g.n_ext, d_, s_, c_);
    *n = d_.size();
    for (size_t i = 0; i < std::min(cap, *n); i++) {
        if (dst) dst[i] = dev_to_ref(d_[i], g.H);
        if (src) src[i] = dev_to_ref(s_[i], g.H);
        if (code) code[i] = c_[i];
    }
    if (shift) for (int k = 0; k < 27; k++) push_shift(dom, k, shift[k]);
    return MISA_B200_OK;
}

extern "C" int misa_b200_create(const misa_b200_domain *dom, misa_b200_ctx **out) {
    REQ(dom && out, MISA_B200_EINVAL, "misa_b200_create: null argument");
    if (g_device < 0) TRY(misa_b200_env_init(-1));
    misa_b200_ctx *c = new misa_b200_ctx();
    c->dom = *dom;
    c->device = g_device;
    Geo &g = c->geo;
    if (geo_from_domain(dom, g) != 0) {
        delete c;
        return fail(MISA_B200_EINVAL, "misa_b200_create: bad sub-box sizes");
    }
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const size_t n = (size_t)g.n_ext;
    // x, y, z and df -- the fields the stencil kernels gather from neighbours -- share one allocation so that one
    // texture handle reaches all four (field stride padded to 512 B, the linear-texture base alignment)
    c->xyzd_stride = (long long)((n + 63) / 64 * 64);
    TRY(dmalloc(&c->d_xyzd, 4 * (size_t)c->xyzd_stride));
    CU(cudaMemset(c->d_xyzd, 0, 4 * (size_t)c->xyzd_stride * 8));
    for (int k = 0; k < 3; k++) c->s.x[k] = c->d_xyzd + (size_t)k * c->xyzd_stride;
    c->s.df = c->d_xyzd + 3 * (size_t)c->xyzd_stride;
    for (int k = 0; k < 3; k++) { TRY(dmalloc(&c->s.v[k], n)); TRY(dmalloc(&c->s.f[k], n)); }
    TRY(dmalloc(&c->s.rho, n)); TRY(dmalloc(&c->s.type, n)); TRY(dmalloc(&c->s.id, n)); TRY(dmalloc(&c->s.ulev, n));
    CU(cudaMemset(c->s.ulev, 0xff, n));
    for (int k = 0; k < 3; k++) c->s.sx[k] = nullptr;
    if (c->opt_vac_sentinel) for (int k = 0; k < 3; k++) { TRY(dmalloc(&c->s.sx[k], n)); CU(cudaMemset(c->s.sx[k], 0, n * 8)); }
    {   // hot-cell map (ctx.h): the template marks every cell whose stencil reaches into the ghost shell (ghost levels are not kept)
        const Geo &g = c->geo;
        std::vector<unsigned char> tmpl((size_t)g.H, 1);
        for (int z = 2 * g.gz; z < g.sz - 2 * g.gz; z++)
            for (int y = 2 * g.gy; y < g.sy - 2 * g.gy; y++)
                for (int x = 2 * g.gx; x < g.sxc - 2 * g.gx; x++) tmpl[((size_t)z * g.sy + y) * g.sxc + x] = 0;
        TRY(dmalloc(&c->d_hot, (size_t)g.H)); TRY(dmalloc(&c->d_hot_init, (size_t)g.H));
        CU(cudaMemcpy(c->d_hot_init, tmpl.data(), (size_t)g.H, cudaMemcpyHostToDevice));
        CU(cudaMemset(c->d_hot, 0, (size_t)g.H));
        c->s.hot = c->d_hot;
    }
    for (int k = 0; k < 3; k++) { CU(cudaMemset(c->s.v[k], 0, n * 8)); CU(cudaMemset(c->s.f[k], 0, n * 8)); }
    CU(cudaMemset(c->s.rho, 0, n * 8)); CU(cudaMemset(c->s.type, 0xff, n)); CU(cudaMemset(c->s.id, 0, n * 8));
    {
        cudaResourceDesc rd;
        cudaTextureDesc td;
        memset(&rd, 0, sizeof rd);
        memset(&td, 0, sizeof td);
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = c->d_xyzd;
        rd.res.linear.desc = cudaCreateChannelDesc<int2>();
        rd.res.linear.sizeInBytes = 4 * (size_t)c->xyzd_stride * sizeof(double);
        td.readMode = cudaReadModeElementType;
        if (4 * c->xyzd_stride >= (1LL << 31) || cudaCreateTextureObject(&c->tex_all, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); c->tex_all = 0; }
        double *fields[4] = {c->s.x[0], c->s.x[1], c->s.x[2], c->s.df};
        cudaTextureObject_t *objs[4] = {&c->tex_x[0], &c->tex_x[1], &c->tex_x[2], &c->tex_df};
        for (int k = 0; k < 4; k++) {
            memset(&rd, 0, sizeof rd);
            memset(&td, 0, sizeof td);
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = fields[k];
            rd.res.linear.desc = cudaCreateChannelDesc<int2>();
            rd.res.linear.sizeInBytes = n * sizeof(double);
            td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(objs[k], &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); *objs[k] = 0; }
        }
    }
    // the step's words and the counters share one allocation: [0] activity, [1] dmax2 bits, [2] marking atoms, then the 16
    // int counters -- so that ONE 28-byte memset in front of k_verlet1 zeroes everything a step accumulates
    TRY(dmalloc(&c->d_stepinfo, 3 + 8));
    CU(cudaMemset(c->d_stepinfo, 0, (3 + 8) * sizeof(unsigned long long)));
    c->d_counters = reinterpret_cast<int *>(c->d_stepinfo + 3);
    CU(cudaHostAlloc((void **)&c->h_counters, 16 * sizeof(int), cudaHostAllocMapped));
    memset(c->h_counters, 0, 16 * sizeof(int));
    CU(cudaHostGetDevicePointer((void **)&c->hd_counters, c->h_counters, 0));
    TRY(dmalloc(&c->d_stepinfo_g, 4));
    CU(cudaMemset(c->d_stepinfo_g, 0, 4 * sizeof(unsigned long long)));
    TRY(dmalloc(&c->d_stepinfo_n, 4));
    CU(cudaMemset(c->d_stepinfo_n, 0, 4 * sizeof(unsigned long long)));
    CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    for (cudaEvent_t *e : {&c->ev_v1, &c->ev_act, &c->ev_hx, &c->ev_rho, &c->ev_hdf}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    CU(cudaHostAlloc((void **)&c->h_stepinfo, 4 * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(c->h_stepinfo, 0, 4 * sizeof(unsigned long long));
    CU(cudaHostGetDevicePointer((void **)&c->hd_stepinfo, c->h_stepinfo, 0));
    TRY(dmalloc(&c->d_reduce, 8));
    CU(cudaMallocHost((void **)&c->h_reduce, 8 * sizeof(double)));

    // halo lists
    std::vector<int> send[3][2], recv[3][2];
    build_halo_lists_host(c->geo, send, recv);
    c->all_self = true;
    size_t max_n = 0;
    for (int d = 0; d < 3; d++)
        for (int dir = 0; dir < 2; dir++) {
            HaloList &h = c->halo[d][dir];
            h.n = (int)send[d][dir].size();
            max_n = std::max(max_n, (size_t)h.n);
            TRY(dmalloc(&h.d_send, (size_t)h.n)); TRY(dmalloc(&h.d_recv, (size_t)h.n));
            CU(cudaMemcpy(h.d_send, send[d][dir].data(), h.n * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(h.d_recv, recv[d][dir].data(), h.n * sizeof(int), cudaMemcpyHostToDevice));
            // periodic image shift, reference src/pack/lat_particle_packer.cpp:22-32
            if (dom->grid_coord[d] == 0 && dir == 0) h.shift[d] = dom->meas_global_length[d];
            if (dom->grid_coord[d] == dom->grid_size[d] - 1 && dir == 1) h.shift[d] = -dom->meas_global_length[d];
            if (dom->grid_size[d] != 1) c->all_self = false;
        }
    c->halo_buf_elems = max_n * 4;
    for (int dir = 0; dir < 2; dir++) { TRY(dmalloc(&c->d_sendbuf[dir], c->halo_buf_elems)); TRY(dmalloc(&c->d_recvbuf[dir], c->halo_buf_elems)); }
    {
        // 1x1x1 grid: the origin of every ghost is this sub-box itself (k_ghost_fill_*); otherwise the map read backwards is
        // what this sub-box pushes into its neighbours (p2p.cuh)
        std::vector<int> dst_list, srcs;
        std::vector<int8_t> codes;
        compose_push_map(send, recv, n, dst_list, srcs, codes);
        if (c->all_self) {
            c->h_ghost_dst = dst_list; c->h_ghost_src = srcs; c->h_ghost_code = codes;
            c->n_ghost_map = (int)dst_list.size();
            TRY(dmalloc(&c->d_ghost_dst, dst_list.size())); TRY(dmalloc(&c->d_ghost_src, dst_list.size())); TRY(dmalloc(&c->d_ghost_shift, dst_list.size()));
            CU(cudaMemcpy(c->d_ghost_dst, dst_list.data(), dst_list.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(c->d_ghost_src, srcs.data(), srcs.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(c->d_ghost_shift, codes.data(), codes.size(), cudaMemcpyHostToDevice));
        } else {
            for (int8_t k : codes) c->push_code_used[k] = true;
            c->n_push = (int)dst_list.size();
            TRY(dmalloc(&c->d_push_dst, dst_list.size())); TRY(dmalloc(&c->d_push_src, dst_list.size())); TRY(dmalloc(&c->d_push_code, dst_list.size()));
            CU(cudaMemcpy(c->d_push_dst, dst_list.data(), dst_list.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(c->d_push_src, srcs.data(), srcs.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(c->d_push_code, codes.data(), codes.size(), cudaMemcpyHostToDevice));
        }
    }
    TRY(inter_alloc(c, 1 << 16));
    TRY(dmalloc(&c->d_census, MISA_MAX_TYPES));
    CU(cudaMallocHost((void **)&c->h_census, MISA_MAX_TYPES * sizeof(unsigned long long)));
    CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    CU(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
    TRY(smem_kernels_init(c->smem_optin));
    misa_b200_set_timestep(c, 0.001);
    if (const char *env = getenv("MISA_B200_OPTS")) {   // "name=value,name=value": A/B switches for benches (same library, same kernels)
        std::string e(env);
        size_t pos = 0;
        while (pos < e.size()) {
            const size_t end = std::min(e.find(',', pos), e.size()), eq = e.find('=', pos);
            if (eq != std::string::npos && eq < end) misa_b200_set_option(c, e.substr(pos, eq - pos).c_str(), atoi(e.substr(eq + 1, end - eq - 1).c_str()));
            pos = end + 1;
        }
    }
    *out = c;
    return MISA_B200_OK;
}

extern "C" int misa_b200_destroy(misa_b200_ctx *c) {
    if (!c) return MISA_B200_OK;
    cudaStreamSynchronize(c->stream);
    misa_b200_comm_destroy(c);
    for (int k = 0; k < 3; k++) if (c->tex_x[k]) cudaDestroyTextureObject(c->tex_x[k]);
    if (c->tex_df) cudaDestroyTextureObject(c->tex_df);
    if (c->tex_all) cudaDestroyTextureObject(c->tex_all);
    cudaFree(c->d_xyzd);
    for (int k = 0; k < 3; k++) { cudaFree(c->s.v[k]); cudaFree(c->s.f[k]); }
    cudaFree(c->s.rho); cudaFree(c->s.type); cudaFree(c->s.id); cudaFree(c->s.ulev); cudaFree(c->d_hot); cudaFree(c->d_hot_init);
    for (int k = 0; k < 3; k++) cudaFree(c->s.sx[k]);
    cudaFree(c->d_aos); cudaFree(c->d_off_full); cudaFree(c->d_off_levels); cudaFree(c->d_off_full_addr); cudaFree(c->d_off_levels_addr);
    cudaFree(c->d_stepinfo); cudaFreeHost(c->h_stepinfo); cudaFree(c->d_stepinfo_g); cudaFree(c->d_stepinfo_n);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    for (cudaEvent_t e : {c->ev_v1, c->ev_act, c->ev_hx, c->ev_rho, c->ev_hdf}) if (e) cudaEventDestroy(e);
    cudaFree(c->d_elec); cudaFree(c->d_embed); cudaFree(c->d_phi); cudaFree(c->d_herm); cudaFree(c->d_c34); cudaFree(c->d_c56); cudaFree(c->d_mono); cudaFree(c->d_ea); cudaFree(c->d_es);
    cudaFree(c->d_census); cudaFreeHost(c->h_census); cudaFree(c->d_pmax); cudaFree(c->d_ptmp);
    for (int d = 0; d < 3; d++) for (int dir = 0; dir < 2; dir++) { cudaFree(c->halo[d][dir].d_send); cudaFree(c->halo[d][dir].d_recv); }
    for (int dir = 0; dir < 2; dir++) { cudaFree(c->d_sendbuf[dir]); cudaFree(c->d_recvbuf[dir]); }
    cudaFree(c->d_ghost_dst); cudaFree(c->d_ghost_src); cudaFree(c->d_ghost_shift);
    cudaFree(c->d_slab_dst); cudaFree(c->d_slab_src); cudaFree(c->d_slab_code); cudaFree(c->d_aos_out);
    if (c->s_up) cudaStreamDestroy(c->s_up);
    if (c->s_dn) cudaStreamDestroy(c->s_dn);
    for (cudaEvent_t e : c->ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_out) cudaEventDestroy(e);
    cudaFreeHost(c->h_counters); cudaFree(c->d_reduce); cudaFreeHost(c->h_reduce);
    cudaFree(c->d_minor); cudaFree(c->d_minor_count); cudaFree(c->d_mcount); cudaFree(c->d_mentry);
    cudaFree(c->d_lo_tab); cudaFree(c->d_pair);
    cudaFree(c->d_push_dst); cudaFree(c->d_push_src); cudaFree(c->d_push_code); cudaFree(c->d_flags);
    if (c->h_p2p_err) cudaFreeHost(c->h_p2p_err);
    cudaFree(c->d_dump); cudaFree(c->d_dump_base); cudaFree(c->d_dump_total); cudaFree(c->d_dump_count); cudaFreeHost(c->h_dump_total);
    inter_free(c);
    for (int k = 0; k < MISA_B200_K_COUNT; k++) for (auto e : c->prof_ev[k]) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
    return MISA_B200_OK;
}

// -------------------------------------------------------------------------------------------------
// neighbour stencil
// -------------------------------------------------------------------------------------------------
// reference offset (linear, doubled-x) of parity p -> device-index offset; see ctx.h
static int ref_off_to_dev(long long off, int p, long long H) {
    if (p == 0) return (int)((off >> 1) + (off & 1) * H);
    if ((off & 1) == 0) return (int)(off / 2);
    return (int)((off + 1) / 2 - H);
}
// decode a reference offset into (dx, dy, dz) and return the squared SITE separation in units of a^2
static void off_decode(long long off, const Geo &g, long long &dx, long long &dy, long long &dz) {
    const long long sx = 2LL * g.sxc, sy = g.sy;
    dx = ((off % sx) + sx + sx / 2) % sx - sx / 2;
    const long long r = (off - dx) / sx;
    dy = ((r % sy) + sy + sy / 2) % sy - sy / 2;
    dz = (r - dy) / sy;
}
static double off_site_r2(long long off, int p, const Geo &g) {
    long long dx, dy, dz;
    off_decode(off, g, dx, dy, dz);
    const double half = (dx & 1) ? (p == 0 ? 0.5 : -0.5) : 0.0; // odd dx switches sub-lattice
    const double X = 0.5 * (double)dx, Y = (double)dy + half, Z = (double)dz + half;
    return X * X + Y * Y + Z * Z;
}

// Everything the stencil kernels need from the reference's four offset vectors, planned on the HOST (no device): the
// distance-sorted full list per parity, its near / half split, the prefix lengths per pair displacement level, the
// lower-neighbour table of the pair-symmetric passes and the second generation's per-level lists.
struct StencilPlan {
    int n_full = 0, near_full = 0, n_half = 0;
    std::vector<int> full;                  // [2][n_full] device-index offsets, sorted
    std::vector<long long> sorted_ref[2];   // the same entries in the reference's index space
    std::vector<double> site_r2[2];         // squared SITE separation of each entry, units of a^2
    int prefix_n[misa_b200_ctx::kPairLevels] = {0};
    std::vector<int2> lo;                   // [2][n_half]
    int sym_lo[3] = {0, 0, 0}, sym_hi[3] = {0, 0, 0};
    std::vector<int> levels;
    size_t level_ofs[misa_b200_ctx::kLevels] = {0};
    int level_n[misa_b200_ctx::kLevels] = {0}, level_near[misa_b200_ctx::kLevels] = {0};
};
static int plan_stencil(const Geo &g, const double crf, const std::vector<int64_t> ref_off[2], StencilPlan &sp) {
    REQ(ref_off[0].size() == ref_off[1].size() && !ref_off[0].empty(), MISA_B200_EINVAL,
        "neighbour offsets: even/odd lists must be non-empty and of equal length");
    const int n_full = sp.n_full = (int)ref_off[0].size();
    sp.full.assign(2 * (size_t)n_full, 0);
    // device lists are PARTITIONED by site separation (the sums do not depend on the order; the reference order stays
    // in ref_off for the ABI and the inter-atom kernels): the leading `near` entries are in range for practically every
    // atom and are evaluated without a branch (eam_fast.cuh)
    {
        // near: shells that reach inside the cutoff at thermal displacements -- up to 0.05a OUTSIDE it, which takes in the
        // <200> shell of bcc (2.0a against crf = 1.961: 16 % of those pairs are in range at 300 K, i.e. practically every
        // warp-wide vote of the far loop was true for them anyway)
        const double near_lim = crf + 0.05;
        int near[2] = {0, 0}, half[2] = {0, 0};
        for (int p = 0; p < 2; p++) {
            std::vector<int> perm(n_full);
            std::vector<double> r2(n_full);
            for (int q = 0; q < n_full; q++) { perm[q] = q; r2[q] = off_site_r2(ref_off[p][q], p, g); }
            // near group first; inside a group keep the reference's order (ascending memory offset: consecutive
            // iterations then touch neighbouring lines, which is what keeps the L1 hit rate up)
            // ... and inside the near group the UPPER half first: the pair-symmetric passes (eam_sym.cuh) evaluate only
            // those. "Upper" is the sign of the real-space site separation (z, then y, then x) -- the same vectors for
            // both sub-lattices of the Bravais lattice, whereas the reference's index-order half list
            // (neighbour_index.inl:79-92) splits the near shells 21 : 37 between even and odd x. Any antisymmetric
            // choice visits every pair exactly once; the per-atom sums do not depend on it.
            auto upper = [&](int q) {
                long long dx, dy, dz;
                off_decode(ref_off[p][q], g, dx, dy, dz);
                const int h2 = (dx & 1) ? (p == 0 ? 1 : -1) : 0;             // twice the half-cell shift of an odd dx
                const long long X = dx, Y = 2 * dy + h2, Z = 2 * dz + h2;     // twice the separation in units of a
                return Z > 0 || (Z == 0 && (Y > 0 || (Y == 0 && X > 0)));
            };
            auto cls = [&](int q) { return r2[q] < near_lim * near_lim ? (upper(q) ? 0 : 1) : 2; };
            // far group: by site separation (shell by shell, reference order inside a shell), so that every pruned list is
            // a prefix of this one (eam_smem.cuh:list_len)
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
                if (cls(a) != cls(b)) return cls(a) < cls(b);
                return cls(a) == 2 && r2[a] < r2[b];
            });
            sp.site_r2[p].resize(n_full);
            sp.sorted_ref[p].resize(n_full);
            for (int q = 0; q < n_full; q++) {
                sp.full[(size_t)p * n_full + q] = ref_off_to_dev(ref_off[p][perm[q]], p, g.H);
                sp.site_r2[p][q] = r2[perm[q]];
                sp.sorted_ref[p][q] = ref_off[p][perm[q]];
                if (cls(perm[q]) < 2) near[p]++;
                if (cls(perm[q]) == 0) half[p]++;
            }
        }
        sp.near_full = std::min(near[0], near[1]);
        // pair-symmetric passes: both parities must agree on the split, every lower near offset of parity pj must be the
        // mirror of an upper near offset of the neighbour's parity, and the sites around the owned box that own such a
        // pair must exist in the ghost shell
        sp.n_half = 0;
        sp.lo.clear();
        if (near[0] == near[1] && half[0] == half[1] && 2 * half[0] == near[0] && half[0] > 0) {
            const int nh = half[0];
            std::vector<int2> lo(2 * (size_t)nh);
            bool ok = true;
            int elo[3] = {0, 0, 0}, ehi[3] = {0, 0, 0};
            for (int pj = 0; pj < 2 && ok; pj++)
                for (int m = 0; m < nh && ok; m++) {
                    const long long o = sp.sorted_ref[pj][nh + m];         // lower near entry: i = j + o
                    const int pi = pj ^ (int)(o & 1);
                    int slot = -1;
                    for (int k = 0; k < nh; k++) if (sp.sorted_ref[pi][k] == -o) slot = k;
                    if (slot < 0) { ok = false; break; }
                    lo[(size_t)pj * nh + m] = make_int2(sp.full[(size_t)pj * n_full + nh + m], slot);
                    long long dx, dy, dz;
                    off_decode(o, g, dx, dy, dz);
                    const long long dc[3] = {(pj + dx) >> 1, dy, dz};      // cells from j to its lower neighbour
                    for (int k = 0; k < 3; k++) { elo[k] = (int)std::max<long long>(elo[k], -dc[k]); ehi[k] = (int)std::max<long long>(ehi[k], dc[k]); }
                }
            const int gh[3] = {g.gx, g.gy, g.gz};
            for (int k = 0; k < 3; k++) ok = ok && elo[k] <= gh[k] && ehi[k] <= gh[k];
            if (ok) {
                sp.n_half = nh;
                for (int k = 0; k < 3; k++) { sp.sym_lo[k] = elo[k]; sp.sym_hi[k] = ehi[k]; }
                sp.lo = lo;
            }
        }
    }
    // prefix lengths of the sorted full list by PAIR displacement level: sites closer than (crf + 0.01 L) a
    for (int L = 0; L < misa_b200_ctx::kPairLevels; L++) {
        const double lim = crf + 0.01 * L + 1e-9;
        int n[2] = {0, 0};
        for (int p = 0; p < 2; p++)
            for (int q = 0; q < n_full; q++)
                if (q < sp.near_full || sp.site_r2[p][q] < lim * lim) n[p] = q + 1;
        sp.prefix_n[L] = std::max(n[0], n[1]);
    }
    // pairs of LATTICE atoms can only be within the cutoff if their sites are closer than crf + 2*dmax/a;
    // atom::decide keeps dmax <= 0.2a (reference src/atom.cpp:42), the device measures the actual value.
    sp.levels.clear();
    for (int L = 0; L < misa_b200_ctx::kLevels; L++) {
        const double lim = crf + 0.02 * L + 1e-9;
        int n[2] = {0, 0};
        sp.level_ofs[L] = sp.levels.size();
        for (int p = 0; p < 2; p++)
            for (int q = 0; q < n_full; q++)
                if (sp.site_r2[p][q] < lim * lim) { sp.levels.push_back(sp.full[(size_t)p * n_full + q]); n[p]++; }
        REQ(n[0] == n[1], MISA_B200_EINVAL, "neighbour offsets: pruned lists differ in length");
        sp.level_n[L] = n[0];
        sp.level_near[L] = std::min(sp.near_full, n[0]);
    }
    return MISA_B200_OK;
}
static int upload_offsets(misa_b200_ctx *c) {
    StencilPlan sp;
    const std::vector<int64_t> ref[2] = {c->ref_off[0], c->ref_off[1]};
    TRY(plan_stencil(c->geo, c->dom.cutoff_radius_factor, ref, sp));
    c->n_full = sp.n_full; c->near_full = sp.near_full; c->n_half = sp.n_half;
    for (int k = 0; k < 3; k++) { c->sym_lo[k] = sp.sym_lo[k]; c->sym_hi[k] = sp.sym_hi[k]; }
    for (int L = 0; L < misa_b200_ctx::kPairLevels; L++) c->prefix_n[L] = sp.prefix_n[L];
    for (int L = 0; L < misa_b200_ctx::kLevels; L++) { c->level_ofs[L] = sp.level_ofs[L]; c->level_n[L] = sp.level_n[L]; c->level_near[L] = sp.level_near[L]; }
    cudaFree(c->d_lo_tab); c->d_lo_tab = nullptr;
    if (sp.n_half > 0) {
        TRY(dmalloc(&c->d_lo_tab, sp.lo.size()));
        CU(cudaMemcpy(c->d_lo_tab, sp.lo.data(), sp.lo.size() * sizeof(int2), cudaMemcpyHostToDevice));
    }
    cudaFree(c->d_off_full); cudaFree(c->d_off_levels);
    c->d_off_full = c->d_off_levels = nullptr;
    TRY(dmalloc(&c->d_off_full, sp.full.size()));
    TRY(dmalloc(&c->d_off_levels, sp.levels.size()));
    CU(cudaMemcpy(c->d_off_full, sp.full.data(), sp.full.size() * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_off_levels, sp.levels.data(), sp.levels.size() * sizeof(int), cudaMemcpyHostToDevice));
    {   // address-ordered copies (LevelSel::levels_addr): same segments, each sorted ascending
        std::vector<int> fa(sp.full), la(sp.levels);
        for (int p = 0; p < 2; p++) std::sort(fa.begin() + (size_t)p * sp.n_full, fa.begin() + (size_t)(p + 1) * sp.n_full);
        for (int L = 0; L < misa_b200_ctx::kLevels; L++)
            for (int p = 0; p < 2; p++)
                std::sort(la.begin() + sp.level_ofs[L] + (size_t)p * sp.level_n[L], la.begin() + sp.level_ofs[L] + (size_t)(p + 1) * sp.level_n[L]);
        cudaFree(c->d_off_full_addr); cudaFree(c->d_off_levels_addr);
        c->d_off_full_addr = c->d_off_levels_addr = nullptr;
        TRY(dmalloc(&c->d_off_full_addr, fa.size()));
        TRY(dmalloc(&c->d_off_levels_addr, la.size()));
        CU(cudaMemcpy(c->d_off_full_addr, fa.data(), fa.size() * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_off_levels_addr, la.data(), la.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    c->have_off = true;
    return MISA_B200_OK;
}

extern "C" int misa_b200_set_neighbour_offsets(misa_b200_ctx *c, const int64_t *even, size_t n_even, const int64_t *odd,
                                               size_t n_odd, const int64_t *half_even, size_t n_half_even,
                                               const int64_t *half_odd, size_t n_half_odd) {
    REQ(c && even && odd, MISA_B200_EINVAL, "misa_b200_set_neighbour_offsets: null argument");
    c->ref_off[0].assign(even, even + n_even);
    c->ref_off[1].assign(odd, odd + n_odd);
    if (half_even) c->ref_off[2].assign(half_even, half_even + n_half_even);
    if (half_odd) c->ref_off[3].assign(half_odd, half_odd + n_half_odd);
    return upload_offsets(c);
}

// NeighbourIndex<T>::make, reference src/atom/neighbour_index.inl:13-76 (C++ % keeps the sign of xIndex)
static void make_offsets_host(const Geo &g, int cut_lattice, double crf, std::vector<int64_t> out[4]) {
    const long long sx = 2LL * g.sxc, sy = g.sy;
    const double lim = crf + 2 * 0.5; // config::nei_lat_cutoff, reference src/md_building_config.h.in:24-29
    for (int v = 0; v < 4; v++) out[v].clear();
    for (int p = 0; p < 2; p++) {
        const double sgn = p == 0 ? 1.0 : -1.0;
        for (long long zi = -cut_lattice - 1; zi <= cut_lattice + 1; zi++)
            for (long long yi = -cut_lattice - 1; yi <= cut_lattice + 1; yi++)
                for (long long xi = -2 * cut_lattice - 2; xi <= 2 * cut_lattice + 2; xi++) {
                    const double z = (double)zi + sgn * (((double)(xi % 2)) / 2);
                    const double y = (double)yi + sgn * (((double)(xi % 2)) / 2);
                    const double x = ((double)xi) / 2;
                    const double r = x * x + y * y + z * z;
                    if (r < lim * lim && r > 0) {
                        const bool neg_odd = xi < 0 && xi % 2 != 0;
                        const long long iy = neg_odd ? yi - (long long)sgn : yi, iz = neg_odd ? zi - (long long)sgn : zi;
                        const long long off = (iz * sy + iy) * sx + xi;
                        out[p].push_back(off);
                        const bool positive = z > 0 || (z == 0 && (y > 0 || (y == 0 && x > 0)));
                        if (positive) out[2 + p].push_back(off);
                    }
                }
    }
}

extern "C" int misa_b200_make_neighbour_offsets(misa_b200_ctx *c, int cut_lattice, double crf) {
    REQ(c, MISA_B200_EINVAL, "null ctx");
    make_offsets_host(c->geo, cut_lattice, crf, c->ref_off);
    return upload_offsets(c);
}

// ---- host-only planning entry points (no CUDA device needed): the integer side of the path, exposed so that
//      lattice indexing and halo ownership can be checked bit-exactly on any machine -------------------------
extern "C" int misa_b200_plan_offsets(const misa_b200_domain *dom, int cut_lattice, double crf, int which, int64_t *out,
                                      size_t cap, size_t *n) {
    REQ(dom && n && which >= 0 && which < 4, MISA_B200_EINVAL, "misa_b200_plan_offsets: bad argument");
    Geo g;
    REQ(geo_from_domain(dom, g) == 0, MISA_B200_EINVAL, "misa_b200_plan_offsets: bad sub-box sizes");
    std::vector<int64_t> v[4];
    make_offsets_host(g, cut_lattice, crf, v);
    *n = v[which].size();
    if (out) for (size_t i = 0; i < std::min(cap, *n); i++) out[i] = v[which][i];
    return MISA_B200_OK;
}

extern "C" int misa_b200_plan_stencil(const misa_b200_domain *dom, int cut_lattice, double crf, int parity, int64_t *sorted, double *site_r2,
                                      size_t cap, size_t *n, int32_t *n_near, int32_t *n_half, int32_t prefix[41], int32_t *lower_slot) {
    REQ(dom && n && parity >= 0 && parity < 2, MISA_B200_EINVAL, "misa_b200_plan_stencil: bad argument");
    Geo g;
    REQ(geo_from_domain(dom, g) == 0, MISA_B200_EINVAL, "misa_b200_plan_stencil: bad sub-box sizes");
    std::vector<int64_t> v[4];
    make_offsets_host(g, cut_lattice, crf, v);
    const std::vector<int64_t> ref[2] = {v[0], v[1]};
    StencilPlan sp;
    TRY(plan_stencil(g, crf, ref, sp));
    *n = (size_t)sp.n_full;
    for (size_t i = 0; i < std::min(cap, *n); i++) {
        if (sorted) sorted[i] = sp.sorted_ref[parity][i];
        if (site_r2) site_r2[i] = sp.site_r2[parity][i];
    }
    if (n_near) *n_near = sp.near_full;
    if (n_half) *n_half = sp.n_half;
    if (prefix) for (int L = 0; L < misa_b200_ctx::kPairLevels; L++) prefix[L] = sp.prefix_n[L];
    if (lower_slot) for (int m = 0; m < sp.n_half; m++) lower_slot[m] = sp.lo[(size_t)parity * sp.n_half + m].y;
    return MISA_B200_OK;
}
extern "C" int misa_b200_plan_halo(const misa_b200_domain *dom, int dim, int dir, int64_t *send, int64_t *recv, size_t cap,
                                   size_t *n, double shift[3]) {
    REQ(dom && n && dim >= 0 && dim < 3 && dir >= 0 && dir < 2, MISA_B200_EINVAL, "misa_b200_plan_halo: bad argument");
    Geo g;
    REQ(geo_from_domain(dom, g) == 0, MISA_B200_EINVAL, "misa_b200_plan_halo: bad sub-box sizes");
    std::vector<int> s[3][2], r[3][2];
    build_halo_lists_host(g, s, r);
    *n = s[dim][dir].size();
    for (size_t i = 0; i < std::min(cap, *n); i++) {
        if (send) send[i] = dev_to_ref(s[dim][dir][i], g.H);
        if (recv) recv[i] = dev_to_ref(r[dim][dir][i], g.H);
    }
    if (shift) { // periodic image shift, reference src/pack/lat_particle_packer.cpp:22-32
        shift[0] = shift[1] = shift[2] = 0.0;
        if (dom->grid_coord[dim] == 0 && dir == 0) shift[dim] = dom->meas_global_length[dim];
        if (dom->grid_coord[dim] == dom->grid_size[dim] - 1 && dir == 1) shift[dim] = -dom->meas_global_length[dim];
    }
    return MISA_B200_OK;
}

extern "C" int misa_b200_get_neighbour_offsets(misa_b200_ctx *c, int which, int64_t *out, size_t cap, size_t *n) {
    REQ(c && which >= 0 && which < 4 && n, MISA_B200_EINVAL, "misa_b200_get_neighbour_offsets: bad argument");
    *n = c->ref_off[which].size();
    if (out) for (size_t i = 0; i < std::min(cap, *n); i++) out[i] = c->ref_off[which][i];
    return MISA_B200_OK;
}

// -------------------------------------------------------------------------------------------------
// potential tables
// -------------------------------------------------------------------------------------------------
static int upload_tables(const misa_b200_table *t, int count, double **dptr, int *n_out, double *inv_out, const char *what) {
    const int n = t[0].n;
    for (int i = 0; i < count; i++)
        REQ(t[i].spline && t[i].n == n && t[i].inv_dx == t[0].inv_dx, MISA_B200_EINVAL,
            std::string("misa_b200_set_potential: ") + what + " tables must share one grid");
    std::vector<double> host((size_t)count * (n + 1) * MISA_ROW, 0.0);
    for (int i = 0; i < count; i++)
        for (int m = 0; m <= n; m++)
            for (int k = 0; k < 7; k++) host[((size_t)i * (n + 1) + m) * MISA_ROW + k] = t[i].spline[(size_t)m * 7 + k];
    cudaFree(*dptr);
    *dptr = nullptr;
    TRY(dmalloc(dptr, host.size()));
    CU(cudaMemcpy(*dptr, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
    *n_out = n;
    *inv_out = t[0].inv_dx;
    return MISA_B200_OK;
}

// Hermite copies (value, knot slope) = columns 6 and 5 of every r-table row, and the host-side proof that the
// caller's rows are what eam_smem.cuh rebuilds from them: s4 = 3dv - 2d0 - d1, s3 = d0 + d1 - 2dv (rows 1..n-1),
// s2 = s5/dx, s1 = 2 s4/dx, s0 = 3 s3/dx (oracle/pot.c:table_build / LAMMPS array2spline).
static int build_hermite(misa_b200_ctx *c, const misa_b200_table *elec, const misa_b200_table *phi) {
    const int nt = c->tab.n_types, n = c->tab.n_r, ntab = nt + nt * nt;
    std::vector<double> h((size_t)ntab * (n + 1) * 2, 0.0);
    bool ok = true;
    for (int t = 0; t < ntab; t++) {
        const double *sp = (t < nt ? elec[t] : phi[t - nt]).spline;
        const double inv_dx = c->tab.inv_dr;
        double scale = 0.0;
        for (int m = 1; m <= n; m++) scale = std::max(scale, fabs(sp[(size_t)m * 7 + 6]));
        const double tol = 1e-12 * std::max(scale, 1e-300);
        for (int m = 0; m <= n; m++) {
            h[((size_t)t * (n + 1) + m) * 2 + 0] = sp[(size_t)m * 7 + 6];
            h[((size_t)t * (n + 1) + m) * 2 + 1] = sp[(size_t)m * 7 + 5];
        }
        for (int m = 1; m <= n - 1 && ok; m++) {
            const double *a = sp + (size_t)m * 7, *b = a + 7;
            const double dv = b[6] - a[6];
            const double s4 = 3.0 * dv - 2.0 * a[5] - b[5], s3 = a[5] + b[5] - 2.0 * dv;
            if (fabs(s4 - a[4]) > tol || fabs(s3 - a[3]) > tol || fabs(a[2] - a[5] * inv_dx) > tol * inv_dx ||
                fabs(a[1] - 2.0 * a[4] * inv_dx) > tol * inv_dx || fabs(a[0] - 3.0 * a[3] * inv_dx) > tol * inv_dx)
                ok = false;
        }
    }
    cudaFree(c->d_herm);
    c->d_herm = nullptr;
    TRY(dmalloc(&c->d_herm, h.size() / 2));
    CU(cudaMemcpy(c->d_herm, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    {   // the value cubic of every r table in the reference's own monomial form, as two half rows (eam_fast.cuh, EAM_MONO_RHO)
        std::vector<double> a((size_t)ntab * (n + 1) * 2, 0.0), b(a.size(), 0.0);
        for (int t = 0; t < ntab; t++) {
            const double *sp = (t < nt ? elec[t] : phi[t - nt]).spline;
            for (int m = 0; m <= n; m++) {
                const size_t o = ((size_t)t * (n + 1) + m) * 2;
                a[o] = sp[(size_t)m * 7 + 3]; a[o + 1] = sp[(size_t)m * 7 + 4];
                b[o] = sp[(size_t)m * 7 + 5]; b[o + 1] = sp[(size_t)m * 7 + 6];
            }
        }
        cudaFree(c->d_c34); cudaFree(c->d_c56); cudaFree(c->d_mono); cudaFree(c->d_ea); cudaFree(c->d_es);
        c->d_c34 = c->d_c56 = c->d_ea = nullptr;
        c->d_mono = c->d_es = nullptr;
        {   // slope-only lookups (force kernel, embedding term): (s_m, v_{m+1} - v_m) rows and the dense slopes
            std::vector<double> ea((size_t)ntab * (n + 1) * 2, 0.0), es((size_t)ntab * (n + 1) + 2, 0.0);
            for (int t = 0; t < ntab; t++) {
                const double *sp = (t < nt ? elec[t] : phi[t - nt]).spline;
                for (int m = 0; m <= n; m++) {
                    const size_t o = (size_t)t * (n + 1) + m;
                    ea[2 * o] = sp[(size_t)m * 7 + 5];
                    ea[2 * o + 1] = m < n ? sp[(size_t)(m + 1) * 7 + 6] - sp[(size_t)m * 7 + 6] : 0.0;   // the subtraction the kernels did per pair, same doubles
                    es[o] = sp[(size_t)m * 7 + 5];
                }
            }
            TRY(dmalloc(&c->d_ea, ea.size() / 2));
            TRY(dmalloc(&c->d_es, es.size()));
            CU(cudaMemcpy(c->d_ea, ea.data(), ea.size() * sizeof(double), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(c->d_es, es.data(), es.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        {
            std::vector<double> w((size_t)ntab * (n + 1) * 4, 0.0);
            for (int t = 0; t < ntab; t++) {
                const double *sp = (t < nt ? elec[t] : phi[t - nt]).spline;
                for (int m = 0; m <= n; m++)
                    for (int k = 0; k < 4; k++) w[((size_t)t * (n + 1) + m) * 4 + k] = sp[(size_t)m * 7 + 3 + k];
            }
            TRY(dmalloc(&c->d_mono, w.size()));
            CU(cudaMemcpy(c->d_mono, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        TRY(dmalloc(&c->d_c34, a.size() / 2)); TRY(dmalloc(&c->d_c56, b.size() / 2));
        CU(cudaMemcpy(c->d_c34, a.data(), a.size() * sizeof(double), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_c56, b.data(), b.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    c->hermite_ok = ok;
    return 0;
}

extern "C" int misa_b200_set_potential(misa_b200_ctx *c, int n_types, const misa_b200_table *elec,
                                       const misa_b200_table *embed, const misa_b200_table *phi) {
    REQ(c && elec && embed && phi && n_types >= 1 && n_types <= MISA_MAX_TYPES, MISA_B200_EINVAL,
        "misa_b200_set_potential: bad argument");
    DevTables &tb = c->tab;
    tb.n_types = n_types;
    int n_phi;
    double inv_phi;
    TRY(upload_tables(elec, n_types, &c->d_elec, &tb.n_r, &tb.inv_dr, "electron-density"));
    TRY(upload_tables(embed, n_types, &c->d_embed, &tb.n_rho, &tb.inv_drho, "embedding"));
    TRY(upload_tables(phi, n_types * n_types, &c->d_phi, &n_phi, &inv_phi, "pair"));
    REQ(n_phi == tb.n_r && inv_phi == tb.inv_dr, MISA_B200_EINVAL,
        "misa_b200_set_potential: pair and electron-density tables must share the r grid (setfl format)");
    tb.elec = c->d_elec; tb.embed = c->d_embed; tb.phi = c->d_phi;
    TRY(build_hermite(c, elec, phi));
    c->have_pot = true;
    return MISA_B200_OK;
}

// -------------------------------------------------------------------------------------------------
// transfers
// -------------------------------------------------------------------------------------------------
static int ensure_aos(misa_b200_ctx *c) {
    if (!c->d_aos) TRY(dmalloc(&c->d_aos, (size_t)c->geo.n_ext * 104));
    return 0;
}
// the owned box of a ghost-extended AoS array as one pitched 3-D copy: rows of 2*nx records (20.8 KB at 100 cells)
static int copy_owned_box(misa_b200_ctx *c, void *dst, const void *src, cudaMemcpyKind kind) {
    const Geo &g = c->geo;
    const size_t rec = 104, pitch = 2 * (size_t)g.sxc * rec;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), pitch, pitch, (size_t)g.sy);
    p.dstPtr = make_cudaPitchedPtr(dst, pitch, pitch, (size_t)g.sy);
    p.srcPos = p.dstPos = make_cudaPos(2 * (size_t)g.gx * rec, (size_t)g.gy, (size_t)g.gz);
    p.extent = make_cudaExtent(2 * (size_t)g.nx * rec, (size_t)g.ny, (size_t)g.nz);
    p.kind = kind;
    CU(cudaMemcpy3DAsync(&p, c->stream));
    return 0;
}
static int h2d_aos(misa_b200_ctx *c, const void *atoms, int fields, int owned_only = 0) {
    TRY(ensure_aos(c));
    c->minor_valid = false; // the host may have put any species anywhere
    if (owned_only) TRY(copy_owned_box(c, c->d_aos, atoms, cudaMemcpyHostToDevice));
    else CU(cudaMemcpyAsync(c->d_aos, atoms, (size_t)c->geo.n_ext * 104, cudaMemcpyHostToDevice, c->stream));
    Slot sl(c, MISA_B200_K_XFER);
    k_aos_to_soa<<<nblk(c->geo.n_ext), MISA_BLOCK, 0, c->stream>>>(c->geo.n_ext, c->geo.H, (const unsigned long long *)c->d_aos, c->s, fields, c->geo, owned_only);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
// owned_only: 1 = convert only owned records but copy the whole array back (compat hooks: the host's ghost records pass
// through the staging copy untouched); 2 = also copy only the owned box (the host's ghost records are not written at all)
static int d2h_aos(misa_b200_ctx *c, void *atoms, int fields, int owned_only) {
    TRY(ensure_aos(c));
    {
        Slot sl(c, MISA_B200_K_XFER);
        k_soa_to_aos<<<nblk(c->geo.n_ext), MISA_BLOCK, 0, c->stream>>>(c->geo, (unsigned long long *)c->d_aos, c->s, fields, owned_only);
        c->launches++;
    }
    CU(cudaGetLastError());
    if (owned_only == 2) TRY(copy_owned_box(c, atoms, c->d_aos, cudaMemcpyDeviceToHost));
    else CU(cudaMemcpyAsync(atoms, c->d_aos, (size_t)c->geo.n_ext * 104, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int misa_b200_site_count(misa_b200_ctx *c, size_t *n_sites) {
    REQ(c && n_sites, MISA_B200_EINVAL, "null argument");
    *n_sites = (size_t)c->geo.n_ext;
    return 0;
}
extern "C" int misa_b200_host_register(void *ptr, size_t bytes) {
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return 0;
}
extern "C" int misa_b200_host_unregister(void *ptr) {
    CU(cudaHostUnregister(ptr));
    return 0;
}

extern "C" int misa_b200_upload_atoms(misa_b200_ctx *c, const void *atoms) {
    REQ(c && atoms, MISA_B200_EINVAL, "misa_b200_upload_atoms: null argument");
    TRY(h2d_aos(c, atoms, F_ALL));
    TRY(census_local(c));
    TRY(census_fetch(c));
    c->have_atoms = true;
    c->dmax_valid = false;
    c->pmax_valid = false;
    return 0;
}
extern "C" int misa_b200_download_atoms(misa_b200_ctx *c, void *atoms) {
    REQ(c && atoms, MISA_B200_EINVAL, "misa_b200_download_atoms: null argument");
    REQ(c->have_atoms, MISA_B200_ESTATE, "misa_b200_download_atoms: nothing uploaded");
    // the staging copy may be stale or absent: rebuild every field from the SoA
    TRY(ensure_aos(c));
    CU(cudaMemsetAsync(c->d_aos, 0, (size_t)c->geo.n_ext * 104, c->stream));
    return d2h_aos(c, atoms, F_ALL, 0);
}

// A flag wait of the ghost push timed out (p2p.cuh): tell the caller ONCE, clear the word, and keep the push path off --
// the epochs of the sub-boxes no longer agree -- until misa_b200_comm_init re-synchronises them.
static int p2p_check(misa_b200_ctx *c) {
    if (c->h_p2p_err && *c->h_p2p_err != 0) {
        const unsigned int what = *c->h_p2p_err;
        *c->h_p2p_err = 0;
        c->p2p_fault = true;
        c->p2p_last_error = what;
        return fail(MISA_B200_ENCCL, "ghost push over peer memory: a neighbouring sub-box did not answer within " + std::to_string(c->opt_p2p_timeout_s) +
                                         " s (flag " + std::to_string(what) + "); nothing was stored into its ghosts. Call misa_b200_comm_init again to re-synchronise");
    }
    REQ(!c->p2p_fault, MISA_B200_ESTATE, "ghost push timed out earlier: call misa_b200_comm_init again before stepping");
    return 0;
}
extern "C" int misa_b200_sync(misa_b200_ctx *c) {
    REQ(c, MISA_B200_EINVAL, "null ctx");
    CU(cudaStreamSynchronize(c->stream));
    return p2p_check(c);
}

extern "C" int misa_b200_set_timestep(misa_b200_ctx *c, double dt) {
    REQ(c, MISA_B200_EINVAL, "null ctx");
    c->dt = dt; // NewtonMotion::preComputeDtInv2m, reference src/newton_motion.cpp:19-28
    for (int i = 0; i < MISA_MAX_TYPES; i++) {
        const double dt_halve = 0.5 * dt * kFtm2v;
        c->dt_inv_m[i] = dt_halve / kMass[i];
    }
    return 0;
}

// vacant sites get their stale positions back into x and the side arrays go (Soa::sx)
static int drop_sentinel(misa_b200_ctx *c) {
    if (!c->s.sx[0]) return 0;
    k_unsentinel<<<nblk(c->geo.n_ext), MISA_BLOCK, 0, c->stream>>>(c->geo.n_ext, c->s);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; k++) { cudaFree(c->s.sx[k]); c->s.sx[k] = nullptr; }
    c->opt_vac_sentinel = 0;
    return 0;
}
extern "C" int misa_b200_set_option(misa_b200_ctx *c, const char *name, int value) {
    REQ(c && name, MISA_B200_EINVAL, "null argument");
    if (!strcmp(name, "prune")) c->opt_prune = value;
    else if (!strcmp(name, "fuse")) c->opt_fuse = value;
    else if (!strcmp(name, "smem")) c->opt_smem = value;
    else if (!strcmp(name, "tex")) c->opt_tex = value;
    else if (!strcmp(name, "novac")) c->opt_novac = value;
    else if (!strcmp(name, "fast")) c->opt_fast = value;
    else if (!strcmp(name, "dilute")) c->opt_dilute = value;
    else if (!strcmp(name, "sym")) { c->opt_sym = value; if (value) TRY(drop_sentinel(c)); }   // the pair-symmetric passes read every site's stored position
    else if (!strcmp(name, "p2p")) c->opt_p2p = value;
    else if (!strcmp(name, "mark")) c->opt_mark = value;
    else if (!strcmp(name, "late")) c->opt_late = value;
    else if (!strcmp(name, "dmax_flags")) c->opt_dmax_flags = value;
    else if (!strcmp(name, "p2p_fence")) c->opt_p2p_fence = value;
    else if (!strcmp(name, "push_fused")) c->opt_push_fused = value;
    else if (!strcmp(name, "host_slabs")) c->opt_host_slabs = value;
    else if (!strcmp(name, "low_list")) c->opt_low_list = value;
    else if (!strcmp(name, "vac_sentinel")) {
        if (!value) TRY(drop_sentinel(c));
        else REQ(c->s.sx[0] || !c->have_atoms, MISA_B200_ESTATE, "vac_sentinel: switch on before atoms are uploaded");
        if (value && !c->s.sx[0]) for (int k = 0; k < 3; k++) { TRY(dmalloc(&c->s.sx[k], (size_t)c->geo.n_ext)); CU(cudaMemset(c->s.sx[k], 0, (size_t)c->geo.n_ext * 8)); }
        c->opt_vac_sentinel = value;
    }
    else if (!strcmp(name, "inter_dev")) {
        REQ(c->n_inter_local + c->n_inter_ghost == 0, MISA_B200_ESTATE, "inter_dev: switch before any inter atom exists");
        c->opt_inter_dev = value;
    }
    else if (!strcmp(name, "slab_planes")) { c->slab_T = std::max(3, value); c->n_slabs = 0; }
    else if (!strcmp(name, "p2p_debug")) {
        if (value && !c->d_p2p_dbg) {
            TRY(dmalloc(&c->d_p2p_dbg, 16));
            const unsigned long long init[16] = {~0ULL, 0, 0, 0, 0, 0, 0, 0, ~0ULL, 0, 0, 0, 0, 0, 0, 0};
            CU(cudaMemcpy(c->d_p2p_dbg, init, sizeof init, cudaMemcpyHostToDevice));
        }
        c->opt_p2p_debug = value;
    }
    else if (!strcmp(name, "p2p_timeout_s")) { c->opt_p2p_timeout_s = std::max(1, value); c->p2p.spin_limit = (long long)c->opt_p2p_timeout_s * 2000000000LL; }
    else if (!strcmp(name, "fuse_verlet")) c->opt_fuse_verlet = value;
    else if (!strcmp(name, "pipe")) c->opt_pipe = value;
    else if (!strcmp(name, "reserve")) c->opt_reserve = value;
    else if (!strcmp(name, "overlap")) c->opt_overlap = value;
    else return fail(MISA_B200_EINVAL, std::string("unknown option ") + name);
    return 0;
}

// read-only introspection for tests / benches: "n_off" (offsets the stencil kernels loop right now), "dmax"
// (largest displacement from a site, Angstrom), "n_full", "single" (single-species fast path), "novac", "smem_bytes"
extern "C" int misa_b200_query(misa_b200_ctx *c, const char *name, double *value) {
    REQ(c && name && value, MISA_B200_EINVAL, "null argument");
    const int *offs;
    int n_off;
    pick_list(c, offs, n_off);
    StagePlan sp;
    size_t sb = 0;
    const bool planned = c->have_pot && c->have_off && make_plan(c, sp, sb);
    if (!strcmp(name, "n_off")) *value = n_off;
    else if (!strcmp(name, "n_full")) *value = c->n_full;
    else if (!strcmp(name, "pipe_steps")) *value = (double)c->pipe_steps;
    else if (!strcmp(name, "pipe_redo")) *value = (double)c->pipe_redo;
    else if (!strcmp(name, "host_slab_steps")) *value = (double)c->host_slab_steps;
    else if (!strcmp(name, "host_slab_redo")) *value = (double)c->host_slab_redo;
    else if (!strcmp(name, "dmax")) *value = c->dmax_valid ? sqrt(c->dmax2) : -1.0;
    else if (!strcmp(name, "single")) *value = planned ? sp.single : -2;
    else if (!strcmp(name, "novac")) *value = no_vacancy(c) ? 1 : 0;
    else if (!strcmp(name, "smem_bytes")) *value = planned ? (double)sb : 0.0;
    else if (!strcmp(name, "dilute")) *value = planned && dilute_ok(c, sp, false) ? 1 : 0;
    else if (!strcmp(name, "n_minor")) *value = c->minor_valid ? c->n_minor : -1;
    else if (!strcmp(name, "n_half")) *value = c->n_half;
    else if (!strcmp(name, "p2p")) *value = c->p2p_active && c->opt_p2p ? 1 : 0;
    else if (!strcmp(name, "mark_count")) {   // diagnostic: atoms that marked in the last k_verlet1
        unsigned long long n = 0;
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(&n, c->d_stepinfo + 2, sizeof n, cudaMemcpyDeviceToHost));
        *value = (double)n;
    }
    else if (!strcmp(name, "mark_level")) *value = c->mark_valid ? c->mark_T_used : -1;
    else if (!strncmp(name, "p2p_dbg_", 8)) {   // p2p_dbg_<x|df>_<head|body|tail|all>: mean ns per push since the option was set
        *value = -1.0;
        if (c->d_p2p_dbg) {
            unsigned long long h[16];
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaMemcpy(h, c->d_p2p_dbg, sizeof h, cudaMemcpyDeviceToHost));
            const int w = !strncmp(name + 8, "df_", 3) ? 8 : 0;
            const char *ph = name + 8 + (w ? 3 : 2);
            const int k = !strcmp(ph, "head") ? 3 : !strcmp(ph, "body") ? 4 : !strcmp(ph, "tail") ? 5 : 6;
            *value = h[w + 7] ? (double)h[w + k] / (double)h[w + 7] : 0.0;
        }
    }
    else if (!strcmp(name, "p2p_error")) *value = c->h_p2p_err && *c->h_p2p_err ? (double)*c->h_p2p_err : (double)c->p2p_last_error;
    else if (!strcmp(name, "sym")) *value = planned && sym_active(c, sp) ? 1 : 0;
    else return fail(MISA_B200_EINVAL, std::string("unknown query ") + name);
    return 0;
}

extern "C" int misa_b200_comm_unique_id(void *out128) {
    REQ(out128, MISA_B200_EINVAL, "null argument");
    TRY(nccl_load());
    NC(g_nccl.GetUniqueId((nccl_uid *)out128));
    return 0;
}
extern "C" int misa_b200_comm_init(misa_b200_ctx *c, const void *uid, int rank, int n_ranks) {
    REQ(c && uid, MISA_B200_EINVAL, "null argument");
    TRY(nccl_load());
    nccl_uid id;
    memcpy(&id, uid, sizeof id);
    NC(g_nccl.CommInitRank(&c->nccl_comm, n_ranks, id, rank));
    c->comm_rank = rank;
    c->comm_size = n_ranks;
    return p2p_setup(c);   // ghost exchange by direct stores into the neighbours' HBM when all of them are peer-mapped
}
extern "C" int misa_b200_comm_destroy(misa_b200_ctx *c) {
    if (c) p2p_release(c);
    if (c && c->nccl_comm) {
        g_nccl.CommDestroy(c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    return 0;
}

// -------------------------------------------------------------------------------------------------
// halo exchange: comm::neiSendReceive<T> forward (x -> y -> z), restated on the device
// -------------------------------------------------------------------------------------------------
// width: doubles per site in the message (4 for positions+type, 1 for df)
static int halo_forward(misa_b200_ctx *c, bool positions, cudaStream_t st = nullptr) {
    if (!st) st = c->stream;
    const int width = positions ? 4 : 1;
    if (c->all_self && c->n_ghost_map > 0) {
        const int n = c->n_ghost_map;
        if (positions)
            k_ghost_fill_x<<<nblk(n), MISA_BLOCK, 0, st>>>(n, c->d_ghost_dst, c->d_ghost_src, c->d_ghost_shift, c->s,
                                                                 c->dom.meas_global_length[0], c->dom.meas_global_length[1],
                                                                 c->dom.meas_global_length[2]);
        else
            k_ghost_fill_1<<<nblk(n), MISA_BLOCK, 0, st>>>(n, c->d_ghost_dst, c->d_ghost_src, c->s.df);
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (c->p2p_active && c->opt_p2p) return p2p_exchange(c, positions, st);
    for (int d = 0; d < 3; d++) {
        const bool self = c->dom.grid_size[d] == 1;
        const HaloList &h0 = c->halo[d][0], &h1 = c->halo[d][1];
        Halo2 h;
        h.n0 = h0.n; h.n1 = h1.n;
        h.send0 = h0.d_send; h.send1 = h1.d_send; h.recv0 = h0.d_recv; h.recv1 = h1.d_recv;
        for (int k = 0; k < 3; k++) { h.sh0[k] = h0.shift[k]; h.sh1[k] = h1.shift[k]; }
        const int nb = nblk(h.n0 + h.n1);
        if (self) {
            if (positions) k_copy_x2<<<nb, MISA_BLOCK, 0, st>>>(h, c->s);
            else k_copy_12<<<nb, MISA_BLOCK, 0, st>>>(h, c->s.df);
            c->launches++;
            CU(cudaGetLastError());
            continue;
        }
        REQ(c->nccl_comm, MISA_B200_ESTATE, "halo exchange across sub-boxes needs misa_b200_comm_init");
        if (positions) k_pack_x2<<<nb, MISA_BLOCK, 0, st>>>(h, c->s, c->d_sendbuf[0], c->d_sendbuf[1]);
        else k_pack_12<<<nb, MISA_BLOCK, 0, st>>>(h, c->s.df, c->d_sendbuf[0], c->d_sendbuf[1]);
        c->launches++;
        CU(cudaGetLastError());
        NC(g_nccl.GroupStart());
        for (int dir = 0; dir < 2; dir++) {
            const HaloList &hl = c->halo[d][dir];
            NC(g_nccl.Send(c->d_sendbuf[dir], (size_t)hl.n * width, kNcclDouble, c->dom.rank_id_neighbours[d][dir], c->nccl_comm, st));
            NC(g_nccl.Recv(c->d_recvbuf[dir], (size_t)hl.n * width, kNcclDouble, c->dom.rank_id_neighbours[d][(dir + 1) % 2], c->nccl_comm, st));
        }
        NC(g_nccl.GroupEnd());
        if (positions) k_unpack_x2<<<nb, MISA_BLOCK, 0, st>>>(h, c->s, c->d_recvbuf[0], c->d_recvbuf[1]);
        else k_unpack_12<<<nb, MISA_BLOCK, 0, st>>>(h, c->s.df, c->d_recvbuf[0], c->d_recvbuf[1]);
        c->launches++;
        CU(cudaGetLastError());
    }
    return 0;
}

// -------------------------------------------------------------------------------------------------
// passes
// -------------------------------------------------------------------------------------------------
static int ready(misa_b200_ctx *c) {
    REQ(c, MISA_B200_EINVAL, "null ctx");
    REQ(c->have_off, MISA_B200_ESTATE, "neighbour offsets not set (cuda_nei_offset_init)");
    REQ(c->have_pot, MISA_B200_ESTATE, "potential not set (cuda_pot_init)");
    REQ(c->have_atoms, MISA_B200_ESTATE, "no atoms on the device");
    return 0;
}
static inline bool has_inter(const misa_b200_ctx *c) { return c->n_inter_local + c->n_inter_ghost > 0; }

// The offset list the lattice stencil kernels loop: the smallest pruned level that the measured maximum
// displacement allows, the reference's full list when nothing is known (or pruning is switched off).
static void pick_list(const misa_b200_ctx *c, const int *&offs, int &n_off, int *n_near) {
    offs = c->d_off_full;
    n_off = c->n_full;
    if (n_near) *n_near = c->near_full;
    if (!c->opt_prune || !c->dmax_valid) return;
    const double d = sqrt(c->dmax2) + 1e-6;
    const int L = (int)ceil(d / (0.01 * c->geo.a));
    if (L >= misa_b200_ctx::kLevels || c->level_n[L] <= 0) return;
    offs = c->d_off_levels + c->level_ofs[L];
    n_off = c->level_n[L];
    if (n_near) *n_near = c->level_near[L];
}

// measure dmax over the whole ghost-extended array (positions as they are now, ghosts included) and, for the stencil pruning
// of the serial path, the per-cell partner bound (kernels.cuh:k_pmax_*)
static int measure_displacement(misa_b200_ctx *c) {
    const Geo &g = c->geo;
    c->mark_valid = false;   // levels are re-measured for every site, nothing is marked
    c->pmax_valid = false;
    CU(cudaMemsetAsync(c->d_stepinfo + 1, 0, sizeof(unsigned long long), c->stream));
    k_max_displacement<<<nblk(g.n_ext), MISA_BLOCK, 0, c->stream>>>(g, c->s, c->d_stepinfo + 1);
    c->launches++;
    if (c->opt_prune && c->opt_mark) {
        if (!c->d_pmax) { TRY(dmalloc(&c->d_pmax, (size_t)g.H)); TRY(dmalloc(&c->d_ptmp, (size_t)g.H)); }
        k_pmax_x<<<nblk(g.H), MISA_BLOCK, 0, c->stream>>>(g, c->s.type, c->s.ulev, c->d_pmax);
        k_pmax_axis<<<nblk(g.H), MISA_BLOCK, 0, c->stream>>>(g, 1, g.gy, c->d_pmax, c->d_ptmp);
        k_pmax_axis<<<nblk(g.H), MISA_BLOCK, 0, c->stream>>>(g, 2, g.gz, c->d_ptmp, c->d_pmax);
        c->launches += 3;
        c->pmax_valid = true;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->h_stepinfo + 1, c->d_stepinfo + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(&c->dmax2, c->h_stepinfo + 1, sizeof(double));
    c->dmax_valid = true;
    c->mark_T_next = std::max(0, (int)std::min(ceil((sqrt(c->dmax2) + 1e-6) / (0.01 * c->geo.a)), 1000.0) - 3);
    return 0;
}

extern "C" int misa_b200_pass_halo_x(misa_b200_ctx *c) {
    TRY(ready(c));
    Slot sl(c, MISA_B200_K_HALO_X);
    return halo_forward(c, true);
}
extern "C" int misa_b200_pass_halo_df(misa_b200_ctx *c) {
    TRY(ready(c));
    {
        Slot sl(c, MISA_B200_K_HALO_DF);
        TRY(halo_forward(c, false));
    }
    if (c->inter_active) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_halo_df(c)); }
    return 0;
}
extern "C" int misa_b200_pass_clear(misa_b200_ctx *c) {
    TRY(ready(c));
    k_clear<<<nblk(c->geo.n_ext), MISA_BLOCK, 0, c->stream>>>(c->geo.n_ext, c->s.f[0], c->s.f[1], c->s.f[2], c->s.rho);
    c->launches++;
    CU(cudaGetLastError());
    if (has_inter(c)) TRY(inter_clear(c));
    return 0;
}

// Species census of the ghost-extended array. The set of species is an invariant of a run (atoms are neither
// created nor transmuted), so resident mode takes it once (upload + prepare, summed over sub-boxes); the compat
// hooks take it on every call because the host owns the array between calls.
static int census_local(misa_b200_ctx *c) {
    CU(cudaMemsetAsync(c->d_census, 0, MISA_MAX_TYPES * sizeof(unsigned long long), c->stream));
    k_census<<<std::min(nblk(c->geo.n_ext), 4 * std::max(c->sm_count, 1)), MISA_BLOCK, 0, c->stream>>>(c->geo.n_ext, c->s.type, c->d_census);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
static inline bool has_inter(const misa_b200_ctx *c);
static int census_fetch(misa_b200_ctx *c) {
    CU(cudaMemcpyAsync(c->h_census, c->d_census, MISA_MAX_TYPES * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    unsigned long long valid = 0;
    for (int t = 0; t < MISA_MAX_TYPES; t++) { c->census[t] = c->h_census[t]; valid += c->h_census[t]; }
    // after prepare()'s all-reduce the counts cover every (equal-sized) sub-box
    const unsigned long long boxes = (unsigned long long)std::max(c->census_boxes, 1);
    c->n_valid_sites = valid % boxes == 0 ? (long long)(valid / boxes) : -1;
    c->seen_offlattice = has_inter(c);
    c->census_valid = true;
    return 0;
}

// What the stencil kernels stage in shared memory: the majority species' elec and phi tables for r >= r_lo.
static bool make_plan(const misa_b200_ctx *c, StagePlan &sp, size_t &smem_bytes) {
    if (!c->opt_smem || !c->hermite_ok || !c->census_valid || c->sm_count <= 0) return false;
    const int nt = c->tab.n_types, n = c->tab.n_r;
    int present = 0, maj = 0;
    for (int t = 0; t < MISA_MAX_TYPES; t++) {
        if (c->census[t] && t >= nt) return false; // a species the potential does not cover
        if (c->census[t]) present++;
        if (c->census[t] > c->census[maj]) maj = t;
    }
    memset(&sp, 0, sizeof sp);
    for (int t = 0; t < nt; t++) sp.g_elec[t] = c->d_herm + (size_t)t * (n + 1);
    for (int t = 0; t < nt * nt; t++) sp.g_phi[t] = c->d_herm + (size_t)(nt + t) * (n + 1);
    sp.g_mono = c->d_mono;
    sp.single = present <= 1 ? maj : -1;
    sp.row_lo = std::max(1, std::min(n - 2, (int)(c->stage_r_lo * c->tab.inv_dr) - 1));
    sp.rows_s = n + 1 - sp.row_lo;
    sp.off_bytes = (int)((2 * (size_t)((c->n_full + 3) & ~3) * sizeof(int) + 127) / 128 * 128);   // (each parity's list padded to a multiple of four offsets)
    sp.n_staged = 2;
    sp.staged_id[0] = maj;                              // elec[maj]
    sp.staged_id[1] = MISA_MAX_TYPES + maj * nt + maj;  // phi[maj][maj]
    smem_bytes = (size_t)sp.off_bytes + (size_t)sp.n_staged * sp.rows_s * 16;
    return smem_bytes + 1024 <= (size_t)c->smem_optin;
}

// Force launches of the single-species loop (pure boxes and the dilute-alloy variants): slot 0 as (slope, value difference) rows plus
// the dense slopes of elec[maj] behind the slots (StagePlan::half_src) -- when the extra 8 bytes per row still fit.
// false: it does not fit -- the caller must not launch a single-species k_force_f (the second-generation kernels take over).
static bool plan_elec3(const misa_b200_ctx *c, StagePlan &sp, size_t &smem_bytes) {
#if EAM_ELEC3
    const int n = c->tab.n_r, maj = sp.staged_id[0], r0 = sp.row_lo & ~1;
    const size_t bytes = (((size_t)(n + 1 - r0) * 8 + 15) / 16) * 16;
    if (!c->d_ea || !c->d_es || smem_bytes + bytes + 1024 > (size_t)c->smem_optin) return false;
    sp.src[0] = c->d_ea + (size_t)maj * (n + 1);
    sp.half_src = c->d_es + (size_t)maj * (n + 1) + r0;
    sp.half_bytes = (int)bytes;
    smem_bytes += bytes;
#endif
    return true;
}

// No vacant site anywhere in the ghost-extended array: true once the census counted every site as valid and no
// run-away has been seen since (decide() is the only thing that vacates a site; every rank learns of run-aways
// anywhere through the per-step activity reduction, so a vacancy cannot enter the ghost shell unnoticed).
static inline bool no_vacancy(const misa_b200_ctx *c) {
    return c->opt_novac && c->n_valid_sites == c->geo.n_ext && !c->seen_offlattice;
}
// The stencil kernels may skip the per-neighbour species test: no site is vacant -- or vacant sites are invisible to them
// (their stored position is MISA_VACANT_X, ctx.h Soa::sx). A single-species question: alloys load the species anyway.
static inline bool no_type_test(const misa_b200_ctx *c) { return no_vacancy(c) || (c->opt_novac && c->s.sx[0] != nullptr); }

// Dilute alloy (one species holds >= 90 % of the valid sites): the SINGLE-species loop over the majority tables plus
// the minority-neighbour epilogue (eam_fast.cuh). Needs the static lists of prepare(), nothing off-lattice since, and
// every atom within 0.2a of its site (the lists cover the widest pruned stencil).
static bool dilute_ok(const misa_b200_ctx *c, const StagePlan &sp, bool accum) {
    if (!c->opt_dilute || !c->opt_prune || accum || sp.single >= 0 || !c->minor_valid || c->minor_maj != sp.staged_id[0]) return false;
    if (c->seen_offlattice || c->inter_active || has_inter(c)) return false;
    return c->dmax_valid && sqrt(c->dmax2) + 1e-6 < 0.2 * c->geo.a;
}
static MinorList minor_list(const misa_b200_ctx *c) {
    const int L = misa_b200_ctx::kLevels - 1;
    MinorList ml;
    ml.count = c->d_mcount; ml.entry = c->d_mentry;
    ml.offs = c->d_off_levels + c->level_ofs[L]; ml.n_offs = c->level_n[L];
    ml.n_ext = c->geo.n_ext; ml.maj = c->minor_maj;
    return ml;
}
static int build_minor_lists(misa_b200_ctx *c) {
    c->minor_valid = false;
    StagePlan sp;
    size_t sb;
    const int L = misa_b200_ctx::kLevels - 1;
    if (!c->opt_dilute || !c->have_pot || !make_plan(c, sp, sb) || sp.single >= 0 || c->level_n[L] <= 0 || c->level_n[L] > 128) return 0;
    unsigned long long total = 0;
    for (int t = 0; t < MISA_MAX_TYPES; t++) total += c->census[t];
    if (total == 0 || (double)c->census[sp.staged_id[0]] < 0.9 * (double)total) return 0;
    const Geo &g = c->geo;
    if (!c->d_mcount) {
        TRY(dmalloc(&c->d_mcount, (size_t)g.n_ext));
        TRY(dmalloc(&c->d_mentry, (size_t)g.n_ext * MINOR_CAP));
        TRY(dmalloc(&c->d_minor, (size_t)2 * g.n_cells_owned));
        TRY(dmalloc(&c->d_minor_count, 1));
    }
    const int bpp = nblk(g.n_cells_owned);
    CU(cudaMemsetAsync(c->d_minor_count, 0, sizeof(int), c->stream));
    k_build_minor_lists<<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(g, c->s.type, sp.staged_id[0], bpp, c->d_off_levels + c->level_ofs[L], c->level_n[L],
                                                              c->d_mcount, c->d_mentry, c->d_minor, c->d_minor_count);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&c->n_minor, c->d_minor_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->minor_maj = sp.staged_id[0];
    c->minor_valid = true;
    return 0;
}

// whole sub-box / interior (no ghost site within the stencil reach) / the six boundary slabs around it
static RegionList make_regions(const Geo &g, int which) {
    RegionList rl;
    memset(&rl, 0, sizeof rl);
    auto add = [&](int x0, int y0, int z0, int nx, int ny, int nz) {
        if (nx <= 0 || ny <= 0 || nz <= 0) return;
        Region &r = rl.r[rl.n++];
        r.x0 = x0; r.y0 = y0; r.z0 = z0; r.nx = nx; r.ny = ny; r.nz = nz; r.u0 = rl.units;
        rl.units += ((long long)nx * ny * nz + 31) / 32;
    };
    const bool has_interior = g.nx > 2 * g.gx && g.ny > 2 * g.gy && g.nz > 2 * g.gz;
    if (which == 0 || !has_interior) { if (which != 1) add(0, 0, 0, g.nx, g.ny, g.nz); return rl; }
    const int ix = g.nx - 2 * g.gx, iy = g.ny - 2 * g.gy, iz = g.nz - 2 * g.gz;
    if (which == 1) { add(g.gx, g.gy, g.gz, ix, iy, iz); return rl; }
    if (which == 3) {
        // interior FIRST, then the slabs around it, for the in-kernel wait on the neighbours' ghost push (LateWait). The
        // interior keeps 15 more cells from the x faces: its loads must not touch a 128-byte cache LINE that holds a ghost
        // site (16 doubles along x: the line of a ghost reaches 15 owned cells into the row), or a stale copy of the ghost
        // could sit in L1 / the texture cache when the boundary units read it after the wait
        const int mx = g.gx + 15, jx = g.nx - 2 * mx;
        if (jx <= 0) { add(0, 0, 0, g.nx, g.ny, g.nz); return rl; }
        add(mx, g.gy, g.gz, jx, iy, iz);
        rl.split = rl.units;
        add(0, 0, 0, g.nx, g.ny, g.gz); add(0, 0, g.nz - g.gz, g.nx, g.ny, g.gz);
        add(0, 0, g.gz, g.nx, g.gy, iz); add(0, g.ny - g.gy, g.gz, g.nx, g.gy, iz);
        add(0, g.gy, g.gz, mx, iy, iz); add(g.nx - mx, g.gy, g.gz, mx, iy, iz);
        return rl;
    }
    add(0, 0, 0, g.nx, g.ny, g.gz); add(0, 0, g.nz - g.gz, g.nx, g.ny, g.gz);
    add(0, 0, g.gz, g.nx, g.gy, iz); add(0, g.ny - g.gy, g.gz, g.nx, g.gy, iz);
    add(0, g.gy, g.gz, g.gx, iy, iz); add(g.nx - g.gx, g.gy, g.gz, g.gx, iy, iz);
    return rl;
}
// dmax2: device word holding the bit pattern of the largest squared displacement (null: the host's pick_list choice)
static LevelSel make_levelsel(const misa_b200_ctx *c, const unsigned long long *dmax2) {
    LevelSel ls;
    memset(&ls, 0, sizeof ls);
    ls.host_level = -1;
    ls.n_full = c->n_full; ls.near_full = c->near_full;
    ls.step = 0.01 * c->geo.a;
    ls.ulev = c->s.ulev;
    for (int L = 0; L < misa_b200_ctx::kPairLevels; L++) ls.prefix[L] = c->prefix_n[L];
    ls.H = c->geo.H;
    if (c->mark_valid && c->opt_mark) { ls.hot = c->d_hot; ls.edge = c->d_hot_init; ls.hot_T = c->mark_T_used; ls.hot_epoch = c->mark_epoch; ls.hot_count = c->d_stepinfo + 2; }
    if (c->pmax_valid && c->opt_mark && !dmax2) ls.pmax = c->d_pmax;   // the host-level (serial) path: positions unchanged since measure_displacement
    if (!c->opt_prune) return ls;
    if (!dmax2) {   // the host's knowledge (serial path): same rounding as pick_list
        if (c->dmax_valid) ls.host_level = (int)std::min(ceil((sqrt(c->dmax2) + 1e-6) / ls.step), 1000.0);
        return ls;
    }
    ls.dmax2_bits = dmax2;
    ls.levels = c->d_off_levels; ls.full = c->d_off_full;
    ls.levels_addr = c->d_off_levels_addr; ls.full_addr = c->d_off_full_addr;
    for (int L = 0; L < misa_b200_ctx::kLevels; L++) { ls.n[L] = c->level_n[L]; ls.near_[L] = c->level_near[L]; ls.ofs[L] = (int)c->level_ofs[L]; }
    return ls;
}
// host-only view of the work regions (tests): boxes as x0,y0,z0,nx,ny,nz in owned-cell coordinates, split = warp units of
// the interior box when the launch visits it first (which = 3)
extern "C" int misa_b200_plan_regions(const misa_b200_domain *dom, int which, int32_t boxes[7][6], int32_t *n_boxes, int64_t *units,
                                      int64_t *split) {
    REQ(dom && boxes && n_boxes && which >= 0 && which <= 3, MISA_B200_EINVAL, "misa_b200_plan_regions: bad argument");
    Geo g;
    REQ(geo_from_domain(dom, g) == 0, MISA_B200_EINVAL, "misa_b200_plan_regions: bad sub-box sizes");
    const RegionList rl = make_regions(g, which);
    *n_boxes = rl.n;
    for (int b = 0; b < rl.n; b++) {
        const Region &r = rl.r[b];
        const int v[6] = {r.x0, r.y0, r.z0, r.nx, r.ny, r.nz};
        for (int k = 0; k < 6; k++) boxes[b][k] = v[k];
    }
    if (units) *units = rl.units;
    if (split) *split = rl.split;
    return MISA_B200_OK;
}
// host-only: the (sub-lattice, unit) a launch of plan_regions(which) visits as its u-th warp unit -- the kernels' own mapping
extern "C" int misa_b200_plan_unit_order(const misa_b200_domain *dom, int which, int64_t u, int32_t *parity, int64_t *unit) {
    REQ(dom && parity && unit && which >= 0 && which <= 3, MISA_B200_EINVAL, "misa_b200_plan_unit_order: bad argument");
    Geo g;
    REQ(geo_from_domain(dom, g) == 0, MISA_B200_EINVAL, "misa_b200_plan_unit_order: bad sub-box sizes");
    const RegionList rl = make_regions(g, which);
    REQ(u >= 0 && u < 2 * rl.units, MISA_B200_EINVAL, "misa_b200_plan_unit_order: unit out of range");
    int par;
    long long up;
    unit_split(rl, u, par, up);
    *parity = par; *unit = up;
    return MISA_B200_OK;
}
struct StencilOpt;
static RegionList regions_for(const misa_b200_ctx *c, const StencilOpt &so, bool late);
struct StencilOpt {                 // how one stencil launch deviates from "whole sub-box, host-chosen list, main stream"
    int region = 0;                 // 0 whole, 1 interior, 2 boundary slabs
    const unsigned long long *dmax2 = nullptr;
    int reserve_sms = 0;            // leave this many SMs free (the persistent CTAs would otherwise starve the exchange kernels on stream2)
    bool late = false;              // the ghost push this launch consumes has been issued but not waited for (p2p.cuh)
    bool fused = false;             // that push was done from inside the producing kernel: this launch posts its ARRIVE flags (post_epoch,
    unsigned long long post_epoch = 0;   // post_dmax) and, for rho + df, pushes df from its epilogue
    const unsigned long long *post_dmax = nullptr;
    int z0 = -1, z1 = -1;           // z0 >= 0: only the owned z-planes [z0, z1) (slab-pipelined misa_b200_step_host)
};
static RegionList regions_for(const misa_b200_ctx *c, const StencilOpt &so, bool late) {
    if (so.z0 < 0) return make_regions(c->geo, late ? 3 : so.region);
    RegionList rl;
    memset(&rl, 0, sizeof rl);
    Region &r = rl.r[rl.n++];
    r.x0 = 0; r.y0 = 0; r.z0 = so.z0; r.nx = c->geo.nx; r.ny = c->geo.ny; r.nz = so.z1 - so.z0; r.u0 = 0;
    rl.units = ((long long)r.nx * r.ny * r.nz + 31) / 32;
    return rl;
}
// ---- pair-symmetric passes (eam_sym.cuh) --------------------------------------------------------------------------
// One species (or a dilute alloy, whose main loop is the majority species'), overwrite semantics, the whole sub-box in
// one launch (the interior / boundary split of the overlapped exchange keeps the full-list kernels).
static bool sym_ok(const misa_b200_ctx *c, const StagePlan &sp, bool accum, const StencilOpt &so) {
    if (!c->opt_sym || c->n_half <= 0 || !c->d_lo_tab || accum || so.region != 0) return false;
    return sp.single >= 0 || dilute_ok(c, sp, accum);
}
static bool sym_active(const misa_b200_ctx *c, const StagePlan &sp) { return sym_ok(c, sp, false, StencilOpt()); }
// The wait for the neighbours' push goes INSIDE the stencil kernel when that kernel reads nothing but positions / df of its
// neighbours through the texture path (no per-neighbour type loads: their 32-site sectors would need a 32-cell margin) and
// the sub-box has an interior; otherwise it is a kernel of its own in front of the launch.
static bool late_wait_ok(const misa_b200_ctx *c, const StagePlan &sp, bool planned, bool accum, const StencilOpt &so) {
    if (!so.late || !c->opt_late || !planned || !c->opt_fast || !c->tex_all || accum || so.region != 0 || !no_vacancy(c)) return false;
    if (!(sp.single >= 0 || dilute_ok(c, sp, accum)) || sym_ok(c, sp, accum, so)) return false;
    return make_regions(c->geo, 3).split > 0;
}
// Which displacement word a stencil launch of the sync-free step reads. With the maxima travelling on the push flags
// (dmax_by_flags) a launch that waits inside the kernel starts from the OWN word and widens it at the late wait; a launch
// behind k_p2p_wait_arrive reads the word that kernel folded.
static const unsigned long long *stencil_dmax(const misa_b200_ctx *c, const StencilOpt &so, bool late) {
    if (c->dmax_by_flags && so.late) return late ? c->d_stepinfo + 1 : c->d_stepinfo_n + 1;
    return so.dmax2;
}
static LateWait make_latewait(const misa_b200_ctx *c) {
    LateWait lw = LateWait();
    lw.flags = c->d_flags; lw.epoch = c->p2p_epoch; lw.mask = c->p2p.mask; lw.err = c->d_p2p_err; lw.limit = c->p2p.spin_limit;
    lw.fold_dmax = c->dmax_by_flags ? 1 : 0;
    return lw;
}

// The push hooks of a stencil launch in the fused sync-free step. Waiting inside the kernel (late): its first CTA posts the
// ARRIVE flags. Waiting in front of it: a tiny kernel posts them BEFORE the wait kernel (both neighbours wait for each other's
// flags -- posting from behind the wait would deadlock).
static int fused_hooks(misa_b200_ctx *c, const StencilOpt &so, bool late, bool push_df, LateWait &lw) {
    if (!so.fused) return 0;
    if (late) { lw.post = c->d_p2p_dev; lw.post_epoch = so.post_epoch; lw.post_dmax = so.post_dmax; }
    else {
        k_p2p_arrive<<<1, 32, 0, c->stream>>>(c->p2p, so.post_epoch, c->d_p2p_err, so.post_dmax);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (push_df) lw.push_df = c->d_p2p_dev;
    return 0;
}
static int sym_scratch(misa_b200_ctx *c) {
    const size_t need = (size_t)c->n_half * (((size_t)c->geo.n_ext + 31) / 32 * 32);
    if (c->pair_elems >= need) return 0;
    cudaFree(c->d_pair); c->d_pair = nullptr; c->pair_elems = 0;
    TRY(dmalloc(&c->d_pair, need));
    c->pair_elems = need;
    return 0;
}
// pass-A work: the owned box (region 0) + the slabs of ghost sites around it that own pairs with owned atoms
static RegionList make_regions_sym(const misa_b200_ctx *c) {
    const Geo &g = c->geo;
    const int *lo = c->sym_lo, *hi = c->sym_hi;
    RegionList rl;
    memset(&rl, 0, sizeof rl);
    auto add = [&](int x0, int y0, int z0, int nx, int ny, int nz) {
        if (nx <= 0 || ny <= 0 || nz <= 0) return;
        Region &r = rl.r[rl.n++];
        r.x0 = x0; r.y0 = y0; r.z0 = z0; r.nx = nx; r.ny = ny; r.nz = nz; r.u0 = rl.units;
        rl.units += ((long long)nx * ny * nz + 31) / 32;
    };
    const int wx = g.nx + lo[0] + hi[0], wy = g.ny + lo[1] + hi[1];
    add(0, 0, 0, g.nx, g.ny, g.nz);
    add(-lo[0], -lo[1], -lo[2], wx, wy, lo[2]);
    add(-lo[0], -lo[1], g.nz, wx, wy, hi[2]);
    add(-lo[0], -lo[1], 0, wx, lo[1], g.nz);
    add(-lo[0], g.ny, 0, wx, hi[1], g.nz);
    add(-lo[0], 0, 0, lo[0], g.ny, g.nz);
    add(g.nx, 0, 0, hi[0], g.ny, g.nz);
    return rl;
}
static int sym_b_grid(const misa_b200_ctx *c, const RegionList &rl) {
    const long long blocks = (2 * rl.units + SYM_B_THREADS / 32 - 1) / (SYM_B_THREADS / 32);
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)std::max(c->sm_count, 1) * 16));
}

// Cascade form of the "pair below the staged range" recomputation (eam_fast.cuh:k_low_fix): serial path with inter atoms around,
// single species, whole-box launch
static const int kLowCap = 1 << 20;
static bool low_list_ok(const misa_b200_ctx *c, const StagePlan &sp, bool accum, bool fuse_df, const StencilOpt &so) {
    return c->opt_low_list && has_inter(c) && sp.single >= 0 && !accum && !fuse_df && so.region == 0 && so.z0 < 0 && !so.dmax2 && !so.late;
}
static int low_list_arm(misa_b200_ctx *c, LateWait &lw, int which) {
    if (!c->d_low_list) { TRY(dmalloc(&c->d_low_list, (size_t)kLowCap)); TRY(dmalloc(&c->d_low_count, 2)); }
    CU(cudaMemsetAsync(c->d_low_count + which, 0, sizeof(int), c->stream));
    lw.low_list = c->d_low_list; lw.low_count = c->d_low_count + which; lw.low_cap = kLowCap;
    return 0;
}
static int low_fix_launch(misa_b200_ctx *c, const StagePlan &sp, bool force, bool with_type, int which) {
    const int grid = std::max(1, c->sm_count) * 2;
    if (force) k_low_fix<true><<<grid, 256, 0, c->stream>>>(c->geo, c->s, c->tab, sp.g_mono, with_type ? c->s.type : nullptr, sp.single, c->d_off_full, c->n_full,
                                                           c->d_low_list, c->d_low_count + which, kLowCap, false, false);
    else k_low_fix<false><<<grid, 256, 0, c->stream>>>(c->geo, c->s, c->tab, sp.g_mono, with_type ? c->s.type : nullptr, sp.single, c->d_off_full, c->n_full,
                                                       c->d_low_list, c->d_low_count + which, kLowCap, false, false);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
static int launch_rho(misa_b200_ctx *c, bool fuse_df, bool accum, const StencilOpt &so = StencilOpt()) {
    const Geo &g = c->geo;
    const int bpp = nblk(g.n_cells_owned);
    const int *offs;
    int n_off, n_near;
    pick_list(c, offs, n_off, &n_near);
    const size_t sm = (size_t)n_off * sizeof(int);
    Slot sl(c, MISA_B200_K_RHO);
    StagePlan sp;
    size_t sb;
    const bool planned_any = make_plan(c, sp, sb);
    const bool late = late_wait_ok(c, sp, planned_any, accum, so);
    LateWait lw = late ? make_latewait(c) : LateWait();
    TRY(fused_hooks(c, so, late, fuse_df, lw));
    if (so.late && !late) TRY(p2p_wait(c, c->stream));   // the push this launch consumes: waited for in front of it
    if (c->opt_fast && c->tex_all && planned_any) {
        const int grid = std::max(1, c->sm_count - so.reserve_sms);
        const bool novac = no_type_test(c), single = sp.single >= 0;
        const TexAll tex = {c->tex_all, (int)c->xyzd_stride};
#if EAM_MONO_RHO
        if (!sym_ok(c, sp, accum, so) && (single || dilute_ok(c, sp, accum))) {   // every SINGLE-template variant of k_rho_f below: the reference's monomial rows of elec[maj] in both slots (the pair-symmetric pass A keeps the Hermite rows)
            const int maj = sp.staged_id[0];
            sp.src[0] = c->d_c34 + (size_t)maj * (c->tab.n_r + 1);
            sp.src[1] = c->d_c56 + (size_t)maj * (c->tab.n_r + 1);
        }
#endif
        const RegionList rl = regions_for(c, so, late);
        const LevelSel ls = make_levelsel(c, stencil_dmax(c, so, late));
        if (rl.units == 0) return 0;
        if (sym_ok(c, sp, accum, so)) {
            TRY(sym_scratch(c));
            const bool dil = sp.single < 0;
            const MinorList ml = dil ? minor_list(c) : MinorList();
            const SymPlan sy = {c->d_pair, c->d_lo_tab, c->n_half, g.n_ext};
            const RegionList ra = make_regions_sym(c);
#define RHO_A(N, D) k_rho_a<N, D><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, ra, ls, ml, sy)
            if (novac) { if (dil) RHO_A(true, true); else RHO_A(true, false); }
            else { if (dil) RHO_A(false, true); else RHO_A(false, false); }
#undef RHO_A
            CU(cudaGetLastError());
            const int gb = sym_b_grid(c, rl);
            const size_t lb = 2 * (size_t)c->n_half * sizeof(int2);
#define RHO_B(N, F, D) k_rho_b<N, F, D><<<gb, SYM_B_THREADS, lb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, rl, ls, sy)
#define RHO_BF(N, D) do { if (fuse_df) RHO_B(N, true, D); else RHO_B(N, false, D); } while (0)
            if (novac) { if (dil) RHO_BF(true, true); else RHO_BF(true, false); }
            else { if (dil) RHO_BF(false, true); else RHO_BF(false, false); }
#undef RHO_BF
#undef RHO_B
            c->launches += 2;
            CU(cudaGetLastError());
            return 0;
        }
        if (dilute_ok(c, sp, accum)) {
            const MinorList ml = minor_list(c);
#define RHO_D(N, F) k_rho_f<true, N, F, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, ml, lw)
            if (novac) { if (fuse_df) RHO_D(true, true); else RHO_D(true, false); }
            else { if (fuse_df) RHO_D(false, true); else RHO_D(false, false); }
#undef RHO_D
            c->launches++;
            CU(cudaGetLastError());
            return 0;
        }
#define RHO_F(S, N, F, A) k_rho_f<S, N, F, A><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList(), lw)
#define RHO_FA(S, N) do { if (accum) RHO_F(S, N, false, true); else if (fuse_df) RHO_F(S, N, true, false); else RHO_F(S, N, false, false); } while (0)
        if (low_list_ok(c, sp, accum, fuse_df, so)) {
            TRY(low_list_arm(c, lw, 0));
            if (novac) k_rho_f<true, true, false, false, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList(), lw);
            else k_rho_f<true, false, false, false, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList(), lw);
            c->launches++;
            CU(cudaGetLastError());
            return low_fix_launch(c, sp, false, !novac, 0);
        }
        if (single && novac) RHO_FA(true, true); else if (single) RHO_FA(true, false); else RHO_FA(false, false);
#undef RHO_FA
#undef RHO_F
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (make_plan(c, sp, sb)) {
        const int grid = std::max(1, c->sm_count - so.reserve_sms);
        SoaTex tex;
        for (int k = 0; k < 3; k++) tex.x[k] = c->tex_x[k];
        tex.df = c->tex_df;
        const bool use_tex = c->opt_tex && c->tex_x[0] && c->tex_x[1] && c->tex_x[2] && c->tex_df;
        const bool novac = no_vacancy(c);
        const RegionList rl = regions_for(c, so, false);
        const LevelSel ls = make_levelsel(c, stencil_dmax(c, so, late));
        if (rl.units == 0) return 0;
#define RHO_S(S, F, A) k_rho_s<S, F, A><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, tex, rl, ls)
#define RHO_X(T, N) k_rho_s<true, true, false, T, N><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, tex, rl, ls)
        if (sp.single >= 0 && fuse_df && !accum && (use_tex || novac)) {
            if (use_tex && novac) RHO_X(true, true); else if (use_tex) RHO_X(true, false); else RHO_X(false, true);
        } else
        if (sp.single < 0 && fuse_df && !accum && use_tex)
            k_rho_s<false, true, false, true, false><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, tex, rl, ls);
        else
        if (sp.single >= 0) { if (accum) RHO_S(true, false, true); else if (fuse_df) RHO_S(true, true, false); else RHO_S(true, false, false); }
        else { if (accum) RHO_S(false, false, true); else if (fuse_df) RHO_S(false, true, false); else RHO_S(false, false, false); }
#undef RHO_S
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (accum) k_rho<false, true><<<2 * bpp, MISA_BLOCK, sm, c->stream>>>(g, c->s, c->tab, offs, n_off, bpp);
    else if (fuse_df) k_rho<true, false><<<2 * bpp, MISA_BLOCK, sm, c->stream>>>(g, c->s, c->tab, offs, n_off, bpp);
    else k_rho<false, false><<<2 * bpp, MISA_BLOCK, sm, c->stream>>>(g, c->s, c->tab, offs, n_off, bpp);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
static int launch_df(misa_b200_ctx *c) {
    const int bpp = nblk(c->geo.n_cells_owned);
    Slot sl(c, MISA_B200_K_DF);
    k_df<<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(c->geo, c->s, c->tab, bpp);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
// atoms of the minority species of a dilute alloy: one warp each, pairs from the global monomial block
static int launch_force_minor(misa_b200_ctx *c, const StagePlan &sp, const int *offs, int n_off, const LevelSel &ls, const TexAll &tex) {
    if (c->n_minor <= 0) return 0;
    const Geo &g = c->geo;
    const int nt = c->tab.n_types, maj = sp.staged_id[0];
    const int mgrid = std::min((c->n_minor + 3) / 4, std::max(1, c->sm_count) * 20);   // 128 threads, ~94 registers: five CTAs per SM
    // the host's choice of list, in its address-ordered copy (the kernel re-chooses on the device when the level lives there)
    const int *offs_a = offs == c->d_off_full ? c->d_off_full_addr : c->d_off_levels_addr + (offs - c->d_off_levels);
    k_force_minor<<<mgrid, 128, 0, c->stream>>>(g, c->s, c->tab, sp, offs_a, n_off, ls, c->d_minor, c->n_minor, tex);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
static int launch_force(misa_b200_ctx *c, bool accum, const StencilOpt &so = StencilOpt()) {
    const Geo &g = c->geo;
    const int bpp = nblk(g.n_cells_owned);
    const int *offs;
    int n_off, n_near;
    pick_list(c, offs, n_off, &n_near);
    const size_t sm = (size_t)n_off * sizeof(int);
    Slot sl(c, MISA_B200_K_FORCE);
    StagePlan sp;
    size_t sb;
    const bool planned_any = make_plan(c, sp, sb);
    const bool late = late_wait_ok(c, sp, planned_any, accum, so);
    LateWait lw = late ? make_latewait(c) : LateWait();
    TRY(fused_hooks(c, so, late, false, lw));
    if (so.late && !late) TRY(p2p_wait(c, c->stream));   // the push this launch consumes: waited for in front of it
    if (c->opt_fast && c->tex_all && make_plan(c, sp, sb) && sym_ok(c, sp, accum, so)) {
        TRY(sym_scratch(c));
        const int grid = std::max(1, c->sm_count - so.reserve_sms);
        const TexAll tex = {c->tex_all, (int)c->xyzd_stride};
        const RegionList rl = make_regions(g, 0), ra = make_regions_sym(c);
        const LevelSel ls = make_levelsel(c, stencil_dmax(c, so, late));
        const bool dil = sp.single < 0, novac = no_vacancy(c);
        const MinorList ml = dil ? minor_list(c) : MinorList();
        const SymPlan sy = {c->d_pair, c->d_lo_tab, c->n_half, g.n_ext};
#define FORCE_A(N, D) k_force_a<N, D><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, ra, ls, ml, sy)
        if (novac) { if (dil) FORCE_A(true, true); else FORCE_A(true, false); }
        else { if (dil) FORCE_A(false, true); else FORCE_A(false, false); }
#undef FORCE_A
        CU(cudaGetLastError());
        const int gb = sym_b_grid(c, rl);
        const size_t lb = 2 * (size_t)c->n_half * sizeof(int2);
#define FORCE_B(N, D) k_force_b<N, D><<<gb, SYM_B_THREADS, lb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, rl, ls, sy, sp.staged_id[0])
        if (novac) { if (dil) FORCE_B(true, true); else FORCE_B(true, false); }
        else { if (dil) FORCE_B(false, true); else FORCE_B(false, false); }
#undef FORCE_B
        c->launches += 2;
        CU(cudaGetLastError());
        if (dil) TRY(launch_force_minor(c, sp, offs, n_off, ls, tex));
        return 0;
    }
    if (c->opt_fast && c->tex_all && make_plan(c, sp, sb) && dilute_ok(c, sp, accum) && plan_elec3(c, sp, sb)) {
        const int grid = std::max(1, c->sm_count - so.reserve_sms);
        const TexAll tex = {c->tex_all, (int)c->xyzd_stride};
        const RegionList rl = regions_for(c, so, late);
        const LevelSel ls = make_levelsel(c, stencil_dmax(c, so, late));
        const MinorList ml = minor_list(c);
        if (rl.units > 0) {
            if (no_vacancy(c)) k_force_f<true, true, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, ml, lw);
            else k_force_f<true, false, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, ml, lw);
            c->launches++;
            CU(cudaGetLastError());
        }
        // atoms of a minority species: ALL of them with the launch that runs after the df halo has arrived (the whole-box
        // launch or the boundary one), one warp per atom
        if (so.region != 1) {
            if (c->dmax_by_flags && so.late && late) {   // k_force_minor has no late wait: fold the neighbours' maxima in front of it
                TRY(p2p_wait(c, c->stream));
                TRY(launch_force_minor(c, sp, offs, n_off, make_levelsel(c, c->d_stepinfo_n + 1), tex));
            } else TRY(launch_force_minor(c, sp, offs, n_off, ls, tex));
        }
        return 0;
    }
    // non-dilute alloys: the generic-pointer force variant (three generic row fetches per pair) measured slower than the
    // second generation's staged/divergent one (1.50 vs 1.18 ms at 97:2:1), so multi-species force stays on k_force_s
    if (c->opt_fast && c->tex_all && make_plan(c, sp, sb) && (sp.single >= 0 ? plan_elec3(c, sp, sb) : c->opt_fast > 1)) {   // (the multi-species template reads slot 0 as Hermite rows)
        const int grid = std::max(1, c->sm_count - so.reserve_sms);
        const bool novac = no_type_test(c), single = sp.single >= 0;
        const TexAll tex = {c->tex_all, (int)c->xyzd_stride};
        const RegionList rl = regions_for(c, so, late);
        const LevelSel ls = make_levelsel(c, stencil_dmax(c, so, late));
        if (rl.units == 0) return 0;
#define FORCE_F(S, N) do { if (accum) k_force_f<S, N, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList()); \
                           else k_force_f<S, N, false><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList(), lw); } while (0)
        if (low_list_ok(c, sp, accum, false, so)) {
            TRY(low_list_arm(c, lw, 1));
            if (novac) k_force_f<true, true, false, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList(), lw);
            else k_force_f<true, false, false, false, true><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, c->d_off_full, c->n_full, c->near_full, tex, rl, ls, MinorList(), lw);
            c->launches++;
            CU(cudaGetLastError());
            return low_fix_launch(c, sp, true, !novac, 1);
        }
        if (single && novac) FORCE_F(true, true); else if (single) FORCE_F(true, false); else FORCE_F(false, false);
#undef FORCE_F
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (make_plan(c, sp, sb)) {
        const int grid = std::max(1, c->sm_count - so.reserve_sms);
        SoaTex tex;
        for (int k = 0; k < 3; k++) tex.x[k] = c->tex_x[k];
        tex.df = c->tex_df;
        const bool use_tex = c->opt_tex && c->tex_x[0] && c->tex_x[1] && c->tex_x[2] && c->tex_df;
        const bool novac = no_vacancy(c);
        const RegionList rl = regions_for(c, so, false);
        const LevelSel ls = make_levelsel(c, stencil_dmax(c, so, late));
        if (rl.units == 0) return 0;
#define FORCE_S(S, A) k_force_s<S, A><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, tex, rl, ls)
#define FORCE_X(T, N) k_force_s<true, false, T, N><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, tex, rl, ls)
        if (sp.single >= 0 && !accum && (use_tex || novac)) {
            if (use_tex && novac) FORCE_X(true, true); else if (use_tex) FORCE_X(true, false); else FORCE_X(false, true);
        } else
        if (sp.single < 0 && !accum && use_tex)
            k_force_s<false, false, true, false><<<grid, EAM_THREADS, sb, c->stream>>>(g, c->s, c->tab, sp, offs, n_off, tex, rl, ls);
        else
        if (sp.single >= 0) { if (accum) FORCE_S(true, true); else FORCE_S(true, false); }
        else { if (accum) FORCE_S(false, true); else FORCE_S(false, false); }
#undef FORCE_S
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (accum) k_force<true><<<2 * bpp, MISA_BLOCK, sm, c->stream>>>(g, c->s, c->tab, offs, n_off, bpp);
    else k_force<false><<<2 * bpp, MISA_BLOCK, sm, c->stream>>>(g, c->s, c->tab, offs, n_off, bpp);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int misa_b200_pass_rho(misa_b200_ctx *c) {
    TRY(ready(c));
    TRY(launch_rho(c, false, false));
    if (has_inter(c)) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_rho(c)); }
    return 0;
}
extern "C" int misa_b200_pass_df(misa_b200_ctx *c) {
    TRY(ready(c));
    return launch_df(c);
}
extern "C" int misa_b200_pass_force(misa_b200_ctx *c) {
    TRY(ready(c));
    TRY(launch_force(c, false));
    if (has_inter(c)) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_force(c)); }
    return 0;
}

static VerletPar verlet_par(const misa_b200_ctx *c) {
    VerletPar vp;
    vp.dt = c->dt;
    for (int i = 0; i < MISA_MAX_TYPES; i++) vp.c[i] = c->dt_inv_m[i];
    vp.mark_T = 0; vp.hot = nullptr; vp.epoch = 0; vp.mark_count = nullptr; vp.push = nullptr;
    vp.c_begin = 0; vp.c_end = c->geo.n_cells_owned;
    // level arithmetic of k_verlet1 in single precision, every constant rounded UP (kernels.cuh:disp_level_fast)
    vp.inv100_a = nextafterf((float)(100.0 / c->geo.a), INFINITY);
    vp.lev_slack = nextafterf((float)(2e-4 / c->geo.a), INFINITY);
    return vp;
}

// Agree across all sub-boxes whether any off-lattice atom exists this step (the list exchanges are collective
// between neighbours, so every rank must take the same branch); also brings the step counters to the host.
// Enqueued on `st`; the results are in h_counters / h_stepinfo once `st` (or ev_act) has been waited for, and in
// d_stepinfo_g on the device (kernels of the same step read the global dmax from there).
static int activity_enqueue(misa_b200_ctx *c, cudaStream_t st) {
    const bool reduce = c->comm_size > 1 && c->nccl_comm;
    k_activity<<<1, 1, 0, st>>>(c->d_counters, c->n_inter_local + c->n_inter_ghost, c->d_stepinfo, reduce ? nullptr : c->d_stepinfo_g, c->hd_counters,
                                c->hd_stepinfo);
    c->launches++;
    if (reduce) { // [0] activity, [1] dmax2 bit pattern: MAX over the sub-boxes
        NC(g_nccl.AllReduce(c->d_stepinfo, c->d_stepinfo_g, 2, kNcclUint64, kNcclMax, c->nccl_comm, st));
        k_publish_stepinfo<<<1, 1, 0, st>>>(c->d_stepinfo_g, c->hd_stepinfo);
        c->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}
static int update_activity(misa_b200_ctx *c) {
    TRY(activity_enqueue(c, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->inter_active = c->h_stepinfo[0] > 0;
    if (c->inter_active) c->seen_offlattice = true;
    return 0;
}

// NewtonMotion::firststep + the displacement test of atom::decide, enqueued on the main stream
static int verlet1_enqueue(misa_b200_ctx *c, bool kick2 = false, bool push = false) {
    const Geo &g = c->geo;
    const int bpp = nblk(g.n_cells_owned);
    VerletPar vp = verlet_par(c);
    vp.push = push ? c->d_p2p_dev : nullptr;   // band sites store their new position straight into the neighbours' ghosts
    Slot sl(c, MISA_B200_K_VERLET1);
    c->mark_valid = false;
    c->pmax_valid = false;
    if (c->opt_mark && c->opt_prune) {   // this step's marks carry a fresh epoch byte (1..255): nothing to clear
        c->mark_epoch = c->mark_epoch % 255 + 1;
        c->mark_T_used = c->mark_T_next;
        vp.mark_T = c->mark_T_used; vp.hot = c->d_hot; vp.epoch = (unsigned char)c->mark_epoch; vp.mark_count = c->d_stepinfo + 2;
        c->mark_valid = true;
    }
    CU(cudaMemsetAsync(c->d_stepinfo, 0, 3 * sizeof(unsigned long long) + sizeof(int), c->stream));   // [0..2] + counters[0] (run-aways)
#define V1(K, P) k_verlet1<K, P><<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(g, soa_v(c->s), vp, bpp, c->d_counters, c->d_runaway, c->inter_cap, c->d_stepinfo)
    if (vp.push) { if (kick2) V1(true, true); else V1(false, true); }
    else { if (kick2) V1(true, false); else V1(false, false); }
#undef V1
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
// ... and what the host does once the step's counters are known: the rest of atom::decide and the inter-atom
// exchanges when anything is off-lattice anywhere (reference src/atom.cpp:21-84, src/atom/inter_atom_list.cpp:19-53)
static int verlet1_finish(misa_b200_ctx *c) {
    REQ(c->h_counters[3] == 0, MISA_B200_EOVERFLOW, "run-away list overflow");
    c->last_runaways = c->h_counters[0];
    // k_verlet1 measured every owned atom after the drift and the reduction took the MAX over all sub-boxes, so
    // the value also bounds every ghost. With off-lattice activity anywhere (run-aways leave, inter atoms may
    // re-occupy vacancies far from the site) it is re-measured after the ghost exchange instead.
    memcpy(&c->dmax2, c->h_stepinfo + 1, sizeof(double));
    c->dmax_valid = !c->inter_active;
    // marking level of the NEXT step: three levels under the current global maximum -- a few dozen atoms of millions lie above
    c->mark_T_next = std::max(0, (int)std::min(ceil((sqrt(c->dmax2) + 1e-6) / (0.01 * c->geo.a)), 1000.0) - 3);
    if (c->inter_active) {
        Slot sl(c, MISA_B200_K_INTER);
        TRY(inter_decide(c, c->last_runaways));
        TRY(inter_exchange(c));
        TRY(inter_border(c));
    } else {
        inter_drop_ghosts(c);
    }
    return 0;
}

// NewtonMotion::firststep + atom::decide (+ exchangeInter / borderInter when off-lattice atoms exist)
extern "C" int misa_b200_pass_verlet1(misa_b200_ctx *c) {
    TRY(ready(c));
    const VerletPar vp = verlet_par(c);
    if (c->n_inter_local > 0) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_first_step(c, vp)); }
    TRY(verlet1_enqueue(c));
    TRY(update_activity(c));
    return verlet1_finish(c);
}

extern "C" int misa_b200_pass_verlet2(misa_b200_ctx *c) {
    TRY(ready(c));
    const int bpp = nblk(c->geo.n_cells_owned);
    const VerletPar vp = verlet_par(c);
    {
        Slot sl(c, MISA_B200_K_VERLET2);
        k_verlet2<<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(c->geo, soa_v(c->s), vp, bpp);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (c->n_inter_local > 0) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_second_step(c, vp)); }
    return 0;
}

// atom::computeEam (reference src/atom.cpp:102-149) on the resident state. With the full-list gather the rho
// and force reverse halos of the reference carry nothing and are omitted.
static int compute_eam(misa_b200_ctx *c) {
    const bool inter = has_inter(c);
    if (inter) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_make_index(c)); }
    const bool fuse = c->opt_fuse && !inter;
    TRY(launch_rho(c, fuse, false));
    if (inter) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_rho(c)); }
    if (!fuse) TRY(launch_df(c));
    TRY(misa_b200_pass_halo_df(c));
    TRY(launch_force(c, false));
    if (inter) { Slot sl(c, MISA_B200_K_INTER); TRY(inter_force(c)); }
    return 0;
}

extern "C" int misa_b200_prepare(misa_b200_ctx *c) {
    TRY(ready(c));
    TRY(misa_b200_pass_halo_x(c)); // exchangeAtomFirst: the lists are static, built in misa_b200_create
    TRY(census_local(c));          // ghosts are filled now; sum over sub-boxes so every rank plans alike
    c->census_boxes = 1;
    if (c->comm_size > 1 && c->nccl_comm) {
        NC(g_nccl.AllReduce(c->d_census, c->d_census, MISA_MAX_TYPES, kNcclUint64, kNcclSum, c->nccl_comm, c->stream));
        c->census_boxes = c->comm_size;
    }
    TRY(census_fetch(c));
    c->census_boxes = 1;
    TRY(measure_displacement(c));
    TRY(build_minor_lists(c));     // dilute alloys: species sit on fixed sites until something runs away
    CU(cudaMemsetAsync(c->d_counters, 0, sizeof(int), c->stream));
    TRY(update_activity(c));
    if (c->inter_active) { TRY(inter_exchange(c)); TRY(inter_border(c)); }
    TRY(misa_b200_pass_clear(c));
    TRY(compute_eam(c));
    return 0;
}

// One step of the thermal (no off-lattice atom) case without a host round trip on the critical path, with the ghost
// exchange overlapped with interior-cell compute (the north-star's halo/compute overlap):
//   main  : verlet1 | rho(interior)            | rho(boundary) | force(interior)         | force(boundary) | verlet2
//   stream2:        | activity, halo_x (NCCL)  |               | halo_df (NCCL)          |
// The stencil kernels pick their pruned offset list on the device from the displacement k_verlet1 just measured
// (interior: this sub-box's own maximum -- every neighbour is owned; boundary: the all-reduced one). The host only
// looks at the step's activity word right before verlet2 -- it has long arrived by then -- and, if a run-away
// turned up anywhere, discards the speculative rho/force (they only wrote rho, df, f) and redoes the step serially
// through the off-lattice path. With a 1x1x1 grid the "exchange" is the periodic ghost fill and runs in line.
static bool pipe_ok(const misa_b200_ctx *c) {
    StagePlan sp;
    size_t sb;
    // (profiling keeps the sync-free step: the slot events are recorded on the main stream between its kernels)
    return c->opt_pipe && c->opt_fast && c->opt_prune && c->opt_fuse && c->tex_all && !has_inter(c) && !c->inter_active &&
           c->stream2 && make_plan(c, sp, sb);
}
// The ghost pushes of the sync-free step done from INSIDE the producing kernels (k_verlet1: positions, k_rho_f's epilogue: df),
// ARRIVE posted by the consuming stencil kernel's first CTA: needs the third-generation kernels on both stencil launches.
static bool fused_ok(const misa_b200_ctx *c) {
    StagePlan sp;
    size_t sb;
    if (!c->opt_push_fused || !c->opt_dmax_flags || !c->d_p2p_dev || !c->opt_fast || !c->tex_all || !make_plan(c, sp, sb)) return false;
    if (sym_ok(c, sp, false, StencilOpt())) return false;
    return sp.single >= 0 || dilute_ok(c, sp, false);
}
// kick2_in: the previous step of this call left its second half-kick to this step's k_verlet1<true>;
// defer_out: the caller runs another step right after this one, so this step may do the same (*deferred says it did).
static int step_pipelined(misa_b200_ctx *c, bool &redone, bool kick2_in = false, bool defer_out = false, bool *deferred = nullptr) {
    redone = false;
    if (deferred) *deferred = false;
    // The interior/boundary split pays when the exchange is long against what it costs (8 SMs left to the exchange
    // kernels + two extra launches): measured on B200, 100^3 cells per GPU -- 2x1x1 grid (one NCCL stage, 0.11 ms of
    // exchange): 1.50 ms in line vs 1.59 ms split; 2x2x2 grid (three stages, 0.38 ms): 1.81 ms in line vs 1.61 ms
    // split. "overlap": -1 auto (split when at least two dimensions exchange over NCCL), 0 never, 1 whenever any
    // dimension does, 2 always (tests: also on a 1x1x1 grid).
    int nccl_dims = 0;
    for (int d = 0; d < 3; d++) nccl_dims += c->dom.grid_size[d] > 1;
    // with the direct push (p2p.cuh) an exchange costs tens of microseconds: nothing left to hide
    const bool p2p = c->p2p_active && c->opt_p2p;
    const bool overlap = c->opt_overlap > 1 || (c->opt_overlap == 1 && nccl_dims >= 1) || (c->opt_overlap < 0 && nccl_dims >= 2 && !p2p);
    const bool fused = p2p && !overlap && c->stream2 && fused_ok(c);
    unsigned long long e_x = 0, e_df = 0;
    if (fused) {
        // READY for both exchanges of this step (unless the previous step of this call sent it ahead), then the gate: every
        // destination has freed its ghosts -- k_verlet1 and k_rho_f store into them without a look at the flags
        e_x = c->p2p_epoch + 1; e_df = c->p2p_epoch + 2;
        const unsigned long long post = c->p2p_ready_sent < e_df ? e_df : 0ULL;
        c->p2p.fence_mode = c->opt_p2p_fence;
        k_p2p_ready_gate<<<1, 32, 0, c->stream>>>(c->p2p, post, e_df, c->d_flags, c->d_p2p_err);
        c->launches++;
        CU(cudaGetLastError());
        c->p2p_ready_sent = std::max(c->p2p_ready_sent, e_df);
    }
    TRY(verlet1_enqueue(c, kick2_in, fused));
    StencilOpt whole, interior, boundary;
    whole.dmax2 = c->d_stepinfo_g + 1;
    interior.region = 1; interior.dmax2 = c->d_stepinfo + 1; interior.reserve_sms = c->opt_reserve;
    boundary.region = 2; boundary.dmax2 = c->d_stepinfo_g + 1;
    if (!overlap) {
        if (p2p && c->stream2) {
            // The all-reduce of the activity / displacement words (NCCL, latency-bound) runs on stream2 and only the HOST waits
            // for it, at the end of the step (has anything run away anywhere?). The stencil kernels no longer do: the partner
            // bound of their pruning is the maximum over this sub-box and the 26 around it -- every partner of an owned atom
            // lives there -- and those maxima travel with the position push (P2P_DMAX words in front of ARRIVE). Option
            // "dmax_flags" 0 restores the all-reduce in front of rho.
            CU(cudaEventRecord(c->ev_v1, c->stream));
            CU(cudaStreamWaitEvent(c->stream2, c->ev_v1, 0));
            TRY(activity_enqueue(c, c->stream2));
            CU(cudaEventRecord(c->ev_act, c->stream2));
            c->dmax_by_flags = c->opt_dmax_flags != 0;
            whole.late = true;                      // the wait for the neighbours' pushes moves into the stencil kernels
            if (fused) {
                // k_verlet1 has stored the band sites' positions into the neighbours' ghosts; rho's first CTA posts ARRIVE
                // (+ this sub-box's displacement maximum), its epilogue pushes df; force's first CTA posts that ARRIVE
                whole.fused = true;
                c->p2p_epoch = e_x;
                whole.post_epoch = e_x; whole.post_dmax = c->d_stepinfo + 1;
                TRY(launch_rho(c, true, false, whole));
                c->p2p_epoch = e_df;
                whole.post_epoch = e_df; whole.post_dmax = nullptr;
                TRY(launch_force(c, false, whole));
            } else {
                { Slot sl(c, MISA_B200_K_HALO_X); TRY(p2p_push(c, true, c->stream)); }
                if (!c->dmax_by_flags) CU(cudaStreamWaitEvent(c->stream, c->ev_act, 0));
                TRY(launch_rho(c, true, false, whole));
                { Slot sl(c, MISA_B200_K_HALO_DF); TRY(p2p_push(c, false, c->stream)); }
                TRY(launch_force(c, false, whole));
            }
            c->dmax_by_flags = false;
        } else {
            TRY(activity_enqueue(c, c->stream));
            CU(cudaEventRecord(c->ev_act, c->stream));
            { Slot sl(c, MISA_B200_K_HALO_X); TRY(halo_forward(c, true)); }
            TRY(launch_rho(c, true, false, whole));
            { Slot sl(c, MISA_B200_K_HALO_DF); TRY(halo_forward(c, false)); }
            TRY(launch_force(c, false, whole));
        }
        // ghost x and df are free for the next step's two exchanges -- only when that step follows inside this call:
        // between calls the host may run readers of the ghosts (thermo, dump) that a neighbour's next push must not overtake
        if (defer_out && !fused) TRY(p2p_post_ready(c, 2, c->stream));   // (fused: the next step's gate kernel posts it)
    } else {
        CU(cudaEventRecord(c->ev_v1, c->stream));
        CU(cudaStreamWaitEvent(c->stream2, c->ev_v1, 0));
        TRY(activity_enqueue(c, c->stream2));
        CU(cudaEventRecord(c->ev_act, c->stream2));
        TRY(halo_forward(c, true, c->stream2));
        CU(cudaEventRecord(c->ev_hx, c->stream2));
        TRY(launch_rho(c, true, false, interior));
        CU(cudaStreamWaitEvent(c->stream, c->ev_hx, 0));
        TRY(launch_rho(c, true, false, boundary));
        CU(cudaEventRecord(c->ev_rho, c->stream));
        CU(cudaStreamWaitEvent(c->stream2, c->ev_rho, 0));
        TRY(halo_forward(c, false, c->stream2));
        CU(cudaEventRecord(c->ev_hdf, c->stream2));
        TRY(launch_force(c, false, interior));
        CU(cudaStreamWaitEvent(c->stream, c->ev_hdf, 0));
        TRY(launch_force(c, false, boundary));
    }
    CU(cudaEventSynchronize(c->ev_act));
    TRY(p2p_check(c));
    c->inter_active = c->h_stepinfo[0] > 0;
    c->pipe_steps++;
    if (c->inter_active || c->h_counters[3] != 0) {
        // off-lattice activity somewhere: x and v are as verlet1 left them (rho/force never touch them), so the
        // serial path can take over from exactly there
        c->seen_offlattice = true;
        c->pipe_redo++;
        redone = true;
        CU(cudaStreamSynchronize(c->stream2));
        CU(cudaStreamSynchronize(c->stream));
        TRY(verlet1_finish(c));
        TRY(misa_b200_pass_halo_x(c));
        if (!c->dmax_valid) TRY(measure_displacement(c));
        if (has_inter(c)) TRY(inter_clear(c));
        TRY(compute_eam(c));
    } else {
        TRY(verlet1_finish(c));
        if (defer_out && c->opt_fuse_verlet && pipe_ok(c)) {   // the next k_verlet1 applies this step's second half-kick too
            if (deferred) *deferred = true;
            return 0;
        }
    }
    return misa_b200_pass_verlet2(c);
}

extern "C" int misa_b200_step(misa_b200_ctx *c, int n_steps) {
    TRY(ready(c));
    bool pending_kick2 = false;   // never crosses the end of this call: the state a caller sees is a full step's
    for (int s = 0; s < n_steps; s++) {
        if (pipe_ok(c)) {
            bool redone;
            TRY(step_pipelined(c, redone, pending_kick2, s + 1 < n_steps, &pending_kick2));
            continue;
        }
        if (pending_kick2) { TRY(misa_b200_pass_verlet2(c)); pending_kick2 = false; }
        TRY(misa_b200_pass_verlet1(c));
        TRY(misa_b200_pass_halo_x(c));
        if (!c->dmax_valid) TRY(measure_displacement(c));
        // clearForce is folded into the stencil kernels' stores (owned sites are overwritten)
        if (has_inter(c)) TRY(inter_clear(c));
        TRY(compute_eam(c));
        TRY(misa_b200_pass_verlet2(c));
    }
    if (pending_kick2) TRY(misa_b200_pass_verlet2(c));
    return 0;
}

// Host-buffer form of the step loop: the AoS array the reference driver owns goes in, n steps run on the device and
// the OWNED records come back (every field), so the unmodified host code around the loop (dump, thermo, stage machine
// -- reference frontend/md_simulation.cpp:40-121, all of which read owned sites only) sees current data. Ghost records
// carry no information across a step in the reference either: x / type / df are refilled by exchangeAtom and the df
// halo before they are read, rho / f are cleared (src/atom.cpp:86-146) -- so after the first call (full upload: the
// species census covers the ghost shell) only the owned box crosses PCIe, as one pitched 3-D copy each way, and the
// host's ghost records are left as they were.
// ---- slab-pipelined form of one host-buffer step (single sub-box, thermal case) ------------------------------------------------
// The serial form is PCIe-bound and uses one direction at a time: H2D 208 MB (3.7 ms at 100^3 cells), step (1.1 ms), D2H 208 MB
// (3.7 ms). Here the owned box is cut into z-slabs of slab_T planes (>= the stencil reach of 2.5 cells, so a slab depends on its
// two neighbours only). Uploads run on one copy stream in slab order; as slab k lands it is converted, integrated (k_verlet1 on
// its cell range) and its ghost images filled (the fill map re-sorted by the source site's slab); rho + df of slab k-1 and force +
// second half-kick of slab k-2 follow, whose records then go back D2H on a second copy stream -- both PCIe directions busy at
// once. The periodic wrap (slab 0 needs the last slab) leaves rho of two slabs and force of four for the tail. Partner bound of
// the stencil pruning: the running maximum of the displacements measured so far -- every partner of slab k lies in slabs k-1 ..
// k+1, which have been measured when rho(k) is enqueued. The INPUT records stay intact in their own staging array: if an atom
// turns out to have run away, the state is rebuilt from them and the step redone by the serial path (bit-identical results
// otherwise: tests/test_gpu_parity.py).
static int slab_setup(misa_b200_ctx *c) {
    const Geo &g = c->geo;
    if (c->n_slabs > 0) return 0;
    const int T = std::max(3, c->slab_T);
    int S = g.nz / T;
    if (S < 6) return -1;
    std::vector<int> z0(S + 1);
    for (int k = 0; k < S; k++) z0[k] = k * T;
    z0[S] = g.nz;                                        // the last slab takes the remainder (T .. 2T-1 planes)
    // fill map by source slab
    const size_t n = c->h_ghost_dst.size();
    std::vector<int> slab_of(n), order(n), ofs(S + 1, 0);
    const long long plane = (long long)g.sxc * g.sy;
    for (size_t i = 0; i < n; i++) {
        const long long rem = c->h_ghost_src[i] % g.H;
        const int z = (int)(rem / plane) - g.gz;
        int k = std::min(S - 1, std::max(0, z / T));
        slab_of[i] = k;
        ofs[k + 1]++;
    }
    for (int k = 0; k < S; k++) ofs[k + 1] += ofs[k];
    std::vector<int> fill(ofs.begin(), ofs.end() - 1), dst(n), src(n);
    std::vector<int8_t> code(n);
    for (size_t i = 0; i < n; i++) { const int q = fill[slab_of[i]]++; dst[q] = c->h_ghost_dst[i]; src[q] = c->h_ghost_src[i]; code[q] = c->h_ghost_code[i]; }
    cudaFree(c->d_slab_dst); cudaFree(c->d_slab_src); cudaFree(c->d_slab_code);
    c->d_slab_dst = c->d_slab_src = nullptr; c->d_slab_code = nullptr;
    TRY(dmalloc(&c->d_slab_dst, n)); TRY(dmalloc(&c->d_slab_src, n)); TRY(dmalloc(&c->d_slab_code, n));
    CU(cudaMemcpy(c->d_slab_dst, dst.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_slab_src, src.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_slab_code, code.data(), n, cudaMemcpyHostToDevice));
    if (!c->d_aos_out) {
        TRY(dmalloc(&c->d_aos_out, (size_t)g.n_ext * 104));
        CU(cudaMemset(c->d_aos_out, 0, (size_t)g.n_ext * 104));   // the padding word of the records goes out as zeros
    }
    if (!c->s_up) { CU(cudaStreamCreateWithFlags(&c->s_up, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&c->s_dn, cudaStreamNonBlocking)); }
    while ((int)c->ev_up.size() < S) {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        c->ev_up.push_back(a); c->ev_out.push_back(b);
    }
    c->slab_z0 = z0; c->slab_map_ofs = ofs; c->n_slabs = S;
    return 0;
}
// z-planes [z0, z1) of the owned box of a ghost-extended AoS array as one pitched 3-D copy
static int copy_owned_slab(misa_b200_ctx *c, void *dst, const void *src, cudaMemcpyKind kind, int z0, int z1, cudaStream_t st) {
    const Geo &g = c->geo;
    const size_t rec = 104, pitch = 2 * (size_t)g.sxc * rec;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(src), pitch, pitch, (size_t)g.sy);
    p.dstPtr = make_cudaPitchedPtr(dst, pitch, pitch, (size_t)g.sy);
    p.srcPos = p.dstPos = make_cudaPos(2 * (size_t)g.gx * rec, (size_t)g.gy, (size_t)(g.gz + z0));
    p.extent = make_cudaExtent(2 * (size_t)g.nx * rec, (size_t)g.ny, (size_t)(z1 - z0));
    p.kind = kind;
    CU(cudaMemcpy3DAsync(&p, st));
    return 0;
}
static bool host_slabs_ok(misa_b200_ctx *c) {
    StagePlan sp;
    size_t sb;
    if (!c->opt_host_slabs || !c->all_self || c->comm_size != 1 || c->n_ghost_map <= 0 || c->h_ghost_dst.empty() || c->prof_on) return false;
    if (!pipe_ok(c) || !make_plan(c, sp, sb) || sp.single < 0 || !c->dmax_valid) return false;   // thermal, single species (the minority-atom launch of dilute alloys is per box)
    return slab_setup(c) == 0;
}
static int step_host_slabs(misa_b200_ctx *c, void *atoms, bool &redo) {
    const Geo &g = c->geo;
    const int S = c->n_slabs;
    const long long row = 2LL * g.sxc, plane_recs = row * g.sy, cells_plane = (long long)g.nx * g.ny;
    redo = false;
    // the step's words, marks: as verlet1_enqueue
    VerletPar vp = verlet_par(c);
    c->mark_valid = false;
    c->pmax_valid = false;
    if (c->opt_mark && c->opt_prune) {
        c->mark_epoch = c->mark_epoch % 255 + 1;
        c->mark_T_used = c->mark_T_next;
        vp.mark_T = c->mark_T_used; vp.hot = c->d_hot; vp.epoch = (unsigned char)c->mark_epoch; vp.mark_count = c->d_stepinfo + 2;
        c->mark_valid = true;
    }
    CU(cudaMemsetAsync(c->d_stepinfo, 0, 3 * sizeof(unsigned long long) + sizeof(int), c->stream));
    CU(cudaEventRecord(c->ev_v1, c->stream));
    CU(cudaStreamWaitEvent(c->s_up, c->ev_v1, 0));          // uploads may not overtake whatever the previous call left on the main stream
    CU(cudaStreamWaitEvent(c->s_dn, c->ev_v1, 0));
    for (int k = 0; k < S; k++) {
        TRY(copy_owned_slab(c, c->d_aos, atoms, cudaMemcpyHostToDevice, c->slab_z0[k], c->slab_z0[k + 1], c->s_up));
        CU(cudaEventRecord(c->ev_up[k], c->s_up));
    }
    StencilOpt so;
    so.dmax2 = c->d_stepinfo + 1;                           // running maximum over the slabs integrated so far
    auto slab_in = [&](int k) -> int {                      // records -> SoA, first half-kick + drift, ghost images of the slab's sites
        const int z0 = c->slab_z0[k], z1 = c->slab_z0[k + 1];
        CU(cudaStreamWaitEvent(c->stream, c->ev_up[k], 0));
        const long long i0 = (long long)(g.gz + z0) * plane_recs, i1 = (long long)(g.gz + z1) * plane_recs;
        k_aos_to_soa<<<nblk(i1 - i0), MISA_BLOCK, 0, c->stream>>>(i1, g.H, (const unsigned long long *)c->d_aos, c->s, F_X | F_V | F_F, g, 1, i0);
        VerletPar v = vp;
        v.c_begin = z0 * cells_plane; v.c_end = z1 * cells_plane;
        const int bpp = nblk(v.c_end - v.c_begin);
        k_verlet1<false><<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(g, soa_v(c->s), v, bpp, c->d_counters, c->d_runaway, c->inter_cap, c->d_stepinfo);
        const int m0 = c->slab_map_ofs[k], m = c->slab_map_ofs[k + 1] - m0;
        if (m > 0)
            k_ghost_fill_x<<<nblk(m), MISA_BLOCK, 0, c->stream>>>(m, c->d_slab_dst + m0, c->d_slab_src + m0, c->d_slab_code + m0, c->s, c->dom.meas_global_length[0],
                                                               c->dom.meas_global_length[1], c->dom.meas_global_length[2]);
        c->launches += 3;
        CU(cudaGetLastError());
        return 0;
    };
    auto slab_rho = [&](int k) -> int {
        so.z0 = c->slab_z0[k]; so.z1 = c->slab_z0[k + 1];
        TRY(launch_rho(c, true, false, so));
        const int m0 = c->slab_map_ofs[k], m = c->slab_map_ofs[k + 1] - m0;
        if (m > 0) { k_ghost_fill_1<<<nblk(m), MISA_BLOCK, 0, c->stream>>>(m, c->d_slab_dst + m0, c->d_slab_src + m0, c->s.df); c->launches++; }
        CU(cudaGetLastError());
        return 0;
    };
    auto slab_out = [&](int k) -> int {                     // force, second half-kick, SoA -> records, D2H
        const int z0 = c->slab_z0[k], z1 = c->slab_z0[k + 1];
        so.z0 = z0; so.z1 = z1;
        TRY(launch_force(c, false, so));
        VerletPar v = vp;
        v.c_begin = z0 * cells_plane; v.c_end = z1 * cells_plane;
        const int bpp = nblk(v.c_end - v.c_begin);
        k_verlet2<<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(g, soa_v(c->s), v, bpp);
        const long long i0 = (long long)(g.gz + z0) * plane_recs, i1 = (long long)(g.gz + z1) * plane_recs;
        k_soa_to_aos<<<nblk(i1 - i0), MISA_BLOCK, 0, c->stream>>>(g, (unsigned long long *)c->d_aos_out, c->s, F_ALL, 1, i0, i1);
        c->launches += 2;
        CU(cudaGetLastError());
        CU(cudaEventRecord(c->ev_out[k], c->stream));
        CU(cudaStreamWaitEvent(c->s_dn, c->ev_out[k], 0));
        return copy_owned_slab(c, atoms, c->d_aos_out, cudaMemcpyDeviceToHost, z0, z1, c->s_dn);
    };
    for (int k = 0; k < S; k++) {
        TRY(slab_in(k));
        if (k >= 2) TRY(slab_rho(k - 1));                   // slabs k-2, k-1, k are in
        if (k >= 4) TRY(slab_out(k - 2));                   // rho of k-3, k-2, k-1 is done
    }
    TRY(slab_rho(S - 1));
    TRY(slab_rho(0));
    for (int k : {S - 2, S - 1, 0, 1}) TRY(slab_out(k));
    TRY(activity_enqueue(c, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->s_dn));
    c->inter_active = c->h_stepinfo[0] > 0;
    c->host_slab_steps++;
    if (c->inter_active || c->h_counters[3] != 0) {
        // something ran away: the records that just went back are void. Rebuild the state from the intact input records and let
        // the caller redo the step through the serial path (decide, inter-atom lists).
        c->host_slab_redo++;
        k_aos_to_soa<<<nblk(g.n_ext), MISA_BLOCK, 0, c->stream>>>(g.n_ext, g.H, (const unsigned long long *)c->d_aos, c->s, F_X | F_V | F_F, g, 1);
        c->launches++;
        CU(cudaGetLastError());
        c->inter_active = false;
        redo = true;
        return 0;
    }
    TRY(verlet1_finish(c));
    return 0;
}

extern "C" int misa_b200_step_host(misa_b200_ctx *c, void *atoms, int n_steps) {
    REQ(c && atoms, MISA_B200_EINVAL, "misa_b200_step_host: null argument");
    REQ(c->have_off && c->have_pot, MISA_B200_ESTATE, "misa_b200_step_host: offsets / potential not set");
    const bool first = !c->census_valid || !c->have_atoms;
    if (!first && n_steps == 1 && host_slabs_ok(c)) {
        bool redo = false;
        TRY(step_host_slabs(c, atoms, redo));
        if (!redo) return 0;
        // the state is the caller's input again (rebuilt on the device): serial step + download
        TRY(misa_b200_step(c, 1));
        return d2h_aos(c, atoms, F_ALL, 2);
    }
    TRY(h2d_aos(c, atoms, F_ALL, first ? 0 : 1));
    if (!c->census_valid) { TRY(census_local(c)); TRY(census_fetch(c)); }
    c->have_atoms = true;
    TRY(misa_b200_step(c, n_steps));
    return d2h_aos(c, atoms, F_ALL, first ? 0 : 2);
}

extern "C" int misa_b200_timed_steps(misa_b200_ctx *c, int n_steps, double *ms) {
    TRY(ready(c));
    REQ(ms, MISA_B200_EINVAL, "null argument");
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventRecord(e0, c->stream));
    int rc = misa_b200_step(c, n_steps);
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float f = 0;
    CU(cudaEventElapsedTime(&f, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms = f;
    return rc;
}

extern "C" int misa_b200_stencil_stats(misa_b200_ctx *c, double out[4]) {
    TRY(ready(c));
    REQ(out, MISA_B200_EINVAL, "null argument");
    unsigned long long *d = nullptr, h[4] = {0, 0, 0, 0};
    TRY(dmalloc(&d, 4));
    CU(cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), c->stream));
    const RegionList rl = make_regions(c->geo, 0);
    // the word the sync-free step's kernels read: the global maximum displacement of the last k_verlet1 (host level otherwise)
    const LevelSel ls = make_levelsel(c, pipe_ok(c) && c->pipe_steps > 0 ? c->d_stepinfo_g + 1 : nullptr);
    k_stencil_stats<<<std::max(1, c->sm_count) * 8, 256, 0, c->stream>>>(c->geo, c->s, c->d_off_full, c->n_full, c->near_full, rl, ls, d);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CU(e);
    for (int k = 0; k < 4; k++) out[k] = (double)h[k];
    return 0;
}

// atom::setv, reference src/atom.cpp:475-494
extern "C" int misa_b200_setv(misa_b200_ctx *c, const int32_t lat[4], const double direction[3], double energy) {
    TRY(ready(c));
    const Geo &g = c->geo;
    const long long xl = 2LL * g.lo[0], yl = g.lo[1], zl = g.lo[2];
    if ((lat[0] * 2) >= xl && (lat[0] * 2) < xl + 2LL * g.nx && lat[1] >= yl && lat[1] < yl + g.ny && lat[2] >= zl && lat[2] < zl + g.nz) {
        const long long sx = 2LL * g.sxc;
        const long long kk = ((lat[2] - (zl - g.gz)) * (long long)g.sy + (lat[1] - (yl - g.gy))) * sx + (lat[0] * 2 - (xl - 2LL * g.gx)) + lat[3];
        const long long d = ref_to_dev(kk, g.H);
        int8_t t;
        double v[3];
        CU(cudaMemcpyAsync(&t, c->s.type + d, 1, cudaMemcpyDeviceToHost, c->stream));
        for (int k = 0; k < 3; k++) CU(cudaMemcpyAsync(&v[k], c->s.v[k] + d, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const double mass = t >= 0 ? kMass[t] : 0.0;
        const double v_ = sqrt(2 * energy / mass / kMvv2e);
        const double d_ = sqrt(direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2]);
        for (int k = 0; k < 3; k++) {
            v[k] += v_ * direction[k] / d_;
            CU(cudaMemcpyAsync(c->s.v[k] + d, &v[k], 8, cudaMemcpyHostToDevice, c->stream));
        }
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// simulation::collisionStep, reference src/simulation.cpp:208-217
extern "C" int misa_b200_collision_step(misa_b200_ctx *c, const int32_t lat[4], const double direction[3], double energy) {
    TRY(misa_b200_setv(c, lat, direction, energy));
    if (c->inter_active) { TRY(inter_exchange(c)); TRY(inter_border(c)); }
    TRY(misa_b200_pass_halo_x(c));
    TRY(measure_displacement(c));
    TRY(misa_b200_pass_clear(c));
    return compute_eam(c);
}

extern "C" int misa_b200_thermo(misa_b200_ctx *c, double out[6]) {
    TRY(ready(c));
    REQ(out, MISA_B200_EINVAL, "null argument");
    const Geo &g = c->geo;
    const int bpp = nblk(g.n_cells_owned);
    CU(cudaMemsetAsync(c->d_reduce, 0, 8 * sizeof(double), c->stream));
    k_thermo<<<2 * bpp, MISA_BLOCK, (size_t)c->n_full * sizeof(int), c->stream>>>(g, c->s, c->tab, c->d_off_full, c->n_full, bpp,
                                                                                 kMass[0], kMass[1], kMass[2], c->d_reduce);
    c->launches++;
    CU(cudaGetLastError());
    if (c->n_inter_local > 0) TRY(inter_thermo(c, c->d_reduce));
    CU(cudaMemcpyAsync(c->h_reduce, c->d_reduce, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out[0] = c->h_reduce[0];
    out[1] = c->h_reduce[1];
    out[2] = c->h_reduce[2];
    out[3] = c->n_inter_local;
    out[4] = c->n_inter_ghost;
    out[5] = c->last_runaways;
    return 0;
}

// configuration::rescale, reference src/system_configuration.cpp:86-111 (this rank's atoms; the caller
// supplies the GLOBAL temperature factor through t_set / current T computed from all ranks' thermo[0]).
extern "C" int misa_b200_rescale(misa_b200_ctx *c, double t_set, double t_now) {
    TRY(ready(c));
    REQ(t_now > 0, MISA_B200_EINVAL, "misa_b200_rescale: current temperature must be positive");
    const double fac = sqrt(t_set / t_now);
    k_scale_v<<<nblk(c->geo.n_ext), MISA_BLOCK, 0, c->stream>>>(c->geo.n_ext, c->s.v[0], c->s.v[1], c->s.v[2], fac);
    c->launches++;
    CU(cudaGetLastError());
    if (c->n_inter_local > 0) TRY(inter_scale_v(c, fac));
    (void)kBoltz;
    return 0;
}

// -------------------------------------------------------------------------------------------------
// compat hooks on the HOST AoS array
// -------------------------------------------------------------------------------------------------
static int hook_common(misa_b200_ctx *c, void *atoms, double cutoff_radius) {
    REQ(c && atoms, MISA_B200_EINVAL, "null argument");
    REQ(c->have_off, MISA_B200_ESTATE, "neighbour offsets not set (cuda_nei_offset_init)");
    REQ(c->have_pot, MISA_B200_ESTATE, "potential not set (cuda_pot_init)");
    c->geo.rc2 = cutoff_radius * cutoff_radius; // reference src/atom.cpp:177
    return 0;
}

extern "C" int misa_b200_eam_rho_calc(misa_b200_ctx *c, void *atoms, double cutoff_radius) {
    TRY(hook_common(c, atoms, cutoff_radius));
    TRY(h2d_aos(c, atoms, F_TYPE | F_X | F_RHO));
    c->have_atoms = true;
    TRY(census_local(c));
    TRY(census_fetch(c));
    TRY(measure_displacement(c));
    TRY(launch_rho(c, false, true));
    return d2h_aos(c, atoms, F_RHO, 2);   // only the owned box crosses PCIe on the way back: ghost records are not ours to write
}
extern "C" int misa_b200_eam_df_calc(misa_b200_ctx *c, void *atoms, double cutoff_radius) {
    TRY(hook_common(c, atoms, cutoff_radius));
    TRY(h2d_aos(c, atoms, F_TYPE | F_RHO, c->have_atoms ? 1 : 0));   // latDf reads and writes owned sites only (src/atom.cpp:286-309)
    c->have_atoms = true;
    TRY(launch_df(c));
    return d2h_aos(c, atoms, F_DF, 2);
}
extern "C" int misa_b200_eam_force_calc(misa_b200_ctx *c, void *atoms, double cutoff_radius) {
    TRY(hook_common(c, atoms, cutoff_radius));
    TRY(h2d_aos(c, atoms, F_TYPE | F_X | F_DF | F_F));
    c->have_atoms = true;
    TRY(census_local(c));
    TRY(census_fetch(c));
    TRY(measure_displacement(c));
    TRY(launch_force(c, true));
    return d2h_aos(c, atoms, F_F, 2);
}

// -------------------------------------------------------------------------------------------------
// inter-atom list transfer
// -------------------------------------------------------------------------------------------------
extern "C" int misa_b200_upload_inter(misa_b200_ctx *c, const void *inter_atoms, size_t n) {
    REQ(c, MISA_B200_EINVAL, "null ctx");
    return inter_upload(c, inter_atoms, n);
}
extern "C" int misa_b200_download_inter(misa_b200_ctx *c, void *inter_atoms, size_t cap, size_t *n) {
    REQ(c && n, MISA_B200_EINVAL, "null argument");
    return inter_download(c, inter_atoms, cap, n);
}

// -------------------------------------------------------------------------------------------------
// initial state on the device, by global atom id: WorldBuilder::build (reference src/world_builder.cpp:64-103)
// -------------------------------------------------------------------------------------------------
extern "C" int misa_b200_build_world(misa_b200_ctx *c, uint32_t seed, double t_set, const int32_t ratio[3], uint64_t alloy_seed) {
    REQ(c && ratio, MISA_B200_EINVAL, "misa_b200_build_world: null argument");
    const Geo &g = c->geo;
    WorldPar w;
    w.px = c->dom.phase_space[0]; w.py = c->dom.phase_space[1]; w.pz = c->dom.phase_space[2];
    w.n_global = 2 * w.px * w.py * w.pz;
    w.ratio_total = 0;
    w.single = -1;
    int nonzero = 0;
    for (int i = 0; i < MISA_MAX_TYPES; i++) {
        REQ(ratio[i] >= 0, MISA_B200_EINVAL, "misa_b200_build_world: negative alloy ratio");
        w.ratio[i] = ratio[i];
        w.ratio_total += ratio[i];
        w.mass[i] = kMass[i];
        if (ratio[i] > 0) { nonzero++; w.single = i; }
    }
    REQ(w.ratio_total > 0, MISA_B200_EINVAL, "misa_b200_build_world: alloy ratio sums to zero");
    if (nonzero != 1) w.single = -1;
    w.alloy_seed = alloy_seed;
    w.a = c->dom.lattice_const;
    // the global mt19937 stream: draw 3(id-1)+k is velocity component k of atom id (= the reference on ONE rank)
    unsigned *d_draws = nullptr;
    double *d_part = nullptr, *d_out = nullptr;
    const int nb = (int)std::min<long long>((w.n_global + MISA_BLOCK - 1) / MISA_BLOCK, (long long)std::max(c->sm_count, 1) * 8);
    CU(cudaMalloc((void **)&d_draws, (size_t)w.n_global * 3 * sizeof(unsigned)));
    int rc = 0;
    double h[5] = {0, 0, 0, 0, 0};
    do {
        if ((rc = dmalloc(&d_part, (size_t)nb * 4))) break;
        if ((rc = dmalloc(&d_out, 5))) break;
        k_mt19937_stream<<<1, 256, 0, c->stream>>>(seed, 3 * w.n_global, d_draws);
        k_world_moments<<<nb, MISA_BLOCK, 0, c->stream>>>(w, d_draws, d_part);
        k_sum_partials<4><<<1, MISA_BLOCK, 0, c->stream>>>(d_part, nb, d_out);
        c->launches += 3;
        if (cudaMemcpyAsync(h, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = -1; break; }
        double vcm[3] = {h[0], h[1], h[2]};
        if (h[3] > 0.0) for (int k = 0; k < 3; k++) vcm[k] /= (double)w.n_global;   // world_builder.cpp:84-88
        double factor = 1.0;
        if (t_set != 0.0) {                                                          // :96-101 -> configuration::rescale
            k_world_mvv<<<nb, MISA_BLOCK, 0, c->stream>>>(w, d_draws, vcm[0], vcm[1], vcm[2], d_part);
            k_sum_partials<1><<<1, MISA_BLOCK, 0, c->stream>>>(d_part, nb, d_out + 4);
            c->launches += 2;
            if (cudaMemcpyAsync(h + 4, d_out + 4, sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = -1; break; }
            const unsigned long long dof = 3ull * (unsigned long long)w.n_global - 3ull;
            const double scalar = h[4] * kMvv2e / (dof * kBoltz);                    // configuration::temperature, system_configuration.cpp:56-64
            factor = sqrt(t_set / scalar);                                           // :97
        }
        k_world_fill<<<nblk(g.n_ext), MISA_BLOCK, 0, c->stream>>>(g, c->s, w, d_draws, vcm[0], vcm[1], vcm[2], factor);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = -1; break; }
    } while (0);
    cudaFree(d_draws); cudaFree(d_part); cudaFree(d_out);
    if (rc == -1) return fail(MISA_B200_ENODEV, std::string("misa_b200_build_world: ") + cudaGetErrorString(cudaGetLastError()));
    TRY(rc);
    TRY(inter_upload(c, nullptr, 0));
    TRY(census_local(c));
    TRY(census_fetch(c));
    c->have_atoms = true;
    c->dmax_valid = false;
    c->pmax_valid = false;
    c->minor_valid = false;
    return 0;
}

// -------------------------------------------------------------------------------------------------
// global thermo / rescale (stage machine): configuration::temperature, kineticEnergy, rescale
// (reference src/system_configuration.cpp:26-111), summed over ALL sub-boxes with one NCCL all-reduce
// -------------------------------------------------------------------------------------------------
static int global_mvv(misa_b200_ctx *c, double sums[2]) {
    const Geo &g = c->geo;
    const int bpp = nblk(g.n_cells_owned);
    double *d_part = nullptr;
    TRY(dmalloc(&d_part, (size_t)2 * bpp * 2));
    k_mvv<<<2 * bpp, MISA_BLOCK, 0, c->stream>>>(g, c->s, kMass[0], kMass[1], kMass[2], bpp, d_part);
    k_sum_partials<2><<<1, MISA_BLOCK, 0, c->stream>>>(d_part, 2 * bpp, c->d_reduce);
    c->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_reduce, c->d_reduce, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_part);
    CU(e);
    double mvv = c->h_reduce[0], n = c->h_reduce[1];
    if (c->opt_inter_dev) TRY(idev_sync_host(c));
    for (const HostAtom &a : IH(c)->local) {   // inter list after the lattice, list order (system_configuration.cpp:73-82)
        mvv += (a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]) * kMass[a.type];
        n += 1.0;
    }
    if (c->comm_size > 1) {
        REQ(c->nccl_comm, MISA_B200_ESTATE, "global thermo: communicator not initialised");
        c->h_reduce[0] = mvv; c->h_reduce[1] = n;
        CU(cudaMemcpyAsync(c->d_reduce, c->h_reduce, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        NC(g_nccl.AllReduce(c->d_reduce, c->d_reduce, 2, kNcclDouble, kNcclSum, c->nccl_comm, c->stream));
        CU(cudaMemcpyAsync(c->h_reduce, c->d_reduce, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        mvv = c->h_reduce[0]; n = c->h_reduce[1];
    }
    sums[0] = mvv; sums[1] = n;
    return 0;
}
extern "C" int misa_b200_temperature(misa_b200_ctx *c, uint64_t n_atoms_global, double out[4]) {
    REQ(c && out, MISA_B200_EINVAL, "misa_b200_temperature: null argument");
    REQ(c->have_atoms, MISA_B200_ESTATE, "misa_b200_temperature: no atoms on the device");
    REQ(n_atoms_global > 1, MISA_B200_EINVAL, "misa_b200_temperature: needs more than one atom");
    double sums[2];
    TRY(global_mvv(c, sums));
    const unsigned long long dof = 3ull * n_atoms_global - 3ull;
    out[0] = sums[0];
    out[1] = sums[0] * kMvv2e / (dof * kBoltz);  // configuration::temperature, system_configuration.cpp:56-64
    out[2] = 0.5 * sums[0] * kMvv2e;             // kinetic energy [eV]
    out[3] = sums[1];
    return 0;
}
extern "C" int misa_b200_rescale_to(misa_b200_ctx *c, double t_set, uint64_t n_atoms_global) {
    double th[4];
    TRY(misa_b200_temperature(c, n_atoms_global, th));
    REQ(th[1] > 0, MISA_B200_ESTATE, "misa_b200_rescale_to: current temperature is zero");
    return misa_b200_rescale(c, t_set, th[1]);   // v *= sqrt(T / scalar), system_configuration.cpp:97-110
}

// -------------------------------------------------------------------------------------------------
// dump record stream: AtomDump::dump + BufferedFileWriter::write (reference frontend/io/atom_dump.cpp:39-75,
// frontend/io/buffered_io.cpp:18-36), compacted on the device (dump.cuh)
// -------------------------------------------------------------------------------------------------
struct DumpRecord { // atom_dump::AtomInfoDump, reference frontend/io/atom_info_dump.h:14-22
    unsigned long long id, step;
    int type;
    short inter_type, _pad;
    double x[3], v[3];
};
static_assert(sizeof(DumpRecord) == 72, "AtomInfoDump layout");

extern "C" int misa_b200_dump_records(misa_b200_ctx *c, const int32_t begin[3], const int32_t end[3], uint64_t time_step,
                                      void *records, size_t cap_records, size_t *n_records) {
    REQ(c && n_records, MISA_B200_EINVAL, "misa_b200_dump_records: null argument");
    REQ(c->have_atoms, MISA_B200_ESTATE, "misa_b200_dump_records: no atoms on the device");
    const Geo &g = c->geo;
    DumpRegion r;
    if (begin && end) {
        REQ(begin[0] >= 0 && begin[1] >= 0 && begin[2] >= 0 && end[0] <= 2 * g.sxc && end[1] <= g.sy && end[2] <= g.sz &&
                begin[0] <= end[0] && begin[1] <= end[1] && begin[2] <= end[2],
            MISA_B200_EINVAL, "misa_b200_dump_records: region outside the ghost-extended lattice");
        r = {begin[0], begin[1], begin[2], end[0] - begin[0], end[1] - begin[1], end[2] - begin[2], 0};
    } else { // OutputBaseInterface, reference frontend/io/output_base_interface.h:26-31: the owned sub-box
        r = {2 * g.gx, g.gy, g.gz, 2 * g.nx, g.ny, g.nz, 0};
    }
    r.n = (long long)r.nx * r.ny * r.nz;
    if (c->opt_inter_dev) TRY(idev_sync_host(c));
    const size_t n_inter = IH(c)->local.size();
    const int n_tiles = (int)((r.n + DUMP_TILE - 1) / DUMP_TILE);
    if (!c->d_dump_total) {
        TRY(dmalloc(&c->d_dump_total, 1));
        CU(cudaMallocHost((void **)&c->h_dump_total, sizeof(unsigned long long)));
    }
    if ((size_t)n_tiles > c->dump_tiles_cap) {
        cudaFree(c->d_dump_count); cudaFree(c->d_dump_base);
        c->d_dump_count = nullptr; c->d_dump_base = nullptr; c->dump_tiles_cap = 0;
        TRY(dmalloc(&c->d_dump_count, (size_t)n_tiles));
        TRY(dmalloc(&c->d_dump_base, (size_t)n_tiles));
        c->dump_tiles_cap = (size_t)n_tiles;
    }
    unsigned long long n_lat = 0;
    if (n_tiles > 0) {
        k_dump_count<<<n_tiles, DUMP_TILE, 0, c->stream>>>(g, r, c->s.type, c->d_dump_count);
        k_dump_scan<<<1, 1024, 0, c->stream>>>(c->d_dump_count, c->d_dump_base, n_tiles, c->d_dump_total);
        c->launches += 2;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(c->h_dump_total, c->d_dump_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        n_lat = *c->h_dump_total;
    }
    *n_records = n_inter + (size_t)n_lat;
    if (!records) return 0; // size query
    REQ(cap_records >= *n_records, MISA_B200_EINVAL, "misa_b200_dump_records: record buffer too small");
    DumpRecord *out = static_cast<DumpRecord *>(records);
    for (size_t i = 0; i < n_inter; i++) { // inter atoms first, list order (atom_dump.cpp:59-61)
        const HostAtom &a = IH(c)->local[i];
        DumpRecord rec;
        memset(&rec, 0, sizeof rec);
        rec.id = a.id; rec.step = time_step; rec.type = a.type; rec.inter_type = 0;
        for (int d = 0; d < 3; d++) { rec.x[d] = a.x[d]; rec.v[d] = a.v[d]; }
        out[i] = rec;
    }
    if (n_lat == 0) return 0;
    if ((size_t)n_lat > c->dump_cap) {
        cudaFree(c->d_dump);
        c->d_dump = nullptr; c->dump_cap = 0;
        TRY(dmalloc(&c->d_dump, (size_t)r.n * DUMP_WORDS)); // the whole region: later dumps never reallocate
        c->dump_cap = (size_t)r.n;
    }
    {
        Slot sl(c, MISA_B200_K_XFER);
        k_dump_write<<<n_tiles, DUMP_TILE, 0, c->stream>>>(g, r, c->s, c->d_dump_base, (unsigned long long)time_step, c->d_dump);
        c->launches++;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out + n_inter, c->d_dump, (size_t)n_lat * sizeof(DumpRecord), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// -------------------------------------------------------------------------------------------------
// profiling
// -------------------------------------------------------------------------------------------------
extern "C" int misa_b200_profile_enable(misa_b200_ctx *c, int on) {
    REQ(c, MISA_B200_EINVAL, "null ctx");
    c->prof_on = on != 0;
    return 0;
}
extern "C" int misa_b200_profile_read(misa_b200_ctx *c, double ms_sum[MISA_B200_K_COUNT], int64_t launches[MISA_B200_K_COUNT]) {
    REQ(c && ms_sum && launches, MISA_B200_EINVAL, "null argument");
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < MISA_B200_K_COUNT; k++) {
        auto &ev = c->prof_ev[k];
        for (size_t i = 0; i + 1 < ev.size(); i += 2) {
            float f = 0;
            cudaEventElapsedTime(&f, ev[i], ev[i + 1]);
            c->prof_ms[k] += f;
            c->prof_n[k]++;
        }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        ms_sum[k] = c->prof_ms[k];
        launches[k] = c->prof_n[k];
        c->prof_ms[k] = 0;
        c->prof_n[k] = 0;
    }
    return 0;
}
extern "C" int misa_b200_launch_count(misa_b200_ctx *c, int64_t *n) {
    REQ(c && n, MISA_B200_EINVAL, "null argument");
    *n = c->launches;
    return 0;
}
