// misa_md_b200/csrc/inter.cuh -- the off-lattice ("inter") atom path: atom::decide part 2, InterAtomList,
// InterParticlePacker / InterBorderPacker, atom::interRho / atom::interForce (reference src/atom.cpp:58-82,
// 194-284,360-473; src/atom/inter_atom_list.cpp; src/pack/inter_particle_packer.cpp; src/pack/inter_border_packer.cpp).
//
// Design: the LIST (a std::list + unordered_multimap in the reference) is rare-event control state and is kept
// on the host with the reference's exact ordering rules; every floating-point pair evaluation and every
// access to the lattice arrays is a CUDA kernel. N_inter is 0 for thermal runs and 1e2..1e4 in a cascade.
//
// Pair ownership differs from the reference in ONE deliberate way (it removes both reverse halos): an inter
// atom I -- local or a ghost copy -- adds its contribution to OWNED lattice sites only, and only local inter
// atoms accumulate their own sums. The reference instead updates ghost lattice sites from the owner of I and
// ships them back with RhoPacker / ForcePacker; the set of evaluated pairs is identical.
#pragma once
#include <algorithm>
#include <map>
#include <vector>
#include "ctx.h"
#include "util.cuh"
#include "kernels.cuh"

struct HostAtom { // reference src/atom/atom_element.h:18-41
    unsigned long long id;
    int type, _pad;
    double x[3], v[3], f[3], rho, df;
};
static_assert(sizeof(HostAtom) == 104, "AtomElement layout");

struct InterHost {
    std::vector<HostAtom> local, ghost;                 // InterAtomList::inter_list / inter_ghost_list
    std::vector<int> intersend[6], interrecv[6];        // >= 0 local index, < 0 => ~ghost index
    HostAtom *pin = nullptr;                            // pinned staging
    size_t pin_cap = 0;
    int *pin_idx = nullptr;
    int *d_idx = nullptr;
    HostAtom *d_rec = nullptr;
    double *d_msg[4] = {nullptr, nullptr, nullptr, nullptr};   // NCCL staging: send[2], recv[2]
    double *pin_msg[4] = {nullptr, nullptr, nullptr, nullptr};
};
static std::map<misa_b200_ctx *, InterHost *> g_inter_host;
static InterHost *IH(misa_b200_ctx *c) { return g_inter_host[c]; }
// device-resident list (inter_dev.cuh, the default: ctx.opt_inter_dev); the functions below keep the host-list form as option
// "inter_dev" 0 -- the two are tested against each other, bit for bit (tests/test_gpu_inter.py)
struct VerletPar;
static int idev_alloc(misa_b200_ctx *c, int cap);
static void idev_free(misa_b200_ctx *c);
static int idev_upload(misa_b200_ctx *c, const void *atoms, size_t n);
static int idev_sync_host(misa_b200_ctx *c);
static int idev_first_step(misa_b200_ctx *c, const VerletPar &vp);
static int idev_second_step(misa_b200_ctx *c, const VerletPar &vp);
static int idev_clear(misa_b200_ctx *c);
static int idev_scale_v(misa_b200_ctx *c, double fac);
static int idev_drop_ghosts(misa_b200_ctx *c);
static int idev_decide(misa_b200_ctx *c);
static int idev_exchange(misa_b200_ctx *c);
static int idev_border(misa_b200_ctx *c);
static int idev_halo_df(misa_b200_ctx *c);
static int idev_run_pairs(misa_b200_ctx *c, bool force);
static int idev_thermo(misa_b200_ctx *c, double *d_out);
static int idev_publish(misa_b200_ctx *c);

// ---- device kernels ---------------------------------------------------------------------------------
// stepinfo[0] = this sub-box's off-lattice activity (run-aways of this step + listed inter atoms); the caller
// reduces it (MAX) over the sub-boxes together with stepinfo[1], the dmax2 bit pattern written by k_verlet1
// The step's counters go to the host through MAPPED pinned memory from inside this one-thread kernel (no copy-engine memcpys on
// the step's stream). Single sub-box: the local words ARE the global ones (stepinfo_g, h_stepinfo); otherwise the all-reduce
// that follows produces stepinfo_g and k_publish_stepinfo hands it to the host.
__global__ void k_activity(int *counters, const int n_listed, unsigned long long *stepinfo, unsigned long long *stepinfo_g, int *h_counters,
                           unsigned long long *h_stepinfo) {
    counters[8] = counters[0] + n_listed;
    stepinfo[0] = (unsigned long long)(counters[0] + n_listed);
    for (int k = 0; k < 10; k++) h_counters[k] = counters[k];
    if (stepinfo_g) {
        stepinfo_g[0] = stepinfo[0]; stepinfo_g[1] = stepinfo[1];
        h_stepinfo[0] = stepinfo[0]; h_stepinfo[1] = stepinfo[1];
    }
}
__global__ void k_publish_stepinfo(const unsigned long long *stepinfo_g, unsigned long long *h_stepinfo) {
    h_stepinfo[0] = stepinfo_g[0]; h_stepinfo[1] = stepinfo_g[1];
}
__global__ void k_gather_sites(const int n, const int *__restrict__ sites, const Soa s, HostAtom *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = sites[i];
    HostAtom a;
    a.id = s.id[d]; a.type = s.type[d]; a._pad = 0;
    for (int k = 0; k < 3; k++) { a.x[k] = site_x(s, k, d, a.type); a.v[k] = s.v[k][d]; a.f[k] = s.f[k][d]; }
    a.rho = s.rho[d]; a.df = s.df[d];
    out[i] = a;
}
// atom::decide part 1 tail: site.type = INVALID, v = 0 (reference src/atom.cpp:44-47)
__global__ void k_vacate(const int n, const int *__restrict__ sites, const Soa s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = sites[i];
    s.type[d] = -1;
    site_set_x(s, d, -1, s.x[0][d], s.x[1][d], s.x[2][d]);   // the position stays in the record (Soa::sx), the stencil no longer sees it
    s.v[0][d] = 0.0; s.v[1][d] = 0.0; s.v[2][d] = 0.0;
}
// atom::decide part 2: vacancy re-occupied by an inter atom (reference src/atom.cpp:69-76)
__global__ void k_occupy(const int n, const int *__restrict__ sites, const HostAtom *__restrict__ rec, const Soa s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = sites[i];
    s.id[d] = rec[i].id; s.type[d] = (int8_t)rec[i].type;
    for (int k = 0; k < 3; k++) { s.x[k][d] = rec[i].x[k]; s.v[k][d] = rec[i].v[k]; }
}
__global__ void k_inter_link(const int n, const int *__restrict__ site, int *__restrict__ head, int *__restrict__ next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int st = site[i];
    next[i] = st >= 0 ? atomicExch(&head[st], i) : -1;
}
__global__ void k_inter_unlink(const int n, const int *__restrict__ site, int *__restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && site[i] >= 0) head[site[i]] = -1;
}

struct InterDev { // compact mirror: entries [0,n_local) local, [n_local, n_local+n_ghost) ghost copies
    double *x[3], *f[3], *rho, *df;
    int8_t *type;
    unsigned long long *id;
    int *site, *next;
    int3 *cell; // (x2, y, z) doubled-x lattice coordinate of the site in the ghost-extended array
};

// is the ghost-extended doubled-x coordinate inside the array / inside the owned box
__device__ __forceinline__ bool in_ext(const Geo &g, int x2, int y, int z) {
    return x2 >= 0 && x2 < 2 * g.sxc && y >= 0 && y < g.sy && z >= 0 && z < g.sz;
}
__device__ __forceinline__ bool in_owned(const Geo &g, int x2, int y, int z) {
    return x2 >= 2 * g.gx && x2 < 2 * (g.gx + g.nx) && y >= g.gy && y < g.gy + g.ny && z >= g.gz && z < g.gz + g.nz;
}

// atom::interRho (reference src/atom.cpp:194-284). One WARP per inter atom; lanes stride the 1+n_full sites.
// ref_off: the reference-space full offset lists [2][n_full] decoded to (dx2, dy, dz).
template <bool FORCE>
__global__ void __launch_bounds__(128)
k_inter_pairs(const Geo g, const Soa s, const DevTables tb, const InterDev in, const int n_local, const int n_total,
              const int3 *__restrict__ rel, const int n_full, const int *__restrict__ head, const int ghost_base = -1) {
    const int w_ = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w_ >= n_total) return;
    // list entry of work item w_: ghost copies follow the local atoms directly (host mirror) or start at ghost_base (inter_dev.cuh)
    const int w = (w_ < n_local || ghost_base < 0) ? w_ : ghost_base + (w_ - n_local);
    const int3 cs = in.cell[w];
    const bool local = w_ < n_local;
    const int ti = in.type[w];
    const double xi = in.x[0][w], yi = in.x[1][w], zi = in.x[2][w];
    const double dfi = FORCE ? in.df[w] : 0.0;
    const unsigned long long idi = in.id[w];
    const int par = cs.x & 1;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int q = lane; q <= n_full; q += 32) {
        int x2 = cs.x, y = cs.y, z = cs.z;
        if (q < n_full) { const int3 r = rel[par * n_full + q]; x2 += r.x; y += r.y; z += r.z; }
        if (!in_ext(g, x2, y, z)) continue;
        const long long idx = ((long long)z * g.sy + y) * (2LL * g.sxc) + x2;
        const int d = (int)ref_to_dev(idx, g.H);
        // (a) lattice atom on that site
        const int tl = s.type[d];
        if (tl >= 0) {
            const double dx = xi - s.x[0][d], dy = yi - s.x[1][d], dz = zi - s.x[2][d];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < g.rc2) {
                const bool own = in_owned(g, x2, y, z);
                if (!FORCE) {
                    if (local) a0 += charge_density(tb, tl, d2);
                    if (own) atomicAdd(&s.rho[d], charge_density(tb, ti, d2));
                } else {
                    const double fp = to_force(tb, ti, tl, d2, dfi, s.df[d]);
                    if (local) { a0 += dx * fp; a1 += dy * fp; a2 += dz * fp; }
                    if (own) { atomicAdd(&s.f[0][d], -dx * fp); atomicAdd(&s.f[1][d], -dy * fp); atomicAdd(&s.f[2][d], -dz * fp); }
                }
            }
        }
        // (b) inter atoms bucketed on that site (one-sided, reference src/atom.cpp:244-279,428-471)
        if (local) {
            for (int j = head[d]; j >= 0; j = in.next[j]) {
                if (q == n_full && in.id[j] == idi) continue; // same bucket: not itself
                const double dx = xi - in.x[0][j], dy = yi - in.x[1][j], dz = zi - in.x[2][j];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < g.rc2) {
                    if (!FORCE) a0 += charge_density(tb, in.type[j], d2);
                    else {
                        const double fp = to_force(tb, ti, in.type[j], d2, dfi, in.df[j]);
                        a0 += dx * fp; a1 += dy * fp; a2 += dz * fp;
                    }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, o);
        a1 += __shfl_down_sync(0xffffffffu, a1, o);
        a2 += __shfl_down_sync(0xffffffffu, a2, o);
    }
    if (lane == 0 && local) {
        if (!FORCE) {
            const double rho = in.rho[w] + a0;
            in.rho[w] = rho;
            in.df[w] = d_embed(tb, ti, rho); // reference src/atom.cpp:281-282
        } else {
            in.f[0][w] += a0; in.f[1][w] += a1; in.f[2][w] += a2;
        }
    }
}

// E_pot share of the inter atoms (ours, not in the reference): F(rho_I) + pairs with lattice atoms (counted
// once, from the local inter atom) + half of the inter-inter pairs.
__global__ void __launch_bounds__(128)
k_inter_energy(const Geo g, const Soa s, const DevTables tb, const InterDev in, const int n_local, const int3 *__restrict__ rel,
               const int n_full, const int *__restrict__ head, const double m0, const double m1, const double m2,
               const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz, double *__restrict__ out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_local) return;
    const int3 cs = in.cell[w];
    const int ti = in.type[w];
    const double xi = in.x[0][w], yi = in.x[1][w], zi = in.x[2][w];
    const int par = cs.x & 1;
    double e = 0.0;
    for (int q = lane; q <= n_full; q += 32) {
        int x2 = cs.x, y = cs.y, z = cs.z;
        if (q < n_full) { const int3 r = rel[par * n_full + q]; x2 += r.x; y += r.y; z += r.z; }
        if (!in_ext(g, x2, y, z)) continue;
        const long long idx = ((long long)z * g.sy + y) * (2LL * g.sxc) + x2;
        const int d = (int)ref_to_dev(idx, g.H);
        const int tl = s.type[d];
        if (tl >= 0) {
            const double dx = xi - s.x[0][d], dy = yi - s.x[1][d], dz = zi - s.x[2][d];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < g.rc2) e += pair_energy(tb, ti, tl, d2);
        }
        for (int j = head[d]; j >= 0; j = in.next[j]) {
            if (q == n_full && in.id[j] == in.id[w]) continue;
            const double dx = xi - in.x[0][j], dy = yi - in.x[1][j], dz = zi - in.x[2][j];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < g.rc2) e += 0.5 * pair_energy(tb, ti, in.type[j], d2);
        }
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
    if (lane == 0) {
        const double m = ti == 0 ? m0 : (ti == 1 ? m1 : m2);
        atomicAdd(&out[0], (vx[w] * vx[w] + vy[w] * vy[w] + vz[w] * vz[w]) * m);
        atomicAdd(&out[1], e + embed_energy(tb, ti, in.rho[w]));
    }
}

// ---- host mirror of the Wigner-Seitz helpers (reference src/lattice/ws_utils.cpp:15-163) -------------
static void ws_voronoy(double X, double Y, double Z, double LC, long long out[3]) {
    static const long long offset[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 0, -1}, {1, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {-1, 0, 0}, {1, 0, 0}};
    static const double normal[8][3] = {{-1, -1, -1}, {1, -1, -1}, {-1, 1, -1}, {1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {-1, 1, 1}, {1, 1, 1}};
    long long cx = (long long)lround(X / LC), cy = (long long)lround(Y / LC), cz = (long long)lround(Z / LC);
    const volatile double qx = X / LC, qy = Y / LC, qz = Z / LC; // volatile: keep the reference's rounding, no contraction
    const double dx = qx - cx, dy = qy - cy, dz = qz - cz;
    const unsigned fl = (dz > 0 ? 4u : 0u) | (dy > 0 ? 2u : 0u) | (dx > 0 ? 1u : 0u);
    cx = 2 * cx;
    volatile double acc = normal[fl][0] * dx;
    acc = acc + normal[fl][1] * dy;
    acc = acc + normal[fl][2] * dz;
    acc = acc + (-3.0 / 4.0);
    if (acc >= 0.0) { cx += offset[fl][0]; cy += offset[fl][1]; cz += offset[fl][2]; }
    out[0] = cx; out[1] = cy; out[2] = cz;
}
enum { OUT_XL = 1, OUT_XB = 2, OUT_YL = 4, OUT_YB = 8, OUT_ZL = 16, OUT_ZB = 32 }; // reference src/lattice/box.h:10-20
static unsigned ws_is_out_box(const misa_b200_ctx *c, const double x[3]) { // ws::isOutBox
    const Geo &g = c->geo;
    long long q[3];
    ws_voronoy(x[0], x[1], x[2], g.a, q);
    q[0] -= 2LL * g.lo[0]; q[1] -= g.lo[1]; q[2] -= g.lo[2];
    unsigned fl = 0;
    if (q[0] < 0) fl |= OUT_XL; else if (q[0] >= 2LL * g.nx) fl |= OUT_XB;
    if (q[1] < 0) fl |= OUT_YL; else if (q[1] >= g.ny) fl |= OUT_YB;
    if (q[2] < 0) fl |= OUT_ZL; else if (q[2] >= g.nz) fl |= OUT_ZB;
    return fl;
}
// ws::getNearLatCoord: coordinate in the ghost-extended, doubled-x array
static void ws_near_lat_coord(const misa_b200_ctx *c, const double x[3], long long q[3]) {
    const Geo &g = c->geo;
    ws_voronoy(x[0], x[1], x[2], g.a, q);
    q[0] -= 2LL * (g.lo[0] - g.gx); q[1] -= (g.lo[1] - g.gy); q[2] -= (g.lo[2] - g.gz);
}
// ws::findNearLatIndexInSubBox -> reference linear index or -1
static long long ws_near_index_in_sub_box(const misa_b200_ctx *c, const double x[3]) {
    const Geo &g = c->geo;
    long long q[3];
    ws_voronoy(x[0], x[1], x[2], g.a, q);
    long long j = q[0] - 2LL * g.lo[0], k = q[1] - g.lo[1], l = q[2] - g.lo[2];
    if (j < 0 || k < 0 || l < 0 || j >= 2LL * g.nx || k >= g.ny || l >= g.nz) return -1;
    j += 2LL * g.gx; k += g.gy; l += g.gz;
    return (l * g.sy + k) * (2LL * g.sxc) + j;
}

// ---- allocation --------------------------------------------------------------------------------------
static int3 *g_dummy_rel = nullptr;
struct InterDevBuf { InterDev dv; int cap = 0; int3 *d_rel = nullptr; };
static std::map<misa_b200_ctx *, InterDevBuf *> g_inter_dev;

static int inter_alloc(misa_b200_ctx *c, int cap) {
    c->inter_cap = cap;
    InterHost *h = new InterHost();
    g_inter_host[c] = h;
    InterDevBuf *b = new InterDevBuf();
    g_inter_dev[c] = b;
    b->cap = cap;
    CU(cudaMalloc((void **)&c->d_runaway, cap * sizeof(int)));
    CU(cudaMalloc((void **)&c->d_site_head, (size_t)c->geo.n_ext * sizeof(int)));
    CU(cudaMemset(c->d_site_head, 0xff, (size_t)c->geo.n_ext * sizeof(int)));
    CU(cudaMallocHost((void **)&h->pin, (size_t)cap * sizeof(HostAtom)));
    CU(cudaMallocHost((void **)&h->pin_idx, (size_t)cap * 4 * sizeof(int)));
    h->pin_cap = cap;
    CU(cudaMalloc((void **)&h->d_idx, (size_t)cap * 4 * sizeof(int)));
    CU(cudaMalloc((void **)&h->d_rec, (size_t)cap * sizeof(HostAtom)));
    for (int k = 0; k < 3; k++) { CU(cudaMalloc((void **)&b->dv.x[k], cap * 8)); CU(cudaMalloc((void **)&b->dv.f[k], cap * 8)); }
    CU(cudaMalloc((void **)&b->dv.rho, cap * 8)); CU(cudaMalloc((void **)&b->dv.df, cap * 8));
    CU(cudaMalloc((void **)&b->dv.type, cap)); CU(cudaMalloc((void **)&b->dv.id, cap * 8));
    CU(cudaMalloc((void **)&b->dv.site, cap * 4)); CU(cudaMalloc((void **)&b->dv.next, cap * 4));
    CU(cudaMalloc((void **)&b->dv.cell, cap * sizeof(int3)));
    (void)g_dummy_rel;
    return idev_alloc(c, cap);
}
static void inter_free(misa_b200_ctx *c) {
    idev_free(c);
    InterHost *h = g_inter_host[c];
    InterDevBuf *b = g_inter_dev[c];
    if (h) {
        cudaFreeHost(h->pin); cudaFreeHost(h->pin_idx); cudaFree(h->d_idx); cudaFree(h->d_rec);
        for (int i = 0; i < 4; i++) { cudaFree(h->d_msg[i]); cudaFreeHost(h->pin_msg[i]); }
        delete h;
    }
    if (b) {
        for (int k = 0; k < 3; k++) { cudaFree(b->dv.x[k]); cudaFree(b->dv.f[k]); }
        cudaFree(b->dv.rho); cudaFree(b->dv.df); cudaFree(b->dv.type); cudaFree(b->dv.id); cudaFree(b->dv.site);
        cudaFree(b->dv.next); cudaFree(b->dv.cell); cudaFree(b->d_rel);
        delete b;
    }
    g_inter_host.erase(c); g_inter_dev.erase(c);
    cudaFree(c->d_runaway); cudaFree(c->d_site_head);
}

static int inter_upload(misa_b200_ctx *c, const void *atoms, size_t n) {
    if (c->opt_inter_dev) return idev_upload(c, atoms, n);
    InterHost *h = IH(c);
    REQ((int)n <= c->inter_cap / 2, MISA_B200_EOVERFLOW, "too many inter atoms");
    h->local.assign((const HostAtom *)atoms, (const HostAtom *)atoms + n);
    h->ghost.clear();
    c->n_inter_local = (int)n;
    c->n_inter_ghost = 0;
    if (c->comm_size == 1) c->inter_active = n > 0;
    return 0;
}
static int inter_download(misa_b200_ctx *c, void *atoms, size_t cap, size_t *n) {
    InterHost *h = IH(c);
    if (c->opt_inter_dev) TRY(idev_sync_host(c));
    *n = h->local.size();
    if (atoms) memcpy(atoms, h->local.data(), std::min(cap, *n) * sizeof(HostAtom));
    return 0;
}

// ---- integrator on the list: NewtonMotion::firststep / secondstep inter loops
//      (reference src/newton_motion.cpp:46-54,68-73); host arithmetic, volatile to forbid contraction ----
static int inter_first_step(misa_b200_ctx *c, const VerletPar &vp) {
    if (c->opt_inter_dev) return idev_first_step(c, vp);
    for (HostAtom &a : IH(c)->local)
        for (int d = 0; d < 3; d++) {
            volatile double kick = vp.c[a.type] * a.f[d];
            a.v[d] = a.v[d] + kick;
            volatile double drift = vp.dt * a.v[d];
            a.x[d] += drift;
        }
    return 0;
}
static int inter_second_step(misa_b200_ctx *c, const VerletPar &vp) {
    if (c->opt_inter_dev) return idev_second_step(c, vp);
    for (HostAtom &a : IH(c)->local)
        for (int d = 0; d < 3; d++) {
            volatile double kick = vp.c[a.type] * a.f[d];
            a.v[d] += kick;
        }
    return 0;
}
static void inter_drop_ghosts(misa_b200_ctx *c) {
    if (c->opt_inter_dev) { idev_drop_ghosts(c); return; }
    IH(c)->ghost.clear();
    for (int i = 0; i < 6; i++) { IH(c)->intersend[i].clear(); IH(c)->interrecv[i].clear(); }
    c->n_inter_ghost = 0;
}
static int inter_clear(misa_b200_ctx *c) { // atom::clearForce inter loop, reference src/atom.cpp:94-99
    if (c->opt_inter_dev) return idev_clear(c);
    for (HostAtom &a : IH(c)->local) { a.f[0] = a.f[1] = a.f[2] = 0; a.rho = 0; }
    return 0;
}
static int inter_scale_v(misa_b200_ctx *c, double fac) {
    if (c->opt_inter_dev) return idev_scale_v(c, fac);
    for (HostAtom &a : IH(c)->local) { a.v[0] *= fac; a.v[1] *= fac; a.v[2] *= fac; }
    return 0;
}

// ---- atom::decide (reference src/atom.cpp:21-84) -----------------------------------------------------
static int fetch_sites(misa_b200_ctx *c, const std::vector<int> &dev_sites, std::vector<HostAtom> &out) {
    InterHost *h = IH(c);
    const int n = (int)dev_sites.size();
    out.resize(n);
    if (n == 0) return 0;
    REQ(n <= c->inter_cap, MISA_B200_EOVERFLOW, "inter staging overflow");
    memcpy(h->pin_idx, dev_sites.data(), n * sizeof(int));
    CU(cudaMemcpyAsync(h->d_idx, h->pin_idx, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_gather_sites<<<(n + 127) / 128, 128, 0, c->stream>>>(n, h->d_idx, c->s, h->d_rec);
    c->launches++;
    CU(cudaMemcpyAsync(h->pin, h->d_rec, n * sizeof(HostAtom), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(out.data(), h->pin, n * sizeof(HostAtom));
    return 0;
}

static int inter_decide(misa_b200_ctx *c, int n_runaway) {
    if (c->opt_inter_dev) return idev_decide(c);
    InterHost *h = IH(c);
    const Geo &g = c->geo;
    h->ghost.clear(); // inter_atom_list->clearGhost(), reference src/atom.cpp:22
    c->n_inter_ghost = 0;
    // part 1: run-aways flagged by k_verlet1, in the reference's k,j,i loop order (= ascending reference index)
    if (n_runaway > 0) {
        REQ(n_runaway <= c->inter_cap, MISA_B200_EOVERFLOW, "run-away list overflow");
        std::vector<int> sites(n_runaway);
        CU(cudaMemcpyAsync(h->pin_idx, c->d_runaway, n_runaway * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        std::copy(h->pin_idx, h->pin_idx + n_runaway, sites.begin());
        std::sort(sites.begin(), sites.end(), [&](int a, int b) { return dev_to_ref(a, g.H) < dev_to_ref(b, g.H); });
        std::vector<HostAtom> rec;
        TRY(fetch_sites(c, sites, rec));
        for (const HostAtom &a : rec) h->local.push_back(a); // addInterAtom copies the whole element
        memcpy(h->pin_idx, sites.data(), n_runaway * sizeof(int));
        CU(cudaMemcpyAsync(h->d_idx, h->pin_idx, n_runaway * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        k_vacate<<<(n_runaway + 127) / 128, 128, 0, c->stream>>>(n_runaway, h->d_idx, c->s);
        c->launches++;
        CU(cudaStreamSynchronize(c->stream));
    }
    REQ((int)h->local.size() <= c->inter_cap / 2, MISA_B200_EOVERFLOW, "too many inter atoms");
    // part 2: vacancy-interstitial recombination, in list order
    const size_t n = h->local.size();
    std::vector<long long> near(n);
    std::vector<int> uniq;
    for (size_t i = 0; i < n; i++) {
        near[i] = ws_near_index_in_sub_box(c, h->local[i].x);
        if (near[i] >= 0) uniq.push_back((int)ref_to_dev(near[i], g.H));
    }
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    std::vector<HostAtom> site_rec;
    TRY(fetch_sites(c, uniq, site_rec));
    std::vector<int> occ_sites;
    std::vector<HostAtom> occ_rec, keep;
    keep.reserve(n);
    for (size_t i = 0; i < n; i++) {
        const HostAtom &in = h->local[i];
        bool taken = false;
        if (near[i] >= 0) {
            const int dsite = (int)ref_to_dev(near[i], g.H);
            const size_t k = std::lower_bound(uniq.begin(), uniq.end(), dsite) - uniq.begin();
            HostAtom &site = site_rec[k];
            if (site.type == -1 && ws_is_out_box(c, site.x) == 0) {
                site.id = in.id; site.type = in.type;
                for (int d = 0; d < 3; d++) { site.x[d] = in.x[d]; site.v[d] = in.v[d]; }
                occ_sites.push_back(dsite);
                occ_rec.push_back(site);
                taken = true;
            }
        }
        if (!taken) keep.push_back(in);
    }
    h->local.swap(keep);
    if (!occ_sites.empty()) {
        const int m = (int)occ_sites.size();
        memcpy(h->pin_idx, occ_sites.data(), m * sizeof(int));
        memcpy(h->pin, occ_rec.data(), m * sizeof(HostAtom));
        CU(cudaMemcpyAsync(h->d_idx, h->pin_idx, m * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(h->d_rec, h->pin, m * sizeof(HostAtom), cudaMemcpyHostToDevice, c->stream));
        k_occupy<<<(m + 127) / 128, 128, 0, c->stream>>>(m, h->d_idx, h->d_rec, c->s);
        c->launches++;
        CU(cudaStreamSynchronize(c->stream));
    }
    c->n_inter_local = (int)h->local.size();
    return 0;
}

// ---- staged list exchange: comm::neiSendReceive for InterParticlePacker / InterBorderPacker / the inter
//      part of DfEmbedPacker. Messages between different sub-boxes travel as fixed-capacity NCCL messages. --
static const int kInterMsgCap = 4096; // records per message
static int inter_transport(misa_b200_ctx *c, int dim, std::vector<double> send[2], std::vector<double> recv[2], int width) {
    if (c->dom.grid_size[dim] == 1) { // both neighbours are this sub-box
        recv[0] = send[0];
        recv[1] = send[1];
        return 0;
    }
    REQ(c->nccl_comm, MISA_B200_ESTATE, "inter-atom exchange across sub-boxes needs misa_b200_comm_init");
    InterHost *h = IH(c);
    const size_t cap = 1 + (size_t)kInterMsgCap * 8, count = 1 + (size_t)kInterMsgCap * width;
    for (int i = 0; i < 4; i++)
        if (!h->d_msg[i]) {
            CU(cudaMalloc((void **)&h->d_msg[i], cap * sizeof(double)));
            CU(cudaMallocHost((void **)&h->pin_msg[i], cap * sizeof(double)));
        }
    for (int dir = 0; dir < 2; dir++) {
        REQ(send[dir].size() <= count - 1, MISA_B200_EOVERFLOW, "inter-atom message overflow");
        h->pin_msg[dir][0] = (double)send[dir].size();
        if (!send[dir].empty()) memcpy(h->pin_msg[dir] + 1, send[dir].data(), send[dir].size() * sizeof(double));
        CU(cudaMemcpyAsync(h->d_msg[dir], h->pin_msg[dir], (1 + send[dir].size()) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    NC(g_nccl.GroupStart());
    for (int dir = 0; dir < 2; dir++) {
        NC(g_nccl.Send(h->d_msg[dir], count, kNcclDouble, c->dom.rank_id_neighbours[dim][dir], c->nccl_comm, c->stream));
        NC(g_nccl.Recv(h->d_msg[2 + dir], count, kNcclDouble, c->dom.rank_id_neighbours[dim][(dir + 1) % 2], c->nccl_comm, c->stream));
    }
    NC(g_nccl.GroupEnd());
    for (int dir = 0; dir < 2; dir++)
        CU(cudaMemcpyAsync(h->pin_msg[2 + dir], h->d_msg[2 + dir], count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int dir = 0; dir < 2; dir++) {
        const size_t n = (size_t)h->pin_msg[2 + dir][0];
        REQ(n <= count - 1, MISA_B200_EOVERFLOW, "inter-atom message corrupt");
        recv[dir].assign(h->pin_msg[2 + dir] + 1, h->pin_msg[2 + dir] + 1 + n);
    }
    return 0;
}

static void periodic_shift(const misa_b200_ctx *c, int dim, int dir, double off[3]) { // reference src/pack/inter_particle_packer.cpp:74-81
    off[0] = off[1] = off[2] = 0.0;
    if (c->dom.grid_coord[dim] == 0 && dir == 0) off[dim] = c->dom.meas_global_length[dim];
    if (c->dom.grid_coord[dim] == c->dom.grid_size[dim] - 1 && dir == 1) off[dim] = -c->dom.meas_global_length[dim];
}

// InterAtomList::exchangeInter (reference src/atom/inter_atom_list.cpp:19-25, src/pack/inter_particle_packer.cpp:59-121)
static int inter_exchange(misa_b200_ctx *c) {
    if (c->opt_inter_dev) return idev_exchange(c);
    InterHost *h = IH(c);
    static const unsigned flags[3][2] = {{OUT_XL, OUT_XB}, {OUT_YL, OUT_YB}, {OUT_ZL, OUT_ZB}};
    for (int dim = 0; dim < 3; dim++) {
        std::vector<double> send[2], recv[2];
        for (int dir = 0; dir < 2; dir++) {
            double off[3];
            periodic_shift(c, dim, dir, off);
            std::vector<HostAtom> keep;
            for (const HostAtom &a : h->local) {
                if (ws_is_out_box(c, a.x) & flags[dim][dir]) {
                    // particledata: id, type, r[3], v[3] (64 B) -> 8 doubles
                    double rec[8];
                    memcpy(&rec[0], &a.id, 8);
                    rec[1] = (double)a.type;
                    for (int k = 0; k < 3; k++) { rec[2 + k] = a.x[k] + off[k]; rec[5 + k] = a.v[k]; }
                    send[dir].insert(send[dir].end(), rec, rec + 8);
                } else keep.push_back(a);
            }
            h->local.swap(keep);
        }
        TRY(inter_transport(c, dim, send, recv, 8));
        for (int dir = 0; dir < 2; dir++)
            for (size_t i = 0; i + 8 <= recv[dir].size(); i += 8) {
                HostAtom a;
                memset(&a, 0, sizeof a);
                memcpy(&a.id, &recv[dir][i], 8);
                a.type = (int)recv[dir][i + 1];
                for (int k = 0; k < 3; k++) { a.x[k] = recv[dir][i + 2 + k]; a.v[k] = recv[dir][i + 5 + k]; }
                h->local.push_back(a);
            }
    }
    c->n_inter_local = (int)h->local.size();
    REQ(c->n_inter_local <= c->inter_cap / 2, MISA_B200_EOVERFLOW, "too many inter atoms");
    return 0;
}

// InterAtomList::borderInter (reference src/atom/inter_atom_list.cpp:47-53, src/pack/inter_border_packer.cpp:12-106)
static int inter_border(misa_b200_ctx *c) {
    if (c->opt_inter_dev) { TRY(idev_border(c)); return idev_publish(c); };
    InterHost *h = IH(c);
    const Geo &g = c->geo;
    const int gh[3] = {2 * g.gx, g.gy, g.gz}, bx[3] = {2 * g.nx, g.ny, g.nz}, ex[3] = {2 * g.sxc, g.sy, g.sz};
    for (int i = 0; i < 6; i++) { h->intersend[i].clear(); h->interrecv[i].clear(); }
    for (int dim = 0; dim < 3; dim++) {
        std::vector<double> send[2], recv[2];
        for (int dir = 0; dir < 2; dir++) {
            int lo[3], hi[3]; // comm::fwCommLocalRegion
            for (int k = 0; k < 3; k++) {
                if (k == dim) { if (dir == 0) { lo[k] = gh[k]; hi[k] = 2 * gh[k]; } else { lo[k] = bx[k]; hi[k] = bx[k] + gh[k]; } }
                else if (k < dim) { lo[k] = 0; hi[k] = ex[k]; }
                else { lo[k] = gh[k]; hi[k] = gh[k] + bx[k]; }
            }
            std::vector<int> &sl = h->intersend[2 * dim + dir];
            auto is_in = [&](const double x[3]) {
                long long q[3];
                ws_near_lat_coord(c, x, q);
                return q[0] >= lo[0] && q[0] < hi[0] && q[1] >= lo[1] && q[1] < hi[1] && q[2] >= lo[2] && q[2] < hi[2];
            };
            for (size_t i = 0; i < h->local.size(); i++) if (is_in(h->local[i].x)) sl.push_back((int)i);
            for (size_t i = 0; i < h->ghost.size(); i++) if (is_in(h->ghost[i].x)) sl.push_back(~(int)i);
            double off[3];
            periodic_shift(c, dim, dir, off);
            for (int ref : sl) {
                const HostAtom &a = ref >= 0 ? h->local[ref] : h->ghost[~ref];
                const double rec[4] = {(double)a.type, a.x[0] + off[0], a.x[1] + off[1], a.x[2] + off[2]};
                send[dir].insert(send[dir].end(), rec, rec + 4);
            }
        }
        TRY(inter_transport(c, dim, send, recv, 4));
        for (int dir = 0; dir < 2; dir++)
            for (size_t i = 0; i + 4 <= recv[dir].size(); i += 4) {
                HostAtom e;
                memset(&e, 0, sizeof e);
                e.type = (int)recv[dir][i];
                e.x[0] = recv[dir][i + 1]; e.x[1] = recv[dir][i + 2]; e.x[2] = recv[dir][i + 3];
                h->ghost.push_back(e);
                h->interrecv[2 * dim + dir].push_back(~(int)(h->ghost.size() - 1));
            }
    }
    c->n_inter_ghost = (int)h->ghost.size();
    REQ(c->n_inter_ghost <= c->inter_cap / 2, MISA_B200_EOVERFLOW, "too many ghost inter atoms");
    return 0;
}

// inter part of DfEmbedPacker (reference src/pack/df_embed_packer.cpp:38-43,60-66): df of border inter atoms
static int inter_halo_df(misa_b200_ctx *c) {
    if (c->opt_inter_dev) return idev_halo_df(c);
    InterHost *h = IH(c);
    for (int dim = 0; dim < 3; dim++) {
        std::vector<double> send[2], recv[2];
        for (int dir = 0; dir < 2; dir++)
            for (int ref : h->intersend[2 * dim + dir]) send[dir].push_back(ref >= 0 ? h->local[ref].df : h->ghost[~ref].df);
        TRY(inter_transport(c, dim, send, recv, 1));
        for (int dir = 0; dir < 2; dir++) {
            const std::vector<int> &rl = h->interrecv[2 * dim + dir];
            REQ(recv[dir].size() == rl.size(), MISA_B200_ESTATE, "wrong number of dfembed recv");
            for (size_t i = 0; i < rl.size(); i++) h->ghost[~rl[i]].df = recv[dir][i];
        }
    }
    return 0;
}

// the reference offsets decoded into (dx2, dy, dz), once per context
static int idev_rel(misa_b200_ctx *c) {
    InterDevBuf *b = g_inter_dev[c];
    const Geo &g = c->geo;
    if (b->d_rel) return 0;
    std::vector<int3> rel(2 * (size_t)c->n_full);
    const long long sx = 2LL * g.sxc, sy = g.sy;
    for (int p = 0; p < 2; p++)
        for (int q = 0; q < c->n_full; q++) {
            const long long off = c->ref_off[p][q];
            long long dx = ((off % sx) + sx + sx / 2) % sx - sx / 2;
            long long r = (off - dx) / sx;
            long long dy = ((r % sy) + sy + sy / 2) % sy - sy / 2;
            long long dz = (r - dy) / sy;
            rel[(size_t)p * c->n_full + q] = make_int3((int)dx, (int)dy, (int)dz);
        }
    CU(cudaMalloc((void **)&b->d_rel, rel.size() * sizeof(int3)));
    CU(cudaMemcpy(b->d_rel, rel.data(), rel.size() * sizeof(int3), cudaMemcpyHostToDevice));
    return 0;
}
// ---- device mirror + pair kernels ---------------------------------------------------------------------
static int inter_push_mirror(misa_b200_ctx *c, bool with_df) {
    InterHost *h = IH(c);
    InterDevBuf *b = g_inter_dev[c];
    const Geo &g = c->geo;
    const int nl = (int)h->local.size(), ng = (int)h->ghost.size(), n = nl + ng;
    if (n == 0) return 0;
    REQ(n <= b->cap, MISA_B200_EOVERFLOW, "inter mirror overflow");
    TRY(idev_rel(c));
    if (false) { // (decoded by idev_rel)
        std::vector<int3> rel(2 * (size_t)c->n_full);
        const long long sx = 2LL * g.sxc, sy = g.sy;
        for (int p = 0; p < 2; p++)
            for (int q = 0; q < c->n_full; q++) {
                const long long off = c->ref_off[p][q];
                long long dx = ((off % sx) + sx + sx / 2) % sx - sx / 2;
                long long r = (off - dx) / sx;
                long long dy = ((r % sy) + sy + sy / 2) % sy - sy / 2;
                long long dz = (r - dy) / sy;
                rel[(size_t)p * c->n_full + q] = make_int3((int)dx, (int)dy, (int)dz);
            }
        CU(cudaMalloc((void **)&b->d_rel, rel.size() * sizeof(int3)));
        CU(cudaMemcpy(b->d_rel, rel.data(), rel.size() * sizeof(int3), cudaMemcpyHostToDevice));
    }
    std::vector<double> col(n);
    std::vector<int> site(n);
    std::vector<int3> cell(n);
    std::vector<int8_t> type(n);
    std::vector<unsigned long long> id(n);
    auto at = [&](int i) -> const HostAtom & { return i < nl ? h->local[i] : h->ghost[i - nl]; };
    for (int i = 0; i < n; i++) {
        long long q[3];
        ws_near_lat_coord(c, at(i).x, q); // InterAtomList::makeIndex, reference src/atom/inter_atom_list.cpp:27-45
        const bool inside = q[0] >= 0 && q[0] < 2LL * g.sxc && q[1] >= 0 && q[1] < g.sy && q[2] >= 0 && q[2] < g.sz;
        cell[i] = make_int3((int)q[0], (int)q[1], (int)q[2]);
        site[i] = inside ? (int)ref_to_dev((q[2] * g.sy + q[1]) * (2LL * g.sxc) + q[0], g.H) : -1;
        type[i] = (int8_t)at(i).type;
        id[i] = at(i).id;
    }
    for (int k = 0; k < 3; k++) {
        for (int i = 0; i < n; i++) col[i] = at(i).x[k];
        CU(cudaMemcpyAsync(b->dv.x[k], col.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; i++) col[i] = at(i).f[k];
        CU(cudaMemcpyAsync(b->dv.f[k], col.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    for (int i = 0; i < n; i++) col[i] = at(i).rho;
    CU(cudaMemcpyAsync(b->dv.rho, col.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (with_df) {
        for (int i = 0; i < n; i++) col[i] = at(i).df;
        CU(cudaMemcpyAsync(b->dv.df, col.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    CU(cudaMemcpyAsync(b->dv.type, type.data(), n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b->dv.id, id.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b->dv.site, site.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b->dv.cell, cell.data(), n * sizeof(int3), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

static int inter_make_index(misa_b200_ctx *c) { return 0; } // buckets are (re)built around each pair kernel

static int inter_run_pairs(misa_b200_ctx *c, bool force) {
    if (c->opt_inter_dev) return idev_run_pairs(c, force);
    InterHost *h = IH(c);
    InterDevBuf *b = g_inter_dev[c];
    const int nl = (int)h->local.size(), n = nl + (int)h->ghost.size();
    if (n == 0) return 0;
    TRY(inter_push_mirror(c, force));
    k_inter_link<<<(n + 127) / 128, 128, 0, c->stream>>>(n, b->dv.site, c->d_site_head, b->dv.next);
    const int blocks = (n * 32 + 127) / 128;
    if (force) k_inter_pairs<true><<<blocks, 128, 0, c->stream>>>(c->geo, c->s, c->tab, b->dv, nl, n, b->d_rel, c->n_full, c->d_site_head);
    else k_inter_pairs<false><<<blocks, 128, 0, c->stream>>>(c->geo, c->s, c->tab, b->dv, nl, n, b->d_rel, c->n_full, c->d_site_head);
    k_inter_unlink<<<(n + 127) / 128, 128, 0, c->stream>>>(n, b->dv.site, c->d_site_head);
    c->launches += 3;
    CU(cudaGetLastError());
    if (nl > 0) { // results of the local inter atoms back to the list
        std::vector<double> col(nl);
        if (!force) {
            CU(cudaMemcpyAsync(col.data(), b->dv.rho, nl * 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            for (int i = 0; i < nl; i++) h->local[i].rho = col[i];
            CU(cudaMemcpyAsync(col.data(), b->dv.df, nl * 8, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            for (int i = 0; i < nl; i++) h->local[i].df = col[i];
        } else {
            for (int k = 0; k < 3; k++) {
                CU(cudaMemcpyAsync(col.data(), b->dv.f[k], nl * 8, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
                for (int i = 0; i < nl; i++) h->local[i].f[k] = col[i];
            }
        }
    } else {
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}
static int inter_rho(misa_b200_ctx *c) { return inter_run_pairs(c, false); }
static int inter_force(misa_b200_ctx *c) { return inter_run_pairs(c, true); }

static int inter_thermo(misa_b200_ctx *c, double *d_out) {
    if (c->opt_inter_dev) return idev_thermo(c, d_out);
    InterHost *h = IH(c);
    InterDevBuf *b = g_inter_dev[c];
    const int nl = (int)h->local.size(), n = nl + (int)h->ghost.size();
    if (nl == 0) return 0;
    TRY(inter_push_mirror(c, true));
    // velocities are only needed here: reuse the f columns of the mirror as scratch
    std::vector<double> col(nl);
    double *dv[3] = {b->dv.f[0], b->dv.f[1], b->dv.f[2]};
    for (int k = 0; k < 3; k++) {
        for (int i = 0; i < nl; i++) col[i] = h->local[i].v[k];
        CU(cudaMemcpyAsync(dv[k], col.data(), nl * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    k_inter_link<<<(n + 127) / 128, 128, 0, c->stream>>>(n, b->dv.site, c->d_site_head, b->dv.next);
    k_inter_energy<<<(nl * 32 + 127) / 128, 128, 0, c->stream>>>(c->geo, c->s, c->tab, b->dv, nl, b->d_rel, c->n_full, c->d_site_head,
                                                              55.845, 63.546, 58.6934, dv[0], dv[1], dv[2], d_out);
    k_inter_unlink<<<(n + 127) / 128, 128, 0, c->stream>>>(n, b->dv.site, c->d_site_head);
    c->launches += 3;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
