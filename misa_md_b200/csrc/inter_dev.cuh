// misa_md_b200/csrc/inter_dev.cuh -- the off-lattice ("inter") atom path RESIDENT ON THE DEVICE (round 2).
//
// inter.cuh keeps the reference's list (std::list + unordered_multimap, src/atom/inter_atom_list.h:21-40) on the host and pays
// for it with dozens of stream synchronisations per step; with ~200 inter atoms a 2 M-atom cascade ran 3x slower than the
// thermal step. Here the list is a structure of arrays in HBM whose ORDER is the reference's list order -- that order decides
// which of two interstitials takes a vacancy (atom::decide part 2, src/atom.cpp:58-82), the order of the dump / download
// records and of every packer message -- and every list operation of the step is a kernel:
//
//   Wigner-Seitz mapping    ws::voronoy / isOutBox / getNearLatCoord / findNearLatIndexInSubBox (src/lattice/ws_utils.cpp:15-163)
//                           as __device__ functions, the reference's operation order with explicit round-to-nearest ops
//   NewtonMotion inter loops (src/newton_motion.cpp:46-54,68-73), clearForce inter loop (src/atom.cpp:94-99)
//   atom::decide            part 1: run-away sites (flagged by k_verlet1) sorted into the reference's k,j,i loop order, appended,
//                           vacated; part 2: every inter atom claims its Wigner-Seitz site with atomicMin(list index) -- the
//                           first in list order wins, exactly the sequential loop's outcome -- winners re-occupy, the list is
//                           compacted stably
//   InterParticlePacker     exchangeInter (src/pack/inter_particle_packer.cpp:17-121): per dimension a stable three-way
//                           partition [stay | leave low | leave high]; leavers are appended behind the stayers with the image
//                           shift (own sub-box on that side) or packed into a fixed-capacity NCCL message (another sub-box)
//   InterBorderPacker       borderInter (src/pack/inter_border_packer.cpp:12-106): local, then already received ghost copies
//                           inside the send region -> ghost list, send references kept for the df exchange
//   DfEmbedPacker inter part (src/pack/df_embed_packer.cpp:38-43,60-66)
//   InterAtomList::makeIndex + interRho / interForce: inter.cuh's k_inter_link / k_inter_pairs on these arrays
//
// The host learns two things per step, through mapped memory: the activity word after k_verlet1 (as before) and the list
// sizes after the border exchange. Layout: local atoms [0, n_local), ghost copies [cap/2, cap/2 + n_ghost).
// Single-block kernels do the order-preserving parts (sort, scans): the list holds 1e2 .. 1e4 atoms.
#pragma once
#include "inter.cuh"

#define IDEV_THREADS 1024
// counters (device ints, mirrored to the host by k_idev_publish)
enum { IC_NL = 0, IC_NG = 1, IC_OVERFLOW = 2, IC_MOVED = 3, IC_NSEND = 4 /* [6] */, IC_RSTART = 10 /* [6] */, IC_RN = 16 /* [6] */, IC_MSG = 22 /* [4] */, IC_COUNT = 32 };

struct IdevBuf {
    InterSoa a{}, t{};            // the list and a scratch copy for stable partitions
    int3 *cell = nullptr;         // doubled-x lattice coordinate of the Wigner-Seitz site (ghost-extended array)
    int *cls = nullptr;           // per local atom: partition class of the operation in flight
    int *pos = nullptr;           // per atom: destination index of the partition in flight
    int *send_ref[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // borderInter send lists: >= 0 local index, < 0 ~ghost index
    int *ic = nullptr, *h_ic = nullptr, *hd_ic = nullptr;
    int *sort_keys = nullptr;
    double *msg[4] = {nullptr, nullptr, nullptr, nullptr};   // NCCL messages: send low / high, recv from high / low neighbour
    int cap = 0;
    bool host_stale = false;      // the host copy (InterHost::local) is older than the device list
};
static std::map<misa_b200_ctx *, IdevBuf *> g_idev;
static IdevBuf *ID(misa_b200_ctx *c) { return g_idev[c]; }
static const int kIdevMsgCap = 4096;   // records per NCCL message

// ---- Wigner-Seitz helpers on the device ------------------------------------------------------------------------------------
struct WsGeo { double a; long long lo2x, loy, loz; int nx2, ny, nz, gx2, gy, gz, sx2, sy, sz; long long H; };
__host__ __device__ inline WsGeo ws_geo(const Geo &g) {
    WsGeo w;
    w.a = g.a; w.lo2x = 2LL * g.lo[0]; w.loy = g.lo[1]; w.loz = g.lo[2];
    w.nx2 = 2 * g.nx; w.ny = g.ny; w.nz = g.nz; w.gx2 = 2 * g.gx; w.gy = g.gy; w.gz = g.gz; w.sx2 = 2 * g.sxc; w.sy = g.sy; w.sz = g.sz; w.H = g.H;
    return w;
}
// ws::voronoy through the VORONOY macro (src/lattice/ws_utils.cpp:15-89): nearest cube corner by lround, then the plane test
// towards the body centre of the octant; global doubled-x lattice coordinate out
__device__ __forceinline__ void ws_voronoy_d(const double X, const double Y, const double Z, const double LC, long long q[3]) {
    const double qx = __ddiv_rn(X, LC), qy = __ddiv_rn(Y, LC), qz = __ddiv_rn(Z, LC);
    long long cx = (long long)round(qx), cy = (long long)round(qy), cz = (long long)round(qz);   // lround: halves away from zero
    const double dx = __dadd_rn(qx, -(double)cx), dy = __dadd_rn(qy, -(double)cy), dz = __dadd_rn(qz, -(double)cz);
    const bool px = dx > 0, py = dy > 0, pz = dz > 0;
    // normal = (+-1, +-1, +-1) towards the octant's body centre: the products are exact, the sums rounded in the reference's order
    double acc = px ? dx : -dx;
    acc = __dadd_rn(acc, py ? dy : -dy);
    acc = __dadd_rn(acc, pz ? dz : -dz);
    acc = __dadd_rn(acc, -0.75);
    cx = 2 * cx;
    if (acc >= 0.0) { cx += px ? 1 : -1; cy += py ? 0 : -1; cz += pz ? 0 : -1; }
    q[0] = cx; q[1] = cy; q[2] = cz;
}
enum { D_OUT_XL = 1, D_OUT_XB = 2, D_OUT_YL = 4, D_OUT_YB = 8, D_OUT_ZL = 16, D_OUT_ZB = 32 };   // src/lattice/box.h:10-20
__device__ __forceinline__ unsigned ws_is_out_box_d(const WsGeo &w, const double x, const double y, const double z) {   // ws::isOutBox
    long long q[3];
    ws_voronoy_d(x, y, z, w.a, q);
    q[0] -= w.lo2x; q[1] -= w.loy; q[2] -= w.loz;
    unsigned fl = 0;
    if (q[0] < 0) fl |= D_OUT_XL; else if (q[0] >= w.nx2) fl |= D_OUT_XB;
    if (q[1] < 0) fl |= D_OUT_YL; else if (q[1] >= w.ny) fl |= D_OUT_YB;
    if (q[2] < 0) fl |= D_OUT_ZL; else if (q[2] >= w.nz) fl |= D_OUT_ZB;
    return fl;
}
// ws::getNearLatCoord: coordinate in the ghost-extended doubled-x array (may lie outside it)
__device__ __forceinline__ void ws_near_lat_coord_d(const WsGeo &w, const double x, const double y, const double z, long long q[3]) {
    ws_voronoy_d(x, y, z, w.a, q);
    q[0] -= w.lo2x - w.gx2; q[1] -= w.loy - w.gy; q[2] -= w.loz - w.gz;
}
__device__ __forceinline__ int ws_dev_index(const WsGeo &w, const long long q[3]) {   // device index of an in-array coordinate
    const long long idx = (q[2] * w.sy + q[1]) * (long long)w.sx2 + q[0];
    return (int)((idx >> 1) + (idx & 1) * w.H);
}
// ws::findNearLatIndexInSubBox -> device index of the OWNED site, or -1
__device__ __forceinline__ int ws_near_site_in_sub_box_d(const WsGeo &w, const double x, const double y, const double z) {
    long long q[3];
    ws_voronoy_d(x, y, z, w.a, q);
    q[0] -= w.lo2x; q[1] -= w.loy; q[2] -= w.loz;
    if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= w.nx2 || q[1] >= w.ny || q[2] >= w.nz) return -1;
    q[0] += w.gx2; q[1] += w.gy; q[2] += w.gz;
    return ws_dev_index(w, q);
}

// ---- small pieces --------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void isoa_copy(const InterSoa &d, const int j, const InterSoa &s, const int i) {
#pragma unroll
    for (int k = 0; k < 3; k++) { d.x[k][j] = s.x[k][i]; d.v[k][j] = s.v[k][i]; d.f[k][j] = s.f[k][i]; }
    d.rho[j] = s.rho[i]; d.df[j] = s.df[i]; d.type[j] = s.type[i]; d.id[j] = s.id[i];
}
// exclusive prefix sum over the block of one int per thread (IDEV_THREADS threads); returns the thread's offset, total in *total
__device__ __forceinline__ int block_excl_scan(const int v, int *total) {
    __shared__ int wsum[IDEV_THREADS / 32];
    __shared__ int tot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = lane < IDEV_THREADS / 32 ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        if (lane < IDEV_THREADS / 32) wsum[lane] = s;
        if (lane == 31) tot = s;
    }
    __syncthreads();
    const int base = wid ? wsum[wid - 1] : 0;
    *total = tot;
    __syncthreads();
    return base + x - v;
}

// NewtonMotion::firststep / secondstep inter loops; clearForce inter loop; configuration::rescale inter loop
__global__ void k_idev_first_step(const InterSoa a, const int *__restrict__ ic, const double dt, const double c0, const double c1, const double c2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ic[IC_NL]) return;
    const int t = a.type[i];
    const double cm = t == 0 ? c0 : (t == 1 ? c1 : c2);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double v = __dadd_rn(a.v[k][i], __dmul_rn(cm, a.f[k][i]));
        a.v[k][i] = v;
        a.x[k][i] = __dadd_rn(a.x[k][i], __dmul_rn(dt, v));
    }
}
__global__ void k_idev_second_step(const InterSoa a, const int *__restrict__ ic, const double c0, const double c1, const double c2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ic[IC_NL]) return;
    const int t = a.type[i];
    const double cm = t == 0 ? c0 : (t == 1 ? c1 : c2);
#pragma unroll
    for (int k = 0; k < 3; k++) a.v[k][i] = __dadd_rn(a.v[k][i], __dmul_rn(cm, a.f[k][i]));
}
__global__ void k_idev_clear(const InterSoa a, const int *__restrict__ ic) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ic[IC_NL]) return;
    a.f[0][i] = 0.0; a.f[1][i] = 0.0; a.f[2][i] = 0.0; a.rho[i] = 0.0;
}
__global__ void k_idev_scale_v(const InterSoa a, const int *__restrict__ ic, const double fac) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ic[IC_NL]) return;
    a.v[0][i] *= fac; a.v[1][i] *= fac; a.v[2][i] *= fac;
}
__global__ void k_idev_publish(const int *__restrict__ ic, int *__restrict__ h_ic) {
    if (threadIdx.x < IC_COUNT) h_ic[threadIdx.x] = ic[threadIdx.x];
}
__global__ void k_idev_drop_ghosts(int *__restrict__ ic) {
    if (threadIdx.x == 0) ic[IC_NG] = 0;
    if (threadIdx.x < 6) { ic[IC_NSEND + threadIdx.x] = 0; ic[IC_RSTART + threadIdx.x] = 0; ic[IC_RN + threadIdx.x] = 0; }
}

// ---- atom::decide part 1: the run-away sites of k_verlet1 in the reference's k,j,i order -> appended, vacated ---------------
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_append_runaways(const Geo g, const Soa s, const InterSoa a, int *__restrict__ ic, const int *__restrict__ counters, int *__restrict__ sites,
                       int *__restrict__ keys, const int cap_half, const int site_cap) {
    __shared__ int sk[4096], sv[4096];
    const int n = min(counters[0], site_cap), nl = ic[IC_NL];
    if (n <= 0) return;
    if (nl + n > cap_half) { if (threadIdx.x == 0) atomicOr(&ic[IC_OVERFLOW], 1); return; }
    // sort by the reference's linear index (ascending = k, j, i loop order, src/atom.cpp:28-30)
    int m = 1;
    while (m < n) m <<= 1;
    const bool in_smem = m <= 4096;
    int *K = in_smem ? sk : keys, *V = in_smem ? sv : keys + m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int d = i < n ? sites[i] : 0;
        K[i] = i < n ? (int)dev_to_ref(d, g.H) : 0x7fffffff;
        V[i] = d;
    }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    const int a0 = K[i], a1 = K[p];
                    if ((a0 > a1) == up) { K[i] = a1; K[p] = a0; const int t = V[i]; V[i] = V[p]; V[p] = t; }
                }
            }
            __syncthreads();
        }
    // addInterAtom copies the whole element (src/atom.cpp:43); the site becomes a vacancy with v = 0 (:44-47)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int d = V[i], j = nl + i;
#pragma unroll
        for (int k = 0; k < 3; k++) { a.x[k][j] = s.x[k][d]; a.v[k][j] = s.v[k][d]; a.f[k][j] = s.f[k][d]; }
        a.rho[j] = s.rho[d]; a.df[j] = s.df[d]; a.type[j] = s.type[d]; a.id[j] = s.id[d];
        s.type[d] = -1;
        site_set_x(s, d, -1, s.x[0][d], s.x[1][d], s.x[2][d]);   // the position stays in the record (Soa::sx), the stencil no longer sees it
        s.v[0][d] = 0.0; s.v[1][d] = 0.0; s.v[2][d] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) ic[IC_NL] = nl + n;
}

// ---- atom::decide part 2 (src/atom.cpp:58-82) --------------------------------------------------------------------------------
// claim: the first inter atom in list order whose Wigner-Seitz site is a vacancy inside the box takes it
__global__ void k_idev_claim(const Geo g, const Soa s, const InterSoa a, const int *__restrict__ ic, unsigned int *__restrict__ claim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ic[IC_NL]) return;
    const WsGeo w = ws_geo(g);
    const int d = ws_near_site_in_sub_box_d(w, a.x[0][i], a.x[1][i], a.x[2][i]);
    // near_atom != nullptr && near_atom->isInterElement() && ws::isOutBox(*near_atom) == IN_BOX (the vacancy's own stale position)
    const bool cand = d >= 0 && s.type[d] < 0 && ws_is_out_box_d(w, site_x(s, 0, d, -1), site_x(s, 1, d, -1), site_x(s, 2, d, -1)) == 0;
    a.site[i] = cand ? d : -1;
    if (cand) atomicMin(&claim[d], (unsigned)i);
}
__global__ void k_idev_occupy(const Soa s, const InterSoa a, const int *__restrict__ ic, const unsigned int *__restrict__ claim, int *__restrict__ cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ic[IC_NL]) return;
    const int d = a.site[i];
    const bool win = d >= 0 && claim[d] == (unsigned)i;
    cls[i] = win ? 1 : 0;
    if (win) {
        s.id[d] = a.id[i]; s.type[d] = a.type[i];
#pragma unroll
        for (int k = 0; k < 3; k++) { s.x[k][d] = a.x[k][i]; s.v[k][d] = a.v[k][i]; }
    }
}
// stable removal of the atoms with cls != 0 (the claims are reset on the way); result back in `a`
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_remove(const InterSoa a, const InterSoa t, int *__restrict__ ic, const int *__restrict__ cls, unsigned int *__restrict__ claim) {
    const int nl = ic[IC_NL];
    int kept = 0;
    bool any = false;
    for (int base = 0; base < nl; base += IDEV_THREADS) {
        const int i = base + threadIdx.x;
        const int keep = i < nl && cls[i] == 0;
        if (i < nl && a.site[i] >= 0) claim[a.site[i]] = 0xffffffffu;
        int total;
        const int o = block_excl_scan(keep, &total);
        if (keep) isoa_copy(t, kept + o, a, i);
        any = any || total != min(IDEV_THREADS, nl - base);
        kept += total;
    }
    __syncthreads();
    if (kept == nl) return;                 // nothing removed: `a` is untouched
    for (int i = threadIdx.x; i < kept; i += IDEV_THREADS) isoa_copy(a, i, t, i);
    if (threadIdx.x == 0) ic[IC_NL] = kept;
    (void)any;
}

// ---- exchangeInter, one dimension (src/pack/inter_particle_packer.cpp:59-121) -------------------------------------------------
// Stable three-way partition of the local list: stayers, then the atoms that left through the low face, then the high face.
// self: both neighbours in this dimension are this sub-box -- the leavers come straight back (image-shifted) behind the stayers,
// low-face message first, as the staged exchange delivers them. Otherwise they are packed into msg_lo / msg_hi (record: id, type,
// x + shift, v; word 0 of a message = its record count) and the list is cut to the stayers; k_idev_unpack_particles appends what
// the neighbours sent.
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_exchange_dim(const Geo g, const InterSoa a, const InterSoa t, int *__restrict__ ic, const int dim, const int self, const double shift_lo,
                    const double shift_hi, double *__restrict__ msg_lo, double *__restrict__ msg_hi, const int msg_cap) {
    const WsGeo w = ws_geo(g);
    const int nl = ic[IC_NL];
    const unsigned f_lo = 1u << (2 * dim), f_hi = 2u << (2 * dim);
    int n_stay = 0, n_lo = 0, n_hi = 0;
    __shared__ int s_cnt[3];
    // pass 1: classes and class counts
    int c_stay = 0, c_lo = 0, c_hi = 0;
    for (int i = threadIdx.x; i < nl; i += IDEV_THREADS) {
        const unsigned fl = ws_is_out_box_d(w, a.x[0][i], a.x[1][i], a.x[2][i]);
        // the low-face test runs first and removes its atoms before the high-face test sees the list (packer called per direction)
        const int cl = (fl & f_lo) ? 1 : ((fl & f_hi) ? 2 : 0);
        a.site[i] = cl;
        c_stay += cl == 0; c_lo += cl == 1; c_hi += cl == 2;
    }
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    atomicAdd(&s_cnt[0], c_stay); atomicAdd(&s_cnt[1], c_lo); atomicAdd(&s_cnt[2], c_hi);
    __syncthreads();
    n_stay = s_cnt[0]; n_lo = s_cnt[1]; n_hi = s_cnt[2];
    if (!self && threadIdx.x == 0) {
        if (n_lo > msg_cap || n_hi > msg_cap) atomicOr(&ic[IC_OVERFLOW], 2);
        msg_lo[0] = (double)min(n_lo, msg_cap); msg_hi[0] = (double)min(n_hi, msg_cap);
    }
    if (n_lo + n_hi == 0) return;
    // pass 2: stable placement, chunk by chunk
    int o_stay = 0, o_lo = 0, o_hi = 0;
    for (int base = 0; base < nl; base += IDEV_THREADS) {
        const int i = base + threadIdx.x;
        const int cl = i < nl ? a.site[i] : -1;
        int tot0, tot1, tot2;
        const int p0 = block_excl_scan(cl == 0, &tot0);
        const int p1 = block_excl_scan(cl == 1, &tot1);
        const int p2 = block_excl_scan(cl == 2, &tot2);
        if (cl == 0) isoa_copy(t, o_stay + p0, a, i);
        else if (cl > 0) {
            const double sh = cl == 1 ? shift_lo : shift_hi;
            const int k = cl == 1 ? o_lo + p1 : o_hi + p2;
            if (self) {
                const int j = n_stay + (cl == 1 ? 0 : n_lo) + k;
                isoa_copy(t, j, a, i);
                t.x[dim][j] = __dadd_rn(a.x[dim][i], sh);
                // the packer carries id, type, r, v only (particledata, 64 B): the rest arrives as zeros
                t.f[0][j] = 0.0; t.f[1][j] = 0.0; t.f[2][j] = 0.0; t.rho[j] = 0.0; t.df[j] = 0.0;
            } else if (k < msg_cap) {
                double *r = (cl == 1 ? msg_lo : msg_hi) + 1 + 8 * (size_t)k;
                r[0] = __longlong_as_double((long long)a.id[i]);
                r[1] = (double)a.type[i];
                for (int q = 0; q < 3; q++) { r[2 + q] = q == dim ? __dadd_rn(a.x[q][i], sh) : a.x[q][i]; r[5 + q] = a.v[q][i]; }
            }
        }
        o_stay += tot0; o_lo += tot1; o_hi += tot2;
    }
    __syncthreads();
    const int n_new = self ? nl : n_stay;
    for (int i = threadIdx.x; i < n_new; i += IDEV_THREADS) isoa_copy(a, i, t, i);
    if (threadIdx.x == 0) ic[IC_NL] = n_new;
}
// append the records of two received messages (from the high neighbour's low-face message first -- recv[0] --, then recv[1])
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_unpack_particles(const InterSoa a, int *__restrict__ ic, const double *__restrict__ m0, const double *__restrict__ m1, const int cap_half) {
    const int nl = ic[IC_NL], n0 = (int)m0[0], n1 = (int)m1[0];
    if (nl + n0 + n1 > cap_half) { if (threadIdx.x == 0) atomicOr(&ic[IC_OVERFLOW], 1); return; }
    for (int k = threadIdx.x; k < n0 + n1; k += IDEV_THREADS) {
        const double *r = (k < n0 ? m0 + 1 + 8 * (size_t)k : m1 + 1 + 8 * (size_t)(k - n0));
        const int j = nl + k;
        a.id[j] = (unsigned long long)__double_as_longlong(r[0]);
        a.type[j] = (int8_t)(int)r[1];
        for (int q = 0; q < 3; q++) { a.x[q][j] = r[2 + q]; a.v[q][j] = r[5 + q]; a.f[q][j] = 0.0; }
        a.rho[j] = 0.0; a.df[j] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) ic[IC_NL] = nl + n0 + n1;
}

// ---- borderInter, one dimension (src/pack/inter_border_packer.cpp:12-106) -----------------------------------------------------
// Send lists of both directions: local atoms in list order, then the ghost copies received so far, whose Wigner-Seitz site lies in
// the forwarded slab (comm::fwCommLocalRegion: ghost-wide in `dim`, ghost-inclusive in the dimensions already exchanged). The
// references are kept (send_ref) for the df exchange. self: the copies are appended to the ghost list directly; else packed
// (type, x + shift) into the messages.
struct BorderSlab { int lo[2][3], hi[2][3]; };
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_border_dim(const Geo g, const InterSoa a, int *__restrict__ ic, const int dim, const int self, const BorderSlab sl, const double shift_lo,
                  const double shift_hi, int *__restrict__ ref_lo, int *__restrict__ ref_hi, double *__restrict__ msg_lo, double *__restrict__ msg_hi,
                  const int msg_cap, const int ghost_base, const int cap_half) {
    const WsGeo w = ws_geo(g);
    const int nl = ic[IC_NL], ng = ic[IC_NG];
    int n_sel[2] = {0, 0};
    // candidates in send order: local 0 .. nl-1, then ghost 0 .. ng-1; both directions are selected BEFORE anything is received
    for (int base = 0; base < nl + ng; base += IDEV_THREADS) {
        const int c = base + threadIdx.x;
        int in0 = 0, in1 = 0;
        if (c < nl + ng) {
            const int i = c < nl ? c : ghost_base + (c - nl);
            long long q[3];
            ws_near_lat_coord_d(w, a.x[0][i], a.x[1][i], a.x[2][i], q);
            in0 = q[0] >= sl.lo[0][0] && q[0] < sl.hi[0][0] && q[1] >= sl.lo[0][1] && q[1] < sl.hi[0][1] && q[2] >= sl.lo[0][2] && q[2] < sl.hi[0][2];
            in1 = q[0] >= sl.lo[1][0] && q[0] < sl.hi[1][0] && q[1] >= sl.lo[1][1] && q[1] < sl.hi[1][1] && q[2] >= sl.lo[1][2] && q[2] < sl.hi[1][2];
        }
        int t0, t1;
        const int p0 = block_excl_scan(in0, &t0);
        const int p1 = block_excl_scan(in1, &t1);
        const int ref = c < nl ? c : ~(c - nl);
        if (in0 && n_sel[0] + p0 < cap_half) ref_lo[n_sel[0] + p0] = ref;
        if (in1 && n_sel[1] + p1 < cap_half) ref_hi[n_sel[1] + p1] = ref;
        n_sel[0] += t0; n_sel[1] += t1;
    }
    __syncthreads();
    const int limit = self ? cap_half : msg_cap;
    if (n_sel[0] > limit || n_sel[1] > limit || (self && ng + n_sel[0] + n_sel[1] > cap_half)) { if (threadIdx.x == 0) atomicOr(&ic[IC_OVERFLOW], 4); return; }
    for (int dir = 0; dir < 2; dir++) {
        const int *ref = dir ? ref_hi : ref_lo;
        const double sh = dir ? shift_hi : shift_lo;
        for (int k = threadIdx.x; k < n_sel[dir]; k += IDEV_THREADS) {
            const int r = ref[k], i = r >= 0 ? r : ghost_base + ~r;
            if (self) {
                const int j = ghost_base + ng + (dir ? n_sel[0] : 0) + k;   // recv[0] (= send[0]) first, then recv[1]
                for (int q = 0; q < 3; q++) { a.x[q][j] = q == dim ? __dadd_rn(a.x[q][i], sh) : a.x[q][i]; a.v[q][j] = 0.0; a.f[q][j] = 0.0; }
                a.type[j] = a.type[i]; a.id[j] = 0; a.rho[j] = 0.0; a.df[j] = 0.0;
            } else {
                double *m = (dir ? msg_hi : msg_lo) + 1 + 4 * (size_t)k;
                m[0] = (double)a.type[i];
                for (int q = 0; q < 3; q++) m[1 + q] = q == dim ? __dadd_rn(a.x[q][i], sh) : a.x[q][i];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ic[IC_NSEND + 2 * dim] = n_sel[0]; ic[IC_NSEND + 2 * dim + 1] = n_sel[1];
        if (self) {
            ic[IC_RSTART + 2 * dim] = ng; ic[IC_RN + 2 * dim] = n_sel[0];
            ic[IC_RSTART + 2 * dim + 1] = ng + n_sel[0]; ic[IC_RN + 2 * dim + 1] = n_sel[1];
            ic[IC_NG] = ng + n_sel[0] + n_sel[1];
        } else { msg_lo[0] = (double)n_sel[0]; msg_hi[0] = (double)n_sel[1]; }
    }
}
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_unpack_border(const InterSoa a, int *__restrict__ ic, const int dim, const double *__restrict__ m0, const double *__restrict__ m1, const int ghost_base,
                     const int cap_half) {
    const int ng = ic[IC_NG], n0 = (int)m0[0], n1 = (int)m1[0];
    if (ng + n0 + n1 > cap_half) { if (threadIdx.x == 0) atomicOr(&ic[IC_OVERFLOW], 4); return; }
    for (int k = threadIdx.x; k < n0 + n1; k += IDEV_THREADS) {
        const double *r = (k < n0 ? m0 + 1 + 4 * (size_t)k : m1 + 1 + 4 * (size_t)(k - n0));
        const int j = ghost_base + ng + k;
        a.type[j] = (int8_t)(int)r[0]; a.id[j] = 0;
        for (int q = 0; q < 3; q++) { a.x[q][j] = r[1 + q]; a.v[q][j] = 0.0; a.f[q][j] = 0.0; }
        a.rho[j] = 0.0; a.df[j] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ic[IC_RSTART + 2 * dim] = ng; ic[IC_RN + 2 * dim] = n0;
        ic[IC_RSTART + 2 * dim + 1] = ng + n0; ic[IC_RN + 2 * dim + 1] = n1;
        ic[IC_NG] = ng + n0 + n1;
    }
}
// inter part of DfEmbedPacker, one dimension: df of the send lists -> the ghost copies made from them
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_df_dim(const InterSoa a, const int *__restrict__ ic, const int dim, const int self, const int *__restrict__ ref_lo, const int *__restrict__ ref_hi,
              double *__restrict__ msg_lo, double *__restrict__ msg_hi, const int ghost_base) {
    for (int dir = 0; dir < 2; dir++) {
        const int n = ic[IC_NSEND + 2 * dim + dir];
        const int *ref = dir ? ref_hi : ref_lo;
        const int r0 = ic[IC_RSTART + 2 * dim + dir];
        for (int k = threadIdx.x; k < n; k += IDEV_THREADS) {
            const int r = ref[k], i = r >= 0 ? r : ghost_base + ~r;
            if (self) a.df[ghost_base + r0 + k] = a.df[i];
            else (dir ? msg_hi : msg_lo)[1 + k] = a.df[i];
        }
        if (!self && threadIdx.x == 0) (dir ? msg_hi : msg_lo)[0] = (double)n;
        __syncthreads();   // a later dimension may forward these ghosts
    }
}
__global__ void __launch_bounds__(IDEV_THREADS)
k_idev_df_unpack(const InterSoa a, const int *__restrict__ ic, const int dim, const double *__restrict__ m0, const double *__restrict__ m1, const int ghost_base) {
    for (int dir = 0; dir < 2; dir++) {
        const double *m = dir ? m1 : m0;
        const int n = min((int)m[0], ic[IC_RN + 2 * dim + dir]), r0 = ic[IC_RSTART + 2 * dim + dir];
        for (int k = threadIdx.x; k < n; k += IDEV_THREADS) a.df[ghost_base + r0 + k] = m[1 + k];
    }
}

// InterAtomList::makeIndex (src/atom/inter_atom_list.cpp:27-45): Wigner-Seitz site of every local and ghost inter atom
__global__ void k_idev_index(const Geo g, const InterSoa a, int3 *__restrict__ cell, const int n_local, const int n_total, const int ghost_base) {
    const int w_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (w_ >= n_total) return;
    const int i = w_ < n_local ? w_ : ghost_base + (w_ - n_local);
    const WsGeo w = ws_geo(g);
    long long q[3];
    ws_near_lat_coord_d(w, a.x[0][i], a.x[1][i], a.x[2][i], q);
    const bool inside = q[0] >= 0 && q[0] < w.sx2 && q[1] >= 0 && q[1] < w.sy && q[2] >= 0 && q[2] < w.sz;
    cell[i] = make_int3((int)q[0], (int)q[1], (int)q[2]);
    a.site[i] = inside ? ws_dev_index(w, q) : -1;
}
__global__ void k_idev_link(const int n_local, const int n_total, const int ghost_base, const int *__restrict__ site, int *__restrict__ head, int *__restrict__ next) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_total) return;
    const int i = w < n_local ? w : ghost_base + (w - n_local);
    const int st = site[i];
    next[i] = st >= 0 ? atomicExch(&head[st], i) : -1;
}
__global__ void k_idev_unlink(const int n_local, const int n_total, const int ghost_base, const int *__restrict__ site, int *__restrict__ head) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_total) return;
    const int i = w < n_local ? w : ghost_base + (w - n_local);
    if (site[i] >= 0) head[site[i]] = -1;
}
// list <-> AtomElement records (download / upload / dump / thermo on the host side of the ABI)
__global__ void k_idev_to_records(const InterSoa a, const int n, HostAtom *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HostAtom r;
    r.id = a.id[i]; r.type = a.type[i]; r._pad = 0;
    for (int k = 0; k < 3; k++) { r.x[k] = a.x[k][i]; r.v[k] = a.v[k][i]; r.f[k] = a.f[k][i]; }
    r.rho = a.rho[i]; r.df = a.df[i];
    out[i] = r;
}
__global__ void k_idev_from_records(const InterSoa a, const int n, const HostAtom *__restrict__ in, int *__restrict__ ic) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { ic[IC_NL] = n; ic[IC_NG] = 0; }
    if (i >= n) return;
    a.id[i] = in[i].id; a.type[i] = (int8_t)in[i].type;
    for (int k = 0; k < 3; k++) { a.x[k][i] = in[i].x[k]; a.v[k][i] = in[i].v[k]; a.f[k][i] = in[i].f[k]; }
    a.rho[i] = in[i].rho; a.df[i] = in[i].df;
}

// ============================================ host side =====================================================================
static int idev_alloc(misa_b200_ctx *c, int cap) {
    IdevBuf *b = new IdevBuf();
    g_idev[c] = b;
    b->cap = cap;
    for (InterSoa *q : {&b->a, &b->t}) {
        for (int k = 0; k < 3; k++) { CU(cudaMalloc((void **)&q->x[k], cap * 8)); CU(cudaMalloc((void **)&q->v[k], cap * 8)); CU(cudaMalloc((void **)&q->f[k], cap * 8)); }
        CU(cudaMalloc((void **)&q->rho, cap * 8)); CU(cudaMalloc((void **)&q->df, cap * 8)); CU(cudaMalloc((void **)&q->type, cap));
        CU(cudaMalloc((void **)&q->id, cap * 8)); CU(cudaMalloc((void **)&q->site, cap * 4)); CU(cudaMalloc((void **)&q->next, cap * 4));
    }
    CU(cudaMalloc((void **)&b->cell, cap * sizeof(int3)));
    CU(cudaMalloc((void **)&b->cls, cap * 4)); CU(cudaMalloc((void **)&b->pos, cap * 4));
    CU(cudaMalloc((void **)&b->sort_keys, (size_t)cap * 2 * 2 * 4));
    for (int i = 0; i < 6; i++) CU(cudaMalloc((void **)&b->send_ref[i], cap / 2 * 4));
    CU(cudaMalloc((void **)&b->ic, IC_COUNT * 4));
    CU(cudaMemset(b->ic, 0, IC_COUNT * 4));
    CU(cudaHostAlloc((void **)&b->h_ic, IC_COUNT * 4, cudaHostAllocMapped));
    memset(b->h_ic, 0, IC_COUNT * 4);
    CU(cudaHostGetDevicePointer((void **)&b->hd_ic, b->h_ic, 0));
    return 0;
}
static void idev_free(misa_b200_ctx *c) {
    IdevBuf *b = g_idev[c];
    if (!b) return;
    for (InterSoa *q : {&b->a, &b->t}) {
        for (int k = 0; k < 3; k++) { cudaFree(q->x[k]); cudaFree(q->v[k]); cudaFree(q->f[k]); }
        cudaFree(q->rho); cudaFree(q->df); cudaFree(q->type); cudaFree(q->id); cudaFree(q->site); cudaFree(q->next);
    }
    cudaFree(b->cell); cudaFree(b->cls); cudaFree(b->pos); cudaFree(b->sort_keys); cudaFree(b->ic); cudaFreeHost(b->h_ic);
    for (int i = 0; i < 6; i++) cudaFree(b->send_ref[i]);
    for (int i = 0; i < 4; i++) cudaFree(b->msg[i]);
    delete b;
    g_idev.erase(c);
}
static inline int idev_blocks(int n) { return std::max(1, (n + 255) / 256); }
// list sizes (and the overflow word) to the host: the ONE synchronisation of the off-lattice part of a step
static int idev_publish(misa_b200_ctx *c) {
    IdevBuf *b = ID(c);
    k_idev_publish<<<1, 32, 0, c->stream>>>(b->ic, b->hd_ic);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    REQ((b->h_ic[IC_OVERFLOW] & 1) == 0, MISA_B200_EOVERFLOW, "too many inter atoms");
    REQ((b->h_ic[IC_OVERFLOW] & 2) == 0, MISA_B200_EOVERFLOW, "inter-atom message overflow");
    REQ((b->h_ic[IC_OVERFLOW] & 4) == 0, MISA_B200_EOVERFLOW, "too many ghost inter atoms");
    c->n_inter_local = b->h_ic[IC_NL];
    c->n_inter_ghost = b->h_ic[IC_NG];
    return 0;
}
// the host copy of the list (download, dump, thermo of the ABI): refreshed on demand
static int idev_sync_host(misa_b200_ctx *c) {
    IdevBuf *b = ID(c);
    InterHost *h = IH(c);
    if (!b->host_stale) return 0;
    TRY(idev_publish(c));
    const int n = c->n_inter_local;
    h->local.resize(n);
    if (n > 0) {
        REQ(n <= c->inter_cap, MISA_B200_EOVERFLOW, "inter staging overflow");
        k_idev_to_records<<<idev_blocks(n), 256, 0, c->stream>>>(b->a, n, h->d_rec);
        c->launches++;
        CU(cudaMemcpyAsync(h->pin, h->d_rec, (size_t)n * sizeof(HostAtom), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        memcpy(h->local.data(), h->pin, (size_t)n * sizeof(HostAtom));
    }
    b->host_stale = false;
    return 0;
}
static int idev_upload(misa_b200_ctx *c, const void *atoms, size_t n) {
    IdevBuf *b = ID(c);
    InterHost *h = IH(c);
    REQ((int)n <= c->inter_cap / 2, MISA_B200_EOVERFLOW, "too many inter atoms");
    h->local.assign((const HostAtom *)atoms, (const HostAtom *)atoms + n);
    h->ghost.clear();
    if (n > 0) {
        memcpy(h->pin, atoms, n * sizeof(HostAtom));
        CU(cudaMemcpyAsync(h->d_rec, h->pin, n * sizeof(HostAtom), cudaMemcpyHostToDevice, c->stream));
    }
    k_idev_from_records<<<idev_blocks((int)n), 256, 0, c->stream>>>(b->a, (int)n, h->d_rec, b->ic);
    k_idev_drop_ghosts<<<1, 32, 0, c->stream>>>(b->ic);
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    c->n_inter_local = (int)n;
    c->n_inter_ghost = 0;
    b->host_stale = false;
    if (c->comm_size == 1) c->inter_active = n > 0;
    return 0;
}
static int idev_first_step(misa_b200_ctx *c, const VerletPar &vp) {
    IdevBuf *b = ID(c);
    k_idev_first_step<<<idev_blocks(c->n_inter_local), 256, 0, c->stream>>>(b->a, b->ic, vp.dt, vp.c[0], vp.c[1], vp.c[2]);
    c->launches++;
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_second_step(misa_b200_ctx *c, const VerletPar &vp) {
    IdevBuf *b = ID(c);
    k_idev_second_step<<<idev_blocks(c->n_inter_local), 256, 0, c->stream>>>(b->a, b->ic, vp.c[0], vp.c[1], vp.c[2]);
    c->launches++;
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_clear(misa_b200_ctx *c) {
    IdevBuf *b = ID(c);
    k_idev_clear<<<idev_blocks(c->n_inter_local), 256, 0, c->stream>>>(b->a, b->ic);
    c->launches++;
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_scale_v(misa_b200_ctx *c, double fac) {
    IdevBuf *b = ID(c);
    k_idev_scale_v<<<idev_blocks(c->n_inter_local), 256, 0, c->stream>>>(b->a, b->ic, fac);
    c->launches++;
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_drop_ghosts(misa_b200_ctx *c) {
    if (c->n_inter_ghost == 0) return 0;
    k_idev_drop_ghosts<<<1, 32, 0, c->stream>>>(ID(c)->ic);
    c->launches++;
    c->n_inter_ghost = 0;
    CU(cudaGetLastError());
    return 0;
}
// atom::decide on the device (the host does not look at anything here)
static int idev_decide(misa_b200_ctx *c) {
    IdevBuf *b = ID(c);
    const int cap_half = b->cap / 2;
    k_idev_drop_ghosts<<<1, 32, 0, c->stream>>>(b->ic);   // inter_atom_list->clearGhost(), src/atom.cpp:22
    k_idev_append_runaways<<<1, IDEV_THREADS, 0, c->stream>>>(c->geo, c->s, b->a, b->ic, c->d_counters, c->d_runaway, b->sort_keys, cap_half, c->inter_cap);
    // part 2 over the whole capacity-bounded list: the list length lives on the device
    const int bound = std::min(cap_half, c->n_inter_local + std::max(c->last_runaways, 0));
    if (bound > 0) {
        k_idev_claim<<<idev_blocks(bound), 256, 0, c->stream>>>(c->geo, c->s, b->a, b->ic, reinterpret_cast<unsigned int *>(c->d_site_head));
        k_idev_occupy<<<idev_blocks(bound), 256, 0, c->stream>>>(c->s, b->a, b->ic, reinterpret_cast<const unsigned int *>(c->d_site_head), b->cls);
        k_idev_remove<<<1, IDEV_THREADS, 0, c->stream>>>(b->a, b->t, b->ic, b->cls, reinterpret_cast<unsigned int *>(c->d_site_head));
        c->launches += 3;
    }
    c->launches += 2;
    c->n_inter_ghost = 0;
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_msgs(misa_b200_ctx *c) {
    IdevBuf *b = ID(c);
    const size_t cap = 1 + (size_t)kIdevMsgCap * 8;
    for (int i = 0; i < 4; i++)
        if (!b->msg[i]) { CU(cudaMalloc((void **)&b->msg[i], cap * sizeof(double))); CU(cudaMemset(b->msg[i], 0, cap * sizeof(double))); }
    return 0;
}
// comm::neiSendReceive for one dimension: low-face message to the lower neighbour, high-face to the higher one;
// msg[2] <- what the HIGHER neighbour sent through its low face (recv[0]), msg[3] <- the lower neighbour's high-face message
static int idev_transport(misa_b200_ctx *c, int dim, size_t count) {
    IdevBuf *b = ID(c);
    REQ(c->nccl_comm, MISA_B200_ESTATE, "inter-atom exchange across sub-boxes needs misa_b200_comm_init");
    NC(g_nccl.GroupStart());
    for (int dir = 0; dir < 2; dir++) {
        NC(g_nccl.Send(b->msg[dir], count, kNcclDouble, c->dom.rank_id_neighbours[dim][dir], c->nccl_comm, c->stream));
        NC(g_nccl.Recv(b->msg[2 + dir], count, kNcclDouble, c->dom.rank_id_neighbours[dim][(dir + 1) % 2], c->nccl_comm, c->stream));
    }
    NC(g_nccl.GroupEnd());
    return 0;
}
static void idev_shifts(const misa_b200_ctx *c, int dim, double &lo, double &hi) {   // src/pack/inter_particle_packer.cpp:74-81
    lo = c->dom.grid_coord[dim] == 0 ? c->dom.meas_global_length[dim] : 0.0;
    hi = c->dom.grid_coord[dim] == c->dom.grid_size[dim] - 1 ? -c->dom.meas_global_length[dim] : 0.0;
}
static int idev_exchange(misa_b200_ctx *c) {   // InterAtomList::exchangeInter
    IdevBuf *b = ID(c);
    TRY(idev_msgs(c));
    for (int dim = 0; dim < 3; dim++) {
        const int self = c->dom.grid_size[dim] == 1;
        double lo, hi;
        idev_shifts(c, dim, lo, hi);
        k_idev_exchange_dim<<<1, IDEV_THREADS, 0, c->stream>>>(c->geo, b->a, b->t, b->ic, dim, self, lo, hi, b->msg[0], b->msg[1], kIdevMsgCap);
        c->launches++;
        if (!self) {
            TRY(idev_transport(c, dim, 1 + (size_t)kIdevMsgCap * 8));
            k_idev_unpack_particles<<<1, IDEV_THREADS, 0, c->stream>>>(b->a, b->ic, b->msg[2], b->msg[3], b->cap / 2);
            c->launches++;
        }
    }
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_border(misa_b200_ctx *c) {     // InterAtomList::borderInter
    IdevBuf *b = ID(c);
    const Geo &g = c->geo;
    const int gh[3] = {2 * g.gx, g.gy, g.gz}, bx[3] = {2 * g.nx, g.ny, g.nz}, ex[3] = {2 * g.sxc, g.sy, g.sz};
    TRY(idev_msgs(c));
    k_idev_drop_ghosts<<<1, 32, 0, c->stream>>>(b->ic);
    c->launches++;
    for (int dim = 0; dim < 3; dim++) {
        BorderSlab sl;
        for (int dir = 0; dir < 2; dir++)
            for (int k = 0; k < 3; k++) {   // comm::fwCommLocalRegion
                if (k == dim) { if (dir == 0) { sl.lo[dir][k] = gh[k]; sl.hi[dir][k] = 2 * gh[k]; } else { sl.lo[dir][k] = bx[k]; sl.hi[dir][k] = bx[k] + gh[k]; } }
                else if (k < dim) { sl.lo[dir][k] = 0; sl.hi[dir][k] = ex[k]; }
                else { sl.lo[dir][k] = gh[k]; sl.hi[dir][k] = gh[k] + bx[k]; }
            }
        const int self = c->dom.grid_size[dim] == 1;
        double lo, hi;
        idev_shifts(c, dim, lo, hi);
        k_idev_border_dim<<<1, IDEV_THREADS, 0, c->stream>>>(g, b->a, b->ic, dim, self, sl, lo, hi, b->send_ref[2 * dim], b->send_ref[2 * dim + 1], b->msg[0],
                                                             b->msg[1], kIdevMsgCap, b->cap / 2, b->cap / 2);
        c->launches++;
        if (!self) {
            TRY(idev_transport(c, dim, 1 + (size_t)kIdevMsgCap * 4));
            k_idev_unpack_border<<<1, IDEV_THREADS, 0, c->stream>>>(b->a, b->ic, dim, b->msg[2], b->msg[3], b->cap / 2, b->cap / 2);
            c->launches++;
        }
    }
    CU(cudaGetLastError());
    return 0;
}
static int idev_halo_df(misa_b200_ctx *c) {    // inter part of DfEmbedPacker
    IdevBuf *b = ID(c);
    for (int dim = 0; dim < 3; dim++) {
        const int self = c->dom.grid_size[dim] == 1;
        k_idev_df_dim<<<1, IDEV_THREADS, 0, c->stream>>>(b->a, b->ic, dim, self, b->send_ref[2 * dim], b->send_ref[2 * dim + 1], b->msg[0], b->msg[1], b->cap / 2);
        c->launches++;
        if (!self) {
            TRY(idev_transport(c, dim, 1 + (size_t)kIdevMsgCap));
            k_idev_df_unpack<<<1, IDEV_THREADS, 0, c->stream>>>(b->a, b->ic, dim, b->msg[2], b->msg[3], b->cap / 2);
            c->launches++;
        }
    }
    CU(cudaGetLastError());
    return 0;
}
static InterDev idev_view(misa_b200_ctx *c) {
    IdevBuf *b = ID(c);
    InterDev v;
    for (int k = 0; k < 3; k++) { v.x[k] = b->a.x[k]; v.f[k] = b->a.f[k]; }
    v.rho = b->a.rho; v.df = b->a.df; v.type = b->a.type; v.id = b->a.id; v.site = b->a.site; v.next = b->a.next; v.cell = b->cell;
    return v;
}
static int idev_rel(misa_b200_ctx *c);   // inter.cuh: decoded reference offsets
static int idev_run_pairs(misa_b200_ctx *c, bool force) {   // makeIndex + interRho / interForce
    IdevBuf *b = ID(c);
    const int nl = c->n_inter_local, n = nl + c->n_inter_ghost, gb = b->cap / 2;
    if (n == 0) return 0;
    TRY(idev_rel(c));
    const InterDev v = idev_view(c);
    k_idev_index<<<idev_blocks(n), 256, 0, c->stream>>>(c->geo, b->a, b->cell, nl, n, gb);
    k_idev_link<<<idev_blocks(n), 256, 0, c->stream>>>(nl, n, gb, b->a.site, c->d_site_head, b->a.next);
    const int blocks = (n * 32 + 127) / 128;
    if (force) k_inter_pairs<true><<<blocks, 128, 0, c->stream>>>(c->geo, c->s, c->tab, v, nl, n, g_inter_dev[c]->d_rel, c->n_full, c->d_site_head, gb);
    else k_inter_pairs<false><<<blocks, 128, 0, c->stream>>>(c->geo, c->s, c->tab, v, nl, n, g_inter_dev[c]->d_rel, c->n_full, c->d_site_head, gb);
    k_idev_unlink<<<idev_blocks(n), 256, 0, c->stream>>>(nl, n, gb, b->a.site, c->d_site_head);
    c->launches += 4;
    b->host_stale = true;
    CU(cudaGetLastError());
    return 0;
}
static int idev_thermo(misa_b200_ctx *c, double *d_out) {
    IdevBuf *b = ID(c);
    const int nl = c->n_inter_local, n = nl + c->n_inter_ghost, gb = b->cap / 2;
    if (nl == 0) return 0;
    TRY(idev_rel(c));
    const InterDev v = idev_view(c);
    k_idev_index<<<idev_blocks(n), 256, 0, c->stream>>>(c->geo, b->a, b->cell, nl, n, gb);
    k_idev_link<<<idev_blocks(n), 256, 0, c->stream>>>(nl, n, gb, b->a.site, c->d_site_head, b->a.next);
    k_inter_energy<<<(nl * 32 + 127) / 128, 128, 0, c->stream>>>(c->geo, c->s, c->tab, v, nl, g_inter_dev[c]->d_rel, c->n_full, c->d_site_head, 55.845, 63.546, 58.6934,
                                                              b->a.v[0], b->a.v[1], b->a.v[2], d_out);
    k_idev_unlink<<<idev_blocks(n), 256, 0, c->stream>>>(nl, n, gb, b->a.site, c->d_site_head);
    c->launches += 4;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}
