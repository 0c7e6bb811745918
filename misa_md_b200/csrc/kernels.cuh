// misa_md_b200/csrc/kernels.cuh -- sm_100a kernels of the EAM hot path (fp64, parity-split SoA).
// Each kernel cites the reference CPU loop it replaces. See DESIGN.md section 4 for the per-kernel roofline.
#pragma once
#include "ctx.h"

#define MISA_BLOCK 256
// Position stored for a vacant site (ctx.h Soa::sx): far outside every box (boxes span < 1e4 A) and every cutoff, yet close enough
// that the spline index of such a "pair" (r / dr ~ 1e8) still fits an int -- the branch-free near loop evaluates out-of-range
// lanes along, and an index that overflowed read as "row below the staged range" (the exact, slow recompute of the whole atom:
// measured 1.35x on the stencil kernels of a cascade with 1e30 here)
#define MISA_VACANT_X 1.0e5
// the record's position: for a vacant site the stale one the departed atom left behind
__device__ __forceinline__ double site_x(const Soa &s, const int k, const int d, const int type) { return (type < 0 && s.sx[0]) ? s.sx[k][d] : s.x[k][d]; }
__device__ __forceinline__ void site_set_x(const Soa &s, const int d, const int type, const double x, const double y, const double z) {
    if (type < 0 && s.sx[0]) {
        s.sx[0][d] = x; s.sx[1][d] = y; s.sx[2][d] = z;
        s.x[0][d] = MISA_VACANT_X; s.x[1][d] = MISA_VACANT_X; s.x[2][d] = MISA_VACANT_X;
    } else { s.x[0][d] = x; s.x[1][d] = y; s.x[2][d] = z; }
}

// leave the "vacant sites are invisible" representation: stale positions back into x (option vac_sentinel 0 / sym 1)
__global__ void __launch_bounds__(256) k_unsentinel(const long long n, const Soa s) {
    const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n && s.type[d] < 0) { s.x[0][d] = s.sx[0][d]; s.x[1][d] = s.sx[1][d]; s.x[2][d] = s.sx[2][d]; }
}
// ---- index helpers ------------------------------------------------------------------------------
// owned-cell ordinal c in [0, nx*ny*nz) of sub-lattice p -> device index
__device__ __forceinline__ int owned_cell_to_dev(const Geo &g, int p, long long c, int &cx, int &y, int &z) {
    // device indices are 32-bit throughout (n_ext < 2^31, checked at create): 32-bit divisions, not 64-bit ones
    const unsigned cu = (unsigned)c, nx = (unsigned)g.nx, ny = (unsigned)g.ny;
    const unsigned r = cu / nx;
    cx = (int)(cu - r * nx);
    z = (int)(r / ny);
    y = (int)(r - (unsigned)z * ny);
    return (int)(p * g.H + ((long long)(z + g.gz) * g.sy + (y + g.gy)) * g.sxc + (cx + g.gx));
}
__host__ __device__ __forceinline__ long long ref_to_dev(long long idx, long long H) { return (idx >> 1) + (idx & 1) * H; }
__host__ __device__ __forceinline__ long long dev_to_ref(long long d, long long H) { return d >= H ? 2 * (d - H) + 1 : 2 * d; }

// ---- spline evaluation: libpot InterpolationObject::findSpline + eam::{chargeDensity,dEmbedEnergy,toForce}
//      restated (oracle/pot.c); rows are padded to 8 doubles. ---------------------------------------
__device__ __forceinline__ const double *find_row(const double *__restrict__ tab, int n, double inv_dx, double x, double &p) {
    p = x * inv_dx + 1.0;
    int m = (int)p;
    m = max(1, min(m, n - 1));
    p -= (double)m;
    p = fmin(p, 1.0);
    return tab + (size_t)m * MISA_ROW;
}

__device__ __forceinline__ double charge_density(const DevTables &tb, int tj, double d2) {
    double p;
    const double r = sqrt(d2);
    const double *s = find_row(tb.elec + (size_t)tj * (tb.n_r + 1) * MISA_ROW, tb.n_r, tb.inv_dr, r, p);
    return ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
}

__device__ __forceinline__ double d_embed(const DevTables &tb, int ti, double rho) {
    double p;
    const double *s = find_row(tb.embed + (size_t)ti * (tb.n_rho + 1) * MISA_ROW, tb.n_rho, tb.inv_drho, rho, p);
    return (s[0] * p + s[1]) * p + s[2];
}

__device__ __forceinline__ double embed_energy(const DevTables &tb, int ti, double rho) {
    double p;
    const double *s = find_row(tb.embed + (size_t)ti * (tb.n_rho + 1) * MISA_ROW, tb.n_rho, tb.inv_drho, rho, p);
    return ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
}

__device__ __forceinline__ double to_force(const DevTables &tb, int ti, int tj, double d2, double df_i, double df_j) {
    const double r = sqrt(d2);
    double p;
    const size_t tstride = (size_t)(tb.n_r + 1) * MISA_ROW;
    const double *s = find_row(tb.phi + (size_t)(ti * tb.n_types + tj) * tstride, tb.n_r, tb.inv_dr, r, p);
    const double z2 = ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
    const double z2p = (s[0] * p + s[1]) * p + s[2];
    // elec tables share the r grid with phi: same row index m and fraction p
    const size_t rowoff = (size_t)(s - (tb.phi + (size_t)(ti * tb.n_types + tj) * tstride));
    const double *si = tb.elec + (size_t)ti * tstride + rowoff;
    const double rho_p_from = (si[0] * p + si[1]) * p + si[2];
    double rho_p_to = rho_p_from;
    if (tj != ti) {
        const double *sj = tb.elec + (size_t)tj * tstride + rowoff;
        rho_p_to = (sj[0] * p + sj[1]) * p + sj[2];
    }
    const double recip = 1.0 / r;
    const double phi = z2 * recip;
    const double phip = z2p * recip - phi * recip;
    const double psip = phip + (rho_p_from * df_j + rho_p_to * df_i);
    return -psip * recip;
}

__device__ __forceinline__ double pair_energy(const DevTables &tb, int ti, int tj, double d2) {
    const double r = sqrt(d2);
    double p;
    const size_t tstride = (size_t)(tb.n_r + 1) * MISA_ROW;
    const double *s = find_row(tb.phi + (size_t)(ti * tb.n_types + tj) * tstride, tb.n_r, tb.inv_dr, r, p);
    const double z2 = ((s[3] * p + s[4]) * p + s[5]) * p + s[6];
    return z2 / r;
}

// ---- K8 clear: atom::clearForce (reference src/atom.cpp:86-100) -- all ghost-extended sites -------
__global__ void __launch_bounds__(MISA_BLOCK) k_clear(long long n, double *__restrict__ fx, double *__restrict__ fy,
                                                       double *__restrict__ fz, double *__restrict__ rho) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { fx[i] = 0.0; fy[i] = 0.0; fz[i] = 0.0; rho[i] = 0.0; }
}

// ---- K1 rho (+K2 df fused): atom::latRho / atom::latDf (reference src/atom.cpp:151-192,286-309) ----
// Full-list GATHER: thread = one owned site, loops the parity's offset list, writes the complete sum
// (the reference's half-list scatter + reverse halo, restated; SURVEY.md section 7 "hard part 2").
// ACCUM: add to the existing rho (compat hook semantics, `rho +=`) instead of overwriting.
template <bool FUSE_DF, bool ACCUM>
__global__ void __launch_bounds__(MISA_BLOCK)
k_rho(const Geo g, const Soa s, const DevTables tb, const int *__restrict__ offs, const int n_off, const int blocks_per_parity) {
    extern __shared__ int s_off[];
    const int p = blockIdx.x >= blocks_per_parity;
    const int b = blockIdx.x - p * blocks_per_parity;
    for (int q = threadIdx.x; q < n_off; q += blockDim.x) s_off[q] = offs[p * n_off + q];
    __syncthreads();
    const long long c = (long long)b * blockDim.x + threadIdx.x;
    if (c >= g.n_cells_owned) return;
    int cx, y, z;
    const int d = owned_cell_to_dev(g, p, c, cx, y, z);
    const int ti = s.type[d];
    if (ti < 0) {
        if (!ACCUM) s.rho[d] = 0.0;
        return;
    }
    const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
    double acc = 0.0;
#pragma unroll 4
    for (int q = 0; q < n_off; q++) {
        const int j = d + s_off[q];
        const int tj = s.type[j];
        const double dx = xi - s.x[0][j], dy = yi - s.x[1][j], dz = zi - s.x[2][j];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (tj >= 0 && d2 < g.rc2) acc += charge_density(tb, tj, d2);
    }
    if (ACCUM) acc += s.rho[d];
    s.rho[d] = acc;
    if (FUSE_DF) s.df[d] = d_embed(tb, ti, acc);
}

// ---- K2 df: atom::latDf (reference src/atom.cpp:286-309) -------------------------------------------
__global__ void __launch_bounds__(MISA_BLOCK) k_df(const Geo g, const Soa s, const DevTables tb, const int blocks_per_parity) {
    const int p = blockIdx.x >= blocks_per_parity;
    const int b = blockIdx.x - p * blocks_per_parity;
    const long long c = (long long)b * blockDim.x + threadIdx.x;
    if (c >= g.n_cells_owned) return;
    int cx, y, z;
    const int d = owned_cell_to_dev(g, p, c, cx, y, z);
    const int ti = s.type[d];
    if (ti < 0) return;
    s.df[d] = d_embed(tb, ti, s.rho[d]);
}

// ---- K3 force: atom::latForce (reference src/atom.cpp:311-358) -------------------------------------
// Full-list gather; f_i = sum_j (x_i - x_j) * toForce(t_i, t_j, r^2, df_i, df_j). toForce is symmetric in
// (i,j) so this equals the reference's two-sided half-list update after its force reverse halo.
// FUSE_V2: also apply NewtonMotion::secondstep (v += dt/(2m) f) to the same site (f is complete in-thread).
template <bool ACCUM>
__global__ void __launch_bounds__(MISA_BLOCK)
k_force(const Geo g, const Soa s, const DevTables tb, const int *__restrict__ offs, const int n_off, const int blocks_per_parity) {
    extern __shared__ int s_off[];
    const int p = blockIdx.x >= blocks_per_parity;
    const int b = blockIdx.x - p * blocks_per_parity;
    for (int q = threadIdx.x; q < n_off; q += blockDim.x) s_off[q] = offs[p * n_off + q];
    __syncthreads();
    const long long c = (long long)b * blockDim.x + threadIdx.x;
    if (c >= g.n_cells_owned) return;
    int cx, y, z;
    const int d = owned_cell_to_dev(g, p, c, cx, y, z);
    const int ti = s.type[d];
    if (ti < 0) {
        if (!ACCUM) { s.f[0][d] = 0.0; s.f[1][d] = 0.0; s.f[2][d] = 0.0; }
        return;
    }
    const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
    const double dfi = s.df[d];
    double fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll 4
    for (int q = 0; q < n_off; q++) {
        const int j = d + s_off[q];
        const int tj = s.type[j];
        const double dx = xi - s.x[0][j], dy = yi - s.x[1][j], dz = zi - s.x[2][j];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (tj >= 0 && d2 < g.rc2) {
            const double fp = to_force(tb, ti, tj, d2, dfi, s.df[j]);
            fx += dx * fp; fy += dy * fp; fz += dz * fp;
        }
    }
    if (ACCUM) { fx += s.f[0][d]; fy += s.f[1][d]; fz += s.f[2][d]; }
    s.f[0][d] = fx; s.f[1][d] = fy; s.f[2][d] = fz;
}

// ---- K4 verlet-1 + run-away test: NewtonMotion::firststep (reference src/newton_motion.cpp:30-55) and the
//      displacement test of atom::decide (reference src/atom.cpp:27-52). Operation order is the reference's,
//      with explicit round-to-nearest mul/add (no FMA contraction) so x, v and the run-away decision are
//      bit-exact. Run-aways are appended to `runaway_sites` (device indices); the site is vacated by
//      k_decide_vacate after the list has been sorted into the reference's k,j,i order. ------------------
#define MARK_CAP 4096     // marking atoms per step; beyond it the map is void (stepinfo[2] tells the stencil kernels)
struct VerletPar { double dt; double c[MISA_MAX_TYPES]; int mark_T; unsigned char *hot; unsigned char epoch; unsigned long long *mark_count;
                   float inv100_a, lev_slack;      // 100 / a and 2e-4 / a, both rounded up (disp_level_fast)
                   const P2pPeers *push;           // non-null: band sites store their new position into the neighbours' ghosts (push_site)
                   long long c_begin, c_end; };    // owned-cell ordinals [c_begin, c_end) of each sub-lattice this launch covers (a z-slab; default: all)
// what the integrator kernels touch of the state (a kernel parameter of its own: with the whole Soa -- sixteen pointers -- as the
// parameter the compiler no longer fitted k_verlet1 into 32 registers and spilled freshly loaded values inside the hot path)
struct SoaV { double *x[3], *v[3]; const double *f[3]; const int8_t *type; unsigned char *ulev; };
__host__ __device__ inline SoaV soa_v(const Soa &s) {
    SoaV q;
    for (int k = 0; k < 3; k++) { q.x[k] = s.x[k]; q.v[k] = s.v[k]; q.f[k] = s.f[k]; }
    q.type = s.type; q.ulev = s.ulev;
    return q;
}
// dt / (2 m) of species t WITHOUT indexing the kernel parameter dynamically: `vp.c[t]` made the compiler copy the whole
// parameter struct to local memory in every thread (9 STL + 1 LDL per atom in the SASS of the round-1 kernels)
__device__ __forceinline__ double kick_coef(const VerletPar &vp, const int t) { return t == 0 ? vp.c[0] : (t == 1 ? vp.c[1] : vp.c[2]); }
// max over the warp, then one atomicMax on the bit pattern (non-negative doubles order like unsigned integers)
__device__ __forceinline__ void report_max(double v, unsigned long long *__restrict__ out) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) {
        // the running maximum only grows: a (possibly stale) plain read filters out almost every warp
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        if (bits > *reinterpret_cast<volatile unsigned long long *>(out)) atomicMax(out, bits);
    }
}
// displacement level of an atom: ceil(|x - site| / 0.01a), rounded up like the host's pick_list (+1e-6 A), capped at 255
__device__ __forceinline__ unsigned char disp_level(const double dist2, const double a) {
    const double inv = 100.0 / a;                               // one division per thread, hoisted by the compiler where `a` is uniform
    const double l = ceil(fma(sqrt(dist2), inv, 2e-6 * inv));    // 2e-6 A of slack: never below the host's ceil((d + 1e-6) / (0.01 a))
    return (unsigned char)(l > 255.0 ? 255 : (int)l);
}
// The same level without the fp64 square root (k_verlet1 is a streaming kernel: DSQRT + ceil + the conversions were a third of
// its instructions): single-precision sqrt of the squared displacement, scaled with every rounding pushed UPWARDS -- the level
// is a bound, so "one too high" (probability ~1e-5 per atom, at a level boundary) is safe, "one too low" never happens:
// sqrtf and the conversion err by < 2^-23 relative each, the factor 1 + 2^-20 covers both with room, the slack is the host's.
__device__ __forceinline__ unsigned char disp_level_fast(const double dist2, const float inv100_a, const float slack) {
    const float r = __fsqrt_ru(__double2float_ru(dist2));
    const float l = ceilf(__fmaf_ru(r * 1.00000095f, inv100_a, slack));
    return (unsigned char)(l > 255.0f ? 255 : (int)l);
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// ---- ghost push from INSIDE the producing kernel (p2p.cuh has the protocol; this is its data movement without a kernel of its
//      own). An owned site within the ghost width of a face is the origin of a ghost in up to seven of the 26 surrounding
//      sub-boxes: per dimension it either stays (s = 0) or, from the low band, lands n cells higher in the sub-box below
//      (origin code s = +1) -- from the high band n cells lower in the one above (s = -1); every combination except (0,0,0) is
//      one destination. That arithmetic map equals the composition of the reference's three staged exchanges
//      (misa_b200_plan_push; tests/test_host_logic.py checks the two against each other). Scalars by value: see verlet1_rare.
//      nf fields (3 for positions, 1 for df) at field stride pp->stride from `field0`; positions carry the image shift. ------
struct P2pPeers;
template <int NF>
__device__ __noinline__ void push_site(const P2pPeers *__restrict__ pp, const int d, const int cx, const int y, const int z, const int nx, const int ny,
                                       const int nz, const int gx, const int gy, const int gz, const int sxc, const int sy, const int field0,
                                       const double v0, const double v1, const double v2);
// run-away append (atom::decide's displacement test fired) and hot-cell marking, a few dozen atoms of millions per step.
// Scalars by value on purpose: a reference to a kernel-parameter struct would copy it to local memory in EVERY thread.
__device__ __noinline__ void verlet1_rare(const int d, const long long cell, const int gx, const int gy, const int gz, const int sy, const int sxc,
                                          const bool runaway, const bool mark, unsigned char *__restrict__ hot, const unsigned char epoch,
                                          unsigned long long *__restrict__ mark_count, int *__restrict__ counters, int *__restrict__ runaway_sites,
                                          const int runaway_cap) {
    if (runaway) {
        const int slot = atomicAdd(&counters[0], 1);
        if (slot < runaway_cap) runaway_sites[slot] = d;
        else atomicExch(&counters[3], 1);
    }
    // (a plain look first: once the cap is passed -- thermalisation transients, when most atoms lie above T -- nobody queues on
    // the counter any more; it then stands above MARK_CAP, which is what tells the stencil kernels to ignore the map)
    if (mark && *(volatile unsigned long long *)mark_count <= MARK_CAP && atomicAdd(mark_count, 1ULL) < MARK_CAP) {
        // a far-displaced atom (a few dozen of 2 M at 300 K): every cell within the stencil reach learns that the cheap
        // partner bound mark_T does not hold around it. The mark is the step's epoch byte, so the map is never cleared: a
        // stale byte that aliases 255 steps later only makes a warp keep the global bound (the safe side).
        for (int dz = -gz; dz <= gz; dz++)
            for (int dy = -gy; dy <= gy; dy++)
                for (int dx = -gx; dx <= gx; dx++) hot[cell + ((long long)dz * sy + dy) * sxc + dx] = epoch;
    }
}
// squared displacement of the atom from its ideal site after the drift (0 for vacant sites).
// KICK2: the second half-kick of the step that just finished (NewtonMotion::secondstep, same f) is applied first --
// inside a multi-step call the two streaming passes over v and f become one (bit-identical: the same two rounded adds)
template <bool KICK2, bool PUSH>
__device__ __forceinline__ double verlet1_site(const Geo &g, const SoaV &s, const VerletPar &vp, const int p, const long long c,
                                               int *__restrict__ counters, int *__restrict__ runaway_sites, const int runaway_cap) {
    int cx, y, z;
    const int d = owned_cell_to_dev(g, p, c, cx, y, z);
    const int t = s.type[d];
    if (t < 0) { s.ulev[d] = 0; return 0.0; }

    // all nine loads first (the Soa pointers carry no restrict: interleaved with the stores they were issued in three
    // dependent rounds, and the kernel sat at half of the HBM bandwidth)
    double f[3], v[3], x[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { f[k] = __ldg(s.f[k] + d); v[k] = s.v[k][d]; x[k] = s.x[k][d]; }
    const double cm = kick_coef(vp, t);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double kick = __dmul_rn(cm, f[k]);
        if (KICK2) v[k] = __dadd_rn(v[k], kick);
        v[k] = __dadd_rn(v[k], kick);
        x[k] = __dadd_rn(x[k], __dmul_rn(vp.dt, v[k]));
    }
#pragma unroll
    for (int k = 0; k < 3; k++) { s.v[k][d] = v[k]; s.x[k][d] = x[k]; }
    if (PUSH && (cx < g.gx || cx >= g.nx - g.gx || y < g.gy || y >= g.ny - g.gy || z < g.gz || z >= g.nz - g.gz))
        push_site<3>(vp.push, d, cx, y, z, g.nx, g.ny, g.nz, g.gx, g.gy, g.gz, g.sxc, g.sy, 0, x[0], x[1], x[2]);
    // ideal site, reference src/atom.cpp:34-38: i is the doubled-x sub-box index
    const long long i = 2LL * cx + p;
    const double xt = __dmul_rn(__dmul_rn((double)(i + 2LL * g.lo[0]), 0.5), g.a);
    const double yt = __dmul_rn(__dadd_rn((double)((long long)y + g.lo[1]), (double)(i % 2) * 0.5), g.a);
    const double zt = __dmul_rn(__dadd_rn((double)((long long)z + g.lo[2]), (double)(i % 2) * 0.5), g.a);
    const double ex = __dadd_rn(x[0], -xt), ey = __dadd_rn(x[1], -yt), ez = __dadd_rn(x[2], -zt);
    double dist = __dmul_rn(ex, ex);
    dist = __dadd_rn(dist, __dmul_rn(ey, ey));
    dist = __dadd_rn(dist, __dmul_rn(ez, ez));
    const int lev = disp_level_fast(dist, vp.inv100_a, vp.lev_slack);
    s.ulev[d] = (unsigned char)lev;
    // rare paths out of line (they cost the streaming path 14 registers and a quarter of its occupancy when inlined)
    const bool runaway = dist > g.runaway2, mark = vp.hot && lev > vp.mark_T;
    if (runaway || mark)
        verlet1_rare(d, ((long long)(z + g.gz) * g.sy + (y + g.gy)) * g.sxc + (cx + g.gx), g.gx, g.gy, g.gz, g.sy, g.sxc, runaway, mark, vp.hot, vp.epoch,
                     vp.mark_count, counters, runaway_sites, runaway_cap);
    return dist;
}


// PUSH: band sites also store their new position into the neighbours' ghosts (VerletPar::push) -- a variant of its own so that
// the single-sub-box kernel keeps its 32 registers
template <bool KICK2, bool PUSH = false>
__global__ void __launch_bounds__(MISA_BLOCK, PUSH ? 5 : 8)   // 32 registers (48 with the push): the latency of the nine loads needs every warp it can get
k_verlet1(const Geo g, const SoaV s, const VerletPar vp, const int blocks_per_parity, int *__restrict__ counters,
          int *__restrict__ runaway_sites, const int runaway_cap, unsigned long long *__restrict__ stepinfo) {
    const int p = blockIdx.x >= blocks_per_parity;
    const int b = blockIdx.x - p * blocks_per_parity;
    const long long c = vp.c_begin + (long long)b * blockDim.x + threadIdx.x;
    double dist = 0.0;
    if (c < vp.c_end) dist = verlet1_site<KICK2, PUSH>(g, s, vp, p, c, counters, runaway_sites, runaway_cap);
    // maximum over the BLOCK first: one look at the running maximum per block, not per warp -- 62 500 volatile reads of one word
    // per launch queue up at a single L2 slice (about one per clock: tens of microseconds of a 60-microsecond kernel)
    __shared__ double wmax[MISA_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) dist = fmax(dist, __shfl_xor_sync(0xffffffffu, dist, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = dist;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = wmax[0];
#pragma unroll
        for (int w = 1; w < MISA_BLOCK / 32; w++) m = fmax(m, wmax[w]);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
        if (bits > *reinterpret_cast<volatile unsigned long long *>(&stepinfo[1])) atomicMax(&stepinfo[1], bits);
    }
}
// ---- K5 verlet-2: NewtonMotion::secondstep (reference src/newton_motion.cpp:57-74) -------------------
__global__ void __launch_bounds__(MISA_BLOCK)
k_verlet2(const Geo g, const SoaV s, const VerletPar vp, const int blocks_per_parity) {
    const int p = blockIdx.x >= blocks_per_parity;
    const int b = blockIdx.x - p * blocks_per_parity;
    const long long c = vp.c_begin + (long long)b * blockDim.x + threadIdx.x;
    if (c >= vp.c_end) return;
    int cx, y, z;
    const int d = owned_cell_to_dev(g, p, c, cx, y, z);
    const int t = s.type[d];
    if (t < 0) return;
    const double cm = kick_coef(vp, t);
    double f[3], v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { f[k] = __ldg(s.f[k] + d); v[k] = s.v[k][d]; }   // loads first (no restrict on the Soa pointers)
#pragma unroll
    for (int k = 0; k < 3; k++) s.v[k][d] = __dadd_rn(v[k], __dmul_rn(cm, f[k]));
}

// ---- K6 halo: LatPacker / DfEmbedPacker (reference src/pack/lat_particle_packer.cpp:153-193,
//      src/pack/df_embed_packer.cpp:27-68) as device pack / unpack / local periodic copy -------------------
// message layout for positions: 4 doubles per site {x+shift, y+shift, z+shift, type} (LatParticleData, 32 B)
// Both directions of one exchange stage in ONE launch (the exchange is launch-latency bound: 2.3 MB messages).
// Entry i < n0 belongs to direction 0, the rest to direction 1; the two directions never touch the same site.
struct Halo2 {
    int n0, n1;
    const int *send0, *send1, *recv0, *recv1;
    double sh0[3], sh1[3];
};
__global__ void __launch_bounds__(MISA_BLOCK) k_pack_x2(const Halo2 h, const Soa s, double *__restrict__ buf0, double *__restrict__ buf1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n0 + h.n1) return;
    const bool hi = i >= h.n0;
    if (hi) i -= h.n0;
    const int d = hi ? h.send1[i] : h.send0[i];
    const double *sh = hi ? h.sh1 : h.sh0;
    double4 r;
    const int t = s.type[d];                       // (a vacant site travels with its stale position, like the reference's record)
    r.x = __dadd_rn(site_x(s, 0, d, t), sh[0]);
    r.y = __dadd_rn(site_x(s, 1, d, t), sh[1]);
    r.z = __dadd_rn(site_x(s, 2, d, t), sh[2]);
    r.w = (double)t;
    reinterpret_cast<double4 *>(hi ? buf1 : buf0)[i] = r;
}
__global__ void __launch_bounds__(MISA_BLOCK) k_unpack_x2(const Halo2 h, const Soa s, const double *__restrict__ buf0, const double *__restrict__ buf1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n0 + h.n1) return;
    const bool hi = i >= h.n0;
    if (hi) i -= h.n0;
    const int d = hi ? h.recv1[i] : h.recv0[i];
    const double4 r = reinterpret_cast<const double4 *>(hi ? buf1 : buf0)[i];
    site_set_x(s, d, (int)r.w, r.x, r.y, r.z);
    s.type[d] = (int8_t)(int)r.w;
}
__global__ void __launch_bounds__(MISA_BLOCK) k_copy_x2(const Halo2 h, const Soa s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n0 + h.n1) return;
    const bool hi = i >= h.n0;
    if (hi) i -= h.n0;
    const int a = hi ? h.send1[i] : h.send0[i], b = hi ? h.recv1[i] : h.recv0[i];
    const double *sh = hi ? h.sh1 : h.sh0;
    const int t = s.type[a];
    site_set_x(s, b, t, __dadd_rn(site_x(s, 0, a, t), sh[0]), __dadd_rn(site_x(s, 1, a, t), sh[1]), __dadd_rn(site_x(s, 2, a, t), sh[2]));
    s.type[b] = (int8_t)t;
}
__global__ void __launch_bounds__(MISA_BLOCK) k_pack_12(const Halo2 h, const double *__restrict__ field, double *__restrict__ buf0, double *__restrict__ buf1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n0 + h.n1) return;
    if (i < h.n0) buf0[i] = field[h.send0[i]];
    else buf1[i - h.n0] = field[h.send1[i - h.n0]];
}
__global__ void __launch_bounds__(MISA_BLOCK) k_unpack_12(const Halo2 h, double *__restrict__ field, const double *__restrict__ buf0, const double *__restrict__ buf1) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n0 + h.n1) return;
    if (i < h.n0) field[h.recv0[i]] = buf0[i];
    else field[h.recv1[i - h.n0]] = buf1[i - h.n0];
}
__global__ void __launch_bounds__(MISA_BLOCK) k_copy_12(const Halo2 h, double *__restrict__ field) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h.n0 + h.n1) return;
    if (i < h.n0) field[h.recv0[i]] = field[h.send0[i]];
    else field[h.recv1[i - h.n0]] = field[h.send1[i - h.n0]];
}
// fused periodic ghost fill for a 1x1x1 process grid: the three staged self-exchanges composed into one map
__global__ void __launch_bounds__(MISA_BLOCK)
k_ghost_fill_x(const int n, const int *__restrict__ dst, const int *__restrict__ src, const int8_t *__restrict__ code,
               const Soa s, const double lx, const double ly, const double lz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = src[i], b = dst[i];
    const int cd = code[i]; // (sx+1) + 3*(sy+1) + 9*(sz+1), shifts applied in x,y,z stage order
    const int kx = cd % 3 - 1, ky = (cd / 3) % 3 - 1, kz = cd / 9 - 1;
    const int t = s.type[a];
    double x = site_x(s, 0, a, t), y = site_x(s, 1, a, t), z = site_x(s, 2, a, t);
    if (kx) x = __dadd_rn(x, kx > 0 ? lx : -lx);
    if (ky) y = __dadd_rn(y, ky > 0 ? ly : -ly);
    if (kz) z = __dadd_rn(z, kz > 0 ? lz : -lz);
    site_set_x(s, b, t, x, y, z);
    s.type[b] = (int8_t)t;
}
__global__ void __launch_bounds__(MISA_BLOCK)
k_ghost_fill_1(const int n, const int *__restrict__ dst, const int *__restrict__ src, double *__restrict__ field) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) field[dst[i]] = field[src[i]];
}

// ---- AoS <-> SoA (compat hooks and upload/download): AtomElement is 13 x 8-byte words ---------------
// word: 0 id | 1 type(+pad) | 2-4 x | 5-7 v | 8-10 f | 11 rho | 12 df   (reference src/atom/atom_element.h:18-41)
#define AOS_WORDS 13
enum { F_ID = 1, F_TYPE = 2, F_X = 4, F_V = 8, F_F = 16, F_RHO = 32, F_DF = 64, F_ALL = 127 };

__global__ void __launch_bounds__(MISA_BLOCK)
k_aos_to_soa(const long long n_ext, const long long H, const unsigned long long *__restrict__ aos, const Soa s, const int fields,
             const Geo g = Geo(), const int owned_only = 0, const long long idx0 = 0) {   // records [idx0, n_ext): a z-slab, or everything
    const long long idx = idx0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_ext) return;
    if (owned_only) {
        const int sx = 2 * g.sxc;
        const int x = (int)(idx % sx);
        const long long r = idx / sx;
        const int y = (int)(r % g.sy), z = (int)(r / g.sy);
        if (x < 2 * g.gx || x >= 2 * (g.gx + g.nx) || y < g.gy || y >= g.gy + g.ny || z < g.gz || z >= g.gz + g.nz) return;
    }
    const long long d = ref_to_dev(idx, H);
    const unsigned long long *w = aos + idx * AOS_WORDS;
    if (fields & F_ID) s.id[d] = w[0];
    if (fields & (F_TYPE | F_X)) {
        // the record's position goes where the site's (new) occupancy says: vacant sites keep it aside (Soa::sx)
        const int t_old = s.type[d], t_new = (fields & F_TYPE) ? (int)(int8_t)(int)(unsigned)(w[1] & 0xffffffffull) : t_old;
        const double x = (fields & F_X) ? __longlong_as_double(w[2]) : site_x(s, 0, (int)d, t_old);
        const double y = (fields & F_X) ? __longlong_as_double(w[3]) : site_x(s, 1, (int)d, t_old);
        const double z = (fields & F_X) ? __longlong_as_double(w[4]) : site_x(s, 2, (int)d, t_old);
        if (fields & F_TYPE) s.type[d] = (int8_t)t_new;
        if ((fields & F_X) || (t_old < 0) != (t_new < 0)) site_set_x(s, (int)d, t_new, x, y, z);
    }
    if (fields & F_V) { s.v[0][d] = __longlong_as_double(w[5]); s.v[1][d] = __longlong_as_double(w[6]); s.v[2][d] = __longlong_as_double(w[7]); }
    if (fields & F_F) { s.f[0][d] = __longlong_as_double(w[8]); s.f[1][d] = __longlong_as_double(w[9]); s.f[2][d] = __longlong_as_double(w[10]); }
    if (fields & F_RHO) s.rho[d] = __longlong_as_double(w[11]);
    if (fields & F_DF) s.df[d] = __longlong_as_double(w[12]);
}
// owned_only: write back only sites inside the sub-box (ghost records of the host array stay untouched)
__global__ void __launch_bounds__(MISA_BLOCK)
k_soa_to_aos(const Geo g, unsigned long long *__restrict__ aos, const Soa s, const int fields, const int owned_only, const long long idx0 = 0,
             const long long idx1 = -1) {   // records [idx0, idx1): a z-slab; default everything
    const long long idx = idx0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (idx1 < 0 ? g.n_ext : idx1)) return;
    if (owned_only) {
        const int sx = 2 * g.sxc;
        const int x = (int)(idx % sx);
        const long long r = idx / sx;
        const int y = (int)(r % g.sy), z = (int)(r / g.sy);
        if (x < 2 * g.gx || x >= 2 * (g.gx + g.nx) || y < g.gy || y >= g.gy + g.ny || z < g.gz || z >= g.gz + g.nz) return;
    }
    const long long d = ref_to_dev(idx, g.H);
    unsigned long long *w = aos + idx * AOS_WORDS;
    if (fields & F_ID) w[0] = s.id[d];
    if (fields & F_TYPE) w[1] = (w[1] & 0xffffffff00000000ull) | (unsigned long long)(unsigned)(int)s.type[d];
    if (fields & F_X) {
        const int t = s.type[d];
        w[2] = __double_as_longlong(site_x(s, 0, (int)d, t)); w[3] = __double_as_longlong(site_x(s, 1, (int)d, t)); w[4] = __double_as_longlong(site_x(s, 2, (int)d, t));
    }
    if (fields & F_V) { w[5] = __double_as_longlong(s.v[0][d]); w[6] = __double_as_longlong(s.v[1][d]); w[7] = __double_as_longlong(s.v[2][d]); }
    if (fields & F_F) { w[8] = __double_as_longlong(s.f[0][d]); w[9] = __double_as_longlong(s.f[1][d]); w[10] = __double_as_longlong(s.f[2][d]); }
    if (fields & F_RHO) w[11] = __double_as_longlong(s.rho[d]);
    if (fields & F_DF) w[12] = __double_as_longlong(s.df[d]);
}

// ---- largest squared displacement of any valid atom from its ideal site, over ALL ghost-extended sites (a
//      ghost site's ideal position follows from its extended lattice index; images carry the periodic shift).
//      Sizes the pruned stencil: two lattice atoms can only be within r_c if their SITES are closer than
//      r_c + 2*dmax (atom::decide bounds dmax by 0.2a, reference src/atom.cpp:42). ---------------------------
// hist (optional): 256 bins, how many valid atoms carry each displacement level -- the host picks the marking level from it
__global__ void __launch_bounds__(MISA_BLOCK)
k_max_displacement(const Geo g, const Soa s, unsigned long long *__restrict__ out, unsigned int *__restrict__ hist = nullptr) {
    __shared__ unsigned int sh[256];
    if (hist) { sh[threadIdx.x] = 0; __syncthreads(); }
    const long long d = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double dist = 0.0;
    if (d < g.n_ext && s.type[d] >= 0) {
        const int p = d >= g.H;
        const long long rem = d - (long long)p * g.H;
        const int cx = (int)(rem % g.sxc);
        const long long r = rem / g.sxc;
        const int y = (int)(r % g.sy), z = (int)(r / g.sy);
        const double xt = ((double)(cx - g.gx + g.lo[0]) + 0.5 * p) * g.a;
        const double yt = ((double)(y - g.gy + g.lo[1]) + 0.5 * p) * g.a;
        const double zt = ((double)(z - g.gz + g.lo[2]) + 0.5 * p) * g.a;
        const double ex = s.x[0][d] - xt, ey = s.x[1][d] - yt, ez = s.x[2][d] - zt;
        dist = ex * ex + ey * ey + ez * ez;
    }
    if (d < g.n_ext) {
        const unsigned char lev = disp_level(dist, g.a);
        s.ulev[d] = lev;
        if (hist && s.type[d] >= 0) atomicAdd(&sh[lev], 1u);
    }
    report_max(dist, out);
    if (hist) {
        __syncthreads();
        if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
    }
}
// ---- partner bound of the stencil pruning as a FIELD (serial path: run-aways / inter atoms around, every site's level just
//      re-measured by k_max_displacement, ghosts included): pmax[cell] = the largest displacement level of any valid atom within
//      the stencil reach (+-g cells) of the cell -- a separable running maximum, three passes over the 1.2 M cells. A warp then
//      bounds its partners by the maximum of pmax over its own 32 cells: exact locality, no threshold -- a cascade core keeps
//      its long lists, the thermal rest of the box its short prefixes (with one global bound a single re-occupied site 0.3 a off
//      its position made every warp of the box loop all 228 offsets). ----------------------------------------------------------
__global__ void __launch_bounds__(MISA_BLOCK)
k_pmax_x(const Geo g, const int8_t *__restrict__ type, const unsigned char *__restrict__ ulev, unsigned char *__restrict__ out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.H) return;
    const int cx = (int)(c % g.sxc);
    const long long row = c - cx;
    int m = 0;
    for (int x = max(0, cx - g.gx); x <= min(g.sxc - 1, cx + g.gx); x++) {
        const long long d = row + x;
        if (type[d] >= 0) m = max(m, (int)ulev[d]);
        if (type[d + g.H] >= 0) m = max(m, (int)ulev[d + g.H]);
    }
    out[c] = (unsigned char)m;
}
// running maximum along y (axis 1) or z (axis 2) with radius r
__global__ void __launch_bounds__(MISA_BLOCK)
k_pmax_axis(const Geo g, const int axis, const int r, const unsigned char *__restrict__ in, unsigned char *__restrict__ out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.H) return;
    const long long q = c / g.sxc;
    const int y = (int)(q % g.sy), z = (int)(q / g.sy);
    const int n = axis == 1 ? g.sy : g.sz, p = axis == 1 ? y : z;
    const long long stride = axis == 1 ? g.sxc : (long long)g.sxc * g.sy;
    int m = 0;
    for (int k = max(0, p - r); k <= min(n - 1, p + r); k++) m = max(m, (int)in[c + (long long)(k - p) * stride]);
    out[c] = (unsigned char)m;
}

// ---- diagnostics: sum m v^2 (configuration::mvv, reference src/system_configuration.cpp:61-84), E_pot ---
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[MISA_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < MISA_BLOCK / 32 ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    __syncthreads();
    return v;
}
__global__ void __launch_bounds__(MISA_BLOCK)
k_thermo(const Geo g, const Soa s, const DevTables tb, const int *__restrict__ offs, const int n_off,
         const int blocks_per_parity, const double m0, const double m1, const double m2, double *__restrict__ out) {
    extern __shared__ int s_off[];
    const int p = blockIdx.x >= blocks_per_parity;
    const int b = blockIdx.x - p * blocks_per_parity;
    for (int q = threadIdx.x; q < n_off; q += blockDim.x) s_off[q] = offs[p * n_off + q];
    __syncthreads();
    const long long c = (long long)b * blockDim.x + threadIdx.x;
    double mvv = 0.0, pe = 0.0, cnt = 0.0;
    if (c < g.n_cells_owned) {
        int cx, y, z;
        const int d = owned_cell_to_dev(g, p, c, cx, y, z);
        const int ti = s.type[d];
        if (ti >= 0) {
            const double m = ti == 0 ? m0 : (ti == 1 ? m1 : m2);
            mvv = (s.v[0][d] * s.v[0][d] + s.v[1][d] * s.v[1][d] + s.v[2][d] * s.v[2][d]) * m;
            cnt = 1.0;
            pe = embed_energy(tb, ti, s.rho[d]);
            const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
            double e2 = 0.0;
            for (int q = 0; q < n_off; q++) {
                const int j = d + s_off[q];
                const int tj = s.type[j];
                const double dx = xi - s.x[0][j], dy = yi - s.x[1][j], dz = zi - s.x[2][j];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (tj >= 0 && d2 < g.rc2) e2 += pair_energy(tb, ti, tj, d2);
            }
            pe += 0.5 * e2;
        }
    }
    mvv = block_sum(mvv);
    pe = block_sum(pe);
    cnt = block_sum(cnt);
    if (threadIdx.x == 0) { atomicAdd(&out[0], mvv); atomicAdd(&out[1], pe); atomicAdd(&out[2], cnt); }
}
__global__ void __launch_bounds__(MISA_BLOCK)
k_scale_v(const long long n, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz, const double fac) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { vx[i] *= fac; vy[i] *= fac; vz[i] *= fac; }
}
