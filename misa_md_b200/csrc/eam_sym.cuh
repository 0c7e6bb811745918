// misa_md_b200/csrc/eam_sym.cuh -- PAIR-SYMMETRIC rho / force passes (round 1, fourth generation).
//
// The full-list gather of eam_fast.cuh evaluates every pair from both ends: 2 x 29 spline evaluations per atom on
// a thermal bcc lattice, and the fp64 pipe + the shared-memory table gathers are what bound those kernels. The
// reference's half list (src/atom/neighbour_index.inl:79-92, src/atom.cpp:173-184,333-351) evaluates each pair
// once and SCATTERS into both atoms; a scatter needs atomics (non-deterministic sums) or colouring on a GPU.
// Here each pair is still evaluated once, but its scalar -- rho(r) for latRho, fpair for latForce, both symmetric
// in (i, j) for one species -- goes through a per-pair scratch array in HBM instead of a scatter:
//
//   pass A (k_rho_a / k_force_a): site i loops the UPPER near offsets (one of each +v / -v pair of site separations,
//       restricted to the shells that are practically always in range: 29 of 58), evaluates the pair, adds it to its
//       own sum and stores the scalar to pair[q][i] (coalesced: a warp is 32 consecutive cells of one sub-lattice).
//       The far offsets (rarely in range) stay a full-list test behind the warp vote, as in eam_fast.cuh. Ghost sites
//       within sym_lo / sym_hi cells of the owned box run the near loop too: their upper neighbours are owned atoms.
//   pass B (k_rho_b / k_force_b): owned site j adds pair[slot][j + lower offset] over its LOWER near offsets --
//       29 contiguous 8-byte streams per warp -- (force: times x_j - x_i) to the partial sum of pass A; rho: + df.
//
// Sums are gathers in a fixed order: deterministic, no atomics. Cost: 232 B per site written and read once
// (HBM-bound, ~0.1 ms per pass pair at 2 M atoms) against half of the pair evaluations.
// Species: the single-species loop, and the dilute-alloy loop (every pair from the majority tables, which IS
// symmetric; the minority epilogue of eam_fast.cuh runs per atom in pass A). Vacant sites contribute zeros.
#pragma once
#include "eam_fast.cuh"

// pair[((i >> 5) * n_half + q) * 32 + (i & 31)]: the n_half scalars of 32 consecutive sites are one contiguous 7.4 KB
// chunk, so a warp's 29 streams stay inside one or two DRAM / TLB pages (a [q][n_ext] layout puts each stream in a
// different 2 MB page: pass B ran at 2.7 TB/s with it, profiles/r01w_ncu_sym_first_summary.txt)
__host__ __device__ __forceinline__ size_t pair_index(const long long i, const int q, const int n_half) {
    return (((size_t)i >> 5) * (size_t)n_half + (size_t)q) * 32 + ((size_t)i & 31);
}
struct SymPlan {
    double *pair;        // see pair_index
    const int2 *lo;      // [2][n_half]: (device offset of the lower neighbour, slot of the pair in ITS list)
    int n_half;
    long long n_ext;
};

__device__ __forceinline__ int region_unit_to_dev_b(const Geo &g, const RegionList &rl, const long long u, const int p, const int lane, int &b) {
    b = 0;
    while (b + 1 < rl.n && u >= rl.r[b + 1].u0) b++;
    const Region &r = rl.r[b];
    const long long c = (u - r.u0) * 32 + lane;
    if (c >= (long long)r.nx * r.ny * r.nz) return -1;
    const int cx = (int)(c % r.nx);
    const long long t = c / r.nx;
    const int y = (int)(t % r.ny), z = (int)(t / r.ny);
    return (int)(p * g.H + ((long long)(z + r.z0 + g.gz) * g.sy + (y + r.y0 + g.gy)) * g.sxc + (cx + r.x0 + g.gx));
}

// quiet NaN written as the partial sum of an atom that pass B must recompute generically (dilute lists overflowed)
__device__ __forceinline__ double sym_redo_mark() { return __longlong_as_double(0x7ff8000000000b20LL); }

// ---- out-of-line exact recomputation of one site's pass-A work from the global Hermite tables (a pair below the
//      staged range: close cascade encounters). Rewrites the site's pair scalars; returns its partial sum. ---------
__device__ __noinline__ double slow_rho_half(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                                             const int8_t *__restrict__ type, const int tab_type, const double2 *__restrict__ herm, const int n_r,
                                             const double inv_dr, const double rc2, const int *__restrict__ off, const int n_half, const int n_near,
                                             const int n_off, const int d, double *__restrict__ pairs, const long long n_ext, const bool owned) {
    const double xi = X[d], yi = Y[d], zi = Z[d];
    const bool vi = type ? type[d] >= 0 : true;
    const double2 *tab = herm + (size_t)tab_type * ((size_t)n_r + 1);
    double acc = 0.0;
    for (int q = 0; q < n_off; q++) {
        if (q == n_half) { if (!owned) break; q = n_near; if (q >= n_off) break; }
        const int j = d + off[q];
        const bool vj = type ? type[j] >= 0 : true;
        const double dx = xi - X[j], dy = yi - Y[j], dz = zi - Z[j];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        double v = 0.0;
        if (vi && vj && d2 < rc2) {
            const double r = d2 * rsqrt_fast(d2);
            const Split sx = split_fast(r, inv_dr, n_r - 1, 1);
            v = hval(hbasis(sx.p), __ldg(tab + sx.m), __ldg(tab + sx.m + 1));
        }
        if (q < n_half) pairs[pair_index(d, q, n_half)] = v;
        acc += v;
    }
    return acc;
}
__device__ __noinline__ double3 slow_force_half(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                                                const double *__restrict__ DF, const int8_t *__restrict__ type, const int tab_type,
                                                const double2 *__restrict__ herm, const int nt, const int n_r, const double inv_dr, const double rc2,
                                                const int *__restrict__ off, const int n_half, const int n_near, const int n_off, const int d,
                                                double *__restrict__ pairs, const long long n_ext, const bool owned) {
    const double xi = X[d], yi = Y[d], zi = Z[d], dfi = DF[d];
    const bool vi = type ? type[d] >= 0 : true;
    const size_t tstride = (size_t)n_r + 1;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int q = 0; q < n_off; q++) {
        if (q == n_half) { if (!owned) break; q = n_near; if (q >= n_off) break; }
        const int j = d + off[q];
        const bool vj = type ? type[j] >= 0 : true;
        const double dx = xi - X[j], dy = yi - Y[j], dz = zi - Z[j];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        double fp = 0.0;
        if (vi && vj && d2 < rc2) fp = generic_force_pair_h(herm, tstride, nt, tab_type, tab_type, d2, dfi, DF[j], inv_dr, n_r - 1);
        if (q < n_half) pairs[pair_index(d, q, n_half)] = fp;
        fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
    }
    return make_double3(fx, fy, fz);
}

// ---- rho, pass A ------------------------------------------------------------------------------------------------
// region 0 of `rl` is the owned box (partial sums are stored, far offsets and the dilute epilogue run); the other
// regions are the ghost slabs below it (pair scalars only).
template <bool NOVAC, bool DILUTE>
__global__ void __launch_bounds__(EAM_THREADS, 1)
k_rho_a(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const int n_near_h,
        const TexAll tex, const RegionList rl, const LevelSel ls, const MinorList ml, const SymPlan sy) {
    constexpr bool NEEDTYPE = !NOVAC;
    const int *offs = offs_h;                         // the distance-sorted full list; every warp loops a prefix of it
    const int n_list = n_off_h, n_near = n_near_h;
    const int lg = base_level(ls);
    const bool hot_ok = hot_map_usable(ls);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t mbar;
    const double2 *s_tab = stage_tables(sp, smem, &mbar, offs, 2 * n_list);
    const int *s_off = reinterpret_cast<const int *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = EAM_THREADS / 32;
    const long long upp = rl.units;
    const uint32_t b_el0 = smem_u32(s_tab) - ((uint32_t)sp.row_lo << 4);
    const double rc2 = g.rc2, inv_dr = tb.inv_dr;
    const int n_m1 = tb.n_r - 1, row_lo = sp.row_lo, ns = tex.ns, n_half = sy.n_half;
    const cudaTextureObject_t tx = tex.t;
    const int maj = sp.staged_id[0];
    const double2 *__restrict__ g_herm = sp.g_elec[0];
    const size_t tstride = (size_t)tb.n_r + 1;
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        const int par = u >= upp;
        const long long up = u - (par ? upp : 0);
        int b;
        const int d0 = region_unit_to_dev_b(g, rl, up, par, lane, b);
        const bool owned = b == 0;
        const bool live = d0 >= 0;
        const int d = live ? d0 : region_unit_to_dev_b(g, rl, up, par, 0, b);
        int ti = 0;
        if (NEEDTYPE) ti = s.type[d];
        const int *off = s_off + (par ? n_list : 0);
        const int n_off = list_len(ls, lg, __reduce_max_sync(0xffffffffu, lg >= 0 ? (int)ls.ulev[d] : 0),
                                   __any_sync(0xffffffffu, hot_ok ? cell_hot(ls, d - (par ? ls.H : 0)) : true));
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
        double *__restrict__ prow = sy.pair + pair_index(d, 0, n_half);
        double acc = 0.0;
        int mmin = 0x7fffffff;
        auto pair = [&](const double d2, const bool in) -> double {
            const double r = d2 * rsqrt_fast(d2);
            const Split sx = split_fast(r, inv_dr, n_m1, row_lo);
            mmin = min(mmin, (NEEDTYPE && !in) ? 0x7fffffff : sx.m0);
            double2 r0, r1;
            rows_s(b_el0, sx.m, r0, r1);
            const double v = hval(hbasis(sx.p), r0, r1);
            return in ? v : 0.0;
        };
EAM_UNROLL(EAM_UNROLL_NEAR)
        for (int q = 0; q < n_half; q++) {
            const int j = d + off[q];
            int tj = 0;
            if (NEEDTYPE) tj = s.type[j];
            const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const double v = pair(d2, NEEDTYPE ? (ti >= 0 && tj >= 0 && d2 < rc2) : (d2 < rc2));
            acc += v;
            if (live) prow[q * 32] = v;
        }
        if (owned) {
EAM_UNROLL(EAM_UNROLL_FAR)
            for (int q = n_near; q < n_off; q++) {
                const int j = d + off[q];
                int tj = 0;
                if (NEEDTYPE) tj = s.type[j];
                const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = NEEDTYPE ? (ti >= 0 && tj >= 0 && d2 < rc2) : (d2 < rc2);
                if (__any_sync(0xffffffffu, in)) acc += pair(d2, in);
            }
        }
        bool low = mmin < row_lo;
        bool redo = false;
        if (DILUTE && owned) {
            const int nm = ml.count[d];
            redo = nm == MINOR_OVERFLOW;
            const int maxn = __reduce_max_sync(0xffffffffu, redo ? 0 : nm);
            const int *moff = ml.offs + (par ? ml.n_offs : 0);
EAM_UNROLL(2)
            for (int k = 0; k < maxn; k++) {
                if (k < nm && !redo) {
                    const int e = ml.entry[(size_t)d * MINOR_CAP + k];
                    const int j = d + moff[e & 127];
                    const int tj = minor_species(ml.maj, e >> 7);
                    const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
                    const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    if (d2 < rc2) {
                        const double r = d2 * rsqrt_fast(d2);
                        const Split sx = split_fast(r, inv_dr, n_m1, row_lo);
                        if (sx.m0 < row_lo) redo = true;     // the majority term below was taken from a clamped row
                        const HBasis hb = hbasis(sx.p);
                        double2 r0, r1;
                        rows_s(b_el0, sx.m, r0, r1);
                        const double2 *row = g_herm + (size_t)tj * tstride + sx.m;
                        acc += hval(hb, __ldg(row), __ldg(row + 1)) - hval(hb, r0, r1);
                    }
                }
            }
        }
        if (__any_sync(0xffffffffu, low)) {
            if (low && live)
                acc = slow_rho_half(s.x[0], s.x[1], s.x[2], NEEDTYPE ? s.type : nullptr, maj, g_herm, tb.n_r, inv_dr, rc2, offs + (par ? n_list : 0), n_half,
                                    n_near, n_off, d, sy.pair, sy.n_ext, owned);
            if (DILUTE && low) redo = true;   // its epilogue terms may have used clamped rows too
        }
        if (!live || !owned) continue;
        s.rho[d] = (DILUTE && redo) ? sym_redo_mark() : acc;
    }
}

// ---- rho, pass B (+ df) -----------------------------------------------------------------------------------------
#define SYM_B_THREADS 512    // 16 consecutive warp units per CTA: the lower neighbours' positions are shared through L1
template <bool NOVAC, bool FUSE_DF, bool DILUTE>
__global__ void __launch_bounds__(SYM_B_THREADS, 3)
k_rho_b(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const RegionList rl,
        const LevelSel ls, const SymPlan sy) {
    extern __shared__ int2 s_lo[];
    const int n_half = sy.n_half;
    for (int q = threadIdx.x; q < 2 * n_half; q += blockDim.x) s_lo[q] = sy.lo[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = SYM_B_THREADS / 32;
    const long long upp = rl.units;
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        const int par = u >= upp;
        const long long up = u - (par ? upp : 0);
        const int d = region_unit_to_dev(g, rl, up, par, lane);
        if (d < 0) continue;
        const int ti = s.type[d];
        const int2 *lo = s_lo + (par ? n_half : 0);
        double acc = s.rho[d];
#pragma unroll 8
        for (int m = 0; m < n_half; m++) {
            const int2 e = lo[m];
            acc += __ldcs(sy.pair + pair_index(d + e.x, e.y, n_half));
        }
        if (DILUTE && acc != acc) {   // marked by pass A: the true-species full-list sum from the global tables
            const int *offs = offs_h;
            int n_off = n_off_h, n_near = 0;
            select_list(ls, offs, n_off, n_near);
            acc = slow_rho_atom(s.x[0], s.x[1], s.x[2], s.type, sp.single, sp.g_mono, tb.n_r, tb.inv_dr, g.rc2, offs + (par ? n_off : 0), n_off, d);
        }
        if (!NOVAC && ti < 0) { s.rho[d] = 0.0; continue; }
        s.rho[d] = acc;
        if (FUSE_DF) s.df[d] = d_embed(tb, ti, acc);
    }
}

// ---- force, pass A ----------------------------------------------------------------------------------------------
template <bool NOVAC, bool DILUTE>
__global__ void __launch_bounds__(EAM_THREADS, 1)
k_force_a(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const int n_near_h,
          const TexAll tex, const RegionList rl, const LevelSel ls, const MinorList ml, const SymPlan sy) {
    constexpr bool NEEDTYPE = !NOVAC;
    const int *offs = offs_h;                         // the distance-sorted full list; every warp loops a prefix of it
    const int n_list = n_off_h, n_near = n_near_h;
    const int lg = base_level(ls);
    const bool hot_ok = hot_map_usable(ls);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t mbar;
    const double2 *s_tab = stage_tables(sp, smem, &mbar, offs, 2 * n_list);
    const int *s_off = reinterpret_cast<const int *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = EAM_THREADS / 32;
    const long long upp = rl.units;
    const uint32_t b_el0 = smem_u32(s_tab) - ((uint32_t)sp.row_lo << 4);
    const uint32_t b_ph0 = b_el0 + ((uint32_t)sp.rows_s << 4);
    const double rc2 = g.rc2, inv_dr = tb.inv_dr;
    const int n_m1 = tb.n_r - 1, row_lo = sp.row_lo, ns = tex.ns, n_half = sy.n_half;
    const cudaTextureObject_t tx = tex.t;
    const int maj = sp.staged_id[0];
    const int nt = tb.n_types;
    const double2 *__restrict__ g_herm = sp.g_elec[0];
    const size_t tstride = (size_t)tb.n_r + 1;
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        const int par = u >= upp;
        const long long up = u - (par ? upp : 0);
        int b;
        const int d0 = region_unit_to_dev_b(g, rl, up, par, lane, b);
        const bool owned = b == 0;
        const bool live = d0 >= 0;
        const int d = live ? d0 : region_unit_to_dev_b(g, rl, up, par, 0, b);
        int ti = maj;
        if (NEEDTYPE || DILUTE) ti = s.type[d];
        const int *off = s_off + (par ? n_list : 0);
        const int n_off = list_len(ls, lg, __reduce_max_sync(0xffffffffu, lg >= 0 ? (int)ls.ulev[d] : 0),
                                   __any_sync(0xffffffffu, hot_ok ? cell_hot(ls, d - (par ? ls.H : 0)) : true));
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d], dfi = s.df[d];
        double *__restrict__ prow = sy.pair + pair_index(d, 0, n_half);
        double fx = 0.0, fy = 0.0, fz = 0.0;
        int mmin = 0x7fffffff;
        auto pair = [&](const double d2, const bool in, const int j) -> double {
            const double recip = rsqrt_fast(d2);
            const double dfj = tex_f64(tx, j + 3 * ns);
            const Split sx = split_fast(d2 * recip, inv_dr, n_m1, row_lo);
            mmin = min(mmin, (NEEDTYPE && !in) ? 0x7fffffff : sx.m0);
            const HBasis hb = hbasis(sx.p);
            const HSlope hs = hslope(sx.p);
            double2 r0, r1;
            rows_s(b_ph0, sx.m, r0, r1);
            const double z2 = hval(hb, r0, r1);
            const double z2p = hder(hs, r0, r1);
            rows_s(b_el0, sx.m, r0, r1);
            const double emb = hder(hs, r0, r1) * (dfi + dfj);
            const double fp = -recip * fma(inv_dr, fma(z2p, recip, emb), -(z2 * (recip * recip)));
            return in ? fp : 0.0;
        };
EAM_UNROLL(EAM_UNROLL_NEAR)
        for (int q = 0; q < n_half; q++) {
            const int j = d + off[q];
            int tj = 0;
            if (NEEDTYPE) tj = s.type[j];
            const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const double fp = pair(d2, NEEDTYPE ? (ti >= 0 && tj >= 0 && d2 < rc2) : (d2 < rc2), j);
            fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
            if (live) prow[q * 32] = fp;
        }
        if (owned) {
EAM_UNROLL(EAM_UNROLL_FAR)
            for (int q = n_near; q < n_off; q++) {
                const int j = d + off[q];
                int tj = 0;
                if (NEEDTYPE) tj = s.type[j];
                const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
                const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                const bool in = NEEDTYPE ? (ti >= 0 && tj >= 0 && d2 < rc2) : (d2 < rc2);
                if (__any_sync(0xffffffffu, in)) {
                    const double fp = pair(d2, in, j);
                    fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
                }
            }
        }
        bool low = mmin < row_lo;
        bool redo = false;
        if (DILUTE && owned) {
            const bool mine = ti == ml.maj;       // minority central atoms: k_force_minor writes them
            const int nm = mine ? (int)ml.count[d] : 0;
            redo = mine && nm == MINOR_OVERFLOW;
            const int maxn = __reduce_max_sync(0xffffffffu, nm == MINOR_OVERFLOW ? 0 : nm);
            const int *moff = ml.offs + (par ? ml.n_offs : 0);
EAM_UNROLL(2)
            for (int k = 0; k < maxn; k++) {
                if (k < nm && nm != MINOR_OVERFLOW) {
                    const int e = ml.entry[(size_t)d * MINOR_CAP + k];
                    const int j = d + moff[e & 127];
                    const int tj = minor_species(ml.maj, e >> 7);
                    const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
                    const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    if (d2 < rc2) {
                        const double dfj = tex_f64(tx, j + 3 * ns);
                        const double recip = rsqrt_fast(d2);
                        const Split sx = split_fast(d2 * recip, inv_dr, n_m1, row_lo);
                        if (sx.m0 < row_lo) redo = true;
                        const HBasis hb = hbasis(sx.p);
                        const HSlope hs = hslope(sx.p);
                        double2 r0, r1;
                        rows_s(b_ph0, sx.m, r0, r1);
                        const double z2m = hval(hb, r0, r1), z2pm = hder(hs, r0, r1);
                        rows_s(b_el0, sx.m, r0, r1);
                        const double rho_p_maj = hder(hs, r0, r1);
                        const double fpm = -recip * fma(inv_dr, fma(z2pm, recip, rho_p_maj * (dfi + dfj)), -(z2m * (recip * recip)));
                        const double2 *rp = g_herm + (size_t)(nt + ml.maj * nt + tj) * tstride + sx.m;
                        const double2 *rj = g_herm + (size_t)tj * tstride + sx.m;
                        const double2 p0 = __ldg(rp), p1 = __ldg(rp + 1);
                        const double z2 = hval(hb, p0, p1), z2p = hder(hs, p0, p1);
                        const double emb = fma(rho_p_maj, dfj, hder(hs, __ldg(rj), __ldg(rj + 1)) * dfi);
                        const double fp = -recip * fma(inv_dr, fma(z2p, recip, emb), -(z2 * (recip * recip))) - fpm;
                        fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
                    }
                }
            }
        }
        if (__any_sync(0xffffffffu, low)) {
            if (low && live) {
                const double3 f = slow_force_half(s.x[0], s.x[1], s.x[2], s.df, NEEDTYPE ? s.type : nullptr, maj, g_herm, nt, tb.n_r, inv_dr, rc2,
                                                  offs + (par ? n_list : 0), n_half, n_near, n_off, d, sy.pair, sy.n_ext, owned);
                fx = f.x; fy = f.y; fz = f.z;
            }
            if (DILUTE && low && ti == ml.maj) redo = true;
        }
        if (!live || !owned) continue;
        s.f[0][d] = (DILUTE && redo) ? sym_redo_mark() : fx;
        s.f[1][d] = fy; s.f[2][d] = fz;
    }
}

// ---- force, pass B ----------------------------------------------------------------------------------------------
template <bool NOVAC, bool DILUTE>
__global__ void __launch_bounds__(SYM_B_THREADS, 3)
k_force_b(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const RegionList rl,
          const LevelSel ls, const SymPlan sy, const int maj) {
    extern __shared__ int2 s_lo[];
    const int n_half = sy.n_half;
    for (int q = threadIdx.x; q < 2 * n_half; q += blockDim.x) s_lo[q] = sy.lo[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = SYM_B_THREADS / 32;
    const long long upp = rl.units;
    const double *__restrict__ X = s.x[0], *__restrict__ Y = s.x[1], *__restrict__ Z = s.x[2];
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        const int par = u >= upp;
        const long long up = u - (par ? upp : 0);
        const int d = region_unit_to_dev(g, rl, up, par, lane);
        if (d < 0) continue;
        const int ti = s.type[d];
        if (DILUTE && ti >= 0 && ti != maj) continue;     // k_force_minor
        const int2 *lo = s_lo + (par ? n_half : 0);
        const double xi = X[d], yi = Y[d], zi = Z[d];
        double fx = s.f[0][d], fy = s.f[1][d], fz = s.f[2][d];
#pragma unroll 4
        for (int m = 0; m < n_half; m++) {
            const int2 e = lo[m];
            const int i = d + e.x;
            const double fp = __ldcs(sy.pair + pair_index(i, e.y, n_half));
            fx = fma(xi - X[i], fp, fx); fy = fma(yi - Y[i], fp, fy); fz = fma(zi - Z[i], fp, fz);
        }
        if (DILUTE && fx != fx) {
            const int *offs = offs_h;
            int n_off = n_off_h, n_near = 0;
            select_list(ls, offs, n_off, n_near);
            const double3 f = slow_force_atom(X, Y, Z, s.df, s.type, sp.single, sp.g_mono, tb.n_types, tb.n_r, tb.inv_dr, g.rc2,
                                              offs + (par ? n_off : 0), n_off, d, max(ti, 0));
            fx = f.x; fy = f.y; fz = f.z;
        }
        if (!NOVAC && ti < 0) { s.f[0][d] = 0.0; s.f[1][d] = 0.0; s.f[2][d] = 0.0; continue; }
        s.f[0][d] = fx; s.f[1][d] = fy; s.f[2][d] = fz;
    }
}
