// misa_md_b200/csrc/eam_fast.cuh -- PRODUCTION rho / force kernels (round 1, third generation).
//
// What the ncu capture of the second generation (eam_smem.cuh, profiles/r01j_*) showed: neither the fp64 pipe
// (53 %) nor the L1 data pipe was saturated; the kernels were ISSUE / LATENCY bound -- 100 warp instructions per
// (atom, offset) iteration of which only 35 were fp64, the rest address arithmetic, constant reloads, texture
// handle set-up, reconvergence barriers, the rsqrt slow-path test and a local-memory round trip around the
// rare "row below the staged range" call. This generation removes that overhead:
//   * neighbour fields stay on the TEXTURE pipe (its L1 data path is separate from the LSU path that carries the
//     shared-memory table gathers); a packed {x,y,z,df} copy read with one 256-bit LSU load per neighbour was
//     tried and rejected: it put 1216 more wavefronts per warp on the LSU pipe, which then saturated at 88 %
//     (profiles/r01k_ncu_packed_summary.txt);
//   * the in-range test is warp-uniform (__any_sync -> one vote + one branch, no BSSY/BSYNC): a warp is 32
//     consecutive cells of one sub-lattice, so all lanes see the same lattice shell and agree almost always;
//     lanes that are out of range compute along and their contribution is selected away;
//   * rsqrt is MUFU.RSQ64H + the same third-order Newton step libdevice uses, without its special-case branch
//     (0 < d2 < rc2 here);
//   * rows below the staged range are not handled in the loop at all: the row index is clamped, the lane
//     remembers that it happened, and after the loop such atoms (close cascade encounters only) are recomputed
//     from the global tables by a separate out-of-line routine;
//   * the Hermite cubic is evaluated in basis form (shared between the tables of one pair).
// Multi-species boxes use the same loop: per-lane GENERIC table pointers (shared memory for the staged
// majority tables, global memory for the rest) instead of divergent code paths.
// Arithmetic differs from the reference's operation order only as documented in DESIGN.md 4.3 (<< 1e-10).
#pragma once
#include "eam_smem.cuh"
#ifndef EAM_UNROLL_NEAR
#define EAM_UNROLL_NEAR 4
#endif
#ifndef EAM_UNROLL_FAR
#define EAM_UNROLL_FAR 4
#endif
#ifndef EAM_RHO_G8
#define EAM_RHO_G8 1   // rho kernel: eight near pairs per trip -- twice the neighbour fetches in flight per warp (0.3808 -> 0.3729 ms)
#endif
#ifndef EAM_RHO_G16
#define EAM_RHO_G16 1  // sixteen where the group is long enough (0.3722 -> 0.3666 ms)
#endif
#ifndef EAM_FORCE_G8
#define EAM_FORCE_G8 1 // force kernel: eight near pairs per trip (0.6054 -> 0.5979 ms)
#endif
#ifndef EAM_FORCE_G16
#define EAM_FORCE_G16 0 // (measurement)
#endif
#ifndef EAM_ELEC3
#define EAM_ELEC3 1    // force kernel, single-species loop: slope of elec[maj] from (s_m, dv_m) + s_{m+1} (StagePlan::half_src)
#endif
#ifndef EAM_OFF_V4
#define EAM_OFF_V4 1   // near group: four staged offsets per 16-byte shared-memory load (0: one 4-byte load per pair)
#endif
#define EAM_PRAGMA(x) _Pragma(#x)
// four staged offsets with ONE 16-byte shared-memory load (p 16-byte aligned)
__device__ __forceinline__ int4 ld_off4(const int *p) {
    int4 o;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return o;
}
#define EAM_UNROLL(n) EAM_PRAGMA(unroll n)

// 1/sqrt(a) for 0 < a < inf, normal range: MUFU.RSQ64H seed (rel. error < 2^-20) + one third-order step
//   e = 1 - a y0^2;  y = y0 + y0 e (1/2 + 3/8 e)            -> rel. error ~ 5/16 e^3, below 1 ulp
__device__ __forceinline__ double rsqrt_fast(const double a) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(a));
    const double e = fma(-a, y0 * y0, 1.0);
    return fma(fma(e, 0.375, 0.5), y0 * e, y0);
}

struct Split { int m; double p; int m0; };
// libpot findSpline (oracle/pot.c:table_find): pp = x*inv_dx + 1; m = clamp(int(pp), 1, n-1); p = min(pp - m, 1).
// Additionally m is raised to row_lo (first staged row, >= 1); m0 is the index before that (the caller keeps
// its minimum to learn whether any pair fell below the staged range).
__device__ __forceinline__ Split split_fast(const double x, const double inv_dx, const int n_m1, const int row_lo) {
    Split s;
    const double pp = fma(x, inv_dx, 1.0);
    const double t = __dadd_rd(pp, 4503599627370496.0);   // 2^52 + floor(pp): the integer sits in the low word
    s.m0 = min(__double2loint(t), n_m1);
    s.m = max(s.m0, row_lo);
    const double mf = __hiloint2double(0x43300000, s.m) - 4503599627370496.0;
    const double f = pp - mf;
    s.p = f > 1.0 ? 1.0 : f;
    return s;
}

// Hermite basis on [0,1] (value and d/dp), shared by every table looked up for one pair
struct HBasis { double h01, h10, h11; };          // h00 = 1 - h01
struct HSlope { double g01, g10, g11; };          // g00 = -g01
__device__ __forceinline__ HBasis hbasis(const double p) {
    HBasis b;
    const double q = 1.0 - p, pq = p * q;
    b.h01 = (p * p) * fma(-2.0, p, 3.0);
    b.h10 = pq * q;
    b.h11 = -pq * p;
    return b;
}
__device__ __forceinline__ HSlope hslope(const double p) {
    HSlope g;
    const double q = 1.0 - p;
    g.g01 = 6.0 * (p * q);
    g.g11 = p * fma(3.0, p, -2.0);
    g.g10 = g.g11 + (q - p);                       // 3p^2 - 4p + 1
    return g;
}
__device__ __forceinline__ double hval(const HBasis &b, const double2 r0, const double2 r1) {
    return fma(b.h11, r1.y, fma(b.h10, r0.y, fma(b.h01, r1.x - r0.x, r0.x)));
}
__device__ __forceinline__ double hder(const HSlope &g, const double2 r0, const double2 r1) {
    return fma(g.g11, r1.y, fma(g.g10, r0.y, g.g01 * (r1.x - r0.x)));
}

// rows (m, m+1): shared-space loads for the single-species path, generic loads (shared or global window,
// resolved per lane by the hardware) for the multi-species path
// EAM_EXP_NOCONFLICT (measurement only, WRONG RESULTS, default 0): the low three bits of the row index are replaced by the
// lane's, so that the eight lanes of every quarter-warp hit eight different 16-byte bank groups -- the conflict-free
// shared-memory gather no exact layout can deliver for random rows. Its time is the ceiling of every table-layout trick
// (DESIGN.md section 10.1).
#ifndef EAM_EXP_NOCONFLICT
#define EAM_EXP_NOCONFLICT 0
#endif
__device__ __forceinline__ void rows_s(const uint32_t base, const int m_, double2 &a, double2 &b) {
#if EAM_EXP_NOCONFLICT
    const int m = ((m_ - 7) & ~7) | (int)(threadIdx.x & 7u);   // <= m_, at most 14 rows below it: inside the dynamic shared memory
#else
    const int m = m_;
#endif
    const uint32_t addr = base + ((uint32_t)m << 4);
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "r"(addr));
    asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(b.x), "=d"(b.y) : "r"(addr));
}
__device__ __forceinline__ void rows_g(const unsigned long long base, const int m, double2 &a, double2 &b) {
    const unsigned long long addr = base + ((unsigned long long)(unsigned)m << 4);
    asm("ld.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(addr));
    asm("ld.v2.f64 {%0, %1}, [%2+16];" : "=d"(b.x), "=d"(b.y) : "l"(addr));
}

__device__ __forceinline__ void rows_ldg(const double2 *__restrict__ tab, const int m, double2 &a, double2 &b) {
    a = __ldg(tab + m);
    b = __ldg(tab + m + 1);
}
// ---- pairs that do NOT come from the staged tables (minority species of an alloy, rows below the staged range): the interval's
// cubic in the reference's own monomial form, fetched with ONE 256-bit load (LDG.E.ENL2.256, sm_100) of a 32-byte aligned row.
// The Hermite block needs rows m and m + 1 -- two 128-bit loads that straddle a 32-byte sector every other time; for a scattered
// gather (every lane its own row) the cost is the number of sector requests, which this halves, and value + slope come out of
// five fused multiply-adds (Horner + synthetic division) instead of the 16 of the basis form.
struct Row4 { double c3, c4, c5, c6; };
__device__ __forceinline__ Row4 mono_row(const double *__restrict__ mono, const size_t row) {
    Row4 r;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.c3), "=d"(r.c4), "=d"(r.c5), "=d"(r.c6) : "l"(mono + 4 * row));
    return r;
}
__device__ __forceinline__ double mono_val(const Row4 &r, const double p) { return fma(fma(fma(r.c3, p, r.c4), p, r.c5), p, r.c6); }
__device__ __forceinline__ double mono_der(const Row4 &r, const double p) {      // d/dp = (3 c3 p + 2 c4) p + c5
    const double b2 = fma(r.c3, p, r.c4);
    return fma(fma(r.c3, p, b2), p, fma(b2, p, r.c5));
}
__device__ __forceinline__ void mono_val_der(const Row4 &r, const double p, double &v, double &d) {
    const double b2 = fma(r.c3, p, r.c4), b1 = fma(b2, p, r.c5);
    v = fma(b1, p, r.c6);
    d = fma(fma(r.c3, p, b2), p, b1);
}
// one pair force from the monomial block: phi[ti][tj] value + slope, elec[ti] / elec[tj] slopes (per knot; 1/dr factored out)
__device__ __forceinline__ double mono_fpair(const double *__restrict__ mono, const size_t tstride, const int nt, const int ti, const int tj, const int m,
                                             const double p, const double recip, const double dfi, const double dfj, const double inv_dr) {
    double z2, z2p;
    mono_val_der(mono_row(mono, (size_t)(nt + ti * nt + tj) * tstride + m), p, z2, z2p);
    const double di = mono_der(mono_row(mono, (size_t)ti * tstride + m), p);
    const double dj = tj == ti ? di : mono_der(mono_row(mono, (size_t)tj * tstride + m), p);
    const double emb = fma(di, dfj, dj * dfi);
    return -recip * fma(inv_dr, fma(z2p, recip, emb), -(z2 * (recip * recip)));
}
// EAM_MONO_RHO: the single-species rho kernel evaluates the reference's OWN cubic ((c3 p + c4) p + c5) p + c6 (libpot row columns
// 3-6) from two staged 16-byte half rows instead of rebuilding it from Hermite data: 3 fp64 instructions instead of 11 per pair
// (30 -> 22 in the near loop), the same two LDS.128, the same shared-memory footprint (the rho kernel never needed the staged
// r*phi table).
#ifndef EAM_MONO_RHO
#define EAM_MONO_RHO 1
#endif
__device__ __forceinline__ double mono_val_s(const uint32_t base_a, const uint32_t base_b, const int m_, const double p) {
    double c3, c4, c5, c6;
#if EAM_EXP_NOCONFLICT
    const int m = ((m_ - 7) & ~7) | (int)(threadIdx.x & 7u);   // (measurement only, wrong numbers: see rows_s)
#else
    const int m = m_;
#endif
    const uint32_t o = (uint32_t)m << 4;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c3), "=d"(c4) : "r"(base_a + o));
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(c5), "=d"(c6) : "r"(base_b + o));
    return fma(fma(fma(c3, p, c4), p, c5), p, c6);
}
#ifndef EAM_MULTI_GENERIC
#define EAM_MULTI_GENERIC 0   // 1: per-lane generic pointers (one instruction stream); 0: staged-or-global branch per lane
#endif

// per-CTA directory of generic table base addresses (biased so that base + 16 m is row m):
//   [t] elec[t], [MISA_MAX_TYPES + ti * MISA_MAX_TYPES + tj] phi[ti][tj]
#define EAM_DIR (MISA_MAX_TYPES + MISA_MAX_TYPES * MISA_MAX_TYPES)
__device__ __forceinline__ void build_directory(unsigned long long *dir, const StagePlan &sp, const DevTables &tb, const double2 *s_tab) {
    if (threadIdx.x < EAM_DIR) {
        const int k = threadIdx.x, nt = tb.n_types;
        unsigned long long a = 0;
        int id = -1;
        const double2 *g = nullptr;
        if (k < MISA_MAX_TYPES) { if (k < nt) { g = sp.g_elec[k]; id = k; } }
        else {
            const int ti = (k - MISA_MAX_TYPES) / MISA_MAX_TYPES, tj = (k - MISA_MAX_TYPES) % MISA_MAX_TYPES;
            if (ti < nt && tj < nt) { g = sp.g_phi[ti * nt + tj]; id = MISA_MAX_TYPES + ti * nt + tj; }
        }
        if (g) {
            a = (unsigned long long)g;
            for (int q = 0; q < sp.n_staged; q++)
                if (sp.staged_id[q] == id)
                    a = (unsigned long long)(s_tab + (size_t)q * sp.rows_s) - ((unsigned long long)sp.row_lo << 4); // generic address of the shared copy
        }
        dir[k] = a;
    }
}

// ---- out-of-line recomputation of ONE atom from the global Hermite tables (close encounters: a pair below the
//      staged range). Same arithmetic as the fast loop, no clamping to row_lo. ---------------------------------
// Arguments are scalars / pointers by value on purpose: taking the address of a by-value kernel parameter struct
// would move it (and every access in the hot loop) to local memory.
__device__ __noinline__ double slow_rho_atom(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                                             const int8_t *__restrict__ type, const int single,
                                             const double *__restrict__ mono, const int n_r, const double inv_dr, const double rc2,
                                             const int *__restrict__ off, const int n_off, const int d) {
    const double xi = X[d], yi = Y[d], zi = Z[d];
    const size_t tstride = (size_t)n_r + 1;
    double acc = 0.0;
    for (int q = 0; q < n_off; q++) {
        const int j = d + off[q];
        const int tj = type ? (int)type[j] : single;
        const double dx = xi - X[j], dy = yi - Y[j], dz = zi - Z[j];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        if (tj >= 0 && d2 < rc2) {
            const double r = d2 * rsqrt_fast(d2);
            const Split sx = split_fast(r, inv_dr, n_r - 1, 1);
            acc += mono_val(mono_row(mono, (size_t)tj * tstride + sx.m), sx.p);
        }
    }
    return acc;
}
// mono: the global monomial block, elec[t] at t * (n_r + 1), phi[ti][tj] at (nt + ti * nt + tj) * (n_r + 1) (rows of 4 doubles)
__device__ __noinline__ double3 slow_force_atom(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ Z,
                                                const double *__restrict__ DF, const int8_t *__restrict__ type, const int single,
                                                const double *__restrict__ mono, const int nt, const int n_r, const double inv_dr,
                                                const double rc2, const int *__restrict__ off, const int n_off, const int d, const int ti) {
    const double xi = X[d], yi = Y[d], zi = Z[d], dfi = DF[d];
    const size_t tstride = (size_t)n_r + 1;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int q = 0; q < n_off; q++) {
        const int j = d + off[q];
        const int tj = type ? (int)type[j] : single;
        const double dx = xi - X[j], dy = yi - Y[j], dz = zi - Z[j];
        const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
        if (tj >= 0 && d2 < rc2) {
            const double recip = rsqrt_fast(d2);
            const Split sx = split_fast(d2 * recip, inv_dr, n_r - 1, 1);
            const double fp = mono_fpair(mono, tstride, nt, ti, tj, sx.m, sx.p, recip, dfi, DF[j], inv_dr);
            fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
        }
    }
    return make_double3(fx, fy, fz);
}

// one texture handle for x, y, z and df (one allocation, field stride `ns` doubles; ctx.h d_xyzd)
// (The r*phi rows of the force kernel through this path -- EAM_PHI_TEX, 16-byte texels of the Hermite block -- were measured twice and
// removed: 0.638 -> 0.827 ms in round 1, no gain with every neighbour field on the LSU path in round 2)
struct TexAll { cudaTextureObject_t t; int ns; };
__device__ __forceinline__ double tex_f64(const cudaTextureObject_t t, const int i) {
    const int2 v = tex1Dfetch<int2>(t, i);
    return __hiloint2double(v.y, v.x);
}
// EAM_NB_LDG (measurement only, default 0): bit k set = field k of the NEIGHBOUR (0 x, 1 y, 2 z, 3 df) is read with
// ld.global.nc on the LSU pipe instead of a texture fetch. Measured in round 2 (profiles/r04a_*): every field moved costs
// 2.6 LSU wavefronts per warp-level pair and +0.03 ... +0.05 ms per kernel -- the LSU data pipe (78 % busy in rho, 83 % in force,
// saturating at ~89 %) is the wall of these kernels, the texture front end is half idle (4 cycles per warp-wide TLD, 50 % / 38 %
// busy). DESIGN.md section 4.3d.
#ifndef EAM_NB_LDG
#define EAM_NB_LDG 0
#endif
template <int K>
__device__ __forceinline__ double nb_f64(const cudaTextureObject_t t, const int ns, const double *__restrict__ field, const int j) {
    if ((EAM_NB_LDG >> K) & 1) return __ldg(field + j);
    return tex_f64(t, j + K * ns);
}

// ---- dilute alloys (one species >= 90 % of the sites, e.g. Fe-Cu-Ni 97:2:1) ------------------------------------
// The per-lane staged-or-global choice of the multi-species variants makes nearly every warp-level pair issue BOTH
// paths (86 % of them contain a minority lane at 97:2:1): force 1.16 ms against 0.79 ms for pure Fe. A first attempt
// that listed minority pairs from inside the loop cost more than it saved (+50 % warp instructions under the
// 64-register cap, profiles/r01r_ncu_dilute_inloop_summary.txt). What works: species sit on fixed lattice sites as
// long as nothing runs away, so the sites of minority NEIGHBOURS are a static per-atom list (built once in
// prepare(): up to MINOR_CAP one-byte indices into the widest pruned offset list). The main loop is then the
// SINGLE-species loop, untouched -- every pair evaluated from the staged majority tables -- and an epilogue replaces
// the few listed pairs: + pair(true species) - pair(majority species), both from the global Hermite block.
// Atoms whose OWN species is a minority one are skipped by the force kernel and done by k_force_minor, one warp each.
#define MINOR_CAP 16
#define MINOR_OVERFLOW 255
struct MinorList {
    const unsigned char *count;   // [n_ext] entries of site d, MINOR_OVERFLOW: more than MINOR_CAP (atom recomputed generically)
    const unsigned char *entry;   // [n_ext][MINOR_CAP] bits 0-6: index into offs (n_offs <= 128); bit 7: which of the two minority species.
                                  // One 16-byte row per site: the epilogues fetch a site's whole list with ONE load behind the main loop
                                  // instead of one dependent byte load in front of every trip (the epilogue costs its serial latency)
    const int *offs;              // [2][n_offs] the offset list the indices refer to (widest pruned level)
    int n_offs;
    long long n_ext;
    int maj;
};
// the two species that are not `maj`, in ascending order: which = 0 / 1
__device__ __forceinline__ int entry_byte(const uint4 &e, const int k) {
    const unsigned w = k < 8 ? (k < 4 ? e.x : e.y) : (k < 12 ? e.z : e.w);
    return (int)((w >> (8 * (k & 3))) & 255u);
}
__host__ __device__ __forceinline__ int minor_species(const int maj, const int which) { return which ? (maj == 2 ? 1 : 2) : (maj == 0 ? 1 : 0); }
// one pair from the GLOBAL monomial block (any species)
__device__ __forceinline__ double generic_rho_pair(const double *__restrict__ mono, const size_t tstride, const int tj, const double d2,
                                                   const double inv_dr, const int n_m1) {
    const double r = d2 * rsqrt_fast(d2);
    const Split sx = split_fast(r, inv_dr, n_m1, 1);
    return mono_val(mono_row(mono, (size_t)tj * tstride + sx.m), sx.p);
}
__device__ __forceinline__ double generic_force_pair(const double *__restrict__ mono, const size_t tstride, const int nt, const int ti, const int tj,
                                                     const double d2, const double dfi, const double dfj, const double inv_dr, const int n_m1) {
    const double recip = rsqrt_fast(d2);
    const Split sx = split_fast(d2 * recip, inv_dr, n_m1, 1);
    return mono_fpair(mono, tstride, nt, ti, tj, sx.m, sx.p, recip, dfi, dfj, inv_dr);
}
// the same pair from the global HERMITE block (eam_sym.cuh only: the rejected pair-symmetric passes keep their own arithmetic)
__device__ __forceinline__ double generic_force_pair_h(const double2 *__restrict__ herm, const size_t tstride, const int nt, const int ti, const int tj,
                                                       const double d2, const double dfi, const double dfj, const double inv_dr, const int n_m1) {
    const double recip = rsqrt_fast(d2);
    const Split sx = split_fast(d2 * recip, inv_dr, n_m1, 1);
    const HBasis hb = hbasis(sx.p);
    const HSlope hs = hslope(sx.p);
    const double2 *rp = herm + (size_t)(nt + ti * nt + tj) * tstride + sx.m;
    const double2 *ri = herm + (size_t)ti * tstride + sx.m;
    const double2 *rj = herm + (size_t)tj * tstride + sx.m;
    const double2 p0 = __ldg(rp), p1 = __ldg(rp + 1);
    const double z2 = hval(hb, p0, p1), z2p = hder(hs, p0, p1);
    const double emb = hder(hs, __ldg(ri), __ldg(ri + 1)) * dfj + hder(hs, __ldg(rj), __ldg(rj + 1)) * dfi;
    return -recip * fma(inv_dr, fma(z2p, recip, emb), -(z2 * (recip * recip)));
}

// ---- atoms with a pair below the staged range, cascade form (LOWLIST variants + k_low_fix) ----------------------------------
// A thermal box never has one; there the in-kernel recomputation above (one lane loops, its warp waits) is free. A cascade core
// has thousands, packed into a few hundred warp units: with one lane looping 100+ offsets of dependent global loads per unit the
// stencil kernels of a 5 keV cascade ran 1.35x longer than on a thermal box with the same pair count. The LOWLIST variants --
// launched by the serial (off-lattice) path only -- just append such atoms to a list and store nothing for them; k_low_fix then
// recomputes each listed atom with a whole warp (lanes stride the FULL offset list, fixed-shape butterfly: deterministic).
__device__ __forceinline__ void low_append(const LateWait &lw, const int d) {
    const int k = atomicAdd(lw.low_count, 1);
    if (k < lw.low_cap) lw.low_list[k] = d;
}
template <bool FORCE>
__global__ void __launch_bounds__(256)
k_low_fix(const Geo g, const Soa s, const DevTables tb, const double *__restrict__ herm, const int8_t *__restrict__ type, const int single,
          const int *__restrict__ offs, const int n_list, const int *__restrict__ list, const int *__restrict__ count, const int cap, const bool accum,
          const bool fuse_df) {
    const int lane = threadIdx.x & 31, n = min(*count, cap);
    const size_t tstride = (size_t)tb.n_r + 1;
    const double rc2 = g.rc2, inv_dr = tb.inv_dr;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += (gridDim.x * blockDim.x) >> 5) {
        const int d = list[w], ti = max((int)s.type[d], 0);
        const int *off = offs + (d >= g.H ? n_list : 0);
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d], dfi = FORCE ? s.df[d] : 0.0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int q = lane; q < n_list; q += 32) {
            const int j = d + off[q];
            const int tj = type ? (int)type[j] : single;
            const double dx = xi - s.x[0][j], dy = yi - s.x[1][j], dz = zi - s.x[2][j];
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (tj >= 0 && d2 < rc2) {
                if (!FORCE) a0 += generic_rho_pair(herm, tstride, tj, d2, inv_dr, tb.n_r - 1);
                else {
                    const double fp = generic_force_pair(herm, tstride, tb.n_types, ti, tj, d2, dfi, s.df[j], inv_dr, tb.n_r - 1);
                    a0 = fma(dx, fp, a0); a1 = fma(dy, fp, a1); a2 = fma(dz, fp, a2);
                }
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            if (FORCE) { a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o); }
        }
        if (lane == 0) {
            if (!FORCE) {
                if (accum) a0 += s.rho[d];
                s.rho[d] = a0;
                if (fuse_df) s.df[d] = d_embed(tb, ti, a0);
            } else {
                if (accum) { a0 += s.f[0][d]; a1 += s.f[1][d]; a2 += s.f[2][d]; }
                s.f[0][d] = a0; s.f[1][d] = a1; s.f[2][d] = a2;
            }
        }
    }
}

// ---- K1 rho (+ K2 df fused): atom::latRho / latDf (reference src/atom.cpp:151-192,286-309), full-list gather ----
// SINGLE: every valid site has type sp.single; staged slot 0 = elec[single], slot 1 = phi[single][single].
// NOVAC : the census found no vacant site (ghosts included) -> no per-neighbour type test.
// ACCUM : add to the existing rho (compat hook semantics / inter-atom pass ran first) instead of overwriting.
// The offset list is sorted by site separation; its first n_near entries (sites >= 0.1a inside the cutoff) are
// evaluated without any branch (two independent pairs in flight per thread), the rest behind a warp vote.
// DILUTE: SINGLE loop over the majority tables + the minority-neighbour epilogue (see above).
template <bool SINGLE, bool NOVAC, bool FUSE_DF, bool ACCUM, bool DILUTE = false, bool LOWLIST = false>
__global__ void __launch_bounds__(EAM_THREADS, 1)
k_rho_f(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const int n_near_h,
        const TexAll tex, const RegionList rl, const LevelSel ls, const MinorList ml = MinorList(), const LateWait lw = LateWait()) {
    constexpr bool NEEDTYPE = !SINGLE || !NOVAC;
    const int *offs = offs_h;                         // the distance-sorted full list; every warp loops a prefix of it
    const int n_list = n_off_h, n_near = n_near_h;
    int lg = base_level(ls);                          // with lw.fold_dmax: the own maximum (interior units), widened at the late wait
    const bool hot_ok = hot_map_usable(ls);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ unsigned long long dir[EAM_DIR];
    const int n_pad = (n_list + 3) & ~3;              // stride between the two sub-lattices' staged lists: four offsets per 16-byte load
    const double2 *s_tab = stage_tables(sp, smem, &mbar, offs, 2 * n_list, n_pad);
    if (!SINGLE && EAM_MULTI_GENERIC) { build_directory(dir, sp, tb, s_tab); __syncthreads(); }
    const int *s_off = reinterpret_cast<const int *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = EAM_THREADS / 32;
    const long long upp = rl.units;
    const uint32_t b_el0 = smem_u32(s_tab) - ((uint32_t)sp.row_lo << 4);
    const uint32_t b_el1 = b_el0 + ((uint32_t)sp.rows_s << 4);   // MONO: slot 1 = (c5, c6) rows
    constexpr bool MONO = EAM_MONO_RHO && SINGLE;                 // (the host stages the monomial half rows for exactly these variants)
    const double rc2 = g.rc2, inv_dr = tb.inv_dr;
    const int n_m1 = tb.n_r - 1, row_lo = sp.row_lo, ns = tex.ns;
    const cudaTextureObject_t tx = tex.t;
    const int maj = sp.staged_id[0];                      // species whose tables are staged (== sp.single when SINGLE)
    const int nt = tb.n_types; (void)nt;
    const double2 *__restrict__ g_herm = sp.g_elec[0];
    const size_t tstride = (size_t)tb.n_r + 1;
    bool waited = lw.flags == nullptr;
    post_arrive(lw);
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        int par;
        long long up;
        unit_split(rl, u, par, up);
        if (!waited && u >= 2 * rl.split) {
            const unsigned long long b = late_wait(lw, lane, lw.fold_dmax ? *ls.dmax2_bits : 0ULL);
            if (lw.fold_dmax) lg = level_of_bits(ls, b);
            waited = true;
        }
        const int d0 = region_unit_to_dev(g, rl, up, par, lane);
        const bool live = d0 >= 0;
        const int d = live ? d0 : region_unit_to_dev(g, rl, up, par, 0);   // tail lanes shadow lane 0 (loads stay in bounds), store nothing
        const int ti = s.type[d];
        const int *off = s_off + (par ? n_pad : 0);
        const int n_off = warp_list_len(ls, lg, hot_ok, d, d - (par ? ls.H : 0));
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
        double acc = 0.0;
        int mmin = 0x7fffffff;  // smallest row index of any evaluated pair (out-of-range lanes have large ones)
        auto pair = [&](const double d2, const bool in, const int tj) {
            const double r = d2 * rsqrt_fast(d2);
            const Split sx = split_fast(r, inv_dr, n_m1, row_lo);
            mmin = min(mmin, (NEEDTYPE && !in) ? 0x7fffffff : sx.m0);
            double v;
            if (MONO) v = mono_val_s(b_el0, b_el1, sx.m, sx.p);
            else {
                double2 r0, r1;
                if (SINGLE) rows_s(b_el0, sx.m, r0, r1);
                else if (EAM_MULTI_GENERIC) rows_g(dir[max(tj, 0)], sx.m, r0, r1);
                else if (tj == maj) rows_s(b_el0, sx.m, r0, r1);
                else rows_ldg(g_herm + (size_t)max(tj, 0) * tstride, sx.m, r0, r1);
                v = hval(hbasis(sx.p), r0, r1);
            }
            acc += in ? v : 0.0;
        };
        // the uniform offset loads are LSU wavefronts like any other (5 % of the rho kernel's, 3 % of the force kernel's, on the
        // pipe that bounds both): four offsets per 16-byte load in the branch-free near group
        auto near_pair = [&](const int o) {
            const int j = d + o;
            int tj = 0;
            if (NEEDTYPE) tj = s.type[j];
            const double dx = xi - nb_f64<0>(tx, ns, s.x[0], j), dy = yi - nb_f64<1>(tx, ns, s.x[1], j), dz = zi - nb_f64<2>(tx, ns, s.x[2], j);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            pair(d2, NEEDTYPE ? (tj >= 0 && d2 < rc2) : (d2 < rc2), tj);
        };
        const int n_near4 = EAM_OFF_V4 ? (n_near & ~3) : 0;
        int q = 0;
#if EAM_RHO_G16
        for (; q + 16 <= n_near4; q += 16) {
            const int4 o4 = ld_off4(off + q), o8 = ld_off4(off + q + 4), oc = ld_off4(off + q + 8), og = ld_off4(off + q + 12);
            near_pair(o4.x); near_pair(o4.y); near_pair(o4.z); near_pair(o4.w);
            near_pair(o8.x); near_pair(o8.y); near_pair(o8.z); near_pair(o8.w);
            near_pair(oc.x); near_pair(oc.y); near_pair(oc.z); near_pair(oc.w);
            near_pair(og.x); near_pair(og.y); near_pair(og.z); near_pair(og.w);
        }
#endif
#if EAM_RHO_G8
        for (; q + 8 <= n_near4; q += 8) {                 // eight pairs per trip: twice the fetches in flight per warp
            const int4 o4 = ld_off4(off + q), o8 = ld_off4(off + q + 4);
            near_pair(o4.x); near_pair(o4.y); near_pair(o4.z); near_pair(o4.w);
            near_pair(o8.x); near_pair(o8.y); near_pair(o8.z); near_pair(o8.w);
        }
#endif
        for (; q < n_near4; q += 4) {
            const int4 o4 = ld_off4(off + q);
            near_pair(o4.x); near_pair(o4.y); near_pair(o4.z); near_pair(o4.w);
        }
EAM_UNROLL(EAM_UNROLL_NEAR)
        for (q = n_near4; q < n_near; q++) near_pair(off[q]);
        auto far_pair = [&](const int o) {
            const int j = d + o;
            int tj = 0;
            if (NEEDTYPE) tj = s.type[j];
            const double dx = xi - nb_f64<0>(tx, ns, s.x[0], j), dy = yi - nb_f64<1>(tx, ns, s.x[1], j), dz = zi - nb_f64<2>(tx, ns, s.x[2], j);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool in = NEEDTYPE ? (tj >= 0 && d2 < rc2) : (d2 < rc2);
            if (__any_sync(0xffffffffu, in)) pair(d2, in, tj);
        };
        // (four offsets per load here as in k_force_f was measured: rho 0.3814 -> 0.3836 ms, force 0.6099 -> 0.6060 -- kept there only)
EAM_UNROLL(EAM_UNROLL_FAR)
        for (int qf = n_near; qf < n_off; qf++) far_pair(off[qf]);
        bool low = mmin < row_lo;
        if (DILUTE) {
            const int nm = ml.count[d];
            low = low || nm == MINOR_OVERFLOW;
            const int maxn = __reduce_max_sync(0xffffffffu, nm == MINOR_OVERFLOW ? 0 : nm);
            const uint4 e16 = __ldg(reinterpret_cast<const uint4 *>(ml.entry) + d);
            // What the epilogue costs is its serial latency (entry -> neighbour fields -> table rows -> cubic, trip after trip, while
            // the warp is missing from the main loops that hide each other's latencies): the site's whole list is one 16-byte load and
            // the offsets come from the staged copy of the sorted list (ml.offs is a prefix of it). Requesting the neighbour fields of
            // trip k + 1 before trip k is evaluated was measured and lost (inactive lanes fetch too: force 0.875 -> 0.890 ms).
            // Neighbour fields through the TEX pipe (the LSU pipe is the loaded one), the majority term that is taken
            // back from the staged tables; rows below the staged range make the atom `low` (generic recompute) anyway
EAM_UNROLL(2)
            for (int k = 0; k < maxn; k++) {
                // entries ascend in the sorted offset list and this warp looped only its first n_off offsets: an entry beyond that
                // prefix cannot be in range (the pruning rule) and nothing was evaluated for it -- skip it, and stop when no lane
                // has an entry left inside the prefix
                int e = 0;
                bool act = k < nm && nm != MINOR_OVERFLOW;
                if (act) { e = entry_byte(e16, k); act = (e & 127) < n_off; }   // species are static while the lists are valid: no type gather
                if (!__any_sync(0xffffffffu, act)) break;
                if (act) {
                    const int j = d + off[e & 127];
                    const int tj = minor_species(ml.maj, e >> 7);
                    const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
                    const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    if (d2 < rc2) {
                        const double r = d2 * rsqrt_fast(d2);
                        const Split sx = split_fast(r, inv_dr, n_m1, row_lo);
                        double vm;      // the majority term exactly as the loop evaluated it
                        if (MONO) vm = mono_val_s(b_el0, b_el1, sx.m, sx.p);
                        else { double2 r0, r1; rows_s(b_el0, sx.m, r0, r1); vm = hval(hbasis(sx.p), r0, r1); }
                        acc += mono_val(mono_row(sp.g_mono, (size_t)tj * tstride + sx.m), sx.p) - vm;
                    }
                }
            }
        }
        low = low && ti >= 0;   // (a vacant central site stores zeros whatever its lanes computed)
        if (LOWLIST) {          // cascades: such atoms are listed and recomputed by k_low_fix, one warp each (see there)
            if (low && live) { low_append(lw, d); continue; }
        } else if (__any_sync(0xffffffffu, low)) {
            if (low) acc = slow_rho_atom(s.x[0], s.x[1], s.x[2], (NEEDTYPE || DILUTE) ? s.type : nullptr, sp.single, sp.g_mono, tb.n_r, inv_dr, rc2, offs + (par ? n_list : 0), n_off, d);
        }
        if (!live) continue;
        if (ti < 0) {
            if (!ACCUM) s.rho[d] = 0.0;
            continue;
        }
        if (ACCUM) acc += s.rho[d];
        s.rho[d] = acc;
        if (FUSE_DF) {
            const double dfv = d_embed(tb, ti, acc);
            s.df[d] = dfv;
            if (lw.push_df && (rl.split <= 0 || u >= 2 * rl.split)) {   // interior units hold no band site
                const int rem = d - (par ? (int)g.H : 0), cxe = rem % g.sxc, r2 = rem / g.sxc;
                const int cx = cxe - g.gx, y = r2 % g.sy - g.gy, z = r2 / g.sy - g.gz;
                if (cx < g.gx || cx >= g.nx - g.gx || y < g.gy || y >= g.ny - g.gy || z < g.gz || z >= g.nz - g.gz)
                    push_site<1>(lw.push_df, d, cx, y, z, g.nx, g.ny, g.nz, g.gx, g.gy, g.gz, g.sxc, g.sy, 3, dfv, 0.0, 0.0);
            }
        }
    }
}

// ---- K3 force: atom::latForce (reference src/atom.cpp:311-358), full-list gather ---------------------------
// eam::toForce (oracle/pot.c:pot_to_force): phi = z2/r, phi' = z2'/r - phi/r, fpair = -(phi' + emb)/r with
// emb = rho'_i(r) df_j + rho'_j(r) df_i; z2' and rho' are slopes per knot times 1/dr, factored out:
//   fpair = -(1/r) * ( (1/dr) * (z2'_p / r + emb_p) - z2 / r^2 )
template <bool SINGLE, bool NOVAC, bool ACCUM, bool DILUTE = false, bool LOWLIST = false>
__global__ void __launch_bounds__(EAM_THREADS, 1)
k_force_f(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const int n_near_h,
          const TexAll tex, const RegionList rl, const LevelSel ls, const MinorList ml = MinorList(), const LateWait lw = LateWait()) {
    constexpr bool NEEDTYPE = !SINGLE || !NOVAC;
    const int *offs = offs_h;                         // the distance-sorted full list; every warp loops a prefix of it
    const int n_list = n_off_h, n_near = n_near_h;
    int lg = base_level(ls);                          // with lw.fold_dmax: the own maximum (interior units), widened at the late wait
    const bool hot_ok = hot_map_usable(ls);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t mbar;
    __shared__ unsigned long long dir[EAM_DIR];
    const int n_pad = (n_list + 3) & ~3;              // (see k_rho_f)
    const double2 *s_tab = stage_tables(sp, smem, &mbar, offs, 2 * n_list, n_pad);
    if (!SINGLE && EAM_MULTI_GENERIC) { build_directory(dir, sp, tb, s_tab); __syncthreads(); }
    const int *s_off = reinterpret_cast<const int *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = EAM_THREADS / 32;
    const long long upp = rl.units;
    const uint32_t b_el0 = smem_u32(s_tab) - ((uint32_t)sp.row_lo << 4);
    const uint32_t b_ph0 = b_el0 + ((uint32_t)sp.rows_s << 4);
    // dense slopes of elec[maj] behind the 16-byte slots (staged from slope row_lo & ~1): address of slope m = b_es + 8 m
    constexpr bool e3 = EAM_ELEC3 && SINGLE;              // (the host stages it for exactly these variants: misa_b200.cu:plan_elec3)
    const uint32_t b_es = smem_u32(s_tab + (size_t)sp.n_staged * sp.rows_s) - ((uint32_t)(sp.row_lo & ~1) << 3);
    const double rc2 = g.rc2, inv_dr = tb.inv_dr;
    const int n_m1 = tb.n_r - 1, row_lo = sp.row_lo, ns = tex.ns;
    const cudaTextureObject_t tx = tex.t;
    const int maj = sp.staged_id[0];                      // species whose tables are staged (== sp.single when SINGLE)
    const int nt = tb.n_types; (void)nt;
    const double2 *__restrict__ g_herm = sp.g_elec[0];
    const size_t tstride = (size_t)tb.n_r + 1;
    // slope (per knot) of elec[maj] on interval m: the same three products as hder() -- g11 s_{m+1} + g10 s_m + g01 (v_{m+1} - v_m)
    auto elec_slope = [&](const HSlope &hs, const int m) -> double {
        if (e3) {
            double s0, dv, s1;
            asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(s0), "=d"(dv) : "r"(b_el0 + ((uint32_t)m << 4)));
            asm("ld.shared.f64 %0, [%1+8];" : "=d"(s1) : "r"(b_es + ((uint32_t)m << 3)));
            return fma(hs.g11, s1, fma(hs.g10, s0, hs.g01 * dv));
        }
        double2 r0, r1;
        rows_s(b_el0, m, r0, r1);
        return hder(hs, r0, r1);
    };
    bool waited = lw.flags == nullptr;
    post_arrive(lw);
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        int par;
        long long up;
        unit_split(rl, u, par, up);
        if (!waited && u >= 2 * rl.split) {
            const unsigned long long b = late_wait(lw, lane, lw.fold_dmax ? *ls.dmax2_bits : 0ULL);
            if (lw.fold_dmax) lg = level_of_bits(ls, b);
            waited = true;
        }
        const int d0 = region_unit_to_dev(g, rl, up, par, lane);
        const bool live = d0 >= 0;
        const int d = live ? d0 : region_unit_to_dev(g, rl, up, par, 0);
        const int ti = s.type[d];
        const int tic = max(ti, 0);
        const int *off = s_off + (par ? n_pad : 0);
        const int n_off = warp_list_len(ls, lg, hot_ok, d, d - (par ? ls.H : 0));
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d], dfi = s.df[d];
        unsigned long long d_eli = 0;
        if (!SINGLE && EAM_MULTI_GENERIC) d_eli = dir[tic];
        double fx = 0.0, fy = 0.0, fz = 0.0;
        int mmin = 0x7fffffff;  // smallest row index of any evaluated pair (out-of-range lanes have large ones)
        auto pair = [&](const double dx, const double dy, const double dz, const double d2, const bool in, const int tj, const int j) {
            const double recip = rsqrt_fast(d2);
            const double dfj = nb_f64<3>(tx, ns, s.df, j);
            const Split sx = split_fast(d2 * recip, inv_dr, n_m1, row_lo);
            mmin = min(mmin, (NEEDTYPE && !in) ? 0x7fffffff : sx.m0);
            const HBasis hb = hbasis(sx.p);
            const HSlope hs = hslope(sx.p);
            double z2, z2p, emb;
            double2 r0, r1;
            if (SINGLE) {
                rows_s(b_ph0, sx.m, r0, r1);
                z2 = hval(hb, r0, r1);
                z2p = hder(hs, r0, r1);
                emb = elec_slope(hs, sx.m) * (dfi + dfj);
            } else {
                const int tjc = max(tj, 0);
                double rho_p_from, rho_p_to;
                if (EAM_MULTI_GENERIC) {
                    rows_g(dir[MISA_MAX_TYPES + tic * MISA_MAX_TYPES + tjc], sx.m, r0, r1);
                    z2 = hval(hb, r0, r1);
                    z2p = hder(hs, r0, r1);
                    rows_g(d_eli, sx.m, r0, r1);
                    rho_p_from = hder(hs, r0, r1);
                    rho_p_to = rho_p_from;
                    if (__any_sync(0xffffffffu, tjc != tic)) {
                        rows_g(dir[tjc], sx.m, r0, r1);
                        rho_p_to = hder(hs, r0, r1);
                    }
                } else {
                    const bool mi = tic == maj, mj = tjc == maj;
                    if (mi && mj) rows_s(b_ph0, sx.m, r0, r1);
                    else rows_ldg(g_herm + (size_t)(nt + tic * nt + tjc) * tstride, sx.m, r0, r1);
                    z2 = hval(hb, r0, r1);
                    z2p = hder(hs, r0, r1);
                    if (mi) rows_s(b_el0, sx.m, r0, r1);
                    else rows_ldg(g_herm + (size_t)tic * tstride, sx.m, r0, r1);
                    rho_p_from = hder(hs, r0, r1);
                    rho_p_to = rho_p_from;
                    if (tjc != tic) {
                        if (mj) rows_s(b_el0, sx.m, r0, r1);
                        else rows_ldg(g_herm + (size_t)tjc * tstride, sx.m, r0, r1);
                        rho_p_to = hder(hs, r0, r1);
                    }
                }
                emb = fma(rho_p_from, dfj, rho_p_to * dfi);
            }
            double fp = -recip * fma(inv_dr, fma(z2p, recip, emb), -(z2 * (recip * recip)));
            fp = in ? fp : 0.0;
            fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
        };
        auto near_pair = [&](const int o) {                  // (four offsets per 16-byte load: see k_rho_f)
            const int j = d + o;
            int tj = 0;
            if (NEEDTYPE) tj = s.type[j];
            const double dx = xi - nb_f64<0>(tx, ns, s.x[0], j), dy = yi - nb_f64<1>(tx, ns, s.x[1], j), dz = zi - nb_f64<2>(tx, ns, s.x[2], j);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            pair(dx, dy, dz, d2, NEEDTYPE ? (tj >= 0 && d2 < rc2) : (d2 < rc2), tj, j);
        };
        const int n_near4 = EAM_OFF_V4 ? (n_near & ~3) : 0;
        int q = 0;
#if EAM_FORCE_G16
        for (; q + 16 <= n_near4; q += 16) {
            const int4 o4 = ld_off4(off + q), o8 = ld_off4(off + q + 4), oc = ld_off4(off + q + 8), og = ld_off4(off + q + 12);
            near_pair(o4.x); near_pair(o4.y); near_pair(o4.z); near_pair(o4.w);
            near_pair(o8.x); near_pair(o8.y); near_pair(o8.z); near_pair(o8.w);
            near_pair(oc.x); near_pair(oc.y); near_pair(oc.z); near_pair(oc.w);
            near_pair(og.x); near_pair(og.y); near_pair(og.z); near_pair(og.w);
        }
#endif
#if EAM_FORCE_G8
        for (; q + 8 <= n_near4; q += 8) {
            const int4 o4 = ld_off4(off + q), o8 = ld_off4(off + q + 4);
            near_pair(o4.x); near_pair(o4.y); near_pair(o4.z); near_pair(o4.w);
            near_pair(o8.x); near_pair(o8.y); near_pair(o8.z); near_pair(o8.w);
        }
#endif
        for (; q < n_near4; q += 4) {
            const int4 o4 = ld_off4(off + q);
            near_pair(o4.x); near_pair(o4.y); near_pair(o4.z); near_pair(o4.w);
        }
EAM_UNROLL(EAM_UNROLL_NEAR)
        for (q = n_near4; q < n_near; q++) near_pair(off[q]);
        auto far_pair = [&](const int o) {
            const int j = d + o;
            int tj = 0;
            if (NEEDTYPE) tj = s.type[j];
            const double dx = xi - nb_f64<0>(tx, ns, s.x[0], j), dy = yi - nb_f64<1>(tx, ns, s.x[1], j), dz = zi - nb_f64<2>(tx, ns, s.x[2], j);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool in = NEEDTYPE ? (tj >= 0 && d2 < rc2) : (d2 < rc2);
            if (__any_sync(0xffffffffu, in)) pair(dx, dy, dz, d2, in, tj, j);
        };
        int qf = n_near;
        if (EAM_OFF_V4) {                                  // up to the next multiple of four, then four offsets per load
            for (; (qf & 3) && qf < n_off; qf++) far_pair(off[qf]);
            for (; qf + 4 <= n_off; qf += 4) {
                const int4 o4 = ld_off4(off + qf);
                far_pair(o4.x); far_pair(o4.y); far_pair(o4.z); far_pair(o4.w);
            }
        }
EAM_UNROLL(EAM_UNROLL_FAR)
        for (; qf < n_off; qf++) far_pair(off[qf]);
        bool low = mmin < row_lo;
        if (DILUTE) {
            const bool mine = ti == ml.maj;       // minority central atoms: k_force_minor writes them
            const int nm = mine ? (int)ml.count[d] : 0;
            low = mine && (low || nm == MINOR_OVERFLOW);
            const int maxn = __reduce_max_sync(0xffffffffu, nm == MINOR_OVERFLOW ? 0 : nm);
            const uint4 e16 = __ldg(reinterpret_cast<const uint4 *>(ml.entry) + d);   // (one load for the list, staged offsets: see k_rho_f)
EAM_UNROLL(2)
            for (int k = 0; k < maxn; k++) {
                int e = 0;                                            // (entries beyond the warp's prefix: see k_rho_f)
                bool act = k < nm && nm != MINOR_OVERFLOW;
                if (act) { e = entry_byte(e16, k); act = (e & 127) < n_off; }
                if (!__any_sync(0xffffffffu, act)) break;
                if (act) {
                    const int j = d + off[e & 127];
                    const int tj = minor_species(ml.maj, e >> 7);
                    const double dx = xi - tex_f64(tx, j), dy = yi - tex_f64(tx, j + ns), dz = zi - tex_f64(tx, j + 2 * ns);
                    const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
                    if (d2 < rc2) {
                        const double dfj = tex_f64(tx, j + 3 * ns);
                        const double recip = rsqrt_fast(d2);
                        const Split sx = split_fast(d2 * recip, inv_dr, n_m1, row_lo);
                        const HBasis hb = hbasis(sx.p);
                        const HSlope hs = hslope(sx.p);
                        double2 r0, r1;
                        // majority term as the loop evaluated it (staged rows), to be taken back
                        rows_s(b_ph0, sx.m, r0, r1);
                        const double z2m = hval(hb, r0, r1), z2pm = hder(hs, r0, r1);
                        const double rho_p_maj = elec_slope(hs, sx.m);
                        const double fpm = -recip * fma(inv_dr, fma(z2pm, recip, rho_p_maj * (dfi + dfj)), -(z2m * (recip * recip)));
                        // the pair as it is: phi[maj][tj], rho'_maj * df_j + rho'_tj * df_i (two 256-bit rows of the monomial block)
                        double z2, z2p;
                        mono_val_der(mono_row(sp.g_mono, (size_t)(nt + ml.maj * nt + tj) * tstride + sx.m), sx.p, z2, z2p);
                        const double emb = fma(rho_p_maj, dfj, mono_der(mono_row(sp.g_mono, (size_t)tj * tstride + sx.m), sx.p) * dfi);
                        const double fp = -recip * fma(inv_dr, fma(z2p, recip, emb), -(z2 * (recip * recip))) - fpm;
                        fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
                    }
                }
            }
        }
        low = low && ti >= 0;
        if (LOWLIST) {
            if (low && live) { low_append(lw, d); continue; }
        } else if (__any_sync(0xffffffffu, low)) {
            if (low) {
                const double3 f = slow_force_atom(s.x[0], s.x[1], s.x[2], s.df, (NEEDTYPE || DILUTE) ? s.type : nullptr, sp.single, sp.g_mono, tb.n_types,
                                                  tb.n_r, inv_dr, rc2, offs + (par ? n_list : 0), n_off, d, tic);
                fx = f.x; fy = f.y; fz = f.z;
            }
        }
        if (!live) continue;
        if (DILUTE && ti >= 0 && ti != ml.maj) continue;
        if (ti < 0) {
            if (!ACCUM) { s.f[0][d] = 0.0; s.f[1][d] = 0.0; s.f[2][d] = 0.0; }
            continue;
        }
        if (ACCUM) { fx += s.f[0][d]; fy += s.f[1][d]; fz += s.f[2][d]; }
        s.f[0][d] = fx; s.f[1][d] = fy; s.f[2][d] = fz;
    }
}

// ---- dilute alloys: the static lists (built in prepare(), after the first ghost exchange) and the minority atoms ----
// per owned site: which entries of the widest pruned offset list point at a site holding a minority atom
__global__ void __launch_bounds__(MISA_BLOCK) k_build_minor_lists(const Geo g, const int8_t *__restrict__ type, const int maj, const int blocks_per_parity,
                                                                  const int *__restrict__ offs, const int n_offs, unsigned char *__restrict__ count,
                                                                  unsigned char *__restrict__ entry, int *__restrict__ minor, int *__restrict__ n_minor) {
    const int p = blockIdx.x >= blocks_per_parity;
    const long long c = (long long)(blockIdx.x - p * blocks_per_parity) * MISA_BLOCK + threadIdx.x;
    int d = -1, t = -1;
    if (c < g.n_cells_owned) {
        int cx, y, z;
        d = owned_cell_to_dev(g, p, c, cx, y, z);
        t = type[d];
        const int *off = offs + (p ? n_offs : 0);
        int cnt = 0;
        for (int q = 0; q < n_offs; q++) {
            const int tj = type[d + off[q]];
            if (tj >= 0 && tj != maj) {
                if (cnt < MINOR_CAP) entry[(size_t)d * MINOR_CAP + cnt] = (unsigned char)(q | (tj == minor_species(maj, 1) ? 128 : 0));
                cnt++;
            }
        }
        count[d] = (unsigned char)(cnt > MINOR_CAP ? MINOR_OVERFLOW : cnt);
    }
    // owned atoms of a minority species -> device index list (order irrelevant: entries are independent)
    const bool minor_atom = d >= 0 && t >= 0 && t != maj;
    const unsigned m = __ballot_sync(0xffffffffu, minor_atom);
    if (m == 0) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(n_minor, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (minor_atom) minor[base + __popc(m & ((1u << lane) - 1u))] = d;
}
// force on the minority atoms, one warp each: lanes stride the offsets of the atom's parity, fixed-shape butterfly
// reduction (deterministic per atom; atoms are independent of each other)
__global__ void __launch_bounds__(128) k_force_minor(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h,
                                                     const int n_off_h, const LevelSel ls, const int *__restrict__ list, const int n, const TexAll tex) {
    const int *offs = offs_h;             // (address-ordered copy of the host's choice: launch_force_minor)
    int n_off = n_off_h;
    select_list_addr(ls, offs, n_off);
    const int lane = threadIdx.x & 31;
    const int nt = tb.n_types, n_m1 = tb.n_r - 1;
    const double2 *__restrict__ g_herm = sp.g_elec[0];
    const size_t tstride = (size_t)tb.n_r + 1;
    const double rc2 = g.rc2, inv_dr = tb.inv_dr;
    // The kernel is LATENCY-bound (one warp per atom, 3-4 trips of dependent gathers: offset -> neighbour fields -> table rows, at
    // half occupancy): the lane's offsets of both parities live in registers for the whole kernel (lists of up to 128 offsets: every
    // pruned level of a dilute box), and the neighbour fields of all four trips are requested before the first pair is evaluated.
    int o_reg[2][4];
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int k = 0; k < 4; k++) o_reg[p][k] = lane + 32 * k < n_off ? offs[p * n_off + lane + 32 * k] : 0;
    for (int w = (blockIdx.x * 128 + threadIdx.x) >> 5; w < n; w += (gridDim.x * 128) >> 5) {
        const int d = list[w];
        const int ti = s.type[d];
        if (ti < 0) continue;                 // cannot happen while the lists are valid (no run-away since they were built)
        const bool par = d >= g.H;
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d], dfi = s.df[d];
        double fx = 0.0, fy = 0.0, fz = 0.0;
        // every lane gathers a different neighbour; the list is in address order, so consecutive lanes read runs of neighbouring
        // sites of one row
        int jj[4], tt[4];
        double ex[4], ey[4], ez[4], ed[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            jj[k] = d + (par ? o_reg[1][k] : o_reg[0][k]);       // (a lane beyond the list reads the atom itself: excluded below)
            tt[k] = s.type[jj[k]];
            ex[k] = tex_f64(tex.t, jj[k]); ey[k] = tex_f64(tex.t, jj[k] + tex.ns); ez[k] = tex_f64(tex.t, jj[k] + 2 * tex.ns);
            ed[k] = tex_f64(tex.t, jj[k] + 3 * tex.ns);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double dx = xi - ex[k], dy = yi - ey[k], dz = zi - ez[k];
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (lane + 32 * k < n_off && tt[k] >= 0 && d2 < rc2) {
                const double fp = generic_force_pair(sp.g_mono, tstride, nt, ti, tt[k], d2, dfi, ed[k], inv_dr, n_m1);
                fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
            }
        }
        const int *off = offs + (par ? n_off : 0);
        for (int q = lane + 128; q < n_off; q += 32) {          // lists longer than 128 offsets (no pruning): the plain loop
            const int j = d + off[q];
            const int tj = s.type[j];
            const double dx = xi - tex_f64(tex.t, j), dy = yi - tex_f64(tex.t, j + tex.ns), dz = zi - tex_f64(tex.t, j + 2 * tex.ns);
            const double d2 = fma(dz, dz, fma(dy, dy, dx * dx));
            if (tj >= 0 && d2 < rc2) {
                const double fp = generic_force_pair(sp.g_mono, tstride, nt, ti, tj, d2, dfi, tex_f64(tex.t, j + 3 * tex.ns), inv_dr, n_m1);
                fx = fma(dx, fp, fx); fy = fma(dy, fp, fy); fz = fma(dz, fp, fz);
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o);
            fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o);
        }
        if (lane == 0) { s.f[0][d] = fx; s.f[1][d] = fy; s.f[2][d] = fz; }
    }
}

// (A variant with the minority species' own tables staged in shared memory -- one launch per species, 154 KB of staging per CTA --
// was built, parity-tested, measured slower (force 0.987 against 0.944 ms, profiles/r01ag_*) and removed in round 2.)

// ---- diagnostics for the fp64 view of the roofline (bench.py, not on the step path): what the stencil kernels LOOP on the
//      current state -- offsets per atom (the per-warp prefix), pair evaluations per atom (branch-free near group + the far
//      offsets a warp-wide vote lets through) and pairs actually inside the cutoff. Same list_len / vote logic as k_force_f.
//      out[0] lanes (live atoms), [1] offsets looped, [2] pair evaluations, [3] pairs in range -----------------------------
__global__ void __launch_bounds__(256)
k_stencil_stats(const Geo g, const Soa s, const int *__restrict__ offs, const int n_list, const int n_near, const RegionList rl, const LevelSel ls,
                unsigned long long *__restrict__ out) {
    const int lg = base_level(ls);
    const bool hot_ok = hot_map_usable(ls);
    const int lane = threadIdx.x & 31;
    unsigned long long lanes = 0, looped = 0, evals = 0, inr = 0;
    for (long long u = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < 2 * rl.units; u += ((long long)gridDim.x * blockDim.x) >> 5) {
        int par;
        long long up;
        unit_split(rl, u, par, up);
        const int d0 = region_unit_to_dev(g, rl, up, par, lane);
        const bool live = d0 >= 0;
        const int d = live ? d0 : region_unit_to_dev(g, rl, up, par, 0);
        const int *off = offs + (par ? n_list : 0);
        const int n_off = warp_list_len(ls, lg, hot_ok, d, d - (par ? ls.H : 0));
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
        for (int q = 0; q < n_off; q++) {
            const int j = d + off[q];
            const double dx = xi - s.x[0][j], dy = yi - s.x[1][j], dz = zi - s.x[2][j];
            const bool in = s.type[j] >= 0 && fma(dz, dz, fma(dy, dy, dx * dx)) < g.rc2;
            const bool ev = q < n_near || __any_sync(0xffffffffu, in);
            if (live) { looped++; evals += ev; inr += in && s.type[d] >= 0; }
        }
        lanes += live;
    }
    for (int o = 16; o > 0; o >>= 1) {
        lanes += __shfl_xor_sync(0xffffffffu, lanes, o); looped += __shfl_xor_sync(0xffffffffu, looped, o);
        evals += __shfl_xor_sync(0xffffffffu, evals, o); inr += __shfl_xor_sync(0xffffffffu, inr, o);
    }
    if (lane == 0) { atomicAdd(out, lanes); atomicAdd(out + 1, looped); atomicAdd(out + 2, evals); atomicAdd(out + 3, inr); }
}
