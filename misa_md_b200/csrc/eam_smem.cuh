// misa_md_b200/csrc/eam_smem.cuh -- the production rho / force kernels: spline tables staged in shared
// memory by TMA bulk copies, persistent CTAs (one per SM), warp = 32 consecutive owned cells of ONE sub-lattice.
//
// Why (profiles/r01a_ncu_full_summary.csv): with the 7-coefficient rows in global memory both stencil kernels
// sit at 98 % of the L1 LSU wavefront limit (every pair gathers 2-3 rows of 64 B from 32 different lines per
// warp instruction) while the fp64 pipe idles at 13-17 %. Here the tables live in shared memory in HERMITE
// form -- one double2 (value, knot slope) per knot, i.e. columns 6 and 5 of the reference's 7-coefficient row
// (libpot InterpolationObject / LAMMPS array2spline, restated in oracle/pot.c:table_build) -- and the two cubic
// coefficients are rebuilt in registers with the formulas of table_build:
//     s4 = 3 (v1 - v0) - 2 d0 - d1,   s3 = d0 + d1 - 2 (v1 - v0),
//     value = ((s3 p + s4) p + d0) p + v0,   derivative = ((3 s3 p + 2 s4) p + d0) / dx.
// A pair then costs 2 LDS.128 per table instead of 7 scattered LDG.64, and 16 B per knot lets the r-range a
// solid ever visits ([r_lo, r_c], about 3200 knots) fit in 51 KB per table. Rows below r_lo (close cascade
// encounters) and tables that are not staged are read from the global Hermite copies -- same arithmetic.
// set_potential() verifies on the host that the caller's 7-coefficient rows ARE the Hermite-consistent ones;
// if not, the generic global-table kernels in kernels.cuh are used instead.
//
// Deviations from the reference's operation order (all far below the 1e-10 parity bar):
// r = d2 * rsqrt(d2) instead of sqrt(d2); 1/r = rsqrt(d2); the derivative multiplies by 1/dx.
#pragma once
#include "ctx.h"
#include "kernels.cuh"

// Threads per persistent CTA (one CTA per SM). 768 threads leave 85 registers per thread -- room for four neighbour pairs in
// flight (EAM_UNROLL_NEAR 4 in eam_fast.cuh) -- and measured best on B200 (profiles/r01ai_variants.log: 1024 x unroll 2
// 1.149 ms/step, 1024 x 4 1.122, 768 x 4 1.106, 640 x 4 1.152, 512 x 4 1.197).
#ifndef EAM_THREADS
#define EAM_THREADS 768
#endif
#define EAM_MAX_STAGED 4

struct StagePlan {
    // global Hermite copies, always present: elec[t] and phi[ti * n_types + tj], each (n_r + 1) double2
    const double2 *g_elec[MISA_MAX_TYPES];
    const double2 *g_phi[MISA_MAX_TYPES * MISA_MAX_TYPES];
    int n_staged;                       // tables copied to shared memory by this launch
    int staged_id[EAM_MAX_STAGED];      // < MISA_MAX_TYPES: elec[id]; else phi[id - MISA_MAX_TYPES]
    int row_lo;                         // first staged row
    int rows_s;                         // staged rows per table (row_lo .. n_r)
    int off_bytes;                      // byte offset of the table area inside dynamic smem (after the offsets)
    int single;                         // >= 0: every valid site has this type (single-species fast path)
    // src[k] != null: staged slot k is copied from there instead (rows of 16 bytes like the Hermite tables). Used by the
    // monomial form of the single-species rho kernel: slot 0 = (c3, c4), slot 1 = (c5, c6) of the reference's own
    // 7-coefficient rows of elec[maj] (eam_fast.cuh, EAM_MONO_RHO)
    const double2 *src[EAM_MAX_STAGED];
    // every r table once more as 32-byte rows (c3, c4, c5, c6) of the reference's own 7-coefficient rows, table order as the
    // Hermite block (elec[t], then phi[ti * n_types + tj]), 32-byte aligned: ONE 256-bit load per (table, interval) for every
    // pair that does not come from the staged tables (eam_fast.cuh: mono_row)
    const double *g_mono;
    // force kernel, single-species loop: the embedding term needs only the SLOPE of elec[maj] on the interval -- three numbers
    // (s_m, v_{m+1} - v_m, s_{m+1}) instead of two whole Hermite rows. Slot 0 then holds rows (s_m, v_{m+1} - v_m) (src[0]) and
    // a dense array of the slopes follows the 16-byte slots: one 128-bit + one 64-bit gather per pair instead of two 128-bit
    // ones (8.8 + 6.2 against 17.6 LSU wavefronts). half_src points at slope (row_lo & ~1) (16-byte aligned source), null = off.
    const double *half_src;
    int half_bytes;                     // bytes staged from half_src (multiple of 16)
};

// ---- TMA / mbarrier plumbing (1-D bulk copies; SASS: UBLKCP) -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(b)), "r"(phase)
                     : "memory");
    } while (!ok);
}

// Stage the planned tables and the neighbour offsets; returns the shared-memory table area.
// par_stride > 0: the odd sub-lattice's list starts at that (multiple-of-4) index instead of right behind the even one, the gap
// holds zeros -- so that four consecutive offsets of either list are ONE aligned 16-byte shared-memory load (eam_fast.cuh)
__device__ __forceinline__ double2 *stage_tables(const StagePlan &sp, unsigned char *smem, uint64_t *mbar, const int *__restrict__ offs,
                                                 const int n_off2, const int par_stride = 0) {
    int *s_off = reinterpret_cast<int *>(smem);
    double2 *s_tab = reinterpret_cast<double2 *>(smem + sp.off_bytes);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    if (par_stride > 0) {
        const int n = n_off2 >> 1;
        for (int q = threadIdx.x; q < 2 * par_stride; q += blockDim.x) {
            const int p = q >= par_stride, r = q - (p ? par_stride : 0);
            s_off[q] = r < n ? offs[p * n + r] : 0;
        }
    } else
    for (int q = threadIdx.x; q < n_off2; q += blockDim.x) s_off[q] = offs[q];
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)sp.rows_s * 16u;
        mbar_expect_tx(mbar, bytes * (uint32_t)sp.n_staged + (sp.half_src ? (uint32_t)sp.half_bytes : 0u));
        if (sp.half_src) {
            unsigned char *dst = reinterpret_cast<unsigned char *>(s_tab + (size_t)sp.n_staged * sp.rows_s);
            for (uint32_t o = 0; o < (uint32_t)sp.half_bytes; o += 32768u)
                tma_load_1d(dst + o, reinterpret_cast<const unsigned char *>(sp.half_src) + o, min(32768u, (uint32_t)sp.half_bytes - o), mbar);
        }
        for (int k = 0; k < sp.n_staged; k++) {
            const int id = sp.staged_id[k];
            const double2 *src = (sp.src[k] ? sp.src[k] : (id < MISA_MAX_TYPES ? sp.g_elec[id] : sp.g_phi[id - MISA_MAX_TYPES])) + sp.row_lo;
            unsigned char *dst = reinterpret_cast<unsigned char *>(s_tab + (size_t)k * sp.rows_s);
            for (uint32_t o = 0; o < bytes; o += 32768u)
                tma_load_1d(dst + o, reinterpret_cast<const unsigned char *>(src) + o, min(32768u, bytes - o), mbar);
        }
    }
    mbar_wait(mbar, 0);
    return s_tab;
}

// ---- spline helpers ------------------------------------------------------------------------------------
// libpot findSpline (oracle/pot.c:table_find): p = x*inv_dx + 1; m = clamp(int(p), 1, n-1); p = min(p - m, 1).
// int(p) is taken with a round-down add of 2^52 (p > 0): no F2I / I2F on the XU pipe.
__device__ __forceinline__ int split_index(const double x, const double inv_dx, const int n, double &frac) {
    const double pp = fma(x, inv_dx, 1.0);
    const double t = __dadd_rd(pp, 4503599627370496.0);              // 2^52 + floor(pp): the integer sits in the low word
    const int m = max(1, min(__double2loint(t), n - 1));            // clamp like the reference (off-table arguments)
    const double mf = __hiloint2double(0x43300000, m) - 4503599627370496.0; // (double)m, exact, no conversion instruction
    frac = fmin(pp - mf, 1.0);
    return m;
}
struct Cubic { double s3, s4, d0, v0; };
__device__ __forceinline__ Cubic hermite(const double2 a, const double2 b) {
    const double dv = b.x - a.x;
    Cubic c;
    c.s4 = 3.0 * dv - 2.0 * a.y - b.y;
    c.s3 = a.y + b.y - 2.0 * dv;
    c.d0 = a.y;
    c.v0 = a.x;
    return c;
}
__device__ __forceinline__ double cubic_value(const Cubic &c, const double p) { return ((c.s3 * p + c.s4) * p + c.d0) * p + c.v0; }
// derivative w.r.t. p (multiply by 1/dx for d/dx)
__device__ __forceinline__ double cubic_slope(const Cubic &c, const double p) { return (3.0 * c.s3 * p + 2.0 * c.s4) * p + c.d0; }

// single-species path: rows (m, m+1) of a STAGED table through explicit shared-space loads (2 x LDS.128);
// `base` = 32-bit shared address of the slot, biased by -row_lo rows. Rows below the staged range are rare
// (r < r_lo only in close cascade encounters) and take a real branch to the global copy.
__device__ __noinline__ void fetch_rows_global(const double2 *__restrict__ g_rows, const int m, double2 &a, double2 &b) {
    a = __ldg(g_rows + m);
    b = __ldg(g_rows + m + 1);
}
__device__ __forceinline__ Cubic fetch_cubic_s(const uint32_t base, const double2 *__restrict__ g_rows, const int row_lo, const int m) {
    double2 a, b;
    if (__builtin_expect(m >= row_lo, 1)) {
        const uint32_t addr = base + ((uint32_t)m << 4);
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "r"(addr));
        asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(b.x), "=d"(b.y) : "r"(addr));
    } else {
        fetch_rows_global(g_rows, m, a, b);
    }
    return hermite(a, b);
}

// multi-species path: the majority species' tables are the staged ones (slot 0 = elec[maj], slot 1 = phi[maj][maj]);
// every other table is read from the global Hermite block, table t at g_herm + t * (n_r + 1)
// (t = type for elec, n_types + ti * n_types + tj for phi). Same arithmetic either way.
__device__ __forceinline__ Cubic fetch_cubic_m(const bool staged, const uint32_t base, const double2 *__restrict__ g_rows, const int row_lo,
                                               const int m) {
    double2 a, b;
    if (staged && m >= row_lo) {
        const uint32_t addr = base + ((uint32_t)m << 4);
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "r"(addr));
        asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(b.x), "=d"(b.y) : "r"(addr));
    } else {
        a = __ldg(g_rows + m);
        b = __ldg(g_rows + m + 1);
    }
    return hermite(a, b);
}

// warp work unit u -> owned cell of this lane (or -1)
__device__ __forceinline__ int unit_to_dev(const Geo &g, const long long u, const long long units_per_parity, const int lane) {
    const int p = u >= units_per_parity;
    const long long c = (u - (long long)p * units_per_parity) * 32 + lane;
    if (c >= g.n_cells_owned) return -1;
    int cx, y, z;
    return owned_cell_to_dev(g, p, c, cx, y, z);
}

// ---- work regions: a kernel launch covers a list of boxes of owned cells (the whole sub-box, its interior, or the
//      six boundary slabs), so that interior cells can be computed while the ghost exchange is in flight --------
#define EAM_MAX_REGIONS 7
struct Region { int x0, y0, z0, nx, ny, nz; long long u0; };   // u0: first warp unit of this box (within one parity)
struct RegionList { int n; long long units; Region r[EAM_MAX_REGIONS]; long long split; };  // units: warp units per parity
// split > 0: region 0 is the interior and holds `split` units per parity; the launch then visits BOTH parities' interior units
// before any boundary unit, so that a warp can wait for the neighbours' ghost push as late as possible (LateWait below)
__host__ __device__ __forceinline__ void unit_split(const RegionList &rl, const long long u, int &par, long long &up) {
    if (rl.split <= 0) { par = u >= rl.units; up = u - (par ? rl.units : 0); return; }
    if (u < 2 * rl.split) { par = u >= rl.split; up = u - (par ? rl.split : 0); return; }
    const long long v = u - 2 * rl.split, ub = rl.units - rl.split;
    par = v >= ub;
    up = rl.split + v - (par ? ub : 0);
}
// Ghost push of p2p.cuh consumed INSIDE the stencil kernel: interior units read no ghost site (nor any 32-byte sector that
// holds one), so they run while the neighbours' stores are still in flight; the first boundary unit of a warp spins on the
// arrive flags (ld.acquire.sys). The L1 / texture cache is cold for every sector with a ghost in it until then.
// fold_dmax: the neighbours' displacement maxima arrive with the push (p2p.cuh, P2P_DMAX words, written in front of ARRIVE);
// late_wait then returns max(own, neighbours') as the partner bound of the boundary units. Interior units need only the own
// maximum: every partner of theirs is an owned atom.
// post / post_epoch / post_dmax: the stencil kernel's first CTA posts the ARRIVE flags of the push its PREDECESSOR in the stream
// did from inside (k_verlet1's positions, k_rho_f's df): that kernel is complete, its stores are performed. push_df: k_rho_f's
// epilogue stores the new df of band sites into the neighbours' ghosts (kernels.cuh:push_site).
struct LateWait { const unsigned long long *flags; unsigned long long epoch; unsigned int mask; unsigned int *err; long long limit; int fold_dmax;
                  const P2pPeers *post; unsigned long long post_epoch; const unsigned long long *post_dmax; const P2pPeers *push_df;
                  int *low_list, *low_count; int low_cap; };   // LOWLIST variants (eam_fast.cuh:k_low_fix)
__device__ __forceinline__ void post_arrive(const LateWait &lw) {
    if (!lw.post || blockIdx.x != 0 || threadIdx.x >= 27) return;
    const int k = threadIdx.x;
    const P2pPeers *pp = lw.post;
    if (!((pp->mask >> k) & 1u) || *pp->fault) return;
    __threadfence_system();
    if (lw.post_dmax) *(volatile unsigned long long *)(pp->flags[k] + 64 + k) = *lw.post_dmax;   // P2P_DMAX + code, in front of the release
    st_release_sys(pp->flags[k] + 32 + k, lw.post_epoch);                                         // P2P_ARRIVE + code
}
__device__ __forceinline__ unsigned long long late_wait(const LateWait &lw, const int lane, const unsigned long long own_bits = 0) {
    unsigned long long best = own_bits;
    if (lane < 27 && ((lw.mask >> lane) & 1u)) {
        const unsigned long long *w = lw.flags + 32 + lane;   // P2P_ARRIVE + code
        const long long t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
            if (v < lw.epoch && clock64() - t0 > lw.limit) { *(volatile unsigned int *)lw.err = 200u + lane; break; }
        } while (v < lw.epoch);
        if (lw.fold_dmax) { const unsigned long long n = *(volatile const unsigned long long *)(lw.flags + 64 + lane); best = n > best ? n : best; }   // P2P_DMAX + code
    }
    __syncwarp();
    if (lw.fold_dmax)
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, best, o); best = w > best ? w : best; }
    return best;
}

__device__ __forceinline__ int region_unit_to_dev(const Geo &g, const RegionList &rl, const long long u, const int p, const int lane) {
    int b = 0;
    while (b + 1 < rl.n && u >= rl.r[b + 1].u0) b++;
    const Region &r = rl.r[b];
    const long long c = (u - r.u0) * 32 + lane;
    if (c >= (long long)r.nx * r.ny * r.nz) return -1;
    const unsigned cu = (unsigned)c, t = cu / (unsigned)r.nx;
    const int cx = (int)(cu - t * (unsigned)r.nx);
    const int z = (int)(t / (unsigned)r.ny), y = (int)(t - (unsigned)z * (unsigned)r.ny);
    return (int)(p * g.H + ((long long)(z + r.z0 + g.gz) * g.sy + (y + r.y0 + g.gy)) * g.sxc + (cx + r.x0 + g.gx));
}

// ---- device-side choice of the pruned offset list from the largest displacement measured by k_verlet1 in the
//      SAME step (no host round trip between the integrator and the stencil kernels; host twin: pick_list) ------
#define EAM_LEVELS 21
#define EAM_PAIR_LEVELS 41
struct LevelSel {
    const unsigned long long *dmax2_bits;   // null: use the list the host passed (second-generation kernels) / host_level
    const int *levels, *full;               // d_off_levels, d_off_full
    const int *levels_addr, *full_addr;     // the same lists with every (level, parity) segment in ADDRESS order: kernels whose lanes stride
                                            // one atom's offsets (k_force_minor) then gather runs of neighbouring sites, not 32 sectors
    int n[EAM_LEVELS], near_[EAM_LEVELS], ofs[EAM_LEVELS];
    int n_full, near_full;
    double step;                            // 0.01 a
    // third / fourth generation (eam_fast.cuh, eam_sym.cuh): the full list is sorted by site separation behind the near
    // group, so every pruned list is a PREFIX of it, chosen per warp from the displacement levels of its own 32 atoms
    // plus the global maximum (the partner's bound): sites closer than (crf + 0.01 (lw + lg)) a
    const unsigned char *ulev;
    int prefix[EAM_PAIR_LEVELS];
    int host_level;                         // lg when dmax2_bits is null; -1: no pruning (full list)
    // cheaper partner bound: k_verlet1 marks (with the step's epoch byte) the cells within reach of every atom above level
    // hot_T; `edge` marks the cells whose partners may be ghosts (their levels are not kept). A warp that sees neither
    // bounds its partners by hot_T instead of lg
    const unsigned char *hot, *edge;
    const unsigned long long *hot_count;    // atoms that marked (or wanted to) this step; above MARK_CAP the map is incomplete
    int hot_T, hot_epoch;
    long long H;
    // serial path: per-cell partner bound (kernels.cuh:k_pmax_*), takes precedence over the marks
    const unsigned char *pmax;
};
__device__ __forceinline__ void select_list(const LevelSel &ls, const int *&offs, int &n_off, int &n_near) {
    if (!ls.dmax2_bits) return;
    const double d = sqrt(__longlong_as_double((long long)*ls.dmax2_bits)) + 1e-6;
    const int L = (int)ceil(d / ls.step);
    if (L < EAM_LEVELS && ls.n[min(L, EAM_LEVELS - 1)] > 0) { offs = ls.levels + ls.ofs[L]; n_off = ls.n[L]; n_near = ls.near_[L]; }
    else { offs = ls.full; n_off = ls.n_full; n_near = ls.near_full; }
}
// the same choice from the address-ordered copies (same lengths; no near group there: the order is by address)
__device__ __forceinline__ void select_list_addr(const LevelSel &ls, const int *&offs, int &n_off) {
    if (!ls.dmax2_bits) return;
    const double d = sqrt(__longlong_as_double((long long)*ls.dmax2_bits)) + 1e-6;
    const int L = (int)ceil(d / ls.step);
    if (L < EAM_LEVELS && ls.n[min(L, EAM_LEVELS - 1)] > 0) { offs = ls.levels_addr + ls.ofs[L]; n_off = ls.n[L]; }
    else { offs = ls.full_addr; n_off = ls.n_full; }
}
// global displacement level (the bound on the partner atom of any pair)
__device__ __forceinline__ int level_of_bits(const LevelSel &ls, const unsigned long long bits) {
    const double d = sqrt(__longlong_as_double((long long)bits)) + 1e-6;
    return (int)min(ceil(d / ls.step), 1000.0);
}
__device__ __forceinline__ int base_level(const LevelSel &ls) {
    if (!ls.dmax2_bits) return ls.host_level;
    return level_of_bits(ls, *ls.dmax2_bits);
}
// offsets a warp loops: lw = largest displacement level among its own atoms (warp-uniform)
__device__ __forceinline__ bool hot_map_usable(const LevelSel &ls) { return ls.hot && *ls.hot_count <= MARK_CAP; }
// (edge == null: the marks were stamped by ghost atoms too -- k_mark_levels -- so cells next to the shell need no special case)
__device__ __forceinline__ bool cell_hot(const LevelSel &ls, const long long cell) { return ls.hot[cell] == ls.hot_epoch || (ls.edge && ls.edge[cell] != 0); }
__device__ __forceinline__ int list_len(const LevelSel &ls, const int lg, const int lw, const bool hot) {
    const int L = lw + (hot ? lg : min(lg, ls.hot_T));
    return (lg >= 0 && L < EAM_PAIR_LEVELS) ? ls.prefix[L] : ls.n_full;
}
// the prefix a warp loops: lg / hot_ok are the launch's global level and "is the mark map complete"; d = the lane's site, cell
// = its cell (both sub-lattices share the per-cell maps)
__device__ __forceinline__ int warp_list_len(const LevelSel &ls, const int lg, const bool hot_ok, const int d, const long long cell) {
    const int lw = __reduce_max_sync(0xffffffffu, lg >= 0 ? (int)ls.ulev[d] : 0);
    if (ls.pmax) {
        const int L = lw + __reduce_max_sync(0xffffffffu, (int)ls.pmax[cell]);
        return (lg >= 0 && L < EAM_PAIR_LEVELS) ? ls.prefix[L] : ls.n_full;
    }
    return list_len(ls, lg, lw, __any_sync(0xffffffffu, hot_ok ? cell_hot(ls, cell) : true));
}

// ---- neighbour field access: plain global loads (LSU pipe) or texture fetches (TEX pipe of the same L1) -----
struct SoaTex { cudaTextureObject_t x[3], df; };
template <bool TEX> struct Nbr;
template <> struct Nbr<false> {
    const double *__restrict__ x, *__restrict__ y, *__restrict__ z, *__restrict__ df;
    __device__ __forceinline__ Nbr(const Soa &s, const SoaTex &) : x(s.x[0]), y(s.x[1]), z(s.x[2]), df(s.df) {}
    __device__ __forceinline__ double X(int j) const { return x[j]; }
    __device__ __forceinline__ double Y(int j) const { return y[j]; }
    __device__ __forceinline__ double Z(int j) const { return z[j]; }
    __device__ __forceinline__ double DF(int j) const { return df[j]; }
};
template <> struct Nbr<true> {
    cudaTextureObject_t x, y, z, df;
    __device__ __forceinline__ Nbr(const Soa &, const SoaTex &t) : x(t.x[0]), y(t.x[1]), z(t.x[2]), df(t.df) {}
    static __device__ __forceinline__ double f(cudaTextureObject_t t, int j) {
        const int2 v = tex1Dfetch<int2>(t, j);
        return __hiloint2double(v.y, v.x);
    }
    __device__ __forceinline__ double X(int j) const { return f(x, j); }
    __device__ __forceinline__ double Y(int j) const { return f(y, j); }
    __device__ __forceinline__ double Z(int j) const { return f(z, j); }
    __device__ __forceinline__ double DF(int j) const { return f(df, j); }
};

// ---- K1 rho (+ K2 df fused): atom::latRho / latDf (reference src/atom.cpp:151-192,286-309), full-list gather ----
// SINGLE: every valid site has type sp.single and tables 0 (elec) / 1 (phi) of the staging area are its own.
// NOVAC: the census found no vacant site anywhere (ghosts included), so the per-neighbour type test is dropped.
template <bool SINGLE, bool FUSE_DF, bool ACCUM, bool TEX = false, bool NOVAC = false>
__global__ void __launch_bounds__(EAM_THREADS, 1)
k_rho_s(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const SoaTex tex,
        const RegionList rl, const LevelSel ls) {
    const Nbr<TEX> nb(s, tex);
    const int *offs = offs_h;
    int n_off = n_off_h, n_near_unused = 0;
    select_list(ls, offs, n_off, n_near_unused);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t mbar;
    const double2 *s_tab = stage_tables(sp, smem, &mbar, offs, 2 * n_off);
    const int *s_off = reinterpret_cast<const int *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = EAM_THREADS / 32;
    const long long upp = rl.units;
    const uint32_t b_el0 = smem_u32(s_tab) - ((uint32_t)sp.row_lo << 4); // SINGLE: staged slot 0 = elec[single]
    const int maj = sp.staged_id[0];                 // species whose tables are staged (== sp.single when SINGLE)
    const double2 *g_el0 = sp.g_elec[maj];
    const size_t tstride = (size_t)tb.n_r + 1;
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        const int d = region_unit_to_dev(g, rl, u - (u >= upp ? upp : 0), u >= upp, lane);
        if (d < 0) continue;
        const int ti = s.type[d];
        if (ti < 0) {
            if (!ACCUM) s.rho[d] = 0.0;
            continue;
        }
        const int *off = s_off + (u >= upp ? n_off : 0);
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
        double acc = 0.0;
#pragma unroll 2
        for (int q = 0; q < n_off; q++) {
            const int j = d + off[q];
            const int tj = (SINGLE && NOVAC) ? sp.single : s.type[j];
            const double dx = xi - nb.X(j), dy = yi - nb.Y(j), dz = zi - nb.Z(j);
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (tj >= 0 && d2 < g.rc2) {
                const double r = d2 * rsqrt(d2);
                double p;
                const int m = split_index(r, tb.inv_dr, tb.n_r, p);
                const Cubic c = SINGLE ? fetch_cubic_s(b_el0, g_el0, sp.row_lo, m)
                                       : fetch_cubic_m(tj == maj, b_el0, sp.g_elec[0] + (size_t)tj * tstride, sp.row_lo, m);
                acc += cubic_value(c, p);
            }
        }
        if (ACCUM) acc += s.rho[d];
        s.rho[d] = acc;
        if (FUSE_DF) s.df[d] = d_embed(tb, ti, acc);
    }
}

// ---- K3 force: atom::latForce (reference src/atom.cpp:311-358), full-list gather ---------------------------
template <bool SINGLE, bool ACCUM, bool TEX = false, bool NOVAC = false>
__global__ void __launch_bounds__(EAM_THREADS, 1)
k_force_s(const Geo g, const Soa s, const DevTables tb, const StagePlan sp, const int *__restrict__ offs_h, const int n_off_h, const SoaTex tex,
          const RegionList rl, const LevelSel ls) {
    const Nbr<TEX> nb(s, tex);
    const int *offs = offs_h;
    int n_off = n_off_h, n_near_unused = 0;
    select_list(ls, offs, n_off, n_near_unused);
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint64_t mbar;
    const double2 *s_tab = stage_tables(sp, smem, &mbar, offs, 2 * n_off);
    const int *s_off = reinterpret_cast<const int *>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = EAM_THREADS / 32;
    const long long upp = rl.units;
    const int nt = tb.n_types;
    const uint32_t b_el0 = smem_u32(s_tab) - ((uint32_t)sp.row_lo << 4);      // SINGLE: slot 0 = elec[single]
    const uint32_t b_ph0 = b_el0 + ((uint32_t)sp.rows_s << 4);                // SINGLE: slot 1 = phi[single][single]
    const int maj = sp.staged_id[0];                 // species whose tables are staged (== sp.single when SINGLE)
    const double2 *g_el0 = sp.g_elec[maj];
    const double2 *g_ph0 = sp.g_phi[maj * nt + maj];
    const size_t tstride = (size_t)tb.n_r + 1;
    for (long long u = (long long)blockIdx.x * wpc + warp; u < 2 * upp; u += (long long)gridDim.x * wpc) {
        const int d = region_unit_to_dev(g, rl, u - (u >= upp ? upp : 0), u >= upp, lane);
        if (d < 0) continue;
        const int ti = s.type[d];
        if (ti < 0) {
            if (!ACCUM) { s.f[0][d] = 0.0; s.f[1][d] = 0.0; s.f[2][d] = 0.0; }
            continue;
        }
        const int *off = s_off + (u >= upp ? n_off : 0);
        const double xi = s.x[0][d], yi = s.x[1][d], zi = s.x[2][d];
        const double dfi = s.df[d];
        double fx = 0.0, fy = 0.0, fz = 0.0;
#pragma unroll 2
        for (int q = 0; q < n_off; q++) {
            const int j = d + off[q];
            const int tj = (SINGLE && NOVAC) ? sp.single : s.type[j];
            const double dx = xi - nb.X(j), dy = yi - nb.Y(j), dz = zi - nb.Z(j);
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (tj >= 0 && d2 < g.rc2) {
                const double recip = rsqrt(d2);
                const double r = d2 * recip;
                const double dfj = nb.DF(j);
                double p;
                const int m = split_index(r, tb.inv_dr, tb.n_r, p);
                // eam::toForce (oracle/pot.c:pot_to_force): phi = z2/r, phi' = z2'/r - phi/r, fpair = -(phi' + emb)/r with
                // emb = rho'_i(r) df_j + rho'_j(r) df_i; z2' and rho' are slopes per knot times 1/dr, factored out:
                //   fpair = -(1/r) * ( (1/dr) * (z2'_p / r + emb_p) - z2 / r^2 )
                double z2, z2p, emb;
                if (SINGLE) {
                    const Cubic cp = fetch_cubic_s(b_ph0, g_ph0, sp.row_lo, m);
                    const Cubic ce = fetch_cubic_s(b_el0, g_el0, sp.row_lo, m);
                    z2 = cubic_value(cp, p);
                    z2p = cubic_slope(cp, p);
                    emb = cubic_slope(ce, p) * (dfi + dfj);
                } else {
                    const bool mi = ti == maj, mj = tj == maj;
                    const Cubic cp = fetch_cubic_m(mi && mj, b_ph0, sp.g_phi[0] + (size_t)(ti * nt + tj) * tstride, sp.row_lo, m);
                    const Cubic ci = fetch_cubic_m(mi, b_el0, sp.g_elec[0] + (size_t)ti * tstride, sp.row_lo, m);
                    z2 = cubic_value(cp, p);
                    z2p = cubic_slope(cp, p);
                    const double rho_p_from = cubic_slope(ci, p);
                    double rho_p_to = rho_p_from;
                    if (tj != ti) {
                        const Cubic cj = fetch_cubic_m(mj, b_el0, sp.g_elec[0] + (size_t)tj * tstride, sp.row_lo, m);
                        rho_p_to = cubic_slope(cj, p);
                    }
                    emb = rho_p_from * dfj + rho_p_to * dfi;
                }
                const double t1 = fma(z2p, recip, emb);
                const double fp = -recip * fma(tb.inv_dr, t1, -(z2 * (recip * recip)));
                fx += dx * fp; fy += dy * fp; fz += dz * fp;
            }
        }
        if (ACCUM) { fx += s.f[0][d]; fy += s.f[1][d]; fz += s.f[2][d]; }
        s.f[0][d] = fx; s.f[1][d] = fy; s.f[2][d] = fz;
    }
}

// ---- species census (decides SINGLE vs multi and what to stage) ------------------------------------------
__global__ void __launch_bounds__(MISA_BLOCK) k_census(const long long n, const int8_t *__restrict__ type, unsigned long long *__restrict__ count) {
    __shared__ unsigned int sh[MISA_MAX_TYPES];
    if (threadIdx.x < MISA_MAX_TYPES) sh[threadIdx.x] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int t = type[i];
        if (t >= 0 && t < MISA_MAX_TYPES) atomicAdd(&sh[t], 1u);
    }
    __syncthreads();
    if (threadIdx.x < MISA_MAX_TYPES && sh[threadIdx.x]) atomicAdd(&count[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}
