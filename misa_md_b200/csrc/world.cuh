// misa_md_b200/csrc/world.cuh -- initial state built ON THE DEVICE, by global atom id,
// and the deterministic reductions behind the global thermo / rescale entry points (8f-3, 8f-4).
//
// Replaces WorldBuilder::build (reference src/world_builder.cpp:64-199): createPhaseSpace (:105-131: ids,
// species, perfect bcc positions, velocities (md_rand::random() - 0.5) / mass from std::mt19937(seed) scaled by
// 1/0xFFFFFFFF, src/utils/random/random.cpp:30-37), vcm + zeroMomentum (:75-90,143-178) and
// configuration::rescale (src/system_configuration.cpp:86-111).
//
// The reference draws per rank in sub-box order, so its initial state depends on the process grid (SURVEY.md
// section 8c). Here the mt19937 stream is indexed by GLOBAL id -- draw 3(id-1)+k is velocity component k of atom
// id -- which is exactly what the reference produces on ONE rank, and every sub-box cuts its part out of that one
// global state. mt19937 is sequential by construction; one CTA advances the 624-word state (three dependent
// phases per twist, double-buffered in shared memory) and streams the tempered words to HBM at a few ns per word;
// everything else is one pass over the ids.
// Species for alloys: the reference calls unseeded libc rand() % total (world_builder.cpp:180-199), which has no
// reproducible stream; a counter-based hash of (alloy_seed, id) feeds the same cumulative-ratio rule (host mirror:
// misa_md_b200/synth.py:species_by_id).
#pragma once
#include "kernels.cuh"

#define MT_N 624
#define MT_M 397

__device__ __forceinline__ unsigned mt_twist(const unsigned cur, const unsigned next, const unsigned far) {
    const unsigned y = (cur & 0x80000000u) | (next & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ unsigned mt_temper(unsigned y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
// std::mt19937(seed): n_draws successive outputs into out[]. ONE block of 256 threads.
__global__ void __launch_bounds__(256) k_mt19937_stream(const unsigned seed, const long long n_draws, unsigned *__restrict__ out) {
    __shared__ unsigned st[2][MT_N];
    if (threadIdx.x == 0) {
        unsigned v = seed;
        st[0][0] = v;
        for (int i = 1; i < MT_N; i++) { v = 1812433253u * (v ^ (v >> 30)) + (unsigned)i; st[0][i] = v; }
    }
    __syncthreads();
    int cur = 0;
    const int t = threadIdx.x;
    for (long long base = 0; base < n_draws; base += MT_N) {
        const unsigned *o = st[cur];
        unsigned *n = st[cur ^ 1];
        // phase 1: i in [0, 227) reads old[i], old[i+1], old[i+397]
        if (t < MT_N - MT_M) {
            const unsigned w = mt_twist(o[t], o[t + 1], o[t + MT_M]);
            n[t] = w;
            if (base + t < n_draws) out[base + t] = mt_temper(w);
        }
        __syncthreads();
        // phase 2: i in [227, 454) reads old[i], old[i+1], new[i-227]
        if (t < MT_N - MT_M) {
            const int i = t + (MT_N - MT_M);
            const unsigned w = mt_twist(o[i], o[i + 1], n[t]);
            n[i] = w;
            if (base + i < n_draws) out[base + i] = mt_temper(w);
        }
        __syncthreads();
        // phase 3: i in [454, 624) reads old[i], old[i+1] (new[0] for the last word), new[i-227]
        if (t < MT_N - 2 * (MT_N - MT_M)) {
            const int i = t + 2 * (MT_N - MT_M);
            const unsigned w = mt_twist(o[i], i == MT_N - 1 ? n[0] : o[i + 1], n[i - (MT_N - MT_M)]);
            n[i] = w;
            if (base + i < n_draws) out[base + i] = mt_temper(w);
        }
        __syncthreads();
        cur ^= 1;
    }
}

struct WorldPar {
    long long px, py, pz;          // phase space (cells)
    long long n_global;            // 2 px py pz
    int ratio[MISA_MAX_TYPES];
    int ratio_total, single;       // single >= 0: only that species has a non-zero ratio
    unsigned long long alloy_seed;
    double mass[MISA_MAX_TYPES];
    double a;
};

// species of global id (1-based): splitmix64 finaliser of alloy_seed + id * golden ratio; the top 31 bits stand in
// for libc rand(); then WorldBuilder::randomAtomsType's cumulative rule (reference src/world_builder.cpp:180-199)
__host__ __device__ __forceinline__ int species_by_id(const WorldPar &w, const unsigned long long id) {
    if (w.single >= 0) return w.single;
    unsigned long long z = w.alloy_seed + id * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const int draw = (int)((z >> 33) % (unsigned long long)w.ratio_total);
    int acc = 0;
    for (int i = 0; i < MISA_MAX_TYPES; i++) {
        acc += w.ratio[i];
        if (draw < acc) return i;
    }
    return 0;
}
// (md_rand::random() - 0.5) / mass, reference src/world_builder.cpp:125-127, src/utils/random/random.cpp:30-37
__device__ __forceinline__ double raw_velocity(const unsigned word, const double mass) {
    const double u = __dmul_rn((double)word, 1.0 / 4294967295.0);
    return __ddiv_rn(__dadd_rn(u, -0.5), mass);
}

// deterministic block reduction of NV values per thread -> partial[blockIdx][NV] (fixed tree, no atomics)
template <int NV>
__device__ __forceinline__ void block_partials(double (&v)[NV], double *__restrict__ partial) {
    __shared__ double sh[NV][MISA_BLOCK / 32];
    for (int k = 0; k < NV; k++) {
        double x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0.0;
        for (int w = 0; w < MISA_BLOCK / 32; w++) x += sh[threadIdx.x][w];
        partial[(size_t)blockIdx.x * NV + threadIdx.x] = x;
    }
}
// out[k] = sum over blocks of partial[b][k], fixed order: one block, thread t strides the blocks, then a tree
template <int NV>
__global__ void __launch_bounds__(MISA_BLOCK) k_sum_partials(const double *__restrict__ partial, const int n_blocks, double *__restrict__ out) {
    double v[NV];
    for (int k = 0; k < NV; k++) v[k] = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += MISA_BLOCK)
        for (int k = 0; k < NV; k++) v[k] += partial[(size_t)b * NV + k];
    __shared__ double res[NV];
    block_partials<NV>(v, res); // blockIdx.x == 0
    __syncthreads();
    if (threadIdx.x < NV) out[threadIdx.x] = res[threadIdx.x];
}

// WorldBuilder::vcm over the GLOBAL box (reference src/world_builder.cpp:160-178): p[0..2] = sum v m, p[3] = sum m
__global__ void __launch_bounds__(MISA_BLOCK) k_world_moments(const WorldPar w, const unsigned *__restrict__ draws, double *__restrict__ partial) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long gidx = (long long)blockIdx.x * MISA_BLOCK + threadIdx.x; gidx < w.n_global; gidx += (long long)gridDim.x * MISA_BLOCK) {
        const double m = w.mass[species_by_id(w, (unsigned long long)gidx + 1ull)];
        for (int k = 0; k < 3; k++) v[k] += __dmul_rn(raw_velocity(draws[3 * gidx + k], m), m);
        v[3] += m;
    }
    block_partials<4>(v, partial);
}
// configuration::mvv (reference src/system_configuration.cpp:61-84) of the zero-momentum velocities, GLOBAL box
__global__ void __launch_bounds__(MISA_BLOCK) k_world_mvv(const WorldPar w, const unsigned *__restrict__ draws, const double vcm0, const double vcm1,
                                                          const double vcm2, double *__restrict__ partial) {
    double e[1] = {0.0};
    const double vcm[3] = {vcm0, vcm1, vcm2};
    for (long long gidx = (long long)blockIdx.x * MISA_BLOCK + threadIdx.x; gidx < w.n_global; gidx += (long long)gridDim.x * MISA_BLOCK) {
        const double m = w.mass[species_by_id(w, (unsigned long long)gidx + 1ull)];
        double vv[3];
        for (int k = 0; k < 3; k++) vv[k] = __dadd_rn(raw_velocity(draws[3 * gidx + k], m), -__ddiv_rn(vcm[k], m)); // zeroMomentum :143-158
        e[0] += __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(vv[0], vv[0]), __dmul_rn(vv[1], vv[1])), __dmul_rn(vv[2], vv[2])), m);
    }
    block_partials<1>(e, partial);
}
// createPhaseSpace + zeroMomentum + rescale for every site of THIS sub-box's ghost-extended array: owned sites get
// id / species / ideal position / velocity by global id, ghost sites are INVALID placeholders until the first
// exchange (reference src/atom/atom_list.cpp:25-49); f, rho, df start at zero
__global__ void __launch_bounds__(MISA_BLOCK) k_world_fill(const Geo g, const Soa s, const WorldPar w, const unsigned *__restrict__ draws,
                                                           const double vcm0, const double vcm1, const double vcm2, const double factor) {
    const long long d = (long long)blockIdx.x * MISA_BLOCK + threadIdx.x;
    if (d >= g.n_ext) return;
    const int p = d >= g.H;
    const long long rem = d - (long long)p * g.H;
    const int cx = (int)(rem % g.sxc);
    const long long r = rem / g.sxc;
    const int y = (int)(r % g.sy), z = (int)(r / g.sy);
    const bool own = cx >= g.gx && cx < g.gx + g.nx && y >= g.gy && y < g.gy + g.ny && z >= g.gz && z < g.gz + g.nz;
    double x[3] = {0.0, 0.0, 0.0}, v[3] = {0.0, 0.0, 0.0};
    unsigned long long id = 0;
    int t = -1;
    if (own) {
        const long long gi = 2LL * (cx - g.gx + g.lo[0]) + p, gj = y - g.gy + g.lo[1], gk = z - g.gz + g.lo[2]; // global doubled-x coordinates
        const long long gidx = (gk * w.py + gj) * (2 * w.px) + gi;
        id = (unsigned long long)gidx + 1ull;
        t = species_by_id(w, id);
        const double m = w.mass[t], half = __ddiv_rn(w.a, 2.0), odd = (double)p;
        x[0] = __dmul_rn(__dmul_rn((double)gi, 0.5), w.a);                      // world_builder.cpp:120
        x[1] = __dadd_rn(__dmul_rn((double)gj, w.a), __dmul_rn(odd, half));     // :121-122
        x[2] = __dadd_rn(__dmul_rn((double)gk, w.a), __dmul_rn(odd, half));     // :123-124
        const double vcm[3] = {vcm0, vcm1, vcm2};
        for (int k = 0; k < 3; k++) {
            const double vz = __dadd_rn(raw_velocity(draws[3 * gidx + k], m), -__ddiv_rn(vcm[k], m));
            v[k] = __dmul_rn(vz, factor);                                       // configuration::rescale :100-104
        }
    }
    s.id[d] = id;
    s.type[d] = (int8_t)t;
    for (int k = 0; k < 3; k++) { s.x[k][d] = x[k]; s.v[k][d] = v[k]; s.f[k][d] = 0.0; }
    s.rho[d] = 0.0;
    s.df[d] = 0.0;
}

// configuration::mvv of this sub-box's owned valid sites (reference src/system_configuration.cpp:61-72), partials
__global__ void __launch_bounds__(MISA_BLOCK) k_mvv(const Geo g, const Soa s, const double m0, const double m1, const double m2,
                                                    const int blocks_per_parity, double *__restrict__ partial) {
    const int p = blockIdx.x >= blocks_per_parity;
    const long long c = (long long)(blockIdx.x - p * blocks_per_parity) * MISA_BLOCK + threadIdx.x;
    double e[2] = {0.0, 0.0};
    if (c < g.n_cells_owned) {
        int cx, y, z;
        const int d = owned_cell_to_dev(g, p, c, cx, y, z);
        const int ti = s.type[d];
        if (ti >= 0) {
            const double m = ti == 0 ? m0 : (ti == 1 ? m1 : m2);
            const double vx = s.v[0][d], vy = s.v[1][d], vz = s.v[2][d];
            e[0] = __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)), m);
            e[1] = 1.0;
        }
    }
    block_partials<2>(e, partial);
}
