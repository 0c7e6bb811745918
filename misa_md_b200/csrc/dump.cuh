// misa_md_b200/csrc/dump.cuh -- device-side compaction of the dump record stream.
//
// Replaces, for resident mode, AtomDump::dump + BufferedFileWriter::write (reference frontend/io/atom_dump.cpp:
// 39-75, frontend/io/buffered_io.cpp:18-36): the host loop that walks the owned region in z,y,x order, skips INVALID
// sites and copies {id, step, type, inter_type = 0, x, v} into 72-byte atom_dump::AtomInfoDump records
// (frontend/io/atom_info_dump.h:14-22). Here the valid sites are compacted IN THE SAME ORDER on the device
// (count -> scan -> write) and only n_valid * 72 bytes cross PCIe, instead of the 104-byte ghost-extended array.
// HBM-bound integer/byte work: each tile of DUMP_TILE sites stages its records in shared memory and stores them as
// one contiguous run of 8-byte words.
#pragma once
#include "kernels.cuh"

#define DUMP_TILE 256
#define DUMP_WORDS 9  // 72-byte record = 9 x 8 bytes

struct DumpRegion {   // [begin, end) in ghost-inclusive doubled-x coordinates (AtomList::getAtomEleByGhostIndex)
    int b0, b1, b2;
    int nx, ny, nz;
    long long n;
};

__device__ __forceinline__ long long dump_unit_to_dev(const Geo &g, const DumpRegion &r, const long long u) {
    const int i = (int)(u % r.nx);
    const long long t = u / r.nx;
    const int j = (int)(t % r.ny), k = (int)(t / r.ny);
    const long long idx = ((long long)(r.b2 + k) * g.sy + (r.b1 + j)) * (2LL * g.sxc) + (r.b0 + i);
    return ref_to_dev(idx, g.H);
}

// valid sites per tile
__global__ void __launch_bounds__(DUMP_TILE) k_dump_count(const Geo g, const DumpRegion r, const int8_t *__restrict__ type,
                                                          unsigned *__restrict__ tile_count) {
    const long long u = (long long)blockIdx.x * DUMP_TILE + threadIdx.x;
    const bool valid = u < r.n && type[dump_unit_to_dev(g, r, u)] >= 0;
    const int n = __syncthreads_count(valid);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = (unsigned)n;
}

// exclusive scan of the tile counts by ONE block (n_tiles is a few thousand); total -> *total
__global__ void __launch_bounds__(1024) k_dump_scan(const unsigned *__restrict__ tile_count, unsigned long long *__restrict__ tile_base,
                                                    const int n_tiles, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_count[i] : 0;
        unsigned long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sum[lane], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sum[lane] = wi - w; // exclusive prefix of the warp sums
        }
        __syncthreads();
        const unsigned long long excl = carry + warp_sum[warp] + incl - v;
        if (i < n_tiles) tile_base[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// records of one tile, in reference order, staged in shared memory and stored as one contiguous run
__global__ void __launch_bounds__(DUMP_TILE) k_dump_write(const Geo g, const DumpRegion r, const Soa s, const unsigned long long *__restrict__ tile_base,
                                                          const unsigned long long time_step, unsigned long long *__restrict__ out) {
    __shared__ unsigned long long rec[DUMP_TILE * DUMP_WORDS];
    __shared__ int warp_cnt[DUMP_TILE / 32];
    const long long u = (long long)blockIdx.x * DUMP_TILE + threadIdx.x;
    long long d = 0;
    int t = -1;
    if (u < r.n) {
        d = dump_unit_to_dev(g, r, u);
        t = s.type[d];
    }
    const bool valid = t >= 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < DUMP_TILE / 32; w++) {
        const int c = warp_cnt[w];
        if (w < warp) before += c;
        total += c;
    }
    if (valid) {
        unsigned long long *w = rec + (size_t)(before + __popc(ballot & ((1u << lane) - 1u))) * DUMP_WORDS;
        w[0] = s.id[d];
        w[1] = time_step;
        w[2] = (unsigned long long)(unsigned)t; // type (int) | inter_type = 0 (short) | 2 padding bytes = 0
        w[3] = __double_as_longlong(s.x[0][d]); w[4] = __double_as_longlong(s.x[1][d]); w[5] = __double_as_longlong(s.x[2][d]);
        w[6] = __double_as_longlong(s.v[0][d]); w[7] = __double_as_longlong(s.v[1][d]); w[8] = __double_as_longlong(s.v[2][d]);
    }
    __syncthreads();
    unsigned long long *dst = out + tile_base[blockIdx.x] * DUMP_WORDS;
    for (int q = threadIdx.x; q < total * DUMP_WORDS; q += DUMP_TILE) dst[q] = rec[q];
}
