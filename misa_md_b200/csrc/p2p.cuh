// misa_md_b200/csrc/p2p.cuh -- ghost exchange as ONE push kernel over NVLink peer memory.
//
// The reference moves ghosts in three dependent stages (x, then y, then z: comm::neiSendReceive with LatPacker /
// DfEmbedPacker, src/atom/atom_list.cpp:51-57, src/pack/lat_particle_packer.cpp:97-139); the stage order is what
// carries edge and corner sites to diagonal neighbours. Over NCCL that is 3 x (pack kernel, grouped send/recv,
// unpack kernel) per exchange -- 0.19 ms for positions at 100^3 cells per GPU on a 2x2x2 grid, all latency.
// On one NVSwitch box every GPU can store into every other GPU's HBM, so the composition of the three stages
// is applied directly: each ghost site of a sub-box has exactly ONE owned source site in one of the 26 surrounding
// sub-boxes (misa_b200_create composes the staged send/recv lists into that map; all sub-boxes have the same
// shape, so the map of "what I receive" read backwards is "what I push"). k_p2p_push_* reads the owned source,
// adds the periodic image shift the staged path would have added (same rounded add, same bits) and stores straight
// into the destination's ghost site through an IPC-mapped pointer. Two flag handshakes of system-scope
// release/acquire stores bracket the push:
//   ready : receiver -> origin, "everything I enqueued before this exchange (the kernels that read my ghosts) is done"
//   arrive: origin -> receiver, "my stores into your ghosts are performed"
// Flags carry the exchange epoch (monotonic), so nothing is ever reset. No data-path collective, no staging buffer.
#pragma once
#include <unistd.h>
#include "ctx.h"
#include "util.cuh"
#include "nccl_dl.cuh"

#define P2P_READY 0
#define P2P_ARRIVE 32
#define P2P_DMAX 64          // [64, 91): the origin's largest squared displacement of the step (bit pattern), written before ARRIVE
#define P2P_FLAG_WORDS 128
// A peer that never answers ends the wait with an error, not a hang: P2pPeers::spin_limit SM clocks (default 30 s; option
// "p2p_timeout_s" / MISA_B200_OPTS). A push kernel whose wait for READY timed out stores NOTHING into the neighbours'
// ghosts and does not signal ARRIVE -- the neighbour may still be reading them -- so rank skew beyond the limit is an
// error on every rank concerned, never a data race.


// definition of kernels.cuh:push_site (P2pPeers is complete here)
template <int NF>
__device__ __noinline__ void push_site(const P2pPeers *__restrict__ pp, const int d, const int cx, const int y, const int z, const int nx, const int ny,
                                       const int nz, const int gx, const int gy, const int gz, const int sxc, const int sy, const int field0,
                                       const double v0, const double v1, const double v2) {
    // per dimension: bit 0 = low band (s = +1), bit 1 = high band (s = -1); s = 0 always
    const int bx = (cx < gx ? 1 : 0) | (cx >= nx - gx ? 2 : 0), by = (y < gy ? 1 : 0) | (y >= ny - gy ? 2 : 0), bz = (z < gz ? 1 : 0) | (z >= nz - gz ? 2 : 0);
    const long long stride = pp->stride;
    if (*pp->fault) return;   // the gate in front of this kernel gave up on a neighbour's READY: store nothing (p2p_check reports it)
    for (int iz = 0; iz < 3; iz++) {
        if (iz && !((bz >> (iz - 1)) & 1)) continue;
        const int sz_ = iz == 0 ? 0 : (iz == 1 ? 1 : -1);
        for (int iy = 0; iy < 3; iy++) {
            if (iy && !((by >> (iy - 1)) & 1)) continue;
            const int sy_ = iy == 0 ? 0 : (iy == 1 ? 1 : -1);
            for (int ix = 0; ix < 3; ix++) {
                if (ix && !((bx >> (ix - 1)) & 1)) continue;
                const int sx_ = ix == 0 ? 0 : (ix == 1 ? 1 : -1);
                if (!(sx_ | sy_ | sz_)) continue;
                const int k = (sx_ + 1) + 3 * (sy_ + 1) + 9 * (sz_ + 1);
                const long long b = (long long)d + sx_ * nx + ((long long)sz_ * nz * sy + (long long)sy_ * ny) * sxc;
                double *P = pp->xyzd[k] + (long long)field0 * stride + b;
                if (NF == 3) {
                    double x = v0, yy = v1, zz = v2;
                    if (pp->shift[k][0] != 0.0) x = __dadd_rn(x, pp->shift[k][0]);
                    if (pp->shift[k][1] != 0.0) yy = __dadd_rn(yy, pp->shift[k][1]);
                    if (pp->shift[k][2] != 0.0) zz = __dadd_rn(zz, pp->shift[k][2]);
                    P[0] = x; P[stride] = yy; P[2 * stride] = zz;
                } else {
                    P[0] = v0;
                }
            }
        }
    }
}
template __device__ void push_site<3>(const P2pPeers *, int, int, int, int, int, int, int, int, int, int, int, int, int, double, double, double);
template __device__ void push_site<1>(const P2pPeers *, int, int, int, int, int, int, int, int, int, int, int, int, int, double, double, double);

// code k = (sx+1) + 3 (sy+1) + 9 (sz+1): the ORIGIN of my ghosts with that code sits at sub-box offset +s from me; I push my
// own group k to the sub-box at offset -s. Thread k handles direction k.
__device__ __forceinline__ bool p2p_wait(const unsigned long long *w, const unsigned long long epoch, unsigned int *err, const unsigned what,
                                         const long long limit) {
    const long long t0 = clock64();
    while (ld_acquire_sys(w) < epoch)
        if (clock64() - t0 > limit) { *(volatile unsigned int *)err = what; return false; }
    return true;
}
// READY posted for `epoch_post` (0: nothing to post), then wait until every destination has freed its ghosts up to `epoch_wait`:
// the gate in front of a producing kernel that pushes from inside (k_verlet1 with VerletPar::push, k_rho_f with a df push)
__global__ void k_p2p_ready_gate(const P2pPeers pp, const unsigned long long epoch_post, const unsigned long long epoch_wait,
                                 const unsigned long long *__restrict__ my_flags, unsigned int *__restrict__ err) {
    const int k = threadIdx.x;
    if (k >= 27 || !((pp.mask >> k) & 1u)) return;
    if (epoch_post) { __threadfence_system(); st_release_sys(pp.flags[26 - k] + P2P_READY + k, epoch_post); }
    if (!p2p_wait(my_flags + P2P_READY + k, epoch_wait, err, 1u + k, pp.spin_limit)) *pp.fault = 1u;
}

// READY: I am receiver for code k -> tell its origin, which is my destination for code 26 - k, that everything enqueued
// before this kernel (the readers of my ghosts) is done. `epoch` may lie ahead: "free up to and including exchange epoch".
__global__ void k_p2p_ready(const P2pPeers pp, const unsigned long long epoch) {
    const int k = threadIdx.x;
    if (k >= 27 || !((pp.mask >> k) & 1u)) return;
    __threadfence_system();
    st_release_sys(pp.flags[26 - k] + P2P_READY + k, epoch);
}
// fence_mode 2: the ARRIVE flags as a kernel of their own, stream-ordered behind the push (its stores are complete at the
// kernel boundary)
__global__ void k_p2p_arrive(const P2pPeers pp, const unsigned long long epoch, const unsigned int *__restrict__ err, const unsigned long long *__restrict__ dmax2) {
    const int k = threadIdx.x;
    if (*(volatile const unsigned int *)err != 0) return;
    __threadfence_system();
    if (k < 27 && ((pp.mask >> k) & 1u)) {
        if (dmax2) *(volatile unsigned long long *)(pp.flags[k] + P2P_DMAX + k) = *dmax2;
        st_release_sys(pp.flags[k] + P2P_ARRIVE + k, epoch);
    }
}
// wait until every origin's stores into my ghosts are performed
// fold != null: also leave max(own word, the neighbours' P2P_DMAX words) in fold[1] (bit patterns of non-negative doubles
// order like integers) -- the partner bound for a stencil launch that does not wait inside the kernel
__global__ void k_p2p_wait_arrive(const P2pPeers pp, const unsigned long long epoch, const unsigned long long *__restrict__ my_flags,
                                  unsigned int *__restrict__ err, const unsigned long long *__restrict__ own_dmax2, unsigned long long *__restrict__ fold) {
    const int k = threadIdx.x;
    unsigned long long v = 0;
    if (k < 27 && ((pp.mask >> k) & 1u)) {
        p2p_wait(my_flags + P2P_ARRIVE + k, epoch, err, 100u + k, pp.spin_limit);
        if (fold) v = *(volatile const unsigned long long *)(my_flags + P2P_DMAX + k);
    }
    if (!fold) return;
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    if (k == 0) fold[1] = v > *own_dmax2 ? v : *own_dmax2;
}
// head of a push kernel: every destination has freed its ghosts; tail: the LAST CTA to finish tells every destination
// ---- optional phase timing of the push kernels (option "p2p_debug"; off: dbg == nullptr): %globaltimer stamps, folded over the
//      CTAs with atomics, accumulated by the last CTA. dbg[0] min start, [1] max "READY seen", [2] max "stores issued",
//      [3..6] sums of (head wait, body, tail, whole) in ns, [7] launches -------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// returns false (for the whole CTA) when a destination never freed its ghosts: the caller must not store into them
__device__ __forceinline__ bool p2p_push_head(const P2pPeers &pp, const unsigned long long epoch, const unsigned long long *my_flags, unsigned int *err) {
    __shared__ int timed_out;
    const int k = threadIdx.x;
    if (k == 0) { timed_out = 0; if (pp.dbg) atomicMin(pp.dbg, gtime()); }
    __syncthreads();
    if (k < 27 && ((pp.mask >> k) & 1u) && !p2p_wait(my_flags + P2P_READY + k, epoch, err, 1u + k, pp.spin_limit)) timed_out = 1;
    __syncthreads();
    if (k == 0 && pp.dbg) atomicMax(pp.dbg + 1, gtime());
    return timed_out == 0;
}
// `dmax2`: when non-null, the word k_verlet1 of this step left (this sub-box's largest squared displacement): the last CTA hands
// it to every destination in front of the ARRIVE flag, so that a consumer that has acquired ARRIVE also holds its neighbours'
// maxima -- the partner bound of the stencil pruning without an all-reduce in front of the stencil kernels.
__device__ __forceinline__ void p2p_push_tail(const P2pPeers &pp, const unsigned long long epoch, unsigned int *done, const unsigned int *err,
                                              const unsigned long long *dmax2 = nullptr) {
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (pp.dbg) atomicMax(pp.dbg + 2, gtime());
        // This CTA's stores happen before its ticket (barrier above, release below), every ticket happens before the last CTA's
        // flag stores (its acquire), and those are system-scope releases: causality order carries all CTAs' stores to whoever
        // acquires ARRIVE. fence_mode 0 makes every CTA's fence a system-scope one (the round-1 protocol; measured: the tail of
        // the kernel then takes 25 us -- 300 concurrent MEMBAR.SYS), 1 keeps them at device scope.
        if (pp.fence_mode == 0) __threadfence_system(); else __threadfence();
        last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    const int k = threadIdx.x;
    if (k == 0) *done = 0;                           // next exchange (stream-ordered after this kernel)
    if (pp.fence_mode == 2) return;                  // ARRIVE is posted by k_p2p_arrive, behind the kernel boundary
    __threadfence_system();
    if (*(volatile const unsigned int *)err != 0) return;   // some CTA gave up on READY: its part was not stored, nothing ARRIVEd
    if (k < 27 && ((pp.mask >> k) & 1u)) {
        if (dmax2) *(volatile unsigned long long *)(pp.flags[k] + P2P_DMAX + k) = *dmax2;   // ordered before the release below
        st_release_sys(pp.flags[k] + P2P_ARRIVE + k, epoch);
    }
    if (pp.dbg) {
        __syncthreads();
        if (k == 0) {
            const unsigned long long t3 = gtime(), t0 = pp.dbg[0], t1 = pp.dbg[1], t2 = pp.dbg[2];
            pp.dbg[3] += t1 - t0; pp.dbg[4] += t2 - t1; pp.dbg[5] += t3 - t2; pp.dbg[6] += t3 - t0; pp.dbg[7] += 1;
            pp.dbg[0] = ~0ULL; pp.dbg[1] = 0; pp.dbg[2] = 0;
        }
    }
}

// Persistent grid (at most two CTAs per SM, grid-stride over the map): every CTA pays one READY check (system-scope acquire
// loads) in front and one system-scope fence + ticket behind its stores -- with one CTA per 256 sites (1 500 CTAs at 100^3 cells)
// those fixed costs, not the 12 MB of stores, were the 35-55 us the exchange took on two GPUs (profiles/r02e_*, r02f_*).
// TYPES: also push the species byte. The sync-free step leaves it out: it only runs while nothing is off-lattice anywhere, so no
// site changes its occupant between two of its exchanges (a run-away sends every sub-box through the serial path, whose
// exchange carries the bytes) -- and single-byte stores into a neighbour's HBM are the slowest part of the push.
#define P2P_PUSH_THREADS 512
template <bool TYPES>
__global__ void __launch_bounds__(P2P_PUSH_THREADS)
k_p2p_push_x(const P2pPeers pp, const int n, const int *__restrict__ dst, const int *__restrict__ src, const int8_t *__restrict__ code, const Soa s,
             const unsigned long long epoch, unsigned long long *__restrict__ my_flags, unsigned int *__restrict__ done, unsigned int *__restrict__ err,
             const unsigned long long *__restrict__ dmax2) {
    const bool ok = p2p_push_head(pp, epoch, my_flags, err);
    if (ok)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const int a = src[i], b = dst[i], k = code[i];
            double x = s.x[0][a], y = s.x[1][a], z = s.x[2][a];
            if (pp.shift[k][0] != 0.0) x = __dadd_rn(x, pp.shift[k][0]);
            if (pp.shift[k][1] != 0.0) y = __dadd_rn(y, pp.shift[k][1]);
            if (pp.shift[k][2] != 0.0) z = __dadd_rn(z, pp.shift[k][2]);
            double *P = pp.xyzd[k];
            P[b] = x; P[pp.stride + b] = y; P[2 * pp.stride + b] = z;
            if (TYPES) pp.type[k][b] = s.type[a];
        }
    p2p_push_tail(pp, epoch, done, err, dmax2);
}
__global__ void __launch_bounds__(P2P_PUSH_THREADS)
k_p2p_push_df(const P2pPeers pp, const int n, const int *__restrict__ dst, const int *__restrict__ src, const int8_t *__restrict__ code, const Soa s,
              const unsigned long long epoch, unsigned long long *__restrict__ my_flags, unsigned int *__restrict__ done, unsigned int *__restrict__ err) {
    const bool ok = p2p_push_head(pp, epoch, my_flags, err);
    if (ok)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) pp.xyzd[code[i]][3 * pp.stride + dst[i]] = s.df[src[i]];
    p2p_push_tail(pp, epoch, done, err);
}

// ---- set-up: exchange IPC handles of the position/df block, the type array and the flag words over the NCCL
//      communicator, map the (at most 26) surrounding sub-boxes, agree that EVERY rank succeeded --------------------
struct P2pBlob {
    int coord[3], size[3], device, ok;
    long long pid;
    unsigned long long host;
    cudaIpcMemHandle_t h_xyzd, h_type, h_flags;
    void *p_xyzd, *p_type, *p_flags;
};
static void push_shift(const misa_b200_domain *dom, int code, double shift[3]);   // misa_b200.cu
static unsigned long long p2p_host_hash() {
    char name[256] = {0};
    gethostname(name, sizeof name - 1);
    unsigned long long h = 1469598103934665603ULL;
    for (const char *p = name; *p; p++) h = (h ^ (unsigned char)*p) * 1099511628211ULL;
    // the boot id separates containers that share a hostname
    if (FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r")) {
        int ch;
        while ((ch = fgetc(f)) != EOF) h = (h ^ (unsigned char)ch) * 1099511628211ULL;
        fclose(f);
    }
    return h;
}
static void p2p_release(misa_b200_ctx *c) {
    for (void *p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    c->p2p_opened.clear();
    c->p2p_active = false;
}
static int p2p_setup(misa_b200_ctx *c) {
    c->p2p_active = false;
    if (!c->opt_p2p || c->comm_size <= 1 || !c->nccl_comm || c->n_push <= 0 || !g_nccl.AllGather) return 0;
    const int n = c->comm_size, me = c->comm_rank;
    if (!c->d_flags) {
        TRY(dmalloc(&c->d_flags, P2P_FLAG_WORDS));
        CU(cudaMemset(c->d_flags, 0, P2P_FLAG_WORDS * sizeof(unsigned long long)));
        CU(cudaHostAlloc((void **)&c->h_p2p_err, sizeof(unsigned int), cudaHostAllocMapped));
        *c->h_p2p_err = 0;
        CU(cudaHostGetDevicePointer((void **)&c->d_p2p_err, c->h_p2p_err, 0));
    }
    P2pBlob mine;
    memset(&mine, 0, sizeof mine);
    for (int k = 0; k < 3; k++) { mine.coord[k] = c->dom.grid_coord[k]; mine.size[k] = c->dom.sub_box_lattice_size[k]; }
    mine.device = c->device; mine.pid = (long long)getpid(); mine.host = p2p_host_hash(); mine.ok = 1;
    mine.p_xyzd = c->d_xyzd; mine.p_type = c->s.type; mine.p_flags = c->d_flags;
    if (cudaIpcGetMemHandle(&mine.h_xyzd, c->d_xyzd) != cudaSuccess || cudaIpcGetMemHandle(&mine.h_type, c->s.type) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.h_flags, c->d_flags) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    unsigned char *d_all = nullptr;
    TRY(dmalloc(&d_all, (size_t)n * sizeof(P2pBlob)));
    CU(cudaMemcpyAsync(d_all + (size_t)me * sizeof(P2pBlob), &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(d_all + (size_t)me * sizeof(P2pBlob), d_all, sizeof(P2pBlob), kNcclInt8, c->nccl_comm, c->stream));
    std::vector<P2pBlob> all(n);
    CU(cudaMemcpyAsync(all.data(), d_all, (size_t)n * sizeof(P2pBlob), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_all);

    int ok = 1;
    P2pPeers &pp = c->p2p;
    memset(&pp, 0, sizeof pp);
    pp.stride = c->xyzd_stride;
    pp.spin_limit = (long long)c->opt_p2p_timeout_s * 2000000000LL;
    // a fresh set-up re-synchronises everything a timed-out exchange left behind: epochs restart at 0 on every rank
    CU(cudaMemsetAsync(c->d_flags, 0, P2P_FLAG_WORDS * sizeof(unsigned long long), c->stream));
    *c->h_p2p_err = 0;
    c->p2p_fault = false;
    std::vector<void *> mapped_xyzd(n, nullptr), mapped_type(n, nullptr), mapped_flags(n, nullptr);
    mapped_xyzd[me] = c->d_xyzd; mapped_type[me] = c->s.type; mapped_flags[me] = c->d_flags;
    for (int r = 0; r < n; r++) {
        ok = ok && all[r].ok && all[r].host == mine.host;
        for (int k = 0; k < 3; k++) ok = ok && all[r].size[k] == mine.size[k];
    }
    const int *gs = c->dom.grid_size;
    for (int k = 0; k < 27 && ok; k++) {
        if (k == 13 || !c->push_code_used[k]) continue;
        const int s[3] = {k % 3 - 1, (k / 3) % 3 - 1, k / 9 - 1};
        int dc[3], r = -1;
        for (int d = 0; d < 3; d++) dc[d] = ((mine.coord[d] - s[d]) % gs[d] + gs[d]) % gs[d];
        for (int q = 0; q < n; q++)
            if (all[q].coord[0] == dc[0] && all[q].coord[1] == dc[1] && all[q].coord[2] == dc[2]) r = q;
        if (r < 0) { ok = 0; break; }
        if (!mapped_xyzd[r]) {
            if (all[r].pid == mine.pid) {          // two sub-boxes of one process: plain peer access
                int can = 0;
                cudaDeviceCanAccessPeer(&can, c->device, all[r].device);
                if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0; cudaGetLastError(); }
                if (!can) { ok = 0; break; }
                mapped_xyzd[r] = all[r].p_xyzd; mapped_type[r] = all[r].p_type; mapped_flags[r] = all[r].p_flags;
            } else {
                void *px = nullptr, *pt = nullptr, *pf = nullptr;
                if (cudaIpcOpenMemHandle(&px, all[r].h_xyzd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                    cudaIpcOpenMemHandle(&pt, all[r].h_type, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                    cudaIpcOpenMemHandle(&pf, all[r].h_flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    for (void *p : {px, pt, pf}) if (p) cudaIpcCloseMemHandle(p);
                    ok = 0;
                    break;
                }
                c->p2p_opened.push_back(px); c->p2p_opened.push_back(pt); c->p2p_opened.push_back(pf);
                mapped_xyzd[r] = px; mapped_type[r] = pt; mapped_flags[r] = pf;
            }
        }
        pp.xyzd[k] = (double *)mapped_xyzd[r]; pp.type[k] = (int8_t *)mapped_type[r]; pp.flags[k] = (unsigned long long *)mapped_flags[r];
        pp.mask |= 1u << k;
        push_shift(&c->dom, k, pp.shift[k]);
    }
    // the handshake is symmetric (my origin for code k is my destination for 26 - k): both must be mapped
    for (int k = 0; k < 27 && ok; k++)
        if (((pp.mask >> k) & 1u) && !((pp.mask >> (26 - k)) & 1u)) ok = 0;
    // every rank takes the same path
    int *d_ok = nullptr;
    TRY(dmalloc(&d_ok, 1));
    CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, kNcclInt32, kNcclMin, c->nccl_comm, c->stream));
    CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(d_ok);
    if (!ok) { p2p_release(c); return 0; }
    // device copy of the peer table for the kernels that push from inside (k_verlet1, k_rho_f)
    if (!c->d_p2p_fault) TRY(dmalloc(&c->d_p2p_fault, 1));
    CU(cudaMemsetAsync(c->d_p2p_fault, 0, sizeof(unsigned int), c->stream));
    pp.fault = c->d_p2p_fault;
    pp.fence_mode = c->opt_p2p_fence;
    if (!c->d_p2p_dev) TRY(dmalloc(&c->d_p2p_dev, 1));
    CU(cudaMemcpyAsync(c->d_p2p_dev, &pp, sizeof pp, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->p2p_active = true;
    c->p2p_epoch = 0;
    c->p2p_ready_sent = 0;
    return 0;
}

// one ghost exchange, collective over the sub-boxes like the NCCL path:
//   p2p_push: [ready signal, unless p2p_post_ready sent it ahead] -> push (waits for the destinations' ready, last CTA
//             signals arrive)
//   p2p_wait: wait for every origin's arrive -- as a tiny kernel, or inside the consuming stencil kernel right before its
//             first boundary unit (eam_smem.cuh:LateWait), which hides the neighbours' latency and skew behind the interior
static int p2p_push(misa_b200_ctx *c, bool positions, cudaStream_t st) {
    const unsigned long long e = ++c->p2p_epoch;
    if (c->p2p_ready_sent < e) {
        k_p2p_ready<<<1, 32, 0, st>>>(c->p2p, e);
        c->p2p_ready_sent = e;
        c->launches++;
    }
    const int nb = std::max(1, std::min((c->n_push + P2P_PUSH_THREADS - 1) / P2P_PUSH_THREADS, 2 * std::max(c->sm_count, 1)));
    c->p2p.fence_mode = c->opt_p2p_fence;
    c->p2p.dbg = c->opt_p2p_debug && c->d_p2p_dbg ? c->d_p2p_dbg + (positions ? 0 : 8) : nullptr;
    unsigned int *done = reinterpret_cast<unsigned int *>(c->d_flags + P2P_FLAG_WORDS - 1);
    // inside the sync-free step the push carries this sub-box's displacement maximum along (dmax_by_flags)
    if (positions && c->dmax_by_flags)   // the sync-free step: no species bytes (see k_p2p_push_x)
        k_p2p_push_x<false><<<nb, P2P_PUSH_THREADS, 0, st>>>(c->p2p, c->n_push, c->d_push_dst, c->d_push_src, c->d_push_code, c->s, e, c->d_flags, done, c->d_p2p_err, c->d_stepinfo + 1);
    else if (positions)
        k_p2p_push_x<true><<<nb, P2P_PUSH_THREADS, 0, st>>>(c->p2p, c->n_push, c->d_push_dst, c->d_push_src, c->d_push_code, c->s, e, c->d_flags, done, c->d_p2p_err, nullptr);
    else k_p2p_push_df<<<nb, P2P_PUSH_THREADS, 0, st>>>(c->p2p, c->n_push, c->d_push_dst, c->d_push_src, c->d_push_code, c->s, e, c->d_flags, done, c->d_p2p_err);
    c->launches++;
    if (c->p2p.fence_mode == 2) {
        k_p2p_arrive<<<1, 32, 0, st>>>(c->p2p, e, c->d_p2p_err, positions && c->dmax_by_flags ? c->d_stepinfo + 1 : nullptr);
        c->launches++;
    }
    CU(cudaGetLastError());
    return 0;
}
static int p2p_wait(misa_b200_ctx *c, cudaStream_t st) {
    k_p2p_wait_arrive<<<1, 32, 0, st>>>(c->p2p, c->p2p_epoch, c->d_flags, c->d_p2p_err, c->d_stepinfo + 1, c->dmax_by_flags ? c->d_stepinfo_n : nullptr);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
static int p2p_exchange(misa_b200_ctx *c, bool positions, cudaStream_t st) {
    TRY(p2p_push(c, positions, st));
    return p2p_wait(c, st);
}
// Called right after the last reader of this sub-box's ghosts in a step (the force kernel of the sync-free step): frees
// the ghosts for the next `ahead` exchanges, so that the neighbours' next pushes find the flag already there.
static int p2p_post_ready(misa_b200_ctx *c, int ahead, cudaStream_t st) {
    if (!(c->p2p_active && c->opt_p2p)) return 0;
    const unsigned long long e = c->p2p_epoch + (unsigned long long)ahead;
    if (c->p2p_ready_sent >= e) return 0;
    k_p2p_ready<<<1, 32, 0, st>>>(c->p2p, e);
    c->p2p_ready_sent = e;
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
