// qe_shard.cuh -- the state-range-sharded table (BASELINE config 4) as ONE persistent kernel per GPU over peer memory.
//
// Round 1 ran the sharded step as eager PyTorch + all-to-all rounds (a fixed point over the remote bootstrap values:
// ~5 bulk-synchronous exchange rounds per vector step) and was five times SLOWER on eight GPUs than one GPU on the same
// table.  This file is the pipelined form of qe_pipe.cuh stretched over NVLink instead:
//
//   * rank g owns the states [g * rows, (g + 1) * rows) -- their Q rows, and the writer records / segment bounds of the
//     agents standing on them -- in one slab of device memory that every peer maps (CUDA IPC; plain pointers when the
//     ranks share a process);
//   * agent i lives for good on its HOME rank i / (N / G) (contiguous id ranges: every rank walks its agents in global
//     id order, which is what makes the in-order pipeline deadlock-free across GPUs too) and never migrates: it reads
//     the row of its state, files its writer record, polls the records of the row it bootstraps from and publishes its
//     target with peer loads and stores (one NVLink round trip per dependency level, no host, no collective);
//   * the per-step order is a distributed stable sort: every rank partitions its agents' next states by owner (stable,
//     so the concatenation by source rank is in global id order), stores the pairs into the owners' inboxes, the owners
//     radix-sort them locally and store every agent's position back into its home rank's slab;
//   * phases are separated by a barrier over all ranks: grid barrier + one flag per peer, written and polled over NVLink.
//
// The same kernel runs G "virtual ranks" side by side on one GPU (a group of CTAs per rank, every barrier a grid
// barrier): that is how the one-GPU tests check the multi-rank logic bit for bit against the oracle.
//
// Hash MDP only (the environment of config 4); results are identical to the single-GPU engine and to the reference's
// sequential loop in global agent order.
#pragma once
#include "qe_flow.cuh"

namespace qe {

constexpr int kMaxRanks = 8;
constexpr uint64_t kShardTimeoutNs = 20000000000ull;

struct ShardPeer {           // what rank g exposes to every rank (pointers into its slab, valid in THIS process)
    float* q;                // [rows][ld] table shard
    uint2* rec;              // [N + 8] writer records by sorted position at this owner
    uint2* seg;              // [rows] {start, end} per owned state
    int2* inbox;             // [N] incoming {state - first state of the shard, global agent id} of the distributed sort
    int32_t* pos;            // [n_home] position (at the owner of its current state) of every home agent; written by the owners
    uint4* tw;               // [n_home] per home agent {next state, position, reward bits, action | term << 7 | owner of the current state << 8}
    unsigned int* cin;       // [G] pairs rank r sends to this owner in the current sort
    unsigned int* flag;      // [G] barrier flags, flag[r] written by rank r
};
struct ShardLocal {          // private to a rank
    int32_t* st_a;           // [n_home] states of the home agents
    int32_t* st_b;
    float* ep_ret;           // [n_home]
    int2* kv[2];             // [N] radix ping-pong
    int* ghist;              // [kRadix][blocks per rank]
    int* rowtot;             // [kRadix]
    unsigned int* wcnt;      // [warps per rank][G] pairs per (warp, destination) before this warp inside its block
    unsigned int* bcnt;      // [blocks per rank][G] pairs per (block, destination), scanned in place over the blocks
    unsigned int* ctr;       // [16] 0: chunk claims of phase T; 1: agents in the order kv[] / seg[] still describe; 4: abort flag
    double* ep_sum;
    unsigned long long* ep_count;
    int* err;
    unsigned long long* phase_ns;  // [8 * 16] %globaltimer after {A, T, C, partition counts, scatter, local sort} of the first 16 steps
};
struct ShardArgs {
    int G, first_rank, nlocal, blocks_per_rank, multi_device;
    int A, ld, passes, msd_shift;  // msd_shift: bucket of a state inside its shard = state >> msd_shift (< kRadix)
    int n_total, n_home;     // agents, agents per rank (ceil)
    int64_t S, rows;         // states, states per shard (ceil)
    int steps;
    const uint64_t* eps_thresh;  // device [steps]
    const float* lr;             // device [steps]
    uint32_t stream_seed, t0, env_stream_seed, env_t0, env_seed;
    uint64_t term_thresh;
    int empty_all, use_masks;
    float gamma;
    uint32_t epoch0;         // barrier epoch at launch (the host adds the barriers of the launch afterwards)
    int sorted_valid;        // the order of the current states is already in place
    ShardPeer peer[kMaxRanks];
    ShardLocal loc[kMaxRanks];  // [nlocal] the ranks this launch runs
};

__device__ __forceinline__ U8 ld_relaxed_sys_v8(const uint2* p) {
    U8 r;
    asm volatile("ld.relaxed.sys.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// masked max, tie set and Q[s, a] of a row held by ONE lane (LPR sectors of eight floats)
template <int LPR>
__device__ __forceinline__ void lane_row_max_tie(const F8* v, uint32_t legal, float& m_out, uint32_t& tie_out) {
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < LPR; ++c) m = fmax_plain(m, max8(v[c], (legal >> (8 * c)) & 0xFFu));
    uint32_t t = 0u;
#pragma unroll
    for (int c = 0; c < LPR; ++c) t |= tie8(v[c], (legal >> (8 * c)) & 0xFFu, m) << (8 * c);
    m_out = m;
    tie_out = t;
}

// Barrier over ALL ranks.  One GPU running every rank: a grid barrier.  One rank per GPU: grid barrier, then thread 0
// raises its flag in every peer's slab and waits for theirs (system-scope release / acquire over NVLink), grid barrier.
__device__ __forceinline__ void shard_xsync(cg::grid_group& grid, const ShardArgs& H, uint32_t& epoch) {
    grid.sync();
    if (H.multi_device) {
        ++epoch;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const int me = H.first_rank;
            __threadfence_system();
            for (int g = 0; g < H.G; ++g)
                if (g != me) st_release_sys_u32(H.peer[g].flag + me, epoch);
            const uint64_t t0 = global_ns();
            for (int g = 0; g < H.G; ++g) {
                if (g == me) continue;
                uint32_t spins = 0;
                while ((int32_t)(ld_acquire_sys_u32(H.peer[me].flag + g) - epoch) < 0) {
                    __nanosleep(200);
                    if ((++spins & 1023u) == 0u && (ld_relaxed_u32(H.loc[0].ctr + 4) != 0u || global_ns() - t0 > kShardTimeoutNs)) {
                        atomicExch(H.loc[0].ctr + 4, 1u);
                        atomicOr(H.loc[0].err, kErrTimeout);
                        break;
                    }
                }
            }
            __threadfence_system();
        }
        grid.sync();
    }
}

// The owner's local stable sort of its inbox, MSD first (the bucket sort of qe_flow.cuh with pairs as input, blocks
// counted inside the rank's group of CTAs): ONE pass over the grid partitions the pairs by the top (up to) 10 bits of the
// state -- per-warp digit counts, column scan over the group's blocks, stable scatter --, then every bucket is sorted by
// its low bits by one CTA in shared memory (flow_block_staged), which also writes the segment bounds and stores every
// agent's position into its HOME rank's slab.  (Round 2's first version ran one such grid-wide pass per 10 bits: three
// passes of ~40 us fixed cost each on the 100M-state table, whatever the number of keys.)  The finished order is in kv[0].
template <int WARPS>
__device__ __forceinline__ void shard_sort(cg::grid_group& grid, int (*whist)[kRadix], unsigned char* arena, int arena_bytes, int* s_base, int* s_wsum,
                                           int n, int old_n, const ShardArgs& H, const ShardLocal& L, int me, int b, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const ShardPeer& own = H.peer[me];
    if (old_n > 0) {  // reset the bounds of the previous order (still in kv[0])
        const int2* old = L.kv[0];
        for (int q = b * blockDim.x + threadIdx.x; q < old_n; q += nb * blockDim.x) {
            const int32_t kq = __ldcg(&old[q].x);
            if (q == 0 || __ldcg(&old[q - 1].x) != kq) own.seg[kq] = make_uint2(0u, 0u);
        }
    }
    const int chunk = ((n + nb - 1) / nb + 31) & ~31;
    const int part = ((chunk / 32 + WARPS - 1) / WARPS) * 32;
    const int lo = min(b * chunk + warp * part, n), hi = min(min(b * chunk + (warp + 1) * part, (b + 1) * chunk), n);
    const int shift = H.msd_shift;
    const int2* in = own.inbox;
    int2* out = L.kv[shift ? 1 : 0];
    auto load_pair = [&](int x) {
        int2 e = make_int2(0, 0);
        if (x < hi) e = __ldcg(in + x);
        return e;
    };
    for (int d = lane; d < kRadix; d += 32) whist[warp][d] = 0;
    __syncwarp();
    for (int base = lo; base < hi; base += 256) {
        int2 e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = load_pair(base + 32 * u + lane);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + 32 * u + lane < hi) smem_inc(&whist[warp][((uint32_t)e[u].x >> shift) & (kRadix - 1)]);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) t += whist[w][d];
        L.ghist[(size_t)d * nb + b] = t;
    }
    grid.sync();
    {
        const int per = (nb + 31) / 32;
        for (int d = b + warp * nb; d < kRadix; d += WARPS * nb) {  // (bucket d: warp d / nb of the group's block d % nb)
            int* row = L.ghist + (size_t)d * nb;
            int v[kScanPerLane];
            int sum = 0;
#pragma unroll
            for (int j = 0; j < kScanPerLane; ++j) {
                const int x = lane * per + j;
                v[j] = (j < per && x < nb) ? __ldcg(row + x) : 0;
            }
#pragma unroll
            for (int j = 0; j < kScanPerLane; ++j) sum += v[j];
            const int incl = warp_incl_scan(sum);
            int run = incl - sum;
#pragma unroll
            for (int j = 0; j < kScanPerLane; ++j) {
                const int x = lane * per + j;
                if (j < per && x < nb) row[x] = run;
                run += v[j];
            }
            if (lane == 31) L.rowtot[d] = incl;
        }
    }
    grid.sync();
    {
        const int4 v4 = __ldcg(reinterpret_cast<const int4*>(L.rowtot) + threadIdx.x);
        const int v[4] = {v4.x, v4.y, v4.z, v4.w};
        const int sum = v4.x + v4.y + v4.z + v4.w;
        const int incl = warp_incl_scan(sum);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) before += (w < warp) ? s_wsum[w] : 0;
        int run = before + incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) { s_base[threadIdx.x * 4 + j] = run; run += v[j]; }
        if (threadIdx.x == 255) s_base[kRadix] = run;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
        int run = s_base[d] + __ldcg(L.ghist + (size_t)d * nb + b);
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const int c = whist[w][d];
            whist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int base = lo; base < hi; base += 256) {
        int2 e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = load_pair(base + 32 * u + lane);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int x = base + 32 * u + lane;
            if (base + 32 * u >= hi) break;  // (uniform)
            const bool act = x < hi;
            const uint32_t d = ((uint32_t)e[u].x >> shift) & (kRadix - 1);
            const uint32_t peers = digit_peers(d, act);
            if (act) out[whist[warp][d] + __popc(peers & ((1u << lane) - 1u))] = e[u];
            __syncwarp();
            if (act && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
            __syncwarp();
        }
    }
    grid.sync();
    // ---- the buckets of this block, one after the other
    auto emit = [&](int q, int32_t st, int32_t agent) {
        L.kv[0][q] = make_int2(st, agent);
        const int hr = agent / H.n_home;  // the agent's home rank learns where its record goes
        H.peer[hr].pos[agent - hr * H.n_home] = q;
    };
    const int block_cap = min(flow_block_cap(arena_bytes), 65535);
    const int Lp = (shift + kRadixBits - 1) / kRadixBits;
    for (int d = b; d < kRadix; d += nb) {
        const int blo = s_base[d], bhi = s_base[d + 1];
        if (bhi <= blo) continue;  // (uniform)
        const int32_t key_hi = (int32_t)((uint32_t)d << shift);
        if (bhi - blo <= block_cap) {
            flow_block_staged<WARPS>(arena, arena_bytes, s_wsum, L.kv[shift ? 1 : 0], blo, bhi, shift, key_hi, own.seg, emit);
            continue;
        }
        int src = Lp ? 1 : 0;  // (too large for shared memory: counting passes through global memory)
        for (int ps = 0; ps < Lp; ++ps) {
            const int sh = ps * kRadixBits;
            flow_local_pass<WARPS>(whist, s_wsum, L.kv[src], L.kv[src ^ 1], blo, bhi, sh, min(kRadixBits, shift - sh));
            src ^= 1;
        }
        if (src != 0) {
            for (int q = blo + threadIdx.x; q < bhi; q += blockDim.x) L.kv[0][q] = __ldcg(L.kv[1] + q);
            __syncthreads();
        }
        const int2* fin = L.kv[0];
        for (int q = blo + threadIdx.x; q < bhi; q += blockDim.x) {
            const int2 e = __ldcg(fin + q);
            const int32_t prev = q > blo ? __ldcg(&fin[q - 1].x) : -1, next = q + 1 < bhi ? __ldcg(&fin[q + 1].x) : -1;
            const int hr = e.y / H.n_home;
            H.peer[hr].pos[e.y - hr * H.n_home] = q;
            if (prev != e.x) own.seg[e.x].x = (uint32_t)q;
            if (next != e.x) own.seg[e.x].y = (uint32_t)(q + 1);
        }
        __syncthreads();
    }
}
template <int WARPS>
__device__ __forceinline__ void shard_bounds(const int2* sorted, int n, uint2* seg, int b, int nb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = ((n + nb - 1) / nb + 31) & ~31;
    const int part = ((chunk / 32 + WARPS - 1) / WARPS) * 32;
    const int lo = min(b * chunk + warp * part, n), hi = min(min(b * chunk + (warp + 1) * part, (b + 1) * chunk), n);
    constexpr int U = 8;
    for (int base = lo; base < hi; base += 32 * U) {
        int32_t kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = base + 32 * u + lane;
            kk[u] = x < n ? __ldcg(&sorted[x].x) : -1;
        }
        const int32_t left = base > 0 ? __ldcg(&sorted[base - 1].x) : -1;
        const int32_t right = base + 32 * U < n ? __ldcg(&sorted[base + 32 * U].x) : -1;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = base + 32 * u + lane;
            int32_t prev = __shfl_up_sync(kFull, kk[u], 1), next = __shfl_down_sync(kFull, kk[u], 1);
            const int32_t pl = __shfl_sync(kFull, kk[u > 0 ? u - 1 : 0], 31), nf = __shfl_sync(kFull, kk[u < U - 1 ? u + 1 : U - 1], 0);
            if (lane == 0) prev = u > 0 ? pl : left;
            if (lane == 31) next = u < U - 1 ? nf : right;
            if (x < hi) {
                if (prev != kk[u]) seg[kk[u]].x = (uint32_t)x;
                if (next != kk[u]) seg[kk[u]].y = (uint32_t)(x + 1);
            }
        }
    }
}

// dynamic shared memory: the phase layouts of the pipelined form, and room for the bucket sorts (counters + 8 bytes per key
// of a bucket of ~5600 keys: one rank alone on the 100M-state table has 4096 per bucket) with the bucket starts behind it
__host__ __device__ constexpr size_t shard_smem_bytes(int lpr) {
    return pipe_smem_bytes(lpr) > (size_t)57344 ? pipe_smem_bytes(lpr) : (size_t)57344;
}

template <int LPR>
__global__ void __launch_bounds__(256, 3) shard_kernel(ShardArgs H) {
    cg::grid_group grid = cg::this_grid();
    constexpr int WARPS = 8;
    constexpr int RS = 8 * LPR + 4;
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    __shared__ int s_wsum[WARPS];
    __shared__ unsigned int s_off[kMaxRanks];  // where this rank's pairs start in every owner's inbox
    extern __shared__ __align__(16) float s_mem[];
    int (*s_whist)[kRadix] = reinterpret_cast<int (*)[kRadix]>(s_mem);
    constexpr int kArenaBytes = (int)(shard_smem_bytes(LPR) - sizeof(int) * (kRadix + 8));
    int* s_base = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s_mem) + kArenaBytes);  // [kRadix + 1] bucket starts of the local sort
    float* s_row = s_mem;
    uint32_t* s_touch = reinterpret_cast<uint32_t*>(s_mem) + 8 * LPR * 256;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int vr = blockIdx.x / H.blocks_per_rank;           // which of this launch's ranks the CTA works for
    const int b = blockIdx.x - vr * H.blocks_per_rank, nb = H.blocks_per_rank;
    const int me = H.first_rank + vr;
    const ShardLocal& L = H.loc[vr];
    const ShardPeer& own = H.peer[me];
    const int G = H.G;
    const int rtid = b * blockDim.x + threadIdx.x, rthreads = nb * blockDim.x;  // thread index / count inside the rank's group
    const int rwarp = rtid >> 5, rwarps = rthreads >> 5;
    const int wbase = threadIdx.x & ~31;
    const int home0 = me * H.n_home;                                   // first global id of this rank's agents
    const int nh = max(0, min(H.n_home, H.n_total - home0));           // how many it has
    const uint32_t full = H.A >= 32 ? 0xFFFFFFFFu : ((1u << H.A) - 1u);
    const int64_t rows = H.rows;
    uint32_t epoch = H.epoch0;
    int n_in = (int)ld_relaxed_u32(L.ctr + 1);  // records at this owner in the current order
    int old_n = n_in, n_next = n_in;
    int kstep = -1;
    auto stamp = [&](int slot) {
        if (rtid == 0 && kstep >= 0 && kstep < 16) L.phase_ns[8 * kstep + slot] = global_ns();
    };

    // -------- the distributed stable sort of the agents by state, in pieces.  Every warp owns the contiguous part
    // [part_lo, part_hi) of the home agents (phase A walks the same parts, so it can count the destinations as it goes).
    const int part = ((nh + rwarps - 1) / rwarps + 31) & ~31;
    const int part_lo = min(rwarp * part, nh), part_hi = min(part_lo + part, nh);
    __shared__ unsigned int s_wc[WARPS][kMaxRanks];
    // destinations of one tile of keys -> per-lane counters (lane d < G counts destination d)
    auto count_tile = [&](int d, unsigned int& cnt) {
        for (int g = 0; g < G; ++g) {
            const unsigned int c = __popc(__ballot_sync(kFull, d == g));
            if (lane == g) cnt += c;
        }
    };
    // S1 (end): the warps' counters -> offsets inside the block (wcnt) and block totals (bcnt)
    auto file_counts = [&](unsigned int cnt) {
        if (lane < G) s_wc[warp][lane] = cnt;
        __syncthreads();
        if (threadIdx.x < G) {
            unsigned int run = 0;
            for (int w = 0; w < WARPS; ++w) {
                L.wcnt[(size_t)(b * WARPS + w) * G + threadIdx.x] = run;
                run += s_wc[w][threadIdx.x];
            }
            L.bcnt[(size_t)b * G + threadIdx.x] = run;
        }
        __syncthreads();
    };
    // S2 (after a grid barrier): exclusive scan of the block totals per destination (warp d of the group's first block),
    // totals to the owners
    auto scan_counts = [&]() {
        if (b == 0 && warp < G) {
            const int d = warp;
            const int per = (nb + 31) / 32;
            unsigned int v[kScanPerLane];
            unsigned int sum = 0;
#pragma unroll
            for (int j = 0; j < kScanPerLane; ++j) {
                const int x = lane * per + j;
                v[j] = (j < per && x < nb) ? __ldcg(L.bcnt + (size_t)x * G + d) : 0u;
                sum += v[j];
            }
            const unsigned int incl = (unsigned int)warp_incl_scan((int)sum);
            unsigned int run = incl - sum;
#pragma unroll
            for (int j = 0; j < kScanPerLane; ++j) {
                const int x = lane * per + j;
                if (j < per && x < nb) L.bcnt[(size_t)x * G + d] = run;
                run += v[j];
            }
            if (lane == 31) H.peer[d].cin[me] = incl;
        }
    };
    // S3 (after a barrier over all ranks): where my pairs start at every owner, how many pairs I receive, then the
    // stable scatter of {state inside the shard, agent} into the owners' inboxes
    auto scatter_pairs = [&](const int32_t* keys) {
        if (threadIdx.x < G) {
            unsigned int off = 0;
            for (int r = 0; r < me; ++r) off += ld_relaxed_sys_u32(H.peer[threadIdx.x].cin + r);
            s_off[threadIdx.x] = off + __ldcg(L.bcnt + (size_t)b * G + threadIdx.x);
        }
        {
            unsigned int tot = 0;
            for (int r = 0; r < G; ++r) tot += ld_relaxed_sys_u32(own.cin + r);
            n_next = (int)tot;  // (the current order's records are still in use: n_in changes with the local sort)
        }
        __syncthreads();
        // Rounds of up to 256 agents per warp: the round's pairs are grouped by owner in shared memory first, so that the
        // peer stores of consecutive lanes go to consecutive slots of one inbox (few large NVLink writes instead of one
        // 8-byte write per pair: the fabric is bound by requests, not bytes).
        unsigned int run = lane < G ? __ldcg(L.wcnt + (size_t)rwarp * G + lane) + s_off[lane] : 0u;  // lane d: next free slot at owner d
        int2* stage = reinterpret_cast<int2*>(s_mem) + warp * 256;          // [256] pairs of the round, grouped by owner
        unsigned int* sdst = reinterpret_cast<unsigned int*>(s_mem) + WARPS * 512 + warp * 256;  // [256] their global slots | owner << 28
        constexpr int R = 8;
        for (int base = part_lo; base < part_hi; base += 32 * R) {
            int32_t kk[R];
            int dd[R];
            unsigned int cnt = 0;  // lane d: pairs of the round for owner d
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int j = base + 32 * u + lane;
                kk[u] = j < part_hi ? __ldcg(keys + j) : 0;
                dd[u] = j < part_hi ? (int)(kk[u] / rows) : -1;
                count_tile(dd[u], cnt);
            }
            const unsigned int incl = (unsigned int)warp_incl_scan((int)(lane < G ? cnt : 0u));
            unsigned int lstart = incl - (lane < G ? cnt : 0u);  // lane d: first local slot of owner d in the stage
            const unsigned int total = __shfl_sync(kFull, incl, 31);
            unsigned int lrun = lstart;
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int j = base + 32 * u + lane;
                unsigned int loc = 0, glob = 0;
                for (int g = 0; g < G; ++g) {
                    const uint32_t m = __ballot_sync(kFull, dd[u] == g);
                    const unsigned int ls = __shfl_sync(kFull, lrun, g), gs = __shfl_sync(kFull, run, g), l0 = __shfl_sync(kFull, lstart, g);
                    if (dd[u] == g) {
                        loc = ls + __popc(m & ((1u << lane) - 1u));
                        glob = gs + (loc - l0);
                    }
                    if (lane == g) lrun += __popc(m);
                }
                if (dd[u] >= 0) {
                    stage[loc] = make_int2((int32_t)(kk[u] - (int64_t)dd[u] * rows), home0 + j);
                    sdst[loc] = glob | ((unsigned int)dd[u] << 28);
                }
            }
            __syncwarp();
            for (unsigned int e = lane; e < total; e += 32) {
                const unsigned int w = sdst[e];
                H.peer[w >> 28].inbox[w & 0x0FFFFFFFu] = stage[e];
            }
            __syncwarp();
            if (lane < G) run += cnt;
        }
        __syncthreads();
    };
    // S4 + S5: the owner's local sort, positions to the home ranks, segment bounds
    auto local_sort = [&]() {
        n_in = n_next;
        shard_sort<WARPS>(grid, s_whist, reinterpret_cast<unsigned char*>(s_mem), kArenaBytes, s_base, s_wsum, n_in, old_n, H, L, me, b, nb);
        old_n = n_in;
        if (rtid == 0) L.ctr[1] = (unsigned int)n_in;
    };

    if (!H.sorted_valid) {  // the order of the states this launch starts from
        unsigned int cnt = 0;
        for (int base = part_lo; base < part_hi; base += 32) {
            const int j = base + lane;
            count_tile(j < part_hi ? (int)(__ldcg(L.st_a + j) / rows) : -1, cnt);
        }
        file_counts(cnt);
        grid.sync();
        scan_counts();
        shard_xsync(grid, H, epoch);
        scatter_pairs(L.st_a);
        shard_xsync(grid, H, epoch);
        local_sort();
        shard_xsync(grid, H, epoch);
    }

    for (int k = 0; k < H.steps; ++k) {
        kstep = k;
        stamp(0);
        int32_t* cur = (k & 1) ? L.st_b : L.st_a;
        int32_t* nxt = (k & 1) ? L.st_a : L.st_b;
        const uint32_t t_sel = H.t0 + (uint32_t)k, t_env = H.env_t0 + (uint32_t)k;
        const uint64_t thresh = H.eps_thresh[k];
        const float lr = H.lr[k];
        const int2* sorted = L.kv[0];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + environment step of the home agents; records go to the owners of their states;
        // the destinations of the next states are counted on the way (first stage of the next step's sort)
        // (three tiles in flight per warp: state / position of tile t + 2 and the row of tile t + 1 travel while tile t is finished)
        unsigned int dcnt = 0;
        {
            struct Tile { int s, pos; };
            auto load_sp = [&](int base) {
                Tile t{0, 0};
                if (base + lane < part_hi) { t.s = cur[base + lane]; t.pos = __ldcg(own.pos + base + lane); }
                return t;
            };
            auto load_row = [&](int base, const Tile& t, F8* v) {
#pragma unroll
                for (int c = 0; c < LPR; ++c)
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[c].v[q] = 0.0f;
                if (base + lane < part_hi) {
                    const int o = (int)(t.s / rows);
                    const float* row = H.peer[o].q + (size_t)(t.s - (int64_t)o * rows) * H.ld;
#pragma unroll
                    for (int c = 0; c < LPR; ++c) v[c] = ld_row8(row + 8 * c);
                }
            };
            Tile ta = load_sp(part_lo), tb = load_sp(part_lo + 32);
            F8 va[LPR], vb[LPR];
            load_row(part_lo, ta, va);
            for (int base = part_lo; base < part_hi; base += 32) {
                const Tile tc = load_sp(base + 64);
                load_row(base + 32, tb, vb);
                const int j = base + lane;
                int dnext = -1;
                if (j < part_hi) {
                    const uint32_t gid = (uint32_t)(home0 + j);
                    const int s = ta.s, mypos = ta.pos;
                    const int o = (int)(s / rows);
                    const uint32_t valid = H.use_masks ? mdp_mask((uint32_t)s, H.A, H.env_seed) : full;
                    const bool explore = (uint64_t)stream_u32(H.stream_seed, t_sel, gid, 0u) < thresh;
                    const uint32_t bits1 = stream_u32(H.stream_seed, t_sel, gid, 1u);
                    float mx;
                    uint32_t tie;
                    lane_row_max_tie<LPR>(va, valid, mx, tie);
                    int a = pick_action(H.A, valid, tie, explore, H.empty_all != 0, bits1);
                    if (a < 0) { atomicOr(L.err, kErrEmpty); a = 0; }
                    int32_t s2 = s;
                    float r = 0.0f;
                    bool term = false;
                    mdp_step(s2, a, (uint32_t)H.S, H.A, H.env_seed, H.term_thresh, stream_u32(H.env_stream_seed, t_env, gid, 2u),
                             stream_u32(H.env_stream_seed, t_env, gid, 3u), r, term);
                    nxt[j] = s2;
                    dnext = (int)(s2 / rows);
                    own.tw[j] = make_uint4((uint32_t)s2, (uint32_t)mypos, __float_as_uint(r), (uint32_t)a | (term ? 0x80u : 0u) | ((uint32_t)o << 8));
                    H.peer[o].rec[mypos] = make_uint2(gid | ((uint32_t)a << 24) | ((!term && s2 == s) ? (1u << 29) : 0u),
                                                      term ? __float_as_uint(td_target_s(r, 0.0f, H.gamma)) : kPending);
                    float acc = L.ep_ret[j] + r;
                    if (term) { loc_sum += (double)acc; ++loc_cnt; acc = 0.0f; }
                    L.ep_ret[j] = acc;
                }
                count_tile(dnext, dcnt);
                ta = tb;
                tb = tc;
#pragma unroll
                for (int c = 0; c < LPR; ++c) va[c] = vb[c];
            }
        }
        file_counts(dcnt);
        {
            for (int d = 16; d > 0; d >>= 1) {
                loc_sum += __shfl_xor_sync(kFull, loc_sum, d);
                loc_cnt += __shfl_xor_sync(kFull, loc_cnt, d);
            }
            if (lane == 0) { s_sum[warp] = loc_sum; s_cnt[warp] = loc_cnt; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double bs = 0.0;
                unsigned int bc = 0;
                for (int w = 0; w < WARPS; ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
                if (bc) { atomicAdd(L.ep_sum, bs); atomicAdd(L.ep_count, (unsigned long long)bc); }
            }
            __syncthreads();
        }
        if (rtid == 0) L.ctr[0] = 0u;
        stamp(7);
        grid.sync();
        scan_counts();
        shard_xsync(grid, H, epoch);
        stamp(1);

        // ---------------- (sort, stage 3) the pairs of the next order travel to their owners' inboxes
        scatter_pairs(nxt);
        stamp(4);

        // ---------------- phase T: in-order target pipeline over the home agents (qe_pipe.cuh), rows / records / targets
        // through peer pointers
        {
            unsigned int* claim = L.ctr;
            float* myrow = s_mem + threadIdx.x * RS;
            auto claim_chunk = [&]() {
                int c = 0;
                if (lane == 0) c = (int)atomicAdd(claim, 1u);
                return __shfl_sync(kFull, c, 0) * 32;
            };
            auto load_chunk = [&](int base) {
                uint4 q = make_uint4(0u, 0u, 0u, 0x80u);
                if (base + lane < nh) q = __ldcg(own.tw + base + lane);
                return q;
            };
            int cb = claim_chunk(), cbn = claim_chunk(), cbnn = claim_chunk();
            uint4 pd = load_chunk(cb), pdn = load_chunk(cbn);
            int pn = 0;
            int i = 0;
            float r = 0.0f;
            uint32_t m2 = 0u, p = 0u, pe = 0u;
            const uint2* rec2 = nullptr;     // records of the owner of s'
            uint32_t* mytarget = nullptr;    // where this agent's target goes (the owner of its current state)
            bool busy = false;
            auto row_max = [&]() {
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2 * LPR; ++c) {
                    const float4 v = reinterpret_cast<const float4*>(myrow)[c];
                    m = fmaxf(fmaxf(fmaxf(fmaxf(m, v.x), v.y), v.z), v.w);
                }
                return m;
            };
            const uint64_t t_start = global_ns();
            for (uint32_t spins = 0;; ++spins) {
                const uint32_t freeb = __ballot_sync(kFull, !busy);
                bool fresh = false;
                uint2 sg = make_uint2(0u, 0u);
                F8 rowv[LPR];
                if (cb < nh && freeb != 0u) {  // (latency of the peer accesses, not issue slots, bounds this loop: refill at once)
                    const int cc = min(32, nh - cb);
                    const int src = pn + __popc(freeb & ((1u << lane) - 1u));
                    const uint32_t y2 = __shfl_sync(kFull, pd.x, src & 31), pos2 = __shfl_sync(kFull, pd.y, src & 31);
                    const uint32_t r2 = __shfl_sync(kFull, pd.z, src & 31), at2 = __shfl_sync(kFull, pd.w, src & 31);
                    if (!busy && src < cc && !(at2 & 0x80u)) {
                        i = home0 + cb + src;
                        r = __uint_as_float(r2);
                        const int y = (int)y2;
                        const int o2 = (int)(y / rows);
                        const int64_t yl = y - (int64_t)o2 * rows;
                        mytarget = reinterpret_cast<uint32_t*>(H.peer[(at2 >> 8) & 0xFFu].rec + pos2) + 1;
                        rec2 = H.peer[o2].rec;
                        m2 = H.use_masks ? mdp_mask((uint32_t)y, H.A, H.env_seed) : full;
                        if (m2 == 0u) atomicOr(L.err, kErrEmpty);
                        sg = __ldcg(H.peer[o2].seg + yl);
                        const float* row = H.peer[o2].q + (size_t)yl * H.ld;
#pragma unroll
                        for (int c = 0; c < LPR; ++c) rowv[c] = ld_row8(row + 8 * c);
                        fresh = busy = true;
                    }
                    pn += __popc(freeb);
                    if (pn >= cc) {
                        cb = cbn; pd = pdn; pn = 0;
                        cbn = cbnn;
                        pdn = load_chunk(cbn);
                        cbnn = claim_chunk();
                    }
                }
                if (busy && !fresh) {
                    bool fin = p >= pe;
                    if (!fin) {
                        const uint32_t pa = p & ~3u;
                        const U8 ea = ld_relaxed_sys_v8(rec2 + pa), eb = ld_relaxed_sys_v8(rec2 + pa + 4);
                        bool stop = false;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t ex = j < 4 ? ea.w[2 * j] : eb.w[2 * (j - 4)], et = j < 4 ? ea.w[2 * j + 1] : eb.w[2 * (j - 4) + 1];
                            if (!stop && !fin && pa + j >= p) {
                                if (pa + j >= pe || (int)(ex & 0xFFFFFFu) >= i) {
                                    fin = true;
                                } else {
                                    uint32_t tb = et;
                                    if (tb == kPending && (ex & (1u << 29))) {  // a self loop: derive its target in line
                                        const int ja = (int)(ex & 0xFFFFFFu), hr = ja / H.n_home;
                                        const float rj = __uint_as_float(__ldcg(reinterpret_cast<const uint32_t*>(H.peer[hr].tw + (ja - hr * H.n_home)) + 2));
                                        tb = __float_as_uint(td_target_s(rj, row_max(), H.gamma));
                                    }
                                    if (tb != kPending) {
                                        const uint32_t a2 = (ex >> 24) & 31u;
                                        if ((m2 >> a2) & 1u) {
                                            float* cell = myrow + a2;
                                            *cell = td_from_target_s(*cell, __uint_as_float(tb), lr);
                                        }
                                        ++p;
                                    } else {
                                        stop = true;
                                    }
                                }
                            }
                        }
                        if (!stop && p >= pe) fin = true;
                    }
                    if (fin) {
                        const float tg = td_target_s(r, row_max(), H.gamma);
                        st_relaxed_sys_u32(mytarget, __float_as_uint(tg));
                        busy = false;
                    }
                }
                if (fresh) {
#pragma unroll
                    for (int c = 0; c < LPR; ++c) {
                        float w8[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) w8[j] = ((m2 >> (8 * c + j)) & 1u) ? rowv[c].v[j] : -INFINITY;
                        reinterpret_cast<float4*>(myrow)[2 * c] = make_float4(w8[0], w8[1], w8[2], w8[3]);
                        reinterpret_cast<float4*>(myrow)[2 * c + 1] = make_float4(w8[4], w8[5], w8[6], w8[7]);
                    }
                    p = sg.x; pe = sg.y;
                    if (!(p < pe && pe <= (uint32_t)H.n_total)) {  // nobody stands on s': the untouched row is the answer
                        p = pe = 0u;
                        st_relaxed_sys_u32(mytarget, __float_as_uint(td_target_s(r, row_max(), H.gamma)));
                        busy = false;
                    }
                }
                if (cb >= nh && !__any_sync(kFull, busy)) break;
                if ((spins & 255u) == 255u) {
                    if (ld_relaxed_u32(L.ctr + 4) != 0u || global_ns() - t_start > kShardTimeoutNs) {
                        atomicExch(L.ctr + 4, 1u);
                        atomicOr(L.err, kErrTimeout);
                        break;
                    }
                }
            }
        }
        shard_xsync(grid, H, epoch);
        stamp(2);

        // ---------------- phase C: every owner commits the records it holds (as in qe_pipe.cuh)
        {
            uint2* rec = own.rec;
            const int n = n_in;
            const int ntiles = (n + 31) >> 5;
            for (int tile = rwarp; tile < ntiles; tile += rwarps) {
                const int p = tile * 32 + lane;
                const bool act = p < n;
                uint2 e = make_uint2(0u, 0u);
                uint32_t st = 0xFFFFFFFFu;
                if (act) { e = __ldcg(rec + p); st = (uint32_t)__ldcg(&sorted[p].x); }
                uint32_t prev = __shfl_up_sync(kFull, st, 1);
                if (lane == 0) prev = p > 0 ? (uint32_t)__ldcg(&sorted[p - 1].x) : 0xFFFFFFFFu;
                const bool head = act && (p == 0 || prev != st);
                const uint32_t hb = __ballot_sync(kFull, head);
                const uint32_t below = hb & (0xFFFFFFFFu >> (31 - lane));
                const int hl = below ? 31 - __clz(below) : -1;
                if (act && e.y == kPending) atomicOr(L.err, kErrTimeout);
                if (head) {
                    const float* row = own.q + (size_t)st * H.ld;
#pragma unroll
                    for (int c = 0; c < LPR; ++c) {
                        const F8 v8 = ld_row8(row + 8 * c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) s_row[(8 * c + j) * 256 + threadIdx.x] = v8.v[j];
                    }
                    s_touch[threadIdx.x] = 0u;
                }
                __syncwarp();
                const int off = (act && hl >= 0) ? lane - hl : -1;
                const int maxoff = (int)__reduce_max_sync(kFull, off);
                for (int it = 0; it <= maxoff; ++it) {
                    if (off == it) {
                        const int c = wbase + hl;
                        const uint32_t a = (e.x >> 24) & 31u;
                        float* cell = s_row + a * 256 + c;
                        *cell = td_from_target_s(*cell, __uint_as_float(e.y), lr);
                        s_touch[c] |= 1u << a;
                    }
                    __syncwarp();
                }
                const int hl31 = __shfl_sync(kFull, hl, 31);
                const uint32_t st31 = __shfl_sync(kFull, st, 31);
                if (hl31 >= 0 && tile * 32 + 32 < n) {
                    const int c = wbase + hl31;
                    float v = lane < 8 * LPR ? s_row[lane * 256 + c] : 0.0f;
                    bool touched = false;
                    for (int q = tile * 32 + 32; q < n; q += 32) {
                        uint2 e2 = make_uint2(0u, 0u);
                        uint32_t k2 = 0xFFFFFFFFu;
                        if (q + lane < n) { e2 = __ldcg(rec + q + lane); k2 = (uint32_t)__ldcg(&sorted[q + lane].x); }
                        const uint32_t diff = __ballot_sync(kFull, k2 != st31);
                        const int len = diff ? __ffs(diff) - 1 : 32;
                        for (int j = 0; j < len; ++j) {
                            const uint32_t xa = (__shfl_sync(kFull, e2.x, j) >> 24) & 31u;
                            const float tg = __uint_as_float(__shfl_sync(kFull, e2.y, j));
                            if ((uint32_t)lane == xa) { v = td_from_target_s(v, tg, lr); touched = true; }
                        }
                        if (len < 32) break;
                    }
                    const uint32_t tb = __ballot_sync(kFull, touched);
                    if (lane < 8 * LPR) s_row[lane * 256 + c] = v;
                    if (lane == 0) s_touch[c] |= tb;
                }
                __syncwarp();
                if (head) {
                    float* row = own.q + (size_t)st * H.ld;
                    for (uint32_t bm = s_touch[threadIdx.x]; bm; bm &= bm - 1u) {
                        const int a = __ffs(bm) - 1;
                        row[a] = s_row[a * 256 + threadIdx.x];
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        stamp(3);

        // ---------------- (sort, stages 4 and 5) the owner's local sort of its inbox, positions to the home ranks, bounds
        local_sort();
        shard_xsync(grid, H, epoch);
        stamp(6);
    }
    if (H.steps & 1) {
        for (int j = rtid; j < nh; j += rthreads) L.st_a[j] = L.st_b[j];
    }
}

}  // namespace qe
