// qe_pipe.cuh -- the pipelined form of the fused loop: the exact sequential TD update as an in-order pipeline over
// TARGETS (round 2).
//
// The reference applies the agents of a batch one after the other (QLO:806-817): agent i bootstraps from row s'_i as
// the agents j < i have left it.  Rounds 1's two forms resolve that by waiting for published VALUES (one publish ->
// poll hop per writer of a cell, 56 dependent hops per step on config 3) or by parking row sequencers.  This form uses
// two facts instead:
//
//   1. What agent i contributes to the table is fully described by its TARGET t_i = r_i + gamma * max_a Q_i[s'_i, a]:
//      the value it stores is v = p + lr * (t_i - p) with p the cell as it is at i's turn, and anybody who knows the
//      targets of the writers of a row (in agent order) can replay the row from its value at the start of the step.
//      So an agent never waits for a VALUE -- it reads the untouched row of s'_i, replays the targets of the writers
//      of s'_i that precede it (a prefix of that row's segment in a stable sort of the agents by state), takes the
//      masked max and publishes its own target.  A chain of c writers of one cell costs nothing; only a cross-row
//      read of a row with earlier writers is a dependency (26 levels instead of 56 on config 3, and flat later on).
//      Same floating-point operations in the same order as the reference: the result is bit-identical.
//   2. Dependencies always point to SMALLER agent indices.  Tiles of 32 agents are therefore claimed in increasing
//      order by resident warps which simply poll: a predecessor was claimed earlier, so it is finished or running --
//      no deferred records, no sweeps, no parking, one pass.  Most predecessors were processed long before their
//      dependants are claimed, so most polls succeed at once; a chain costs one L2 round trip per level.
//
//   Writers whose next state is their own row (self loops: the attractors a greedy policy herds the agents into --
//   one row of config 3 holds 1300 agents after 600 steps) would form a chain of one hop per agent; their targets are
//   derived in line during the replay (the replayed row IS the row they bootstrap from), so they cost no hop at all.
//
// Per vector step (three grid barriers):
//   phase A   select + environment step (one lane per agent); every agent writes its writer record
//             {agent, action, target or "pending", reward, state} at its position in the sort of the CURRENT states
//   phase T   in-order target pipeline (above)            ||   phase S  warp 0 of every CTA sorts the NEXT states (known
//             since the end of phase A) for the next step: stable LSD radix sort, 10-bit digits, one warp = one block
//             of the sort (its scattered stores use the store path of all SMs, its instructions 1/8 of their issue slots)
//   phase C   commit: one pass over the sorted records replays every row's segment and stores the cells that changed
//
// Requires the legal-action mask to be a function of the state (true for the device environments).
#pragma once
#include "qe_sorted.cuh"

namespace qe {

constexpr uint32_t kPending = 0xFFFFFFFFu;  // "target not published yet" (a NaN pattern no arithmetic produces)
constexpr uint64_t kPipeTimeoutNs = 3000000000ull;

#ifdef QE_PIPE_STATS  // development counters in ctr[8..] (read with qe_debug_counters)
#define PIPE_STAT(idx, val) atomicAdd(X.ctr + (idx), (unsigned int)(val))
#else
#define PIPE_STAT(idx, val) ((void)0)
#endif

struct PipeScratch {
    uint4* rec[2];        // [cap] writer records by sorted position, by step parity:
                          //   x = agent | action << 24, y = target bits or kPending, z = reward bits, w = state << 2 | term << 1 | self
    uint2* seg[2];        // [S] per state {segment start, segment end} in rec[parity]; stale unless rec[start].w >> 2 == state
    int32_t* pos[2];      // [cap] agent -> sorted position
    int2* kv[2];          // [cap] radix ping-pong {key, agent}
    int* ghist;           // [blocks][kRadix] digit counts per sorting warp, scanned in place
    int* rowtot;          // [kRadix] digit totals
    unsigned int* ctr;    // [64] 0,1: chunk claims by parity; 2: sorter barrier arrivals; 3: its generation; 4: abort flag
    int passes;           // radix passes for the state range
    int parity0;          // parity of the first step of the launch
    int64_t state_base;   // keys are state - state_base (sharded tables)
};

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const unsigned int* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += y;
    }
    return v;
}

// Barrier among the sorting warps (one per CTA; the other warps are busy with phase T).  Generation counter; all CTAs
// are co-resident (cooperative launch).  Gives up when the abort flag is set (a timeout somewhere: the launch is void).
__device__ __forceinline__ void sorter_barrier(unsigned int* ctr, int nb) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        const uint32_t gen = ld_acquire_u32(ctr + 3);
        __threadfence();
        if (atomicAdd(ctr + 2, 1u) == (unsigned int)(nb - 1)) {
            atomicExch(ctr + 2, 0u);
            __threadfence();
            atomicAdd(ctr + 3, 1u);
        } else {
            const uint64_t t0 = global_ns();
            uint32_t spins = 0;
            while (ld_acquire_u32(ctr + 3) == gen) {
                __nanosleep(100);
                if ((++spins & 1023u) == 0u && (ld_relaxed_u32(ctr + 4) != 0u || global_ns() - t0 > kPipeTimeoutNs)) {
                    atomicExch(ctr + 4, 1u);
                    break;
                }
            }
        }
        __threadfence();
    }
    __syncwarp();
}

// lanes of the warp whose 10-bit digit equals this lane's (idle lanes pass d >= kRadix and match nobody that matters).
// match.any costs one hardware iteration per distinct value (~32 here); ten ballots are several times cheaper.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, bool act) {
    uint32_t peers = __ballot_sync(kFull, act);
#pragma unroll
    for (int bit = 0; bit < kRadixBits; ++bit) {
        const bool one = (d >> bit) & 1u;
        const uint32_t b = __ballot_sync(kFull, one);
        peers &= one ? b : ~b;
    }
    return peers;
}

// Stable LSD radix sort of the agents by state, run by ONE WARP PER CTA (b = blockIdx.x of nb = gridDim.x): warp b owns
// the contiguous chunk [lo, hi) of the input of every pass.  Per pass: digit histogram of the chunk (shared memory) ->
// ghist[b][.]; barrier; exclusive scan of every digit's column over the warps (digits dealt round-robin to the warps);
// barrier; first free position per digit = digit base + warps before; the chunk is walked in order, 32 keys at a
// time, position = that + earlier lanes with the same digit; barrier.  Keys and agents travel as one 8-byte pair.  The
// last pass also writes pos[agent]; then every run of equal keys gets its bounds in seg[].  One warp has no latency
// hiding of its own: every stage issues its loads in batches before it consumes them.
__device__ __forceinline__ void pipe_sort_warp(int* hist, const int32_t* states, int n, int par, const PipeScratch& X, int b, int nb, int stat_slot = 16) {
    const int lane = threadIdx.x & 31;
#ifdef QE_PIPE_STATS  // ctr[stat_slot + 8 * pass + stage] = latest time (ns since this warp entered the sort) any warp finished the stage
    const uint64_t t_sort0 = global_ns();
#define PIPE_LAP(stage) do { if (lane == 0 && ps < 2) atomicMax(X.ctr + stat_slot + 8 * ps + (stage), (unsigned int)(global_ns() - t_sort0)); } while (0)
#else
#define PIPE_LAP(stage) ((void)0)
#endif
    (void)stat_slot;
    const int chunk = ((n + nb - 1) / nb + 31) & ~31;
    const int lo = min(b * chunk, n), hi = min(lo + chunk, n);
    const int32_t bias = (int32_t)X.state_base;
    int src = 0;
    for (int ps = 0; ps < X.passes; ++ps) {
        const bool first = ps == 0, last = ps == X.passes - 1;
        const int shift = ps * kRadixBits;
        const int2* in = X.kv[src];
        int2* out = X.kv[src ^ 1];
        auto load_pair = [&](int x) {
            int2 e = make_int2(0, 0);
            if (x < hi) {
                if (first) e = make_int2(__ldcg(states + x) - bias, x);
                else e = __ldcg(in + x);
            }
            return e;
        };
        for (int d = lane; d < kRadix; d += 32) hist[d] = 0;
        __syncwarp();
        for (int base = lo; base < hi; base += 256) {  // eight loads in flight per lane
            int32_t kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int x = base + 32 * u + lane;
                kk[u] = x < hi ? (first ? __ldcg(states + x) - bias : __ldcg(&in[x].x)) : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (base + 32 * u + lane < hi) atomicAdd(&hist[((uint32_t)kk[u] >> shift) & (kRadix - 1)], 1);
        }
        __syncwarp();
        int* mine = X.ghist + (size_t)b * kRadix;
        for (int d = lane; d < kRadix; d += 32) mine[d] = hist[d];
        PIPE_LAP(0);
        sorter_barrier(X.ctr, nb);
        PIPE_LAP(1);
        for (int d = b; d < kRadix; d += nb) {  // exclusive scan of digit d's column over the warps
            int carry = 0;
            for (int x0 = 0; x0 < nb; x0 += 256) {
                int v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u + lane;
                    v[u] = x < nb ? __ldcg(X.ghist + (size_t)x * kRadix + d) : 0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int x = x0 + 32 * u + lane;
                    const int incl = warp_incl_scan(v[u]);
                    if (x < nb) X.ghist[(size_t)x * kRadix + d] = carry + incl - v[u];
                    carry += __shfl_sync(kFull, incl, 31);
                }
            }
            if (lane == 0) X.rowtot[d] = carry;
        }
        PIPE_LAP(2);
        sorter_barrier(X.ctr, nb);
        PIPE_LAP(3);
        {   // first free position per digit for this warp
            int carry = 0;
            for (int d0 = 0; d0 < kRadix; d0 += 256) {
                int v[8], w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    v[u] = __ldcg(X.rowtot + d0 + 32 * u + lane);
                    w[u] = __ldcg(mine + d0 + 32 * u + lane);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int incl = warp_incl_scan(v[u]);
                    hist[d0 + 32 * u + lane] = carry + incl - v[u] + w[u];
                    carry += __shfl_sync(kFull, incl, 31);
                }
            }
        }
        __syncwarp();
        PIPE_LAP(4);
        int2 e[4], en[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) en[u] = load_pair(lo + 32 * u + lane);
        for (int base = lo; base < hi; base += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) { e[u] = en[u]; en[u] = load_pair(base + 128 + 32 * u + lane); }  // the next batch travels under this one
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = base + 32 * u + lane;
                const bool act = x < hi;
                const uint32_t d = ((uint32_t)e[u].x >> shift) & (kRadix - 1);
                const uint32_t peers = digit_peers(d, act);
                if (act) {
                    const int p = hist[d] + __popc(peers & ((1u << lane) - 1u));
                    out[p] = e[u];
                    if (last) X.pos[par][e[u].y] = p;
                }
                __syncwarp();
                if (act && lane == (__ffs(peers) - 1)) hist[d] += __popc(peers);
                __syncwarp();
            }
        }
        PIPE_LAP(5);
        sorter_barrier(X.ctr, nb);
        PIPE_LAP(6);
        src ^= 1;
    }
    // segment bounds: the first and the last position of every run of equal keys (neighbours by shuffle, 256 positions per batch)
    const int2* sorted = X.kv[src];
    uint2* seg = X.seg[par];
    for (int base = lo; base < hi; base += 256) {
        int32_t kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int x = base + 32 * u + lane;
            kk[u] = x < n ? __ldcg(&sorted[x].x) : -1;
        }
        const int32_t left = base > 0 ? __ldcg(&sorted[base - 1].x) : -1;
        const int32_t right = base + 256 < n ? __ldcg(&sorted[base + 256].x) : -1;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int x = base + 32 * u + lane;
            int32_t prev = __shfl_up_sync(kFull, kk[u], 1), next = __shfl_down_sync(kFull, kk[u], 1);
            const int32_t pl = __shfl_sync(kFull, kk[u > 0 ? u - 1 : 0], 31), nf = __shfl_sync(kFull, kk[u < 7 ? u + 1 : 7], 0);
            if (lane == 0) prev = u > 0 ? pl : left;
            if (lane == 31) next = u < 7 ? nf : right;
            if (x < hi) {
                if (prev != kk[u]) seg[kk[u]].x = (uint32_t)x;
                if (next != kk[u]) seg[kk[u]].y = (uint32_t)(x + 1);
            }
        }
    }
}

#ifndef QE_PIPE_MIN_BLOCKS
#define QE_PIPE_MIN_BLOCKS 4
#endif
template <int ENV, int LPR>
__global__ void __launch_bounds__(256, QE_PIPE_MIN_BLOCKS) fused_pipe_kernel(Table T, FusedArgs F, PipeScratch X) {
    cg::grid_group grid = cg::this_grid();
    constexpr int RS = 8 * LPR + 4;                   // words per replayed row in phase T (one row per thread, 16-byte aligned, conflict-free)
    constexpr int kRowWordsT = RS * 256;
    constexpr int kRowWordsC = (8 * LPR + 1) * 256;   // phase C: one column per thread + the mask of cells it changed
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    __shared__ __align__(16) float s_rows[kRowWordsT > kRowWordsC ? kRowWordsT : kRowWordsC];
    __shared__ int s_hist[kRadix];
    float* s_row = s_rows;                                                      // phase C: [8*LPR][256]
    uint32_t* s_touch = reinterpret_cast<uint32_t*>(s_rows) + 8 * LPR * 256;    // phase C: [256]
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;
    const int ntiles = (n + 31) >> 5;
    const int gwarp = tid >> 5, nwarps = nthreads >> 5;
    const bool clk = F.phase_ns != nullptr && tid == 0;
    const int wbase = threadIdx.x & ~31;  // first thread of this warp inside the block
    const bool sorter = threadIdx.x < 32; // warp 0 of every CTA sorts
    if (clk) F.phase_ns[0] = global_ns();

    // the sort of the current states (steady state: done during the previous step's phase T)
    if (sorter) pipe_sort_warp(s_hist, F.st_a, n, X.parity0, X, blockIdx.x, gridDim.x);
    grid.sync();

    for (int k = 0; k < F.steps; ++k) {
        const int par = (X.parity0 + k) & 1;
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        uint4* rec = X.rec[par];
        const int32_t* pos = X.pos[par];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + environment step; every agent files its writer record
        for (int base = (tid & ~31); base < n; base += nthreads) {
            const int i = base + lane;
            const bool active = i < n;
            int s = 0, mypos = 0;
            uint32_t ew = 0u, valid = 0u, bits1 = 0u;
            bool explore = false;
            if (active) {
                s = cur[i];
                mypos = pos[i];
                if (ENV != 0) ew = F.envw[i];
                valid = F.use_masks ? env_mask<ENV>(s, ew, T.A, F.env_seed) : full;
                explore = (uint64_t)U.draw(i, 0) < thresh;
                bits1 = U.draw(i, 1);
            }
            RowGather<LPR> rows;
            rows.issue(T, s, active);
            float mx;
            uint32_t tie;
            rows.row_max_tie(valid, mx, tie);
            int a = pick_action(T.A, valid, tie, explore, F.empty_all != 0, bits1);
            if (active && a < 0) { atomicOr(T.err, kErrEmpty); a = 0; }
            a = max(a, 0);
            if (active) {
                int32_t s2 = s;
                float r = 0.0f;
                bool term = false;
                if (ENV == 0) mdp_step(s2, a, (uint32_t)F.S, T.A, F.env_seed, F.term_thresh, U.draw(i, 2), U.draw(i, 3), r, term);
                else if (ENV == 1) {
                    if (!ttt_step(ew, a, U.draw(i, 2), U.draw(i, 3), U.draw(i, 4), r, term)) atomicOr(T.err, kErrInvalidMove);
                    s2 = ttt_state(ew & 0x3FFFFu);
                } else {
                    r = (float)a;
                    ew += 1u;
                    term = ew >= F.episode_len;
                    if (term) ew = 0u;
                    s2 = 0;
                }
                nxt[i] = s2;
                if (ENV != 0) F.envw[i] = ew;
                F.tr_a[i] = (uint8_t)(a | (term ? 0x80 : 0));
                F.tr_r[i] = r;
                {
                    // a terminated agent bootstraps from nothing: its target is known here (QLO:760-766)
                    const uint32_t tg = term ? __float_as_uint(td_target_s(r, 0.0f, F.gamma)) : kPending;
                    const uint32_t fl = ((uint32_t)s << 2) | (term ? 2u : 0u) | ((!term && s2 == s) ? 1u : 0u);
                    rec[mypos] = make_uint4((uint32_t)i | ((uint32_t)a << 24), tg, __float_as_uint(r), fl);
                }
                float acc = F.ep_ret[i] + r;
                float fin = __int_as_float(0x7FC00000);
                if (term) { fin = acc; loc_sum += (double)acc; ++loc_cnt; acc = 0.0f; }
                F.ep_ret[i] = acc;
                const size_t o = (size_t)k * n + i;
                if (F.trace_actions) F.trace_actions[o] = a;
                if (F.trace_rewards) F.trace_rewards[o] = r;
                if (F.trace_term) F.trace_term[o] = term;
                if (F.trace_next) F.trace_next[o] = s2;
                if (F.trace_epret) F.trace_epret[o] = fin;
            }
            __syncwarp();
        }
        if (F.ep_count) {
            for (int d = 16; d > 0; d >>= 1) {
                loc_sum += __shfl_xor_sync(kFull, loc_sum, d);
                loc_cnt += __shfl_xor_sync(kFull, loc_cnt, d);
            }
            if (lane == 0) { s_sum[threadIdx.x >> 5] = loc_sum; s_cnt[threadIdx.x >> 5] = loc_cnt; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double bs = 0.0;
                unsigned int bc = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
                if (bc) { atomicAdd(F.ep_sum, bs); atomicAdd(F.ep_count, (unsigned long long)bc); }
            }
            __syncthreads();
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[1 + 3 * k] = global_ns();

        // ---------------- phase S (warp 0 of every CTA): the next step's sort, hidden under phase T
        if (sorter && k + 1 < F.steps) {
            const uint64_t ts0 = global_ns();
            pipe_sort_warp(s_hist, nxt, n, par ^ 1, X, blockIdx.x, gridDim.x, 32);
            if (blockIdx.x == 0 && lane == 0) PIPE_STAT(10, (global_ns() - ts0));  // ns spent sorting
            (void)ts0;
        }

        // ---------------- phase T: in-order target pipeline.  Chunks of 32 consecutive agents are claimed in increasing
        // order; a lane keeps its agent until the target is published and then takes the next agent of the warp's current
        // chunk, so one agent that waits for a predecessor holds up one lane, not a tile.  Two chunks are prefetched
        // (their per-agent words arrive while the lanes work); every pass of the loop polls up to four records per lane.
        // The replayed row lives in shared memory with its illegal cells at -inf, so the masked max is a plain max.
        {
            unsigned int* claim = X.ctr + par;
            const uint2* seg = X.seg[par] - X.state_base;
            float* myrow = s_rows + threadIdx.x * RS;
            auto claim_chunk = [&]() {
                int c = 0;
                if (lane == 0) c = (int)atomicAdd(claim, 1u);
                return __shfl_sync(kFull, c, 0) * 32;
            };
            struct Pend { int y, pos; float r; uint32_t at; };  // per-agent words of a chunk: lane L holds agent (base + L)
            auto load_chunk = [&](int base) {
                Pend q;
                const int j = base + lane < n ? base + lane : 0;
                q.y = nxt[j]; q.pos = pos[j]; q.r = F.tr_r[j]; q.at = F.tr_a[j];
                return q;
            };
            int cb = claim_chunk(), cbn = claim_chunk(), cbnn = claim_chunk();
            Pend pd = load_chunk(cb), pdn = load_chunk(cbn);
            int pn = 0;  // agents of the current chunk handed out so far
            int i = 0, y = 0, mypos = 0;
            float r = 0.0f;
            uint32_t m2 = 0u, p = 0u, pe = 0u;
            bool busy = false, first = false;
            uint32_t waits = 0;
            auto row_max = [&]() {
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2 * LPR; ++c) {
                    const float4 v = reinterpret_cast<const float4*>(myrow)[c];
                    m = fmaxf(fmaxf(fmaxf(fmaxf(m, v.x), v.y), v.z), v.w);
                }
                return m;
            };
            auto replay = [&](uint32_t ex, uint32_t tb) {
                const uint32_t a2 = ex >> 24;
                if ((m2 >> a2) & 1u) {  // an illegal cell stays at -inf: it cannot be the masked max
                    float* cell = myrow + a2;
                    *cell = td_from_target_s(*cell, __uint_as_float(tb), lr);
                }
            };
            const uint64_t t_start = global_ns();
            for (uint32_t spins = 0;; ++spins) {
                // ---- hand the next agents of the chunk to the free lanes (in batches: the refill path is divergent code);
                // their row and segment bounds are fetched now and consumed at the end of this pass
                const uint32_t freeb = __ballot_sync(kFull, !busy);
                bool fresh = false;
                uint2 sg = make_uint2(0u, 0u);
                F8 rowv[LPR];
                if (cb < n && (__popc(freeb) >= 8 || (freeb != 0u && (spins & 3u) == 0u))) {
                    const int cc = min(32, n - cb);
                    const int src = pn + __popc(freeb & ((1u << lane) - 1u));
                    const int y2 = __shfl_sync(kFull, pd.y, src & 31), pos2 = __shfl_sync(kFull, pd.pos, src & 31);
                    const float r2 = __shfl_sync(kFull, pd.r, src & 31);
                    const uint32_t at2 = __shfl_sync(kFull, pd.at, src & 31);
                    if (!busy && src < cc && !(at2 & 0x80u)) {  // a terminated agent filed its target in phase A: nothing to do
                        i = cb + src; y = y2; r = r2; mypos = pos2;
                        m2 = F.use_masks ? state_mask<ENV>(y, T.A, F.env_seed, full) : full;
                        if (m2 == 0u) atomicOr(T.err, kErrEmpty);  // np.max of an empty selection (QLO:764)
                        sg = __ldcg(seg + y);
                        const float* row = T.q + (size_t)y * T.ld;
#pragma unroll
                        for (int c = 0; c < LPR; ++c) rowv[c] = ld_row8(row + 8 * c);
                        fresh = busy = true;
                    }
                    pn += __popc(freeb);
                    if (pn >= cc) {  // chunk exhausted: the prefetched one becomes current, the next one starts loading
                        cb = cbn; pd = pdn; pn = 0;
                        cbn = cbnn;
                        pdn = load_chunk(cbn);
                        cbnn = claim_chunk();
                    }
                }
                // ---- poll: up to four records of the segment of s' per pass, in position (= agent) order
                if (busy && !fresh) {
                    bool fin = p >= pe;
                    if (!fin) {
                        const uint32_t ne = min(pe - p, 4u);
                        uint4 e[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) e[j] = ld_relaxed_v4(rec + p + min((uint32_t)j, ne - 1u));
                        bool stop = false;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (!stop && !fin) {
                                if ((uint32_t)j >= ne) {
                                    fin = true;
                                } else if ((first && (e[j].w >> 2) != (uint32_t)y) || (int)(e[j].x & 0xFFFFFFu) >= i) {
                                    fin = true;  // stale bounds (nobody stands on s' in this step), or the writers from here on come after i
                                } else {
                                    uint32_t tb = e[j].y;
                                    if (tb == kPending && (e[j].w & 1u))  // a self loop: the row being replayed is the row it bootstraps from
                                        tb = __float_as_uint(td_target_s(__uint_as_float(e[j].z), row_max(), F.gamma));
                                    if (tb != kPending) {
                                        replay(e[j].x, tb);
                                        ++p;
                                        first = false;
                                    } else {
                                        stop = true;
                                        ++waits;
                                    }
                                }
                            }
                        }
                        if (!stop && p >= pe) fin = true;
                    }
                    if (fin) {
                        const float tg = td_target_s(r, row_max(), F.gamma);
                        st_relaxed_u32(reinterpret_cast<uint32_t*>(rec + mypos) + 1, __float_as_uint(tg));
                        busy = false;
                    }
                }
                // ---- the lanes that took an agent in this pass: masked row into shared memory, cursor at the segment's start
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
                    if (!(p < pe && pe <= (uint32_t)n)) p = pe = 0u;
                    first = true;
                }
                if (cb >= n && !__any_sync(kFull, busy)) break;
                if ((spins & 255u) == 255u) {
                    if (ld_relaxed_u32(X.ctr + 4) != 0u || global_ns() - t_start > kPipeTimeoutNs) {
                        atomicExch(X.ctr + 4, 1u);
                        atomicOr(T.err, kErrTimeout);
                        break;
                    }
                }
            }
#ifdef QE_PIPE_STATS
            {
                const uint32_t wsum = __reduce_add_sync(kFull, waits);
                if (lane == 0) {
                    PIPE_STAT(8, 1);                                   // warps
                    PIPE_STAT(9, (global_ns() - t_start) >> 4);         // time in phase T, 16 ns units
                    PIPE_STAT(11, wsum);                                // failed polls (lanes)
                    atomicMax(X.ctr + 14, (unsigned int)((global_ns() - t_start) >> 4));
                }
            }
#endif
            (void)waits;
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[2 + 3 * k] = global_ns();

        // ---------------- phase C: commit.  One pass over the sorted records: a segment's row lives in the shared-memory
        // column of its head lane, the members take turns in position (= agent) order; what extends beyond the tile is
        // replayed by the whole warp, one lane per action.
        for (int tile = gwarp; tile < ntiles; tile += nwarps) {
            const int p = tile * 32 + lane;
            const bool act = p < n;
            uint4 e = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
            if (act) e = __ldcg(rec + p);
            const uint32_t st = e.w >> 2;
            uint32_t prev = __shfl_up_sync(kFull, st, 1);
            if (lane == 0) prev = p > 0 ? (__ldcg(reinterpret_cast<const uint32_t*>(rec + p - 1) + 3) >> 2) : 0xFFFFFFFFu;
            const bool head = act && (p == 0 || prev != st);
            const uint32_t hb = __ballot_sync(kFull, head);
            const uint32_t below = hb & (0xFFFFFFFFu >> (31 - lane));
            const int hl = below ? 31 - __clz(below) : -1;  // head lane of this lane's segment; -1: the segment began in an earlier tile
            if (act && e.y == kPending) atomicOr(T.err, kErrTimeout);  // cannot happen: phase T published every target
            if (head) {
                const float* row = T.q + (size_t)st * T.ld;
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
                    const uint32_t a = e.x >> 24;
                    float* cell = s_row + a * 256 + c;
                    *cell = td_from_target_s(*cell, __uint_as_float(e.y), lr);
                    s_touch[c] |= 1u << a;
                }
                __syncwarp();
            }
            // the tile's last segment may go on in the following tiles
            const int hl31 = __shfl_sync(kFull, hl, 31);
            const uint32_t st31 = __shfl_sync(kFull, st, 31);
            if (hl31 >= 0 && tile * 32 + 32 < n) {
                const int c = wbase + hl31;
                float v = lane < 8 * LPR ? s_row[lane * 256 + c] : 0.0f;
                bool touched = false;
                for (int q = tile * 32 + 32; q < n; q += 32) {
                    uint4 e2 = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
                    if (q + lane < n) e2 = __ldcg(rec + q + lane);
                    const uint32_t diff = __ballot_sync(kFull, (e2.w >> 2) != st31);
                    const int len = diff ? __ffs(diff) - 1 : 32;
                    for (int j = 0; j < len; ++j) {
                        const uint32_t xa = __shfl_sync(kFull, e2.x, j) >> 24;
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
                float* row = T.q + (size_t)st * T.ld;
                for (uint32_t bm = s_touch[threadIdx.x]; bm; bm &= bm - 1u) {
                    const int a = __ffs(bm) - 1;
                    row[a] = s_row[a * 256 + threadIdx.x];
                }
            }
            __syncwarp();
        }
        if (tid == 0) X.ctr[par] = 0u;  // this parity's claim counter is idle until the step after the next one
        grid.sync();
        if (clk && k < 10) F.phase_ns[3 + 3 * k] = global_ns();
    }
    if (F.steps & 1) {
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
