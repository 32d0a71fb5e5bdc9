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
//   2. Dependencies always point to SMALLER agent indices.  Chunks of 32 agents are claimed in increasing order by
//      resident warps whose lanes simply poll: a predecessor was claimed earlier, so it is finished, in some lane, or
//      next in line for a warp whose lanes all hold smaller agents -- the smallest unfinished agent always makes
//      progress: no deferred records, no sweeps, no parking, one pass.  Most predecessors were processed long before
//      their dependants come up, so most polls succeed at once; a chain costs one L2 round trip per level.
//
//   Writers whose next state is their own row (self loops: the attractors a greedy policy herds the agents into --
//   one row of config 3 holds 1300 agents after 600 steps) would form a chain of one hop per agent; their targets are
//   derived in line during the replay (the replayed row IS the row they bootstrap from), so they cost no hop at all.
//
// Per vector step:
//   phase A   select + environment step (one lane per agent); every agent writes its writer record
//             {agent, action, target or "pending", reward, state} at its position in the sort of the CURRENT states
//   phase T   in-order target pipeline (above)
//   phase C   commit: one pass over the sorted records replays every row's segment and stores the cells that changed
//   phase S   stable LSD radix sort of the NEXT states (10-bit digits, all CTAs): positions and segment bounds per state
//             for the next step.  The order survives the launch: the next launch only checks that the states are
//             still the ones that were sorted.
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
    uint2* rec;           // [cap + 8] writer records by sorted position: x = agent | action << 24 | self << 29, y = target bits or kPending
    uint2* seg;           // [S] per state {segment start, segment end} in rec ({0, 0}: nobody stands on the state; the sort clears
                          //     the bounds of the previous order before it writes the new ones)
    int32_t* pos;         // [cap] agent -> sorted position
    int2* kv[2];          // [cap] radix ping-pong {key, agent}
    uint4* tw;            // [cap] per agent, from phase A to phase T: {next state, sorted position, reward bits, action | term << 7}
    int* ghist;           // [kRadix][blocks] digit counts per block, scanned in place
    int* rowtot;          // [kRadix] digit totals
    unsigned int* ctr;    // [64] 0: chunk claims of phase T; 4: abort flag; 6: "the states are not the sorted ones"; 8..: development counters
    int passes;           // radix passes for the state range
    int sorted_valid;     // pos / seg / kv hold the sort of the states this launch starts from (to be checked)
    int old_n;            // agents of the order kv[passes & 1] and seg[] still describe (0: none)
    int64_t state_base;   // keys are state - state_base (sharded tables)
};

// four 8-byte records with one 256-bit load (32-byte aligned; SASS LDG.E.ENL2.256.STRONG.GPU: one L1 wavefront per lane)
__device__ __forceinline__ U8 ld_relaxed_v8(const uint2* p) {
    U8 r;
    asm volatile("ld.relaxed.gpu.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 16 bytes global -> shared without passing through registers (L2 only: the source is rewritten between phases)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += y;
    }
    return v;
}
// lanes of the warp whose 10-bit digit equals this lane's.  match.any costs one hardware iteration per distinct value
// (~32 here); ten ballots are several times cheaper.
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

// Stable LSD radix sort of the agents by state, all CTAs.  Block b owns the contiguous chunk [b * chunk, ...) of the
// input of every pass, warp w of the block a contiguous part of it.  Per pass: per-warp digit histograms (shared
// memory) -> block counts ghist[digit][block]; grid barrier; exclusive scan of every digit's row over the blocks (one
// warp per digit, every lane a contiguous piece of the row); grid barrier; first free position per (digit, warp) =
// digit base + blocks before + warps before; the part is walked in order, 32 keys at a time, position = that + earlier
// lanes with the same digit; grid barrier.  Keys and agents travel as one 8-byte pair.  The last pass also writes
// pos[agent].  Returns the buffer that holds the sorted pairs; pipe_bounds() then gives every run of equal keys its
// bounds in seg[] (no barrier in between is needed by the caller's next phase unless it reads seg[]).
constexpr int kScanPerLane = 20;  // the column scan keeps ceil(blocks / 32) counts per lane in registers: up to 640 blocks
template <int WARPS>
__device__ __forceinline__ int pipe_sort(cg::grid_group& grid, int (*whist)[kRadix], int* s_base, int* s_wsum, const int32_t* states, int n,
                                         int old_n, const PipeScratch& X) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x, nb = gridDim.x;
    if (old_n > 0) {  // the previous order is still in kv[passes & 1] (nothing writes it before two barriers from here): reset its bounds
        const int2* old = X.kv[X.passes & 1];
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < old_n; q += gridDim.x * blockDim.x) {
            const int32_t kq = __ldcg(&old[q].x);
            if (q == 0 || __ldcg(&old[q - 1].x) != kq) X.seg[kq] = make_uint2(0u, 0u);
        }
    }
    const int chunk = ((n + nb - 1) / nb + 31) & ~31;          // per block, multiple of 32
    const int part = ((chunk / 32 + WARPS - 1) / WARPS) * 32;  // per warp, multiple of 32
    const int lo = min(b * chunk + warp * part, n), hi = min(min(b * chunk + (warp + 1) * part, (b + 1) * chunk), n);
    const int32_t bias = (int32_t)X.state_base;
    int32_t* posout = X.pos;
    int src = 0;
#ifdef QE_PIPE_STATS  // ctr[16 + 8 * pass + stage] += ns block 0 spent up to the end of the stage (since the previous lap)
    uint64_t t_lap = global_ns();
#define PIPE_LAP(ps, stage) do { if (b == 0 && threadIdx.x == 0 && (ps) < 3) { const uint64_t t_ = global_ns(); atomicAdd(X.ctr + 16 + 8 * (ps) + (stage), (unsigned int)(t_ - t_lap)); t_lap = t_; } } while (0)
#else
#define PIPE_LAP(ps, stage) ((void)0)
#endif
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
        for (int d = lane; d < kRadix; d += 32) whist[warp][d] = 0;
        __syncwarp();
        for (int base = lo; base < hi; base += 256) {  // eight loads in flight per lane
            int2 e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) e[u] = load_pair(base + 32 * u + lane);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (base + 32 * u + lane < hi) atomicAdd(&whist[warp][((uint32_t)e[u].x >> shift) & (kRadix - 1)], 1);
        }
        __syncthreads();
        for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += whist[w][d];
            X.ghist[(size_t)d * nb + b] = t;
        }
        PIPE_LAP(ps, 0);
        grid.sync();
        PIPE_LAP(ps, 1);
        {   // exclusive scan of every digit's row of block counts; one warp per digit, lane l owns entries [l * per, (l + 1) * per)
            const int gw = b * WARPS + warp, nw = nb * WARPS;
            const int per = (nb + 31) / 32;
            for (int d = gw; d < kRadix; d += nw) {
                int* row = X.ghist + (size_t)d * nb;
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
                if (lane == 31) X.rowtot[d] = incl;
            }
        }
        PIPE_LAP(ps, 2);
        grid.sync();
        PIPE_LAP(ps, 3);
        {   // digit bases (every block redoes the scan of the digit totals), then this warp's first free position per digit
            static_assert(kRadix == 4 * 256, "four digits per thread");
            const int4 v4 = __ldcg(reinterpret_cast<const int4*>(X.rowtot) + threadIdx.x);
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
        }
        __syncthreads();
        for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
            int run = s_base[d] + __ldcg(X.ghist + (size_t)d * nb + b);
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const int c = whist[w][d];
                whist[w][d] = run;
                run += c;
            }
        }
        __syncthreads();
        PIPE_LAP(ps, 4);
        for (int base = lo; base < hi; base += 256) {
            int2 e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) e[u] = load_pair(base + 32 * u + lane);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int x = base + 32 * u + lane;
                const bool act = x < hi;
                const uint32_t d = ((uint32_t)e[u].x >> shift) & (kRadix - 1);
                const uint32_t peers = digit_peers(d, act);
                if (act) {
                    const int p = whist[warp][d] + __popc(peers & ((1u << lane) - 1u));
                    out[p] = e[u];
                    if (last) posout[e[u].y] = p;
                }
                __syncwarp();
                if (act && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
                __syncwarp();
            }
        }
        PIPE_LAP(ps, 5);
        grid.sync();
        PIPE_LAP(ps, 6);
        src ^= 1;
    }
    return src;
}
// segment bounds: the first and the last position of every run of equal keys (neighbours by shuffle, 320 positions per batch)
template <int WARPS>
__device__ __forceinline__ void pipe_bounds(const int2* sorted, int n, const PipeScratch& X) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x, nb = gridDim.x;
    const int chunk = ((n + nb - 1) / nb + 31) & ~31;
    const int part = ((chunk / 32 + WARPS - 1) / WARPS) * 32;
    const int lo = min(b * chunk + warp * part, n), hi = min(min(b * chunk + (warp + 1) * part, (b + 1) * chunk), n);
    uint2* seg = X.seg;
    constexpr int U = 10;
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

// dynamic shared memory of fused_pipe_kernel<ENV, LPR>: the largest of the three phase layouts (see the kernel)
__host__ __device__ constexpr size_t pipe_smem_bytes(int lpr) {
    const size_t t = (size_t)(8 * lpr + 4) * 256, c = (size_t)(8 * lpr + 1) * 256, s = 8 * (size_t)kRadix + kRadix;
    return sizeof(float) * (t > s ? (t > c ? t : c) : (s > c ? s : c));
}
#ifndef QE_PIPE_MIN_BLOCKS
#define QE_PIPE_MIN_BLOCKS 3
#endif
template <int ENV, int LPR>
__global__ void __launch_bounds__(256, QE_PIPE_MIN_BLOCKS) fused_pipe_kernel(Table T, FusedArgs F, PipeScratch X) {
    cg::grid_group grid = cg::this_grid();
    constexpr int WARPS = 8;
    constexpr int RS = 8 * LPR + 4;                   // words per replayed row in phase T (one row per thread, 16-byte aligned, conflict-free)
    constexpr int kWordsT = RS * 256;
    constexpr int kWordsC = (8 * LPR + 1) * 256;      // phase C: one column per thread + the mask of cells it changed
    constexpr int kWordsS = WARPS * kRadix + kRadix;  // phase S: per-warp digit counters + digit bases
    constexpr int kWords = kWordsT > kWordsS ? (kWordsT > kWordsC ? kWordsT : kWordsC) : (kWordsS > kWordsC ? kWordsS : kWordsC);
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    __shared__ int s_wsum[WARPS];
    extern __shared__ __align__(16) float s_mem[];  // pipe_smem_bytes(LPR) of dynamic shared memory
    static_assert(kWords * sizeof(float) == pipe_smem_bytes(LPR), "host and device disagree about the shared memory size");
    int (*s_whist)[kRadix] = reinterpret_cast<int (*)[kRadix]>(s_mem);
    int* s_base = reinterpret_cast<int*>(s_mem) + WARPS * kRadix;
    float* s_row = s_mem;                                                     // phase C: [8*LPR][256]
    uint32_t* s_touch = reinterpret_cast<uint32_t*>(s_mem) + 8 * LPR * 256;   // phase C: [256]
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;
    const int ntiles = (n + 31) >> 5;
    const int gwarp = tid >> 5, nwarps = nthreads >> 5;
    const bool clk = F.phase_ns != nullptr && tid == 0;
    const int wbase = threadIdx.x & ~31;  // first thread of this warp inside the block
    const int32_t sbase = (int32_t)X.state_base;
    uint2* rec = X.rec;
    const int32_t* pos = X.pos;
    int old_n = X.old_n;
    if (clk) F.phase_ns[0] = global_ns();

    // ---------------- the order of the states this launch starts from: left behind by the previous launch (checked:
    // every agent must sit at its recorded position with its current state), else sorted now
    {
        const int2* sorted = X.kv[X.passes & 1];
        bool bad = !X.sorted_valid;
        if (!bad)
            for (int i = tid; i < n; i += nthreads) {
                const uint32_t q = (uint32_t)__ldcg(pos + i);
                if (q >= (uint32_t)n) { bad = true; continue; }
                const int2 e = __ldcg(sorted + q);
                bad |= e.y != i || e.x != __ldcg(F.st_a + i) - sbase;
            }
        if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(X.ctr + 6, 1u);
        grid.sync();
        if (ld_relaxed_u32(X.ctr + 6) != 0u) {  // (the same answer in every block)
            const int src = pipe_sort<WARPS>(grid, s_whist, s_base, s_wsum, F.st_a, n, old_n, X);
            pipe_bounds<WARPS>(X.kv[src], n, X);
            grid.sync();
        }
        old_n = n;
    }
    const int2* sorted = X.kv[X.passes & 1];  // {state - state_base, agent} by position, for the current step

    for (int k = 0; k < F.steps; ++k) {
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + environment step; every agent files its writer record.  (The state and the
        // position of the next tile are fetched while this one is processed.)
        {
            int base = tid & ~31;
            int s_nx = 0, pos_nx = 0;
            if (base + lane < n) { s_nx = cur[base + lane]; pos_nx = pos[base + lane]; }
            for (; base < n; base += nthreads) {
                const int i = base + lane;
                const bool active = i < n;
                const int s = s_nx, mypos = pos_nx;
                if (base + nthreads + lane < n) { s_nx = cur[base + nthreads + lane]; pos_nx = pos[base + nthreads + lane]; }
                uint32_t ew = 0u, valid = 0u, bits1 = 0u;
                bool explore = false;
                if (active) {
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
                    X.tw[i] = make_uint4((uint32_t)s2, (uint32_t)mypos, __float_as_uint(r), (uint32_t)a | (term ? 0x80u : 0u));
                    // a terminated agent bootstraps from nothing: its target is known here (QLO:760-766); bit 29: self loop
                    rec[mypos] = make_uint2((uint32_t)i | ((uint32_t)a << 24) | ((!term && s2 == s) ? (1u << 29) : 0u),
                                            term ? __float_as_uint(td_target_s(r, 0.0f, F.gamma)) : kPending);
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
        if (tid == 0) X.ctr[0] = 0u;  // phase T's chunk counter (idle since the last barrier of the previous step)
        grid.sync();
        if (clk && k < 10) F.phase_ns[1 + 3 * k] = global_ns();

        // ---------------- phase T: in-order target pipeline.  Chunks of 32 consecutive agents are claimed in increasing
        // order; a lane keeps its agent until the target is published and then takes the next agent of the warp's current
        // chunk, so one agent that waits for a predecessor holds up one lane, not a tile.  Two chunks are prefetched
        // (their per-agent words arrive while the lanes work); every pass of the loop polls up to four records per lane.
        // The replayed row lives in shared memory with its illegal cells at -inf, so the masked max is a plain max.
        {
            unsigned int* claim = X.ctr;
            const uint2* seg = X.seg - sbase;
            float* myrow = s_mem + threadIdx.x * RS;
            auto claim_chunk = [&]() {
                int c = 0;
                if (lane == 0) c = (int)atomicAdd(claim, 1u);
                return __shfl_sync(kFull, c, 0) * 32;
            };
            // per-agent words of a chunk ({s', position, reward, action | term << 7}): lane L holds agent (base + L)
            auto load_chunk = [&](int base) {
                uint4 q = make_uint4(0u, 0u, 0u, 0x80u);
                if (base + lane < n) q = __ldcg(X.tw + base + lane);
                return q;
            };
            int cb = claim_chunk(), cbn = claim_chunk(), cbnn = claim_chunk();
            uint4 pd = load_chunk(cb), pdn = load_chunk(cbn);
            int pn = 0;  // agents of the current chunk handed out so far
            int i = 0, y = 0, mypos = 0;
            float r = 0.0f;
            uint32_t m2 = 0u, p = 0u, pe = 0u;
            bool busy = false;
            uint32_t waits = 0, spins_total = 0;
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
                const uint32_t a2 = (ex >> 24) & 31u;
                if ((m2 >> a2) & 1u) {  // an illegal cell stays at -inf: it cannot be the masked max
                    float* cell = myrow + a2;
                    *cell = td_from_target_s(*cell, __uint_as_float(tb), lr);
                }
            };
            const uint64_t t_start = global_ns();
            for (uint32_t spins = 0;; ++spins) {
                spins_total = spins;
                // ---- hand the next agents of the chunk to the free lanes (in batches: the refill path is divergent code);
                // their row and segment bounds are fetched now and consumed at the end of this pass
                const uint32_t freeb = __ballot_sync(kFull, !busy);
                bool fresh = false;
                uint2 sg = make_uint2(0u, 0u);
                F8 rowv[LPR];
                if (cb < n && (__popc(freeb) >= 8 || (freeb != 0u && (spins & 3u) == 0u))) {
                    const int cc = min(32, n - cb);
                    const int src = pn + __popc(freeb & ((1u << lane) - 1u));
                    const uint32_t y2 = __shfl_sync(kFull, pd.x, src & 31), pos2 = __shfl_sync(kFull, pd.y, src & 31);
                    const uint32_t r2 = __shfl_sync(kFull, pd.z, src & 31), at2 = __shfl_sync(kFull, pd.w, src & 31);
                    if (!busy && src < cc && !(at2 & 0x80u)) {  // a terminated agent filed its target in phase A: nothing to do
                        i = cb + src; y = (int)y2; r = __uint_as_float(r2); mypos = (int)pos2;
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
                // ---- poll: the (up to eight) records of the two aligned 32-byte groups at the cursor, in position (= agent) order
                if (busy && !fresh) {
                    bool fin = p >= pe;
                    if (!fin) {
                        const uint32_t pa = p & ~3u;
                        const U8 ea = ld_relaxed_v8(rec + pa), eb = ld_relaxed_v8(rec + pa + 4);
                        bool stop = false;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t ex = j < 4 ? ea.w[2 * j] : eb.w[2 * (j - 4)], et = j < 4 ? ea.w[2 * j + 1] : eb.w[2 * (j - 4) + 1];
                            if (!stop && !fin && pa + j >= p) {
                                if (pa + j >= pe || (int)(ex & 0xFFFFFFu) >= i) {
                                    fin = true;  // the segment ends here, or the writers from here on come after i
                                } else {
                                    uint32_t tb = et;
                                    if (tb == kPending && (ex & (1u << 29)))  // a self loop: the row being replayed is the row it bootstraps from
                                        tb = __float_as_uint(td_target_s(__uint_as_float(__ldcg(reinterpret_cast<const uint32_t*>(X.tw + (ex & 0xFFFFFFu)) + 2)),
                                                                         row_max(), F.gamma));
                                    if (tb != kPending) {
                                        replay(ex, tb);
                                        ++p;
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
                    PIPE_STAT(12, spins_total);                         // passes of the loop
                    atomicMax(X.ctr + 14, (unsigned int)((global_ns() - t_start) >> 4));
                }
            }
#endif
            (void)waits; (void)spins_total;
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[2 + 3 * k] = global_ns();

        // ---------------- phase C: commit.  One pass over the sorted records: a segment's row lives in the shared-memory
        // column of its head lane, the members take turns in position (= agent) order; what extends beyond the tile is
        // replayed by the whole warp, one lane per action.
        for (int tile = gwarp; tile < ntiles; tile += nwarps) {
            const int p = tile * 32 + lane;
            const bool act = p < n;
            uint2 e = make_uint2(0u, 0u);
            uint32_t st = 0xFFFFFFFFu;  // key (state - state_base) of this position
            if (act) { e = __ldcg(rec + p); st = (uint32_t)__ldcg(&sorted[p].x); }
            uint32_t prev = __shfl_up_sync(kFull, st, 1);
            if (lane == 0) prev = p > 0 ? (uint32_t)__ldcg(&sorted[p - 1].x) : 0xFFFFFFFFu;
            const bool head = act && (p == 0 || prev != st);
            const uint32_t hb = __ballot_sync(kFull, head);
            const uint32_t below = hb & (0xFFFFFFFFu >> (31 - lane));
            const int hl = below ? 31 - __clz(below) : -1;  // head lane of this lane's segment; -1: the segment began in an earlier tile
            if (act && e.y == kPending) atomicOr(T.err, kErrTimeout);  // cannot happen: phase T published every target
            if (head) {
                const float* row = T.q + ((size_t)st + (size_t)sbase) * T.ld;
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
            // the tile's last segment may go on in the following tiles
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
                float* row = T.q + ((size_t)st + (size_t)sbase) * T.ld;
                for (uint32_t bm = s_touch[threadIdx.x]; bm; bm &= bm - 1u) {
                    const int a = __ffs(bm) - 1;
                    row[a] = s_row[a * 256 + threadIdx.x];
                }
            }
            __syncwarp();
        }
        if (clk && k < 10) F.phase_ns[32 + k] = global_ns();  // (this thread's end of phase C)

        // ---------------- phase S: the next step's order.  Its first stage only reads the next states, so no barrier is
        // needed after phase C; its last barrier orders pos[] before the next phase A, the bounds are needed by phase T.
        __syncthreads();  // the histograms share their shared memory with phase C's rows
        {
            const int src = pipe_sort<WARPS>(grid, s_whist, s_base, s_wsum, nxt, n, old_n, X);
            pipe_bounds<WARPS>(X.kv[src], n, X);
        }
        if (clk && k < 10) F.phase_ns[3 + 3 * k] = global_ns();
    }
    grid.sync();
    if (F.steps & 1) {
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
