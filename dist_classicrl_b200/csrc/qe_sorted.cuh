// qe_sorted.cuh -- the sorted form of the fused loop: the exact sequential TD update as segments of a per-step sort.
//
// Why.  The bundled environments are deterministic in (s, a), so a greedy policy herds agents: rows with hundreds of
// writers appear as the table is learned (C3 after 800 vector steps: half of the agents share a row with more than
// four others, one row has 3333 writers).  Per-agent scans of per-row writer lists are then quadratic in the crowd and
// same-cell chains resolve one publish-poll round trip per agent.  Sorting the agents by state every step turns each
// row's writers into one contiguous segment in agent order:
//
//   phase A   select + environment step (one lane per agent; nothing is registered anywhere)
//   phase S   stable LSD radix sort of (state -> agent) : perm[] (agents by (state, agent)), rank[] (its inverse),
//             segment bounds per row
//   phase R   readers, one lane per agent: the bootstrap of agent i from row y = s'_i is the masked row max of y after
//             all writers of y that precede i.  No earlier writer (binary search in y's segment) -> the table row is
//             still untouched, take its masked max and deposit the target t_i = r_i + gamma*m at i's sorted position.
//             Otherwise the reader is deferred: it needs mhist[p], the row max after the writer at sorted position p.
//   phase Q   sequencers, one lane per row: walk the segment in order, v = Q[s,a] + lr*(t - Q[s,a]) from the deposited
//             targets, publish the row max after every writer in mhist[], park when a target is missing; deferred
//             readers poll mhist[] and deposit.  Both run as non-blocking sweeps until everything has drained.
//
// A chain of c writers of one cell costs its sequencer c steps from registers / shared memory instead of c round trips
// through L2, and a reader costs O(log c).  Requires the legal-action mask to be a function of the state (true for the
// device environments: hash MDP, TicTacToe, bandit); the general per-agent-mask update keeps the writer-list kernels.
#pragma once
#include "qe_kernels.cuh"

namespace qe {

constexpr int kRadixBits = 10;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortStride = 4;        // every 4th block streams the keys of the sort
constexpr int kSortMaxBlocks = 256;   // sorting blocks: the digit scan keeps one row of block counts in a warp's registers

struct SortedScratch {
    int32_t* key[2];      // [cap] ping-pong keys (states)
    int32_t* val[2];      // [cap] ping-pong values (agent indices)
    int* ghist;           // [kRadix][grid] digit counts per block, scanned in place
    int* rowtot;          // [kRadix] digit totals
    int32_t* rank;        // [cap] sorted position of every agent
    uint64_t* targ;       // [cap] by sorted position: (epoch << 6 | action) << 32 | target bits
    uint64_t* mhist;      // [cap] by sorted position: epoch << 32 | masked row max after this writer
    uint32_t* seg;        // [S][4] per state: {segment start, segment end, epoch of both, -}
    uint32_t* rrec;       // [cap][4] deferred readers: {mhist position, own sorted position, reward bits, action}
    uint32_t* rmask;      // [cap/32] per agent tile: deferred readers
    uint32_t* hmask;      // [cap/32] per tile of sorted positions: segment heads whose sequencer has not finished
    uint32_t* hrec;       // [cap][4] at a segment's head position: {state, segment end, sequencer progress, -}
    int passes;           // radix passes needed for the state range
    unsigned long long* dbg;  // [8] development counters (phase Q), accumulated over the launch
};

__device__ __forceinline__ uint32_t ttt_mask_of_state(int s) {  // empty cells of the board a base-3 state id encodes
    uint32_t m = 0;
#pragma unroll
    for (int c = 8; c >= 0; --c) {
        if (s % 3 == 0) m |= 1u << c;
        s /= 3;
    }
    return m;
}
__device__ __forceinline__ float td_target_s(float r, float m, float gamma) { return __fadd_rn(r, __fmul_rn(gamma, m)); }
__device__ __forceinline__ float td_from_target_s(float p, float target, float lr) {  // QLO:766-768 after the target
    return __fadd_rn(p, __fmul_rn(lr, __fsub_rn(target, p)));
}
template <int ENV>
__device__ __forceinline__ uint32_t state_mask(int s, int A, uint32_t env_seed, uint32_t full) {
    if (ENV == 0) return mdp_mask((uint32_t)s, A, env_seed);
    if (ENV == 1) return ttt_mask_of_state(s);
    return full;
}

// ------------------------------------------------------------------ phase S: stable LSD radix sort, inside the grid
// Every block owns one contiguous chunk of the input, every warp one contiguous part of the chunk.  Per pass: per-warp
// digit histograms (shared memory) -> block counts to global; grid barrier; per-digit exclusive scan over the blocks
// (one warp per digit); grid barrier; every warp walks its part in order, 32 keys at a time: position = digit base +
// blocks before + warps before + earlier keys of the part + earlier lanes with the same digit (match.any).
template <int WARPS>
__device__ __forceinline__ void radix_pass(cg::grid_group& grid, int (*whist)[kRadix], int* s_base, const int32_t* kin, const int32_t* vin,
                                           int32_t* kout, int32_t* vout, int32_t* rank_out, int n, int shift, int* ghist, int* rowtot,
                                           bool iota_vals) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // One block in kSortStride streams the keys (the matrix of per-block digit counts, which is touched with a stride
    // three times per pass, stays small); all blocks help with the digit scan.
    const bool sorter = (blockIdx.x % kSortStride) == 0;
    const int nb = (gridDim.x + kSortStride - 1) / kSortStride, b = blockIdx.x / kSortStride;
    const int chunk = ((n + nb - 1) / nb + 31) & ~31;          // per sorting block, multiple of 32
    const int part = ((chunk / 32 + WARPS - 1) / WARPS) * 32;  // per warp, multiple of 32
    int lo = min(b * chunk + warp * part, n), hi = min(min(b * chunk + (warp + 1) * part, (b + 1) * chunk), n);
    if (!sorter) lo = hi = 0;
    for (int d = lane; d < kRadix; d += 32) whist[warp][d] = 0;
    __syncwarp();
    for (int base = lo; base < hi; base += 128) {  // four loads in flight per lane
        int32_t kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int x = base + 32 * u + lane;
            kk[u] = x < hi ? kin[x] : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (base + 32 * u + lane < hi) atomicAdd(&whist[warp][((uint32_t)kk[u] >> shift) & (kRadix - 1)], 1);
    }
    __syncthreads();
    if (sorter)
        for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) t += whist[w][d];
            ghist[d * nb + b] = t;
        }
    grid.sync();
    {   // exclusive scan of every digit's row of block counts; one warp per digit, the whole row in registers
        const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
        constexpr int kPerLane = kSortMaxBlocks / 32;
        const int per = (nb + 31) / 32;
        for (int d = gw; d < kRadix; d += nw) {
            int* row = ghist + d * nb;
            int v[kPerLane];
            int sum = 0;
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) {
                const int x = lane * per + j;
                v[j] = (j < per && x < nb) ? row[x] : 0;
            }
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) sum += v[j];
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += y;
            }
            int run = incl - sum;
#pragma unroll
            for (int j = 0; j < kPerLane; ++j) {
                const int x = lane * per + j;
                if (j < per && x < nb) row[x] = run;
                run += v[j];
            }
            if (lane == 31) rowtot[d] = incl;
        }
    }
    grid.sync();
    // digit bases (every block redoes the scan of the digit totals), then this warp's first free position per digit
    {
        constexpr int kPer = kRadix / 256;  // digits per thread (blockDim.x == 256)
        int v[kPer];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { v[j] = rowtot[threadIdx.x * kPer + j]; sum += v[j]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += y;
        }
        __shared__ int s_wsum[WARPS];
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) before += (w < warp) ? s_wsum[w] : 0;
        int run = before + incl - sum;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { s_base[threadIdx.x * kPer + j] = run; run += v[j]; }
    }
    __syncthreads();
    if (sorter)
    for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
        int run = s_base[d] + ghist[d * nb + b];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const int c = whist[w][d];
            whist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int base = lo; base < hi; base += 128) {
        int32_t kk[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int x = base + 32 * u + lane;
            kk[u] = x < hi ? kin[x] : 0;
            vv[u] = x < hi ? (iota_vals ? x : vin[x]) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int x = base + 32 * u + lane;
            const bool act = x < hi;
            const uint32_t d = act ? (((uint32_t)kk[u] >> shift) & (kRadix - 1)) : (uint32_t)kRadix + lane;  // idle lanes match nobody
            const uint32_t peers = __match_any_sync(kFull, d);
            if (act) {
                const int pos = whist[warp][d] + __popc(peers & ((1u << lane) - 1u));
                kout[pos] = kk[u];
                vout[pos] = vv[u];
                if (rank_out) rank_out[vv[u]] = pos;
            }
            __syncwarp();
            if (act && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
            __syncwarp();
        }
    }
    grid.sync();
}

// development aid: cost of one grid barrier at the fused loop's launch shape
static __global__ void __launch_bounds__(256, 4) gridsync_probe_kernel(int iters, uint64_t* out) {
    cg::grid_group grid = cg::this_grid();
    const uint64_t t0 = global_ns();
    for (int i = 0; i < iters; ++i) grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = global_ns() - t0;
}

// ------------------------------------------------------------------ the kernel
template <int ENV, int LPR>
__global__ void __launch_bounds__(256, 4) fused_sorted_kernel(Table T, FusedArgs F, SortedScratch X) {
    cg::grid_group grid = cg::this_grid();
    constexpr int WARPS = 8;
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    // the sort (per-warp digit counters) and the sequencers (one row per thread) never run at the same time
    constexpr int kSortWords = WARPS * kRadix, kValWords = 8 * LPR * 256;
    __shared__ int s_union[kSortWords > kValWords ? kSortWords : kValWords];
    __shared__ int s_base[kRadix];
    int (*s_whist)[kRadix] = reinterpret_cast<int (*)[kRadix]>(s_union);
    float* s_vals = reinterpret_cast<float*>(s_union);
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;
    const int ntiles = (n + 31) >> 5;
    const int gwarp = tid >> 5, nwarps = nthreads >> 5;
    const bool clk = F.phase_ns != nullptr && tid == 0;
    float* vals = s_vals + threadIdx.x;
    if (clk) F.phase_ns[0] = global_ns();

    for (int k = 0; k < F.steps; ++k) {
        const uint32_t epoch = F.step0 + (uint32_t)k + 1u;
        const uint32_t etag = epoch & 0x3FFFFFFu;  // 26 bits travel with the action in targ[]
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + environment step
        for (int base = (tid & ~31); base < n; base += nthreads) {
            const int i = base + lane;
            const bool active = i < n;
            int s = 0;
            uint32_t ew = 0u, valid = 0u, bits1 = 0u;
            bool explore = false;
            if (active) {
                s = cur[i];
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
        if (F.evaluate) {
            grid.sync();
            if (clk && k < 10) F.phase_ns[1 + 3 * k] = F.phase_ns[2 + 3 * k] = F.phase_ns[3 + 3 * k] = global_ns();
            continue;
        }

        // ---------------- phase S: sort the agents by state (stable: agents of one row stay in agent order).  The sort
        // reads cur[] only, which phase A does not write; its own barriers also order phase A before the readers.
        int src = 0;
        for (int p = 0; p < X.passes; ++p) {
            const bool last = p == X.passes - 1;
            radix_pass<WARPS>(grid, s_whist, s_base, p == 0 ? cur : X.key[src], X.val[src], X.key[src ^ 1], X.val[src ^ 1], last ? X.rank : nullptr, n,
                              p * kRadixBits, X.ghist, X.rowtot, p == 0);
            src ^= 1;
        }
        const int32_t* skey = X.key[src];
        const int32_t* perm = X.val[src];
        // segment bounds per row: the head of every segment finds its end (gallop + binary search over the sorted keys)
        for (int base = (tid & ~31); base < n; base += nthreads) {
            const int p = base + lane;
            bool head = false;
            if (p < n) {
                const int s = skey[p];
                head = p == 0 || skey[p - 1] != s;
                if (head) {
                    int lo = p + 1, step = 1;  // invariant: skey[lo - 1] == s
                    while (lo < n && __ldcg(skey + min(lo + step - 1, n - 1)) == s && lo + step - 1 < n) { lo += step; step <<= 1; }
                    int hi = min(lo + step - 1, n);  // first position known (or assumed, at n) to differ is in (lo-1, hi]
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (__ldcg(skey + mid) == s) lo = mid + 1; else hi = mid;
                    }
                    *reinterpret_cast<uint4*>(X.seg + (size_t)s * 4) = make_uint4((uint32_t)p, (uint32_t)lo, epoch, 0u);
                    *reinterpret_cast<uint4*>(X.hrec + (size_t)p * 4) = make_uint4((uint32_t)s, (uint32_t)lo, (uint32_t)p, 0u);
                }
            }
            const uint32_t hm = __ballot_sync(kFull, head);
            if (lane == 0) X.hmask[base >> 5] = hm;
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[1 + 3 * k] = global_ns();

        // ---------------- phase R: readers
        for (int base = (tid & ~31); base < n; base += nthreads) {
            const int i = base + lane;
            const bool active = i < n;
            const int ii = active ? i : 0;
            const uint8_t at = F.tr_a[ii];
            const bool term = (at & 0x80) != 0;
            const int a = at & 0x7F;
            const float r = F.tr_r[ii];
            const int y = nxt[ii];
            const int q = X.rank[ii];
            const bool boot = active && !term;
            uint4 sg = make_uint4(0, 0, 0, 0);
            if (boot) sg = __ldcg(reinterpret_cast<const uint4*>(X.seg + (size_t)y * 4));
            RowGather<LPR> rows;
            rows.issue(T, y, boot);
            const bool has_writers = boot && sg.z == epoch;
            int before = 0;  // writers of y that precede agent i
            if (has_writers) {
                int lo = (int)sg.x, hi = (int)sg.y;  // first position in [lo, hi) whose agent is >= i
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldcg(perm + mid) < i) lo = mid + 1; else hi = mid;
                }
                before = lo - (int)sg.x;
            }
            const uint32_t ew = (ENV == 1 && active) ? F.envw[ii] : 0u;
            const uint32_t m2 = boot ? (F.use_masks ? env_mask<ENV>(y, ew, T.A, F.env_seed) : full) : 0u;
            if (boot && m2 == 0u) atomicOr(T.err, kErrEmpty);
            const float m_row = rows.row_max((boot && before == 0) ? m2 : 0u);
            bool deferred = false;
            if (active) {
                if (!boot || before == 0) {
                    const float t = td_target_s(r, term ? 0.0f : m_row, F.gamma);
                    st_relaxed_u64(X.targ + q, ((uint64_t)((etag << 6) | (uint32_t)a) << 32) | (uint64_t)__float_as_uint(t));
                } else {
                    uint4 rec = make_uint4((uint32_t)((int)sg.x + before - 1), (uint32_t)q, __float_as_uint(r), (uint32_t)a);
                    *reinterpret_cast<uint4*>(X.rrec + (size_t)i * 4) = rec;
                    deferred = true;
                }
            }
            const uint32_t dm = __ballot_sync(kFull, deferred);
            if (lane == 0) X.rmask[base >> 5] = dm;
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[2 + 3 * k] = global_ns();

        // ---------------- phase Q: sequencers and deferred readers, non-blocking sweeps over statically owned tiles
        // Warp w owns tiles w, w + nwarps, ...; lane t keeps the pending masks of the t-th owned tile in registers and
        // the pending jobs of all owned tiles are compacted into batches of 32 (one job per lane), so a sweep costs as
        // many round trips as there are batches, not as there are tiles.  (A warp that owns more than 32 tiles walks
        // them one by one with the masks in memory.)
        {
            const int owned = (ntiles - gwarp + nwarps - 1) / nwarps;  // may be <= 0
            const bool in_regs = owned <= 32;
            uint32_t my_h = 0u, my_r = 0u;
            if (in_regs && lane < owned) {
                my_h = __ldcg(X.hmask + gwarp + lane * nwarps);
                my_r = __ldcg(X.rmask + gwarp + lane * nwarps);
            }
            // the position every deferred reader of the first kRposTiles owned tiles polls, in shared memory (the part
            // of the union the sort used): an unsuccessful poll then costs one round trip, not two
            constexpr int kRposTiles = 8;
            uint32_t* rpos = reinterpret_cast<uint32_t*>(s_union) + kValWords + (threadIdx.x >> 5);  // [tile][lane][warp]
            const bool rpos_cached = in_regs && (kValWords + kRposTiles * 32 * 8 <= (kSortWords > kValWords ? kSortWords : kValWords));
            if (rpos_cached) {
                for (int t = 0; t < min(owned, kRposTiles); ++t) {
                    const uint32_t dm = __shfl_sync(kFull, my_r, t);
                    if ((dm >> lane) & 1u) rpos[(t * 32 + lane) * 8] = __ldcg(X.rrec + (size_t)((gwarp + t * nwarps) * 32 + lane) * 4);
                }
                __syncwarp();
            }
            // A sequencer walks the segment of one row in agent order while targets are there.
            struct Seq {
                int s, en, p, amx;
                uint32_t legal, touched;
                float mx;
                bool loaded;
            };
            auto seq_load_row = [&](Seq& q) {
                const float* row = T.q + (size_t)q.s * T.ld;
#pragma unroll
                for (int c = 0; c < LPR; ++c) {
                    const F8 v8 = ld_row8(row + 8 * c);
#pragma unroll
                    for (int j = 0; j < 8; ++j) vals[(8 * c + j) * 256] = v8.v[j];
                }
                q.mx = -INFINITY;  // running masked row max and one action that attains it
                q.amx = 0;
                for (uint32_t bm = q.legal; bm; bm &= bm - 1u) {
                    const int a2 = __ffs(bm) - 1;
                    const float x = vals[a2 * 256];
                    if (x > q.mx) { q.mx = x; q.amx = a2; }
                }
                q.loaded = true;
            };
            // consume the run of deposited targets that starts with w (= targ[q.p], known to be there); true = segment done
            auto seq_advance = [&](Seq& q, uint64_t w) -> bool {
                bool more = true;
                while (more) {
                    uint64_t wv[8];  // up to eight targets per round trip
                    wv[0] = w;
#pragma unroll
                    for (int j = 1; j < 8; ++j) wv[j] = (q.p + j < q.en) ? ld_relaxed_u64(X.targ + q.p + j) : 0ull;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (!more) break;
                        if (j > 0 && (q.p == q.en || (uint32_t)(wv[j] >> 38) != etag)) { more = false; break; }
                        const int a = (int)((wv[j] >> 32) & 63u);
                        const float v = td_from_target_s(vals[a * 256], __uint_as_float((uint32_t)wv[j]), lr);
                        vals[a * 256] = v;
                        q.touched |= 1u << a;
                        if (v >= q.mx) { q.mx = v; q.amx = a; }
                        else if (a == q.amx) {  // the cell that held the max went down: look again
                            q.mx = -INFINITY;
                            for (uint32_t bm = q.legal; bm; bm &= bm - 1u) {
                                const int a2 = __ffs(bm) - 1;
                                const float x = vals[a2 * 256];
                                if (x > q.mx) { q.mx = x; q.amx = a2; }
                            }
                        }
                        st_relaxed_u64(X.mhist + q.p, ((uint64_t)epoch << 32) | (uint64_t)__float_as_uint(q.mx));
                        ++q.p;
                    }
                    if (more) {  // all eight consumed: is there a ninth?
                        if (q.p == q.en) break;
                        w = ld_relaxed_u64(X.targ + q.p);
                        if ((uint32_t)(w >> 38) != etag) break;
                    }
                }
                return q.p == q.en;
            };
            auto seq_store_row = [&](Seq& q) {  // commit (or park) the cells that changed
                float* row = T.q + (size_t)q.s * T.ld;
                for (uint32_t bm = q.touched; bm; bm &= bm - 1u) {
                    const int a = __ffs(bm) - 1;
                    row[a] = vals[a * 256];
                }
                q.touched = 0u;
            };
            auto seq_open = [&](int p0) -> Seq {
                const uint4 h4 = __ldcg(reinterpret_cast<const uint4*>(X.hrec + (size_t)p0 * 4));
                Seq q;
                q.s = (int)h4.x; q.en = (int)h4.y; q.p = (int)h4.z;
                q.legal = F.use_masks ? state_mask<ENV>(q.s, T.A, F.env_seed, full) : full;
                q.touched = 0u; q.mx = 0.0f; q.amx = 0; q.loaded = false;
                return q;
            };
            // one visit of a swept sequencer job
            auto sequencer = [&](int p0) -> bool {
                const uint64_t w0 = ld_relaxed_u64(X.targ + p0);  // most segments are visited once: fetch their first target
                Seq q = seq_open(p0);                              // together with the head record
                const uint64_t w = q.p == p0 ? w0 : ld_relaxed_u64(X.targ + q.p);
                if ((uint32_t)(w >> 38) != etag) return false;  // the next target is not there: nothing to do yet
                seq_load_row(q);
                const bool fin = seq_advance(q, w);
                seq_store_row(q);
                if (!fin) __stcg(X.hrec + (size_t)p0 * 4 + 2, (uint32_t)q.p);
                return fin;
            };
            // one deferred reader: the row max it waits for has been published -> deposit the target
            auto reader = [&](int i) -> bool {
                const int t_own = (i / 32 - gwarp) / nwarps;  // which of this warp's tiles
                const uint32_t mpos = (rpos_cached && t_own < kRposTiles) ? rpos[(t_own * 32 + (i & 31)) * 8]
                                                                          : __ldcg(X.rrec + (size_t)i * 4);
                const uint64_t w = ld_relaxed_u64(X.mhist + mpos);
                if ((uint32_t)(w >> 32) != epoch) return false;
                const uint4 rec = __ldcg(reinterpret_cast<const uint4*>(X.rrec + (size_t)i * 4));
                const float tg = td_target_s(__uint_as_float(rec.z), __uint_as_float((uint32_t)w), F.gamma);
                st_relaxed_u64(X.targ + rec.y, ((uint64_t)((etag << 6) | rec.w) << 32) | (uint64_t)__float_as_uint(tg));
                return true;
            };
            // one sweep over the compacted pending jobs of `mine`; returns the updated mask register
            auto sweep = [&](uint32_t mine, auto job) -> uint32_t {
                DeferredGroup grp;
                grp.init(mine);
                for (int b0 = 0; b0 < grp.total; b0 += 32) {
                    const int rank = b0 + lane;
                    const bool act = rank < grp.total;
                    const int rel = grp.agent_of(act ? rank : 0);  // (owned tile index) * 32 + lane-in-tile
                    bool fin = false;
                    if (act) fin = job((gwarp + (rel >> 5) * nwarps) * 32 + (rel & 31));
                    const int t_lo = __shfl_sync(kFull, rel >> 5, 0);
                    const int t_hi = __shfl_sync(kFull, rel >> 5, min(grp.total - b0, 32) - 1);
                    for (int t = t_lo; t <= t_hi; ++t) {
                        const uint32_t bits = __reduce_or_sync(kFull, (fin && (rel >> 5) == t) ? (1u << (rel & 31)) : 0u);
                        if (lane == t) mine &= ~bits;
                    }
                }
                return mine;
            };
            bool seq_left = true, rd_left = true;
            const uint64_t t_start = global_ns();
            for (uint32_t spins = 0; seq_left || rd_left; ++spins) {
                if (in_regs) {
                    if (seq_left) { my_h = sweep(my_h, sequencer); seq_left = __any_sync(kFull, my_h != 0u); }
                    if (rd_left) { my_r = sweep(my_r, reader); rd_left = __any_sync(kFull, my_r != 0u); }
                    // Few jobs left: they move into the lanes (sequencers first, then readers) and are polled back to
                    // back -- a sequencer keeps its row in shared memory, a reader its two positions in registers.
                    const int nh = (int)__reduce_add_sync(kFull, (uint32_t)__popc(my_h)), nr = (int)__reduce_add_sync(kFull, (uint32_t)__popc(my_r));
                    if (nh + nr > 0 && nh + nr <= 32) {
                        const uint64_t t_res = global_ns();
                        if (lane == 0 && X.dbg) {
                            atomicAdd(X.dbg + 0, (unsigned long long)spins + 1ull);
                            atomicAdd(X.dbg + 1, (unsigned long long)(t_res - t_start));
                            atomicMax(X.dbg + 2, (unsigned long long)(t_res - t_start));
                            atomicAdd(X.dbg + 5, 1ull);
                            atomicAdd(X.dbg + 6, (unsigned long long)(nh + nr));
                        }
                        DeferredGroup gh, gr;
                        gh.init(my_h);
                        gr.init(my_r);
                        const int relh = gh.agent_of(lane < nh ? lane : 0), relr = gr.agent_of((lane >= nh && lane < nh + nr) ? lane - nh : 0);
                        int kind = lane < nh ? 2 : (lane < nh + nr ? 1 : 0);
                        Seq q;
                        q.s = q.en = q.p = q.amx = 0; q.legal = q.touched = 0u; q.mx = 0.0f; q.loaded = false;
                        uint32_t mpos = 0u, tpos = 0u, rbits = 0u, ract = 0u, wmpos = 0xFFFFFFFFu;
                        if (kind == 2) q = seq_open((gwarp + (relh >> 5) * nwarps) * 32 + (relh & 31));
                        if (kind == 1) {
                            const uint4 rec = __ldcg(reinterpret_cast<const uint4*>(X.rrec + (size_t)((gwarp + (relr >> 5) * nwarps) * 32 + (relr & 31)) * 4));
                            mpos = rec.x; tpos = rec.y; rbits = rec.z; ract = rec.w;
                        }
                        for (uint32_t it = 0; __any_sync(kFull, kind != 0); ++it) {
                            if (kind == 1) {
                                const uint64_t w = ld_relaxed_u64(X.mhist + mpos);
                                if ((uint32_t)(w >> 32) == epoch) {
                                    const float tg = td_target_s(__uint_as_float(rbits), __uint_as_float((uint32_t)w), F.gamma);
                                    st_relaxed_u64(X.targ + tpos, ((uint64_t)((etag << 6) | ract) << 32) | (uint64_t)__float_as_uint(tg));
                                    kind = 0;
                                }
                            } else if (kind == 2) {
                                uint64_t w = ld_relaxed_u64(X.targ + q.p);
                                if ((uint32_t)(w >> 38) != etag) {
                                    // The writer at q.p is a deferred reader.  Do not wait for its own deposit: fetch its
                                    // record once and poll the row max it waits for -- one hop per level instead of two.
                                    if (wmpos == 0xFFFFFFFFu) {
                                        const uint4 rec = __ldcg(reinterpret_cast<const uint4*>(X.rrec + (size_t)__ldcg(perm + q.p) * 4));
                                        wmpos = rec.x; rbits = rec.z; ract = rec.w;
                                    }
                                    const uint64_t mw = ld_relaxed_u64(X.mhist + wmpos);
                                    if ((uint32_t)(mw >> 32) == epoch) {
                                        const float tg = td_target_s(__uint_as_float(rbits), __uint_as_float((uint32_t)mw), F.gamma);
                                        w = ((uint64_t)((etag << 6) | ract) << 32) | (uint64_t)__float_as_uint(tg);
                                    }
                                }
                                if ((uint32_t)(w >> 38) == etag) {
                                    wmpos = 0xFFFFFFFFu;
                                    if (!q.loaded) seq_load_row(q);
                                    if (seq_advance(q, w)) { seq_store_row(q); kind = 0; }
                                }
                            }
                            if ((it & 1023u) == 1023u && global_ns() - t_start > kTimeoutNs) { atomicOr(T.err, kErrTimeout); break; }
                        }
                        seq_left = rd_left = false;
                        if (lane == 0 && X.dbg) {
                            atomicAdd(X.dbg + 3, (unsigned long long)(global_ns() - t_res));
                            atomicMax(X.dbg + 4, (unsigned long long)(global_ns() - t_res));
                        }
                    }
                } else {
                    if (seq_left) {
                        seq_left = false;
                        for (int tile = gwarp; tile < ntiles; tile += nwarps) {
                            const uint32_t hm = __ldcg(X.hmask + tile);
                            if (hm == 0u) continue;
                            const bool fin = ((hm >> lane) & 1u) ? sequencer(tile * 32 + lane) : false;
                            const uint32_t keep = hm & ~__ballot_sync(kFull, fin);
                            if (lane == 0 && keep != hm) __stcg(X.hmask + tile, keep);
                            seq_left |= keep != 0u;
                        }
                    }
                    if (rd_left) {
                        rd_left = false;
                        for (int tile = gwarp; tile < ntiles; tile += nwarps) {
                            const uint32_t dm = __ldcg(X.rmask + tile);
                            if (dm == 0u) continue;
                            const bool fin = ((dm >> lane) & 1u) ? reader(tile * 32 + lane) : false;
                            const uint32_t keep = dm & ~__ballot_sync(kFull, fin);
                            if (lane == 0 && keep != dm) __stcg(X.rmask + tile, keep);
                            rd_left |= keep != 0u;
                        }
                    }
                }
                if ((spins & 63u) == 63u && global_ns() - t_start > kTimeoutNs) { atomicOr(T.err, kErrTimeout); break; }
            }
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[3 + 3 * k] = global_ns();
    }
    if (F.steps & 1) {
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
