// qe_flow.cuh -- the one-pass form of the fused loop (round 2): select, environment step and the exact sequential TD
// update of an agent happen back to back in ONE in-order pass over the agents, four grid barriers per vector step.
//
// It keeps the two ideas of the target pipeline (qe_pipe.cuh): what an agent contributes to the table is its TARGET
// t_i = r_i + gamma * max_a Q_i[s'_i, a] (QLO:760-768), and anybody who knows the targets of the earlier writers of a
// row (a prefix of the row's segment in a stable sort of the agents by state) can replay the row from its value at
// the start of the step; and chunks of 32 agents are claimed in increasing order, so a dependency always points to an
// agent that was claimed earlier.  What changes:
//
//   * select + environment step (round 2's phase A) and the target pipeline (phase T) are one phase.  The table is only
//     written by the commit, so "select" of agent i may run while the targets of agents j < i are still being resolved:
//     the row gathers of select (throughput) and the publish -> poll hops of the targets (latency) overlap, the
//     per-agent hand-over array between the phases (16 B per agent each way) and one grid barrier are gone.  A warp
//     claims a chunk, selects and steps its 32 agents and leaves them in a queue in shared memory; a free lane takes the
//     next queue entry, fetches its bootstrap row into shared memory (cp.async) beside the first poll of the writer
//     records of s'_i, and keeps polling until it has published its target -- a waiting agent holds up a lane, not a chunk.
//   * a writer record is one 64-bit word {agent | action << 24 | flags, value}: the sort of the previous step leaves
//     {agent, 0} at every position, the agent's warp overwrites it once with FINAL (value = target; terminated agents
//     at once, QLO:760-766) and before that, for a self loop (s' = s), with SELF (value = reward: the reader derives
//     the target from the row it is replaying, which IS the row the writer bootstraps from).  One 8-byte store, one
//     256-bit load per four records: no separate "pending" pass, no barrier between filing and reading.
//   * the order of the next states is an MSD bucket sort instead of LSD passes over the grid: the top (up to) 10 bits
//     are a stable partition over all CTAs (digit counts per CTA taken in the tail of the in-order pass, column scan
//     beside the commit, scatter), the remaining bits are sorted bucket by bucket inside one CTA -- by a team of two
//     warps (one digit of low bits), by the whole CTA in shared memory (more bits, herded buckets) or by one warp in
//     registers (tiny buckets of small batches) -- which also writes position, segment bounds and the initial writer
//     records of its bucket.
//
// Per vector step:   [ select + step + targets | digit counts ]  B  [ column scan, commit ]  B  [ scatter ]  B
//                    [ bucket sorts, positions, bounds, records ]  B
// Same floating-point operations in the same order as the reference: bit-identical to the oracle and to the other forms.
// Requires the legal-action mask to be a function of the state (true for the device environments).
#pragma once
#include "qe_pipe.cuh"

namespace qe {

constexpr uint32_t kRecSelf = 1u << 29;   // value = reward of a writer whose next state is its own row
constexpr uint32_t kRecFinal = 1u << 30;  // value = target
constexpr uint32_t kRecAgent = 0xFFFFFFu;

struct FlowScratch {
    uint64_t* rec;        // [cap + 8] writer records by sorted position (see above)
    uint2* seg;           // [S] per state {segment start, segment end} ({0, 0}: nobody stands on the state)
    int32_t* pos;         // [cap] agent -> sorted position
    int2* kv[2];          // [cap] {state, agent} by position; the finished order is always in kv[0] (the scatter leaves the buckets in kv[1])
    int* ghist;           // [kRadix][blocks] bucket counts per block, scanned in place
    int* rowtot;          // [kRadix] bucket totals
    unsigned int* ctr;    // [64] 0: chunk claims; 1: chunks selected + stepped; 2: commit tile claims; 4: abort; 6: order invalid
    int msd_shift;        // bucket = state >> msd_shift (< kRadix)
    int local_passes;     // LSD passes (kRadixBits each) over the low msd_shift bits inside a bucket
    int sorted_valid;     // pos / seg / kv[0] / rec describe the states this launch starts from (to be checked)
    int old_n;            // agents of the order kv[0] and seg[] still describe (0: none, seg[] is all-empty)
    int flags;            // development / test switches (QE_FLOW_FLAGS): 1 = no teams of two warps in the bucket sorts (whole-block path for
                          // every bucket), 4 = the shared-memory block path only takes buckets of <= 64 keys (the rest: global-memory passes)
};

// counter += 1 in shared memory, address space stated (through a pointer the compiler cannot trace to shared memory a plain
// atomicAdd becomes a generic atomic, which crawls when the 32 lanes hit one address: a herded row in a bucket)
__device__ __forceinline__ void smem_inc(int* p) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
}
__device__ __forceinline__ void st_relaxed_rec(uint64_t* p, uint32_t x, uint32_t y) {
    st_relaxed_u64(p, (uint64_t)x | ((uint64_t)y << 32));
}

// static split of [0, n) in whole tiles of 32: block b of nb, warp w of WARPS -> [lo, hi) (sizes differ by at most one tile)
template <int WARPS>
__device__ __forceinline__ void flow_part(int n, int b, int nb, int w, int& lo, int& hi) {
    const int tiles = (n + 31) >> 5;
    const int per = tiles / nb, rem = tiles % nb;
    const int t0 = b * per + min(b, rem), bt = per + (b < rem ? 1 : 0);
    const int pw = bt / WARPS, rw = bt % WARPS;
    const int w0 = t0 + w * pw + min(w, rw), w1 = w0 + pw + (w < rw ? 1 : 0);
    lo = min(w0 * 32, n);
    hi = min(w1 * 32, n);
}

// ---- bucket counts of this block's part of states[] -> whist (kept for the scatter) and ghist[bucket][block]
template <int WARPS>
__device__ __forceinline__ void flow_hist(int (*whist)[kRadix], const int32_t* states, int n, const FlowScratch& X) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x, nb = gridDim.x;
    int lo, hi;
    flow_part<WARPS>(n, b, nb, warp, lo, hi);
    for (int d = lane; d < kRadix; d += 32) whist[warp][d] = 0;
    __syncwarp();
    const int sh = X.msd_shift;
    for (int base = lo; base < hi; base += 256) {  // eight loads in flight per lane
        int32_t kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) kk[u] = base + 32 * u + lane < hi ? __ldcg(states + base + 32 * u + lane) : 0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + 32 * u + lane < hi) smem_inc(&whist[warp][(uint32_t)kk[u] >> sh]);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadix; d += blockDim.x) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) t += whist[w][d];
        X.ghist[(size_t)d * nb + b] = t;
    }
}

// ---- exclusive scan of every bucket's row of block counts (one warp per bucket), bucket totals
template <int WARPS>
__device__ __forceinline__ void flow_scan(const FlowScratch& X) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = gridDim.x;
    const int per = (nb + 31) / 32;
    // bucket d belongs to warp d / nb of block d % nb: one to three warps per block scan a row each while the block's other
    // warps go on to the commit (all rows on the first kRadix / WARPS blocks: those blocks reach the commit ~10 us late)
    for (int d = blockIdx.x + warp * nb; d < kRadix; d += WARPS * nb) {
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

// ---- stable scatter of this block's part into the buckets; s_base[0 .. kRadix] = bucket starts (kept for the bucket sorts)
template <int WARPS>
__device__ __forceinline__ void flow_scatter(int (*whist)[kRadix], int* s_base, int* s_wsum, const int32_t* states, int n, int2* out,
                                             const FlowScratch& X) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x, nb = gridDim.x;
    static_assert(kRadix == 4 * 256, "four buckets per thread");
    {
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
        if (threadIdx.x == 255) s_base[kRadix] = run;
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
    int lo, hi;
    flow_part<WARPS>(n, b, nb, warp, lo, hi);
    const int sh = X.msd_shift;
    for (int base = lo; base < hi; base += 256) {
        int32_t kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) kk[u] = base + 32 * u + lane < hi ? __ldcg(states + base + 32 * u + lane) : 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int x = base + 32 * u + lane;
            if (base + 32 * u >= hi) break;  // (uniform)
            const bool act = x < hi;
            const uint32_t d = (uint32_t)kk[u] >> sh;
            const uint32_t peers = digit_peers(d, act);
            if (act) out[whist[warp][d] + __popc(peers & ((1u << lane) - 1u))] = make_int2(kk[u], x);
            __syncwarp();
            if (act && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
            __syncwarp();
        }
    }
}

// ---- one stable counting pass over [lo, hi) of src -> dst by the digit (key >> shift) & (2^bits - 1), whole block
template <int WARPS>
__device__ __forceinline__ void flow_local_pass(int (*whist)[kRadix], int* s_wsum, const int2* src, int2* dst, int lo, int hi, int shift,
                                                int bits) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nd = 1 << bits;
    const uint32_t dm = (uint32_t)nd - 1u;
    const int tiles = (hi - lo + 31) >> 5;
    const int pw = tiles / WARPS, rw = tiles % WARPS;
    const int w0 = warp * pw + min(warp, rw), w1 = w0 + pw + (warp < rw ? 1 : 0);
    const int wlo = min(lo + w0 * 32, hi), whi = min(lo + w1 * 32, hi);
    for (int d = lane; d < nd; d += 32) whist[warp][d] = 0;
    __syncwarp();
    for (int base = wlo; base < whi; base += 32) {
        const int x = base + lane;
        if (x < whi) smem_inc(&whist[warp][((uint32_t)__ldcg(&src[x].x) >> shift) & dm]);
    }
    __syncthreads();
    {   // digit totals, exclusive scan over the digits (thread t owns `per` consecutive digits), first free position per (digit, warp)
        const int per = nd >= 256 ? nd / 256 : 1;
        int v[4] = {0, 0, 0, 0};
        int sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = threadIdx.x * per + j;
            if (j < per && d < nd) {
#pragma unroll
                for (int w = 0; w < WARPS; ++w) v[j] += whist[w][d];
            }
            sum += v[j];
        }
        const int incl = warp_incl_scan(sum);
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int before = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) before += (w < warp) ? s_wsum[w] : 0;
        int run = lo + before + incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int d = threadIdx.x * per + j;
            if (j < per && d < nd) {
                int r2 = run;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) {
                    const int c = whist[w][d];
                    whist[w][d] = r2;
                    r2 += c;
                }
                run += v[j];
            }
        }
    }
    __syncthreads();
    for (int base = wlo; base < whi; base += 32) {
        const int x = base + lane;
        const bool act = x < whi;
        int2 e = make_int2(0, 0);
        if (act) e = __ldcg(src + x);
        const uint32_t d = ((uint32_t)e.x >> shift) & dm;
        const uint32_t peers = digit_peers(d, act);
        if (act) dst[whist[warp][d] + __popc(peers & ((1u << lane) - 1u))] = e;
        __syncwarp();
        if (act && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
}

// ---- TWO warps, one bucket, one pass: as flow_warp_staged, the bucket split in halves.  A bucket sort is one long chain of
// dependent instructions (count, scan, rank, write: ~50 cycles per element for a single warp), so a second warp nearly
// halves it.  Stable: the first warp's half precedes the second's within every digit (own counters per warp, the scan
// gives the second warp's elements of a digit the places behind the first's).  The team meets at a named barrier.
__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void flow_team_staged(int tw, int bar_id, int* cnt2, int* s_tot, uint32_t* sagent, uint16_t* sdigit, uint16_t* order,
                                                 const int2* src, int2* dst, int lo, int hi, int bits, int32_t key_hi, const FlowScratch& X) {
    const int lane = threadIdx.x & 31;
    const int nd = 1 << bits, m = hi - lo;
    const uint32_t dm = (uint32_t)nd - 1u;
    const int h = min(((m + 63) >> 6) << 5, m);          // the first warp's share (whole groups of 32)
    const int x0 = tw ? h : 0, x1 = tw ? m : h;
    int* cnt = cnt2 + tw * kRadix;                         // this warp's counters; the other's: cnt2 + (tw ^ 1) * kRadix
    for (int d = lane; d < nd; d += 32) cnt[d] = 0;
    for (int base = x0; base < x1; base += 256) {
        int2 e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = base + 32 * u + lane < x1 ? __ldcg(src + lo + base + 32 * u + lane) : make_int2(0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + 32 * u + lane < x1) {
                sagent[base + 32 * u + lane] = (uint32_t)e[u].y;
                sdigit[base + 32 * u + lane] = (uint16_t)((uint32_t)e[u].x & dm);
            }
    }
    __syncwarp();
    for (int x = x0 + lane; x < x1; x += 32) smem_inc(&cnt[sdigit[x]]);
    team_bar(bar_id);
    // exclusive scan in digit order: this warp does the digits [tw * nd / 2, ...); a digit is a state: its bounds
    const int* c0 = cnt2;
    const int* c1 = cnt2 + kRadix;
    const int dh = nd >> 1, d0 = tw ? dh : 0, d1 = tw ? nd : dh;
    {
        int part = 0;
        for (int d = d0 + lane; d < d1; d += 32) part += c0[d] + c1[d];
        part = __reduce_add_sync(kFull, part);
        if (lane == 0) s_tot[tw] = part;
    }
    team_bar(bar_id);
    int run = tw ? s_tot[0] : 0;
    for (int j = 0; j < dh; j += 32) {  // (a warp reads and rewrites the counters of its own digits only)
        const int d = d0 + j + lane;
        const bool in = d < d1;
        const int va = in ? c0[d] : 0, vb = in ? c1[d] : 0, v = va + vb;
        const int incl = warp_incl_scan(v);
        const int first = run + incl - v;
        if (in) { cnt2[d] = first; cnt2[kRadix + d] = first + va; }
        if (v > 0) X.seg[key_hi | d] = make_uint2((uint32_t)(lo + first), (uint32_t)(lo + first + v));
        run += __shfl_sync(kFull, incl, 31);
    }
    team_bar(bar_id);
    for (int base = x0; base < x1; base += 32) {
        const bool act = base + lane < x1;
        const uint32_t d = act ? (uint32_t)sdigit[base + lane] : 0u;
        const uint32_t peers = digit_peers(d, act);
        if (act) order[cnt[d] + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)(base + lane);
        __syncwarp();
        if (act && lane == (__ffs(peers) - 1)) cnt[d] += __popc(peers);
        __syncwarp();
    }
    team_bar(bar_id);
    for (int x = lane + 32 * tw; x < m; x += 64) {  // (only agent -> position is scattered)
        const int o = order[x];
        const uint32_t ag = sagent[o];
        const int q = lo + x;
        dst[q] = make_int2(key_hi | (int32_t)sdigit[o], (int32_t)ag);
        X.rec[q] = (uint64_t)ag;
        X.pos[ag] = q;
    }
}

// ---- the WHOLE block, one bucket whose low `bits` bits are still unsorted (several digits: large tables; or a bucket that
// herding made too large for a team): LSD passes of 8 bits entirely in shared memory.  Staged: the low bits of every key
// (4 bytes) and two permutations (2 bytes each) -- the agents stay in global memory and are gathered once, at the end.
// emit(position, state, agent) files one element of the finished order; the bounds come from the neighbours.
constexpr int kBlockDigitBits = 8;
__device__ __forceinline__ uint32_t digit_peers_bits(uint32_t d, bool act, int nbits) {
    uint32_t peers = __ballot_sync(kFull, act);
    for (int bit = 0; bit < nbits; ++bit) {
        const bool one = (d >> bit) & 1u;
        const uint32_t bb = __ballot_sync(kFull, one);
        peers &= one ? bb : ~bb;
    }
    return peers;
}
__host__ __device__ constexpr int flow_block_cap(int arena_bytes) {
    return ((arena_bytes - 8 * (1 << kBlockDigitBits) * (int)sizeof(int)) / 8) & ~31;
}
template <int WARPS, class Emit>
__device__ __forceinline__ void flow_block_staged(unsigned char* arena, int arena_bytes, int* s_wsum, const int2* src, int lo, int hi, int bits,
                                                  int32_t key_hi, uint2* seg, Emit emit) {
    constexpr int ND = 1 << kBlockDigitBits;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = hi - lo, cap = min(flow_block_cap(arena_bytes), 65535);
    int (*wc)[ND] = reinterpret_cast<int (*)[ND]>(arena);                       // [WARPS][ND] per-warp digit counters
    uint32_t* key = reinterpret_cast<uint32_t*>(arena + sizeof(int) * WARPS * ND);  // [cap] low bits of the keys, by arrival
    uint16_t* oa = reinterpret_cast<uint16_t*>(key + cap);                          // [cap] permutations (ping-pong)
    uint16_t* ob = oa + cap;
    const uint32_t mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    for (int x0 = 0; x0 < m; x0 += 4 * 256) {  // four loads in flight per thread
        int32_t kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) kk[u] = x0 + 256 * u + (int)threadIdx.x < m ? __ldcg(&src[lo + x0 + 256 * u + threadIdx.x].x) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (x0 + 256 * u + (int)threadIdx.x < m) key[x0 + 256 * u + threadIdx.x] = (uint32_t)kk[u] & mask;
    }
    const int tiles = (m + 31) >> 5;
    const int pw = tiles / WARPS, rw = tiles % WARPS;
    const int t0 = warp * pw + min(warp, rw), t1 = t0 + pw + (warp < rw ? 1 : 0);
    const int w0 = min(t0 * 32, m), w1 = min(t1 * 32, m);  // this warp's slice of the current order
    uint16_t* cur = oa;
    uint16_t* nxt = ob;
    const int P = (bits + kBlockDigitBits - 1) / kBlockDigitBits;
    __syncthreads();
    for (int ps = 0; ps < P; ++ps) {
        const int shift = ps * kBlockDigitBits, nb_bits = min(kBlockDigitBits, bits - shift);
        const uint32_t dm = (1u << nb_bits) - 1u;
        for (int d = lane; d < ND; d += 32) wc[warp][d] = 0;
        __syncwarp();
        for (int x = w0 + lane; x < w1; x += 32) smem_inc(&wc[warp][(key[ps == 0 ? x : (int)cur[x]] >> shift) & dm]);
        __syncthreads();
        {   // first free place per (digit, warp): thread t owns digit t
            const int t = threadIdx.x;
            int v = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) v += wc[w][t];
            const int incl = warp_incl_scan(v);
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            int run = incl - v;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) run += (w < warp) ? s_wsum[w] : 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const int c = wc[w][t];
                wc[w][t] = run;
                run += c;
            }
        }
        __syncthreads();
        for (int base = w0; base < w1; base += 32) {
            const bool act = base + lane < w1;
            const int idx = act ? (ps == 0 ? base + lane : (int)cur[base + lane]) : 0;
            const uint32_t d = act ? ((key[idx] >> shift) & dm) : 0u;
            const uint32_t peers = digit_peers_bits(d, act, nb_bits);
            if (act) nxt[wc[warp][d] + __popc(peers & ((1u << lane) - 1u))] = (uint16_t)idx;
            __syncwarp();
            if (act && lane == (__ffs(peers) - 1)) wc[warp][d] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        uint16_t* t = cur;
        cur = nxt;
        nxt = t;
    }
    for (int x = threadIdx.x; x < m; x += 256) {
        const int o = P ? (int)cur[x] : x;
        const uint32_t k = key[o];
        const uint32_t prev = x > 0 ? key[P ? (int)cur[x - 1] : x - 1] : 0xFFFFFFFFu;
        const int32_t agent = __ldcg(&src[lo + o].y);
        const int32_t st = key_hi | (int32_t)k;
        const int q = lo + x;
        emit(q, st, agent);
        if (prev != k) {  // first of its state: one 8-byte store of the bounds (on a 100M-state table every such store is a DRAM access)
            int e = x + 1;
            while (e < m && key[P ? (int)cur[e] : e] == k) ++e;
            seg[st] = make_uint2((uint32_t)q, (uint32_t)(lo + e));
        }
    }
    __syncthreads();
}

// ---- one warp, one bucket of at most 32 keys (small batches: most buckets of a grid of a few blocks): every lane ranks its
// key against the other lanes' (stable: equal keys keep their order), no shared memory, no barrier.
template <class Emit>
__device__ __forceinline__ void flow_warp_tiny(const int2* src, int lo, int hi, int bits, int32_t key_hi, uint2* seg, Emit emit) {
    const int lane = threadIdx.x & 31;
    const int m = hi - lo;
    const uint32_t mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    int2 e = make_int2(0, 0);
    if (lane < m) e = __ldcg(src + lo + lane);
    const uint32_t k = lane < m ? ((uint32_t)e.x & mask) : 0xFFFFFFFFu;
    int rank = 0, before = 0, equal = 0;
    for (int j = 0; j < m; ++j) {
        const uint32_t kj = __shfl_sync(kFull, k, j);
        rank += (kj < k) ? 1 : 0;
        before += (kj == k && j < lane) ? 1 : 0;
        equal += (kj == k) ? 1 : 0;
    }
    __syncwarp();  // (src may be the buffer the order is written to: every lane has its element by now)
    if (lane < m) {
        const int32_t st = key_hi | (int32_t)k;
        const int q = lo + rank + before;
        emit(q, st, e.y);
        if (before == 0) seg[st] = make_uint2((uint32_t)q, (uint32_t)(q + equal));
    }
}

// ---- the buckets of this block: sort by the low bits, then positions, segment bounds and fresh writer records.
//   * the low bits are ONE digit and the block has at most three buckets (the usual case on a table of <= 2^20 states): a
//     team of two warps per bucket, all buckets of the grid in flight at once (flow_team_staged);
//   * everything else -- more low bits (larger tables), buckets that herding made too large for a team, small grids -- is
//     done bucket by bucket by the whole block in shared memory (flow_block_staged), and what does not even fit there by
//     counting passes through global memory (flow_local_pass).
constexpr int kWarpBucketMax = 3072;
template <int WARPS>
__device__ __forceinline__ void flow_buckets(int (*whist)[kRadix], unsigned char* arena, int arena_bytes, const int* s_base, int* s_wsum,
                                             const FlowScratch& X) {
    const int L = X.local_passes;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x, nb = gridDim.x;
    const bool teams = (kRadix - 1) / nb < 3 && L == 1 && !(X.flags & 1);
    const int slot_bytes = (arena_bytes / 3) & ~15;
    const int stage_cap = min(min(((slot_bytes - 2 * (int)sizeof(int) * kRadix - 16) / 8) & ~7, 65535), kWarpBucketMax);
    bool rest = !teams;
    if (teams) {  // teams 2|3, 4|5, 6|7 (warps 0 and 1 stay out -- warp 0 runs several times slower here)
        const int team = warp >= 2 ? (warp - 2) >> 1 : -1, tw = warp & 1;
        int* cnt = reinterpret_cast<int*>(arena + (size_t)(team >= 0 ? team : 0) * slot_bytes);
        int* s_tot = cnt + 2 * kRadix;  // [2] (+ 2 words of padding)
        uint32_t* sagent = reinterpret_cast<uint32_t*>(s_tot + 4);
        uint16_t* sdigit = reinterpret_cast<uint16_t*>(sagent + stage_cap);
        uint16_t* order = sdigit + stage_cap;
        const int d = team * nb + b;
        if (team >= 0 && d < kRadix) {
            const int lo = s_base[d], hi = s_base[d + 1];
            if (hi - lo > stage_cap) rest = true;
            else if (hi > lo)
                flow_team_staged(tw, 1 + team, cnt, s_tot, sagent, sdigit, order, X.kv[1], X.kv[0], lo, hi, X.msd_shift,
                                 (int32_t)((uint32_t)d << X.msd_shift), X);
        }
    }
    auto emit = [&](int q, int32_t st, int32_t agent) {
        X.kv[0][q] = make_int2(st, agent);
        X.pos[agent] = q;
        X.rec[q] = (uint64_t)(uint32_t)agent;
    };
    if (!teams) {  // buckets of at most 32 keys: one warp each (a grid of a few blocks has hundreds of them per block)
        rest = false;
        for (int j = warp; j * nb + b < kRadix; j += WARPS) {
            const int d = j * nb + b;
            const int lo = s_base[d], hi = s_base[d + 1];
            if (hi - lo > 32) rest = true;
            else if (hi > lo) flow_warp_tiny(X.kv[L ? 1 : 0], lo, hi, X.msd_shift, (int32_t)((uint32_t)d << X.msd_shift), X.seg, emit);
        }
    }
    if (!__syncthreads_or(rest)) return;
    const int block_cap = (X.flags & 4) ? 64 : min(flow_block_cap(arena_bytes), 65535);  // (QE_FLOW_FLAGS & 4: tests push ordinary buckets through the global-memory passes)
    for (int d = b; d < kRadix; d += nb) {
        const int lo = s_base[d], hi = s_base[d + 1];
        if (hi <= lo || (teams ? hi - lo <= stage_cap : hi - lo <= 32)) continue;  // (uniform; the teams / single warps have done theirs)
        const int32_t key_hi = (int32_t)((uint32_t)d << X.msd_shift);
        if (hi - lo <= block_cap) {  // the scatter left the bucket in kv[1] (kv[0] when there are no low bits: nothing moves then)
            flow_block_staged<WARPS>(arena, arena_bytes, s_wsum, X.kv[L ? 1 : 0], lo, hi, X.msd_shift, key_hi, X.seg, emit);
            continue;
        }
        int src = L ? 1 : 0;
        for (int ps = 0; ps < L; ++ps) {
            const int shift = ps * kRadixBits;
            flow_local_pass<WARPS>(whist, s_wsum, X.kv[src], X.kv[src ^ 1], lo, hi, shift, min(kRadixBits, X.msd_shift - shift));
            src ^= 1;
        }
        if (src != 0) {  // (an even number of passes ends in kv[1])
            for (int q = lo + threadIdx.x; q < hi; q += blockDim.x) X.kv[0][q] = __ldcg(X.kv[1] + q);
            __syncthreads();
        }
        const int2* fin = X.kv[0];
        for (int q = lo + threadIdx.x; q < hi; q += blockDim.x) {
            const int2 e = __ldcg(fin + q);
            const int32_t prev = q > lo ? __ldcg(&fin[q - 1].x) : -1, next = q + 1 < hi ? __ldcg(&fin[q + 1].x) : -1;
            X.pos[e.y] = q;
            X.rec[q] = (uint64_t)(uint32_t)e.y;
            if (prev != e.x) X.seg[e.x].x = (uint32_t)q;
            if (next != e.x) X.seg[e.x].y = (uint32_t)(q + 1);
        }
        __syncthreads();
    }
}

constexpr size_t kFlowExtraBytes = 0;  // (3 x 63.5 KB fit the 196 KB shared-memory carve-out; one step more and the L1 shrinks: scatter +2 us, in-order pass +3 us)
__host__ __device__ constexpr int flow_row_words(int lpr) { return 8 * lpr + 4; }  // one replayed row per thread, 16-byte aligned, conflict-free
__host__ __device__ constexpr size_t flow_smem_bytes(int lpr) {
    return sizeof(int) * ((size_t)8 * kRadix + kRadix + 8) + (sizeof(uint4) + sizeof(uint2)) * 8 * 32 + sizeof(float) * (size_t)flow_row_words(lpr) * 256 +
           kFlowExtraBytes;
}
#ifndef QE_FLOW_CLAIM
#define QE_FLOW_CLAIM 1
#endif
#ifndef QE_FLOW_MIN_BLOCKS
#define QE_FLOW_MIN_BLOCKS 3
#endif
template <int ENV, int LPR>
__global__ void __launch_bounds__(256, QE_FLOW_MIN_BLOCKS) fused_flow_kernel(Table T, FusedArgs F, FlowScratch X) {
    cg::grid_group grid = cg::this_grid();
    constexpr int WARPS = 8;
    constexpr int RS = flow_row_words(LPR);
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    __shared__ int s_wsum[WARPS];
    extern __shared__ __align__(16) unsigned char s_raw[];  // flow_smem_bytes(LPR)
    int* s_base = reinterpret_cast<int*>(s_raw);                                  // [kRadix + 1] bucket starts
    unsigned char* s_arena = s_raw + sizeof(int) * (kRadix + 8);                  // everything below: carved up anew by the bucket sorts
    int (*s_whist)[kRadix] = reinterpret_cast<int (*)[kRadix]>(s_arena);          // [8][kRadix] per-warp digit counters
    uint4* s_queue = reinterpret_cast<uint4*>(s_arena + sizeof(int) * WARPS * kRadix);  // [8][32] per-warp queues of stepped agents
    uint2* s_qseg = reinterpret_cast<uint2*>(s_queue + WARPS * 32);               // [8][32] segment bounds of the queued agents' next states
    float* s_rows = reinterpret_cast<float*>(s_qseg + WARPS * 32);                // in-order pass: [256][RS]; commit: [8*LPR][256] + [256]
    float* s_row = s_rows;
    uint32_t* s_touch = reinterpret_cast<uint32_t*>(s_rows) + 8 * LPR * 256;
    static_assert((8 * LPR + 1) * 256 <= RS * 256, "the commit's columns fit in the rows region");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;
    const int ntiles = (n + 31) >> 5;
    const bool clk = F.phase_ns != nullptr && tid == 0;
    const int wbase = threadIdx.x & ~31;
    uint64_t* rec = X.rec;
    const int2* sorted = X.kv[0];
    if (clk) F.phase_ns[0] = global_ns();

    // ---------------- the order of the states this launch starts from: left behind by the previous launch (checked), else made now
    {
        bool bad = !X.sorted_valid;
        if (!bad)
            for (int i = tid; i < n; i += nthreads) {
                const uint32_t q = (uint32_t)__ldcg(X.pos + i);
                if (q >= (uint32_t)n) { bad = true; continue; }
                const int2 e = __ldcg(sorted + q);
                bad |= e.y != i || e.x != __ldcg(F.st_a + i) || __ldcg(rec + q) != (uint64_t)(uint32_t)i;
            }
        if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(X.ctr + 6, 1u);
        grid.sync();
        if (ld_relaxed_u32(X.ctr + 6) != 0u) {  // (the same answer in every block)
            for (int q = tid; q < X.old_n; q += nthreads) {  // the bounds of the order kv[0] still describes
                const int32_t kq = __ldcg(&sorted[q].x);
                if (q == 0 || __ldcg(&sorted[q - 1].x) != kq) X.seg[kq] = make_uint2(0u, 0u);
            }
            flow_hist<WARPS>(s_whist, F.st_a, n, X);
            grid.sync();
            flow_scan<WARPS>(X);
            grid.sync();
            flow_scatter<WARPS>(s_whist, s_base, s_wsum, F.st_a, n, X.kv[X.local_passes ? 1 : 0], X);
            grid.sync();
            flow_buckets<WARPS>(s_whist, s_arena, (int)(flow_smem_bytes(LPR) - sizeof(int) * (kRadix + 8)), s_base, s_wsum, X);
            grid.sync();
        }
    }

    for (int k = 0; k < F.steps; ++k) {
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- the in-order pass: select + environment step + target, chunk by chunk
        {
            unsigned int* claim = X.ctr;
            float* myrow = s_rows + threadIdx.x * RS;
            // chunks are claimed kClaim at a time (one: two per atomic halve the traffic on the one counter but make the last wave
            // of the pass -- the tail everybody waits for -- coarser: ~1.5 us slower) and the claim after the current one is
            // always in flight: nobody waits for the counter's round trip
            constexpr int kClaim = QE_FLOW_CLAIM;
            auto claim_raw = [&]() {  // lane 0's answer; nobody waits for it before it is needed
                int c = 0;
                if (lane == 0) c = (int)atomicAdd(claim, (unsigned int)kClaim);
                return c;
            };
            int ch_next = 0, ch_left = 0, ch_raw = claim_raw();
            auto next_chunk = [&]() {  // first agent of the warp's next chunk (>= n: no more)
                if (ch_left == 0) {
                    ch_next = __shfl_sync(kFull, ch_raw, 0) * 32;
                    ch_left = kClaim;
                    ch_raw = claim_raw();
                }
                --ch_left;
                const int c = ch_next;
                ch_next += 32;
                return c;
            };
            uint32_t m2 = 0u;  // legal actions of the row this lane is replaying (illegal cells never reach the max)
            auto row_max = [&]() {
                float m = -INFINITY;
#pragma unroll
                for (int c = 0; c < 2 * LPR; ++c) {
                    const float4 v = reinterpret_cast<const float4*>(myrow)[c];
                    const uint32_t mm = m2 >> (4 * c);
                    m = fmaxf(m, (mm & 1u) ? v.x : -INFINITY);
                    m = fmaxf(m, (mm & 2u) ? v.y : -INFINITY);
                    m = fmaxf(m, (mm & 4u) ? v.z : -INFINITY);
                    m = fmaxf(m, (mm & 8u) ? v.w : -INFINITY);
                }
                return m;
            };
            // Lanes and chunks are decoupled: the warp selects and steps a whole chunk at a time (all 32 lanes, whatever they
            // hold) and leaves its agents in a 32-entry queue in shared memory; a lane keeps its agent until the target is
            // published and then takes the next one from the queue, so an agent that waits for a predecessor holds up one
            // lane, not a chunk.
            uint4* myq = s_queue + warp * 32;  // {next state, sorted position, reward bits, action | done << 7}
            uint2* myqs = s_qseg + warp * 32;  // {start, end} of the segment of the next state
            int cb = next_chunk();
            int s_nx = 0, pos_nx = 0;
            float ep_nx = 0.0f;
            if (cb + lane < n) { s_nx = cur[cb + lane]; pos_nx = X.pos[cb + lane]; ep_nx = F.ep_ret[cb + lane]; }
            int cbn = cb >= n ? n : next_chunk();
            const uint64_t t_start = global_ns();
            if (cb >= n && lane == 0) atomicAdd(X.ctr + 1, 1u);  // a warp without work: its "all my chunks are stepped" arrival
            int q_base = 0, q_next = 0, q_rem = 0;
            bool busy = false;
#ifdef QE_FLOW_STATS  // development: lane-passes by what the lane did (ctr[24..29], launch totals)
            uint32_t st_pass = 0, st_busy = 0, st_prog = 0, st_block = 0, st_fresh = 0, st_prod = 0;
#define FLOW_STAT(x) x
#else
#define FLOW_STAT(x)
#endif
            int i = 0, mypos = 0;
            float r = 0.0f;
            uint32_t p = 0u, pe = 0u, head = 0u;
            bool fresh = false;  // took its agent in the previous pass: the bootstrap row is on its way into shared memory
            for (uint32_t spins = 0;; ++spins) {
                // ---- produce: the queue is empty -> select + environment step of the next chunk
                FLOW_STAT(++st_pass; st_busy += busy ? 1u : 0u;)
                if (q_rem == 0 && cb < n) {
                    FLOW_STAT(++st_prod;)
                    const int ia = cb + lane;
                    const bool active = ia < n;
                    const int s = s_nx, pos_a = pos_nx;
                    const float ep_a = ep_nx;
                    if (cbn + lane < n) { s_nx = cur[cbn + lane]; pos_nx = X.pos[cbn + lane]; ep_nx = F.ep_ret[cbn + lane]; }
                    uint32_t ew = 0u, valid = 0u, bits1 = 0u;
                    bool explore = false;
                    if (active) {
                        if (ENV != 0) ew = F.envw[ia];
                        valid = F.use_masks ? env_mask<ENV>(s, ew, T.A, F.env_seed) : full;
                        explore = (uint64_t)U.draw(ia, 0) < thresh;
                        bits1 = U.draw(ia, 1);
                    }
                    int32_t s2 = s;
                    float ra = 0.0f;
                    bool term = false;
                    uint2 sg2 = make_uint2(0u, 0u);
                    int a;
                    {
                        RowGather<LPR> rows;
                        rows.issue(T, s, active);
                        float mx;
                        uint32_t tie;
                        rows.row_max_tie(valid, mx, tie);
                        a = pick_action(T.A, valid, tie, explore, F.empty_all != 0, bits1);
                    }
                    if (active && a < 0) { atomicOr(T.err, kErrEmpty); a = 0; }
                    a = max(a, 0);
                    if (active) {
                        if (ENV == 0) mdp_step(s2, a, (uint32_t)F.S, T.A, F.env_seed, F.term_thresh, U.draw(ia, 2), U.draw(ia, 3), ra, term);
                        else if (ENV == 1) {
                            if (!ttt_step(ew, a, U.draw(ia, 2), U.draw(ia, 3), U.draw(ia, 4), ra, term)) atomicOr(T.err, kErrInvalidMove);
                            s2 = ttt_state(ew & 0x3FFFFu);
                        } else {
                            ra = (float)a;
                            ew += 1u;
                            term = ew >= F.episode_len;
                            if (term) ew = 0u;
                            s2 = 0;
                        }
                        nxt[ia] = s2;
                        if (!term) sg2 = __ldcg(X.seg + s2);  // the writers of the row this agent bootstraps from (in flight until the queue entry is written)
                        if (ENV != 0) F.envw[ia] = ew;
                        float acc = ep_a + ra;
                        float fin = __int_as_float(0x7FC00000);
                        if (term) { fin = acc; loc_sum += (double)acc; ++loc_cnt; acc = 0.0f; }
                        F.ep_ret[ia] = acc;
                        const size_t o = (size_t)k * n + ia;
                        if (F.trace_actions) F.trace_actions[o] = a;
                        if (F.trace_rewards) F.trace_rewards[o] = ra;
                        if (F.trace_term) F.trace_term[o] = term;
                        if (F.trace_next) F.trace_next[o] = s2;
                        if (F.trace_epret) F.trace_epret[o] = fin;
                        // a terminated agent bootstraps from nothing (QLO:760-766): final at once; a self loop says so
                        const uint32_t ha = (uint32_t)ia | ((uint32_t)a << 24);
                        if (term) st_relaxed_rec(rec + pos_a, ha | kRecFinal, __float_as_uint(td_target_s(ra, 0.0f, F.gamma)));
                        else if (s2 == s) st_relaxed_rec(rec + pos_a, ha | kRecSelf, __float_as_uint(ra));
                    }
                    myq[lane] = make_uint4((uint32_t)s2, (uint32_t)pos_a, __float_as_uint(ra), (uint32_t)a | ((term || !active) ? 0x80u : 0u));
                    if (!(sg2.x < sg2.y && sg2.y <= (uint32_t)n)) sg2 = make_uint2(0u, 0u);
                    myqs[lane] = sg2;
                    __syncwarp();
                    // the warp's last chunk: all its next states are in place (the bucket counts wait for every warp's arrival)
                    if (cbn >= n && lane == 0) { __threadfence(); atomicAdd(X.ctr + 1, 1u); }
                    q_base = cb; q_next = 0; q_rem = min(32, n - cb);
                    cb = cbn;
                    cbn = next_chunk();
                }
                // ---- consume: the free lanes take the next agents of the queue, in order.  The bootstrap row travels to the
                // lane's shared-memory slot asynchronously (cp.async, no registers) beside the first poll of the writer records
                // (the segment bounds came with the queue entry): an agent without unfinished predecessors is done in this pass.
                const uint32_t freeb = __ballot_sync(kFull, !busy);
                if (q_rem > 0 && freeb != 0u) {
                    const int idx = q_next + __popc(freeb & ((1u << lane) - 1u));
                    if (!busy && idx < q_next + q_rem) {
                        const uint4 d = myq[idx];
                        if (!(d.w & 0x80u)) {
                            const uint2 sg = myqs[idx];
                            i = q_base + idx; mypos = (int)d.y; r = __uint_as_float(d.z);
                            head = (uint32_t)i | ((d.w & 31u) << 24);
                            const int y = (int)d.x;
                            m2 = F.use_masks ? state_mask<ENV>(y, T.A, F.env_seed, full) : full;
                            if (m2 == 0u) atomicOr(T.err, kErrEmpty);  // np.max of an empty selection (QLO:764)
                            p = sg.x; pe = sg.y;
                            const float* row = T.q + (size_t)y * T.ld;
#pragma unroll
                            for (int c = 0; c < 2 * LPR; ++c) cp_async16(myrow + 4 * c, row + 4 * c);
                            fresh = busy = true;
                        }
                    }
                    const int took = min(__popc(freeb), q_rem);
                    q_next += took;
                    q_rem -= took;
                }
                if (busy) {
                    // ---- poll: the (up to four) records of the aligned 32-byte group at the cursor, in position (= agent) order
                    bool go = p < pe, fin = !go;  // nobody (left) on s': the row is as the replay has it
                    const uint32_t pa = p & ~3u;
                    FLOW_STAT(const uint32_t p0 = p;)
                    U8 e4, e5;  // (a long prefix -- a herded row -- is read two groups per pass: both loads travel together)
#pragma unroll
                    for (int j = 0; j < 8; ++j) e4.w[j] = e5.w[j] = 0u;
                    if (go) e4 = ld_relaxed_v8(reinterpret_cast<const uint2*>(rec + pa));
                    if (go && pa + 4u < pe) e5 = ld_relaxed_v8(reinterpret_cast<const uint2*>(rec + pa + 4));
                    if (fresh) {
                        cp_async_wait_all();
                        fresh = false;
                        FLOW_STAT(++st_fresh;)
                    }
                    auto replay4 = [&](const U8& g, const uint32_t base) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t ex = g.w[2 * j], ey = g.w[2 * j + 1];
                            if (go && base + j >= p) {
                                if (base + j >= pe || (int)(ex & kRecAgent) >= i) {
                                    fin = true;  // the segment ends here, or the writers from here on come after i
                                    go = false;
                                } else if (ex & (kRecFinal | kRecSelf)) {
                                    const float tg = (ex & kRecFinal) ? __uint_as_float(ey) : td_target_s(__uint_as_float(ey), row_max(), F.gamma);
                                    const uint32_t a2 = (ex >> 24) & 31u;
                                    float* cell = myrow + a2;
                                    *cell = td_from_target_s(*cell, tg, lr);
                                    ++p;
                                } else {
                                    go = false;  // an earlier writer that has not published yet
                                }
                            }
                        }
                    };
                    replay4(e4, pa);
                    if (go && p == pa + 4u && p < pe) replay4(e5, pa + 4u);  // (no vote here: this code runs under `if (busy)`, not all lanes are present)
                    if (go && p >= pe) fin = true;
                    FLOW_STAT(st_prog += (fin || p != p0) ? 1u : 0u; st_block += (!fin && p == p0) ? 1u : 0u;)
                    if (fin) {
                        st_relaxed_rec(rec + mypos, head | kRecFinal, __float_as_uint(td_target_s(r, row_max(), F.gamma)));
                        busy = false;
                    }
                }
                if (cb >= n && q_rem == 0 && !__any_sync(kFull, busy)) break;
                if ((spins & 255u) == 255u) {  // (all lanes take the same way out)
                    if (__any_sync(kFull, ld_relaxed_u32(X.ctr + 4) != 0u || global_ns() - t_start > kPipeTimeoutNs)) {
                        atomicExch(X.ctr + 4, 1u);
                        atomicOr(T.err, kErrTimeout);
                        break;
                    }
                }
            }
#ifdef QE_FLOW_STATS
            {
                const uint32_t a0 = __reduce_add_sync(kFull, st_busy), a1 = __reduce_add_sync(kFull, st_prog), a2 = __reduce_add_sync(kFull, st_block),
                               a3 = __reduce_add_sync(kFull, st_fresh);
                if (lane == 0) {
                    atomicAdd(X.ctr + 24, st_pass); atomicAdd(X.ctr + 25, a0 >> 5); atomicAdd(X.ctr + 26, a1 >> 5); atomicAdd(X.ctr + 27, a2 >> 5);
                    atomicAdd(X.ctr + 28, a3 >> 5); atomicAdd(X.ctr + 29, st_prod);
                }
            }
#endif
            // ---- tail: once every chunk has been selected and stepped, the bucket counts of this block's part of the next states
            __syncthreads();
            if (threadIdx.x == 0) {
                for (uint32_t spins = 0; ld_relaxed_u32(X.ctr + 1) < (unsigned int)(nthreads >> 5); ++spins) {
                    __nanosleep(64);
                    if ((spins & 255u) == 255u && (ld_relaxed_u32(X.ctr + 4) != 0u || global_ns() - t_start > kPipeTimeoutNs)) {
                        atomicExch(X.ctr + 4, 1u);
                        atomicOr(T.err, kErrTimeout);
                        break;
                    }
                }
                __threadfence();
            }
            __syncthreads();
            flow_hist<WARPS>(s_whist, nxt, n, X);  // (after a timeout the counts are garbage and so is everything else: the error flag says so)
        }
        if (F.ep_count) {
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
                if (bc) { atomicAdd(F.ep_sum, bs); atomicAdd(F.ep_count, (unsigned long long)bc); }
            }
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[1 + 3 * k] = global_ns();

        // ---------------- column scan of the bucket counts (the first kRadix warps) and the commit: one pass over the sorted
        // records, tiles claimed dynamically; a segment's row lives in the shared-memory column of its head lane, the
        // members take turns in position (= agent) order; what extends beyond the tile is replayed by the whole warp, one
        // lane per action.  The bounds of the segment are cleared on the way (the bucket sorts write the next ones).
        // The loop is software-pipelined, because a tile is a chain of three round trips (claim, records, rows) and little
        // else: while tile A is applied, the rows of tile B and the records of tile C are on their way and the claim of the
        // pair after that is in flight.  A warp's first pair is its own index (no 3552-way rush on the counter).
        flow_scan<WARPS>(X);
        {
            constexpr int kGroup = 2;  // tiles per claim (8 per claim: 40 us instead of 31 -- tiles differ a lot in cost, small claims balance them)
            const int nw2 = kGroup * gridDim.x * WARPS;
            auto claim_raw2 = [&]() {
                int c = 0;
                if (lane == 0) c = (int)atomicAdd(X.ctr + 2, (unsigned int)kGroup);
                return c;
            };
            int nx_tile = kGroup * (blockIdx.x * WARPS + warp), nx_left = kGroup;
            int nx_pair = nw2 + __shfl_sync(kFull, claim_raw2(), 0);
            int raw = claim_raw2();
            auto next_tile = [&]() {
                if (nx_left == 0) {
                    nx_tile = nx_pair;
                    nx_left = kGroup;
                    nx_pair = nw2 + __shfl_sync(kFull, raw, 0);
                    raw = claim_raw2();
                }
                --nx_left;
                return nx_tile++;
            };
            struct TileIn { int tile; uint64_t e; uint32_t st, pv; uint64_t e2; uint32_t k2; };  // e2 / k2: the same lane of the following tile
            auto load_in = [&](int tile) {
                TileIn t{tile, 0ull, 0xFFFFFFFFu, 0xFFFFFFFFu, 0ull, 0xFFFFFFFFu};
                const int p = tile * 32 + lane;
                if (tile < ntiles && p < n) { t.e = __ldcg(rec + p); t.st = (uint32_t)__ldcg(&sorted[p].x); }
                if (tile < ntiles && p + 32 < n) { t.e2 = __ldcg(rec + p + 32); t.k2 = (uint32_t)__ldcg(&sorted[p + 32].x); }
                if (tile < ntiles && lane == 0 && p > 0) t.pv = (uint32_t)__ldcg(&sorted[p - 1].x);
                return t;
            };
            auto is_head = [&](const TileIn& t) {
                const int p = t.tile * 32 + lane;
                uint32_t prev = __shfl_up_sync(kFull, t.st, 1);
                if (lane == 0) prev = t.pv;
                return t.tile < ntiles && p < n && (p == 0 || prev != t.st);
            };
            struct TileRows { F8 v[LPR]; };
            auto load_rows = [&](const TileIn& t) {
                TileRows r;
                const bool head = is_head(t);
                const float* row = T.q + (size_t)(head ? t.st : 0u) * T.ld;
#pragma unroll
                for (int c = 0; c < LPR; ++c) {
                    if (head) r.v[c] = ld_row8(row + 8 * c);
                    else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) r.v[c].v[j] = 0.0f;
                    }
                }
                return r;
            };
            TileIn tA = load_in(next_tile());
            TileIn tB = load_in(next_tile());
            TileRows rA = load_rows(tA);
        for (;;) {
            if (tA.tile >= ntiles) break;  // (the tiles of a warp increase)
            const TileRows rB = load_rows(tB);
            const TileIn tC = load_in(next_tile());
            {
                const int tile = tA.tile;
                const int p = tile * 32 + lane;
                const bool act = p < n;
                const uint64_t e = tA.e;
                const uint32_t st = tA.st;  // state of this position
                const uint32_t ex = (uint32_t)e, ey = (uint32_t)(e >> 32);
                const bool head = is_head(tA);
                const uint32_t hb = __ballot_sync(kFull, head);
                const uint32_t below = hb & (0xFFFFFFFFu >> (31 - lane));
                const int hl = below ? 31 - __clz(below) : -1;  // head lane of this lane's segment; -1: the segment began in an earlier tile
                if (act && !(ex & kRecFinal)) atomicOr(T.err, kErrTimeout);  // cannot happen: the in-order pass published every target
                if (head) {
#pragma unroll
                    for (int c = 0; c < LPR; ++c) {
                        const F8 v8 = rA.v[c];

#pragma unroll
                        for (int j = 0; j < 8; ++j) s_row[(8 * c + j) * 256 + threadIdx.x] = v8.v[j];
                    }
                    s_touch[threadIdx.x] = 0u;
                    X.seg[st] = make_uint2(0u, 0u);
                }
                __syncwarp();
                const int off = (act && hl >= 0) ? lane - hl : -1;
                const int maxoff = (int)__reduce_max_sync(kFull, off);
                for (int it = 0; it <= maxoff; ++it) {
                    if (off == it) {
                        const int c = wbase + hl;
                        const uint32_t a = (ex >> 24) & 31u;
                        float* cell = s_row + a * 256 + c;
                        *cell = td_from_target_s(*cell, __uint_as_float(ey), lr);
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
                        uint64_t e2 = tA.e2;  // (the following tile came with this one: two out of three tiles need it)
                        uint32_t k2 = tA.k2;
                        if (q > tile * 32 + 32) {
                            e2 = 0ull;
                            k2 = 0xFFFFFFFFu;
                            if (q + lane < n) { e2 = __ldcg(rec + q + lane); k2 = (uint32_t)__ldcg(&sorted[q + lane].x); }
                        }
                        const uint32_t diff = __ballot_sync(kFull, k2 != st31);
                        const int len = diff ? __ffs(diff) - 1 : 32;
                        for (int j = 0; j < len; ++j) {
                            const uint32_t xa = (__shfl_sync(kFull, (uint32_t)e2, j) >> 24) & 31u;
                            const float tg = __uint_as_float(__shfl_sync(kFull, (uint32_t)(e2 >> 32), j));
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
            tA = tB;
            tB = tC;
            rA = rB;
        }
        }
        grid.sync();
        if (clk && k < 10) F.phase_ns[2 + 3 * k] = global_ns();

        // ---------------- the next step's order: stable scatter into the buckets, then every bucket inside one block
        if (tid == 0) { X.ctr[0] = 0u; X.ctr[1] = 0u; X.ctr[2] = 0u; }  // (idle since the last barrier; next used after two more)
        flow_scatter<WARPS>(s_whist, s_base, s_wsum, nxt, n, X.kv[X.local_passes ? 1 : 0], X);
        grid.sync();
        if (clk && k < 10) F.phase_ns[32 + k] = global_ns();
        flow_buckets<WARPS>(s_whist, s_arena, (int)(flow_smem_bytes(LPR) - sizeof(int) * (kRadix + 8)), s_base, s_wsum, X);
        grid.sync();
        if (clk && k < 10) F.phase_ns[3 + 3 * k] = global_ns();
    }
    if (F.steps & 1) {
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
