// qe_small.cuh -- the fused loop for SMALL batches (BASELINE configs 1 and 2: 1 .. 256 agents): one CTA, no grid barrier.
//
// A cooperative launch over 148 SMs, grid barriers and publish/poll through L2 cost ~10 us per vector step however few
// agents there are (round 1: 31 us per 128-agent TicTacToe step -- slower than the 16-core C port).  Here one CTA holds the
// whole batch (one thread per agent), the phases are separated by __syncthreads(), and the exact sequential order is
// resolved in shared memory with the same algebra as qe_pipe.cuh:
//   phase 1  select + environment step (row of s from L2, everything else in registers); terminated agents know their
//            target at once;
//   phase 2  every other agent fetches the row of s', lists the EARLIER agents that stand on s' (their writes come first
//            in the reference's loop, QLO:806-817) and, round by round (a round = one __syncthreads), replays their
//            targets in agent order as soon as all of them are known, takes the masked max and files its own target;
//            earlier agents whose own bootstrap row is this very row (self loops) are derived in line;
//   phase 3  the first writer of every cell replays the targets of all writers of that cell in agent order from the value
//            the cell had before the step, and stores the result.
// Same floating-point operations in the same order as the reference: bit-identical results.  K vector steps per launch.
#pragma once
#include "qe_sorted.cuh"

namespace qe {

constexpr int kSmallMaxAgents = 256;

// bits j (j < words * 32) with keys[j] == key: every thread compares against the whole (broadcast) array, four at a time
template <int WORDS>
__device__ __forceinline__ void match_mask(const int* keys, int key, int words, uint32_t* out) {
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
        uint32_t m = 0u;
        if (w < words) {
#pragma unroll
            for (int b = 0; b < 32; b += 4) {
                const int4 v = *reinterpret_cast<const int4*>(keys + w * 32 + b);
                m |= (v.x == key ? 1u : 0u) << b | (v.y == key ? 2u : 0u) << b | (v.z == key ? 4u : 0u) << b | (v.w == key ? 8u : 0u) << b;
            }
        }
        out[w] = m;
    }
}

template <int ENV, int LPR>
__global__ void __launch_bounds__(kSmallMaxAgents) fused_small_kernel(Table T, FusedArgs F) {
    constexpr int N = kSmallMaxAgents, WORDS = N / 32;
    __shared__ __align__(16) int s_s[N];     // state the agent stands on (the row it writes); -2: no agent
    __shared__ __align__(16) int s_key[N];   // cell it writes: state * 32 + action; -2: no agent
    __shared__ int s_s2[N];                  // row it bootstraps from, -1: terminated
    __shared__ uint32_t s_a[N];              // action
    __shared__ float s_r[N];                 // reward
    __shared__ float s_tgt[N];               // target, valid once s_ok
    __shared__ volatile int s_ok[N];
    __shared__ unsigned long long s_thr[2];  // exploration threshold / learning rate of the next step (prefetched)
    __shared__ float s_lr[2];
    __shared__ float s_col[32][N];           // this thread's replayed row of s' (illegal cells at -inf): s_col[a][thread]
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    const int i = threadIdx.x, lane = i & 31;
    const int n = F.n;
    const int words = (n + 31) >> 5;
    const bool act = i < n;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int A = T.A;
    const bool clk = F.phase_ns != nullptr && i == 0;
    if (clk) F.phase_ns[0] = global_ns();
    int s = act ? F.st_a[i] : 0;
    uint32_t ew = (act && ENV != 0) ? F.envw[i] : 0u;
    float epret = act ? F.ep_ret[i] : 0.0f;
    double loc_sum = 0.0;
    unsigned int loc_cnt = 0;
    // bits below / above this thread's own position
    uint32_t below[WORDS], above[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; ++w) {
        below[w] = w < (i >> 5) ? 0xFFFFFFFFu : (w == (i >> 5) ? ((1u << (i & 31)) - 1u) : 0u);
        above[w] = w > (i >> 5) ? 0xFFFFFFFFu : (w == (i >> 5) ? ((i & 31) == 31 ? 0u : (0xFFFFFFFFu << ((i & 31) + 1))) : 0u);
    }
    // the row of state x into this thread's column (one 256-bit load per sector), cells outside `keep` at -inf
    auto load_row = [&](int x, uint32_t keep) {
        const float* row = T.q + (size_t)x * T.ld;
#pragma unroll
        for (int c = 0; c < LPR; ++c) {
            const F8 v = ld_row8(row + 8 * c);
#pragma unroll
            for (int q = 0; q < 8; ++q) s_col[8 * c + q][i] = ((keep >> (8 * c + q)) & 1u) ? v.v[q] : -INFINITY;
        }
    };
    auto col_max = [&]() {  // (cells beyond A hold -inf or are never legal: the whole padded row can be scanned)
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < 8 * LPR; ++c) m = fmax_plain(m, s_col[c][i]);
        return m;
    };
    // the counter stream U[t][agent][k] = fmix32(pre[k] + t * gold + add): the inner hash does not change from step to step
    uint32_t pre[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) pre[q] = fmix32((((uint32_t)F.agent0 + (uint32_t)i) * 8u + (uint32_t)q) ^ ((q < 2 ? F.stream_seed : F.env_stream_seed) * kGold));
    uint32_t valid = act ? (F.use_masks ? env_mask<ENV>(s, ew, A, F.env_seed) : full) : 0u;  // mask of the current state (next step: the mask of s')

    if (i == 0) { s_thr[0] = F.eps_thresh[0]; s_lr[0] = F.lr[0]; }
    __syncthreads();
    for (int k = 0; k < F.steps; ++k) {
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = s_thr[k & 1];
        const float lr = s_lr[k & 1];
        auto draw = [&](int q) -> uint32_t {
            if (F.uniforms) return U.draw(i, q);
            return fmix32(pre[q] + (q < 2 ? U.t : U.env_t) * kGold + kStreamAdd);
        };
        if (i == 0 && k + 1 < F.steps) { s_thr[(k + 1) & 1] = F.eps_thresh[k + 1]; s_lr[(k + 1) & 1] = F.lr[k + 1]; }  // off the critical path
        // ---------------- phase 1: select + environment step
        int a = 0, s2 = s;
        float r = 0.0f, p = 0.0f;
        bool term = false;
        uint32_t m2 = 0u;
        if (act) {
            F8 v[LPR];  // the row of s in registers
            {
                const float* row = T.q + (size_t)s * T.ld;
#pragma unroll
                for (int c = 0; c < LPR; ++c) v[c] = ld_row8(row + 8 * c);
            }
            const bool explore = (uint64_t)draw(0) < thresh;
            const uint32_t bits1 = draw(1), bits2 = draw(2), bits3 = draw(3);
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < LPR; ++c) mx = fmax_plain(mx, max8(v[c], (valid >> (8 * c)) & 0xFFu));
            uint32_t tie = 0u;
#pragma unroll
            for (int c = 0; c < LPR; ++c) tie |= tie8(v[c], (valid >> (8 * c)) & 0xFFu, mx) << (8 * c);
            a = pick_action(A, valid, tie, explore, F.empty_all != 0, bits1);
            if (a < 0) { atomicOr(T.err, kErrEmpty); a = 0; }
#pragma unroll
            for (int c = 0; c < LPR; ++c)
                if ((a >> 3) == c) p = sel8(v[c], a);
            if (ENV == 0) mdp_step(s2, a, (uint32_t)F.S, A, F.env_seed, F.term_thresh, bits2, bits3, r, term);
            else if (ENV == 1) {
                if (!ttt_step(ew, a, bits2, bits3, draw(4), r, term)) atomicOr(T.err, kErrInvalidMove);
                s2 = ttt_state(ew & 0x3FFFFu);
            } else {
                r = (float)a;
                ew += 1u;
                term = ew >= F.episode_len;
                if (term) ew = 0u;
                s2 = 0;
            }
            m2 = F.use_masks ? env_mask<ENV>(s2, ew, A, F.env_seed) : full;
            valid = m2;
            if (!term && !F.evaluate) load_row(s2, m2);  // the untouched row of s' (nothing is committed before phase 3) travels under the rest of the phase
            float acc = epret + r;
            float fin = __int_as_float(0x7FC00000);
            if (term) { fin = acc; loc_sum += (double)acc; ++loc_cnt; acc = 0.0f; }
            epret = acc;
            const size_t o = (size_t)k * n + i;
            if (F.trace_actions) F.trace_actions[o] = a;
            if (F.trace_rewards) F.trace_rewards[o] = r;
            if (F.trace_term) F.trace_term[o] = term;
            if (F.trace_next) F.trace_next[o] = s2;
            if (F.trace_epret) F.trace_epret[o] = fin;
        }
        if (F.evaluate) {  // BaseRuntime.evaluate_*: the table is not touched
            s = s2;
            continue;
        }
        s_s[i] = act ? s : -2;
        s_key[i] = act ? s * 32 + a : -2;
        s_s2[i] = (act && !term) ? s2 : -1;
        s_a[i] = (uint32_t)a;
        s_r[i] = r;
        s_ok[i] = (act && term) ? 1 : 0;
        s_tgt[i] = td_target_s(r, 0.0f, F.gamma);  // (only read where s_ok: the terminated agents)
        __syncthreads();

        // ---------------- phase 2: targets.  dm = the EARLIER agents that stand on s' (their writes come first)
        bool pending = act && !term;
        uint32_t dm[WORDS];
        if (pending) {
            if (m2 == 0u) atomicOr(T.err, kErrEmpty);  // np.max of an empty selection (QLO:764)
            match_mask<WORDS>(s_s, s2, words, dm);
#pragma unroll
            for (int w = 0; w < WORDS; ++w) dm[w] &= below[w];
        } else {
#pragma unroll
            for (int w = 0; w < WORDS; ++w) dm[w] = 0u;
        }
        // Rounds are per WARP (a round ends at a vote of the warp, not at a block barrier): a warp re-reads the flags of the
        // agents it waits for -- volatile shared memory, written target first, flag second -- until none of its lanes is
        // pending.  Dependencies point to smaller agent indices, so the warp that holds the smallest pending agent never
        // waits for anybody who waits: no deadlock; the bound only catches a broken invariant.
        for (int round = 0; round < (1 << 22); ++round) {
            bool now = false;
            if (pending) {
                // ready when every earlier writer of s' has its target, or is a self loop on s' (derived in line below)
                bool ready = true;
#pragma unroll
                for (int w = 0; w < WORDS; ++w)
                    for (uint32_t m = dm[w]; m; m &= m - 1u) {
                        const int j = w * 32 + __ffs(m) - 1;
                        ready = ready && (s_ok[j] != 0 || s_s2[j] == s2);
                    }
                if (ready) {
#pragma unroll
                    for (int w = 0; w < WORDS; ++w)
                        for (uint32_t m = dm[w]; m; m &= m - 1u) {  // ascending j: the reference's order
                            const int j = w * 32 + __ffs(m) - 1;
                            float tj;
                        if (s_ok[j]) { __threadfence_block(); tj = *reinterpret_cast<volatile float*>(s_tgt + j); }
                        else tj = td_target_s(s_r[j], col_max(), F.gamma);  // self loop: this row IS its bootstrap row
                            const uint32_t aj = s_a[j];
                            if ((m2 >> aj) & 1u) s_col[aj][i] = td_from_target_s(s_col[aj][i], tj, lr);
                        }
                    now = true;
                }
            }
            if (now) {  // target first, flag second: a reader that sees the flag in this very round reads the final target
                s_tgt[i] = td_target_s(r, col_max(), F.gamma);
                __threadfence_block();
                s_ok[i] = 1;
                pending = false;
            }
            if (!__any_sync(kFull, pending)) break;
        }
        if (pending) atomicOr(T.err, kErrTimeout);
        __syncthreads();  // every target is out before the commit reads them

        // ---------------- phase 3: commit, the first writer of every cell replays all of its writers in agent order
        if (act) {
            uint32_t cm[WORDS];
            match_mask<WORDS>(s_key, s * 32 + a, words, cm);
            bool head = true;
#pragma unroll
            for (int w = 0; w < WORDS; ++w) head = head && (cm[w] & below[w]) == 0u;
            if (head) {
                float v = td_from_target_s(p, s_tgt[i], lr);
#pragma unroll
                for (int w = 0; w < WORDS; ++w)
                    for (uint32_t m = cm[w] & above[w]; m; m &= m - 1u) v = td_from_target_s(v, s_tgt[w * 32 + __ffs(m) - 1], lr);
                T.q[(size_t)s * T.ld + a] = v;
            }
        }
        s = s2;
        __threadfence_block();
        __syncthreads();  // the next step's select sees the commits (rows are read from L2)
        if (clk && k < 10) { F.phase_ns[1 + 3 * k] = F.phase_ns[2 + 3 * k] = F.phase_ns[3 + 3 * k] = global_ns(); }
    }
    if (act) {
        F.st_a[i] = s;
        if (ENV != 0) F.envw[i] = ew;
        F.ep_ret[i] = epret;
    }
    if (F.ep_count) {
        for (int d = 16; d > 0; d >>= 1) {
            loc_sum += __shfl_xor_sync(kFull, loc_sum, d);
            loc_cnt += __shfl_xor_sync(kFull, loc_cnt, d);
        }
        if (lane == 0) { s_sum[i >> 5] = loc_sum; s_cnt[i >> 5] = loc_cnt; }
        __syncthreads();
        if (i == 0) {
            double bs = 0.0;
            unsigned int bc = 0;
            for (int w = 0; w < N / 32; ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
            if (bc) { atomicAdd(F.ep_sum, bs); atomicAdd(F.ep_count, (unsigned long long)bc); }
        }
    }
}

}  // namespace qe
