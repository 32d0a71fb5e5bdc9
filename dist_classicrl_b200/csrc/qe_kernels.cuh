// qe_kernels.cuh -- CUDA kernels of the engine (sm_100a): select, exact sequential TD update, env steps, fused loop.
//
// Exact sequential TD update in parallel (DESIGN.md "TD update"):
//   The reference applies agents 0..N-1 one after the other (QLO:806-817), so agent i must see every write of
//   agents j<i -- to its own cell Q[s_i,a_i] AND to the row Q[s'_i,:] it bootstraps from.
//   Phase 1 (insert)  : every agent captures p_i = Q0[s_i,a_i] and pushes itself on a per-state writer list
//                       (head[s] atomicExch, node[i] = {next, action}).
//   -- grid-wide barrier --
//   Phase 2 (resolve) : agent i walks the (short) writer lists of s_i and s'_i.  For every cell it needs it finds
//                       the latest writer j<i; the value "just before i" is slot[j] if such a writer exists, the
//                       first writer's captured p (== Q0, the table may already hold a later commit) if the cell is
//                       written only by agents >= i, and the table itself if nobody writes the cell this step.
//                       v_i = p + lr*((r + gamma*m) - p) is published in slot[i] = {epoch, v_i} (one 64-bit store);
//                       waiting agents poll their predecessors' slots.  The dependency graph is a DAG in agent
//                       order and every group processes its agents in increasing order, so the smallest unresolved
//                       agent can always proceed (cooperative launch => all CTAs are co-resident).
//                       The last writer of a cell commits v to the table.
//   Epoch-stamped heads/slots need no per-step clearing (heads are wiped every 255 steps).
#pragma once
#include <cooperative_groups.h>

#include "qe_common.cuh"

namespace qe {
namespace cg = cooperative_groups;

constexpr uint32_t kNone = 0xFFFFFFu;       // list terminator (24-bit agent index)
constexpr int kErrInvalidMove = 1, kErrEmpty = 2, kErrTimeout = 4;
constexpr uint32_t kSpinLimit = 1u << 22;

struct Table {
    float* q;          // [S][ld]
    int ld;            // floats per row (multiple of 4)
    int A;             // actions
    uint32_t* head;    // [S]  (tag8 << 24) | agent24
    uint32_t* node;    // [cap] (action << 24) | next24
    uint64_t* slot;    // [cap] (epoch << 32) | float bits
    float* tr_p;       // [cap] Q0[s_i, a_i] captured before any commit of this step
    int* err;          // device error flags
};

// ------------------------------------------------------------------ phase 1: list insert (one lane per agent)
__device__ __forceinline__ void list_insert(const Table& T, int i, int s, int a, float p, uint32_t tag) {
    T.tr_p[i] = p;
    const uint32_t old = atomicExch(T.head + s, (tag << 24) | (uint32_t)i);
    T.node[i] = ((uint32_t)a << 24) | (((old >> 24) == tag) ? (old & kNone) : kNone);
}

// ------------------------------------------------------------------ phase 2: resolve + commit (group-cooperative)
template <int LPA>
__device__ __forceinline__ void learn_resolve(const Table& T, int i, int s, int a, float r, float p, int s2, bool term,
                                              uint32_t mask2, float lr, float gamma, uint32_t tag, uint32_t epoch,
                                              uint32_t gm) {
    const int l = threadIdx.x & (LPA - 1);
    // own cell: latest earlier writer (predecessor) and whether a later writer exists
    int pj = -1;
    bool later = false;
    {
        uint32_t j = __ldcg(T.head + s) & kNone;  // tag matches: agent i itself is on this list
        while (j != kNone) {
            const uint32_t nd = __ldcg(T.node + j);
            if ((int)(nd >> 24) == a) {
                if ((int)j < i) pj = max(pj, (int)j);
                else if ((int)j > i) later = true;
            }
            j = nd & kNone;
        }
    }
    // bootstrap row s2: lane l owns actions 4l..4l+3
    float m_static = -INFINITY;
    int bj0 = -1, bj1 = -1, bj2 = -1, bj3 = -1;  // latest writer j<i per owned action
    if (!term) {
        const float4 rv = ld_row4(T.q + (size_t)s2 * T.ld + 4 * l);
        int fj0 = INT_MAX, fj1 = INT_MAX, fj2 = INT_MAX, fj3 = INT_MAX;  // first writer per owned action
        const uint32_t w = __ldcg(T.head + s2);
        if ((w >> 24) == tag) {
            uint32_t j = w & kNone;
            while (j != kNone) {
                const uint32_t nd = __ldcg(T.node + j);
                const int k = (int)(nd >> 24) - 4 * l;
                const int ji = (int)j;
                if (k == 0) { if (ji < i) bj0 = max(bj0, ji); fj0 = min(fj0, ji); }
                if (k == 1) { if (ji < i) bj1 = max(bj1, ji); fj1 = min(fj1, ji); }
                if (k == 2) { if (ji < i) bj2 = max(bj2, ji); fj2 = min(fj2, ji); }
                if (k == 3) { if (ji < i) bj3 = max(bj3, ji); fj3 = min(fj3, ji); }
                j = nd & kNone;
            }
        }
        const uint32_t my = (mask2 >> (4 * l)) & 0xFu;
        // cells without an earlier writer: Q0 -- from the table if nobody writes them this step, else from the first
        // writer's captured p (the table may already hold that writer's commit)
        if ((my & 1u) && bj0 < 0) m_static = fmax_plain(m_static, fj0 == INT_MAX ? rv.x : __ldcg(T.tr_p + fj0));
        if ((my & 2u) && bj1 < 0) m_static = fmax_plain(m_static, fj1 == INT_MAX ? rv.y : __ldcg(T.tr_p + fj1));
        if ((my & 4u) && bj2 < 0) m_static = fmax_plain(m_static, fj2 == INT_MAX ? rv.z : __ldcg(T.tr_p + fj2));
        if ((my & 8u) && bj3 < 0) m_static = fmax_plain(m_static, fj3 == INT_MAX ? rv.w : __ldcg(T.tr_p + fj3));
        if (!(my & 1u)) bj0 = -1;
        if (!(my & 2u)) bj1 = -1;
        if (!(my & 4u)) bj2 = -1;
        if (!(my & 8u)) bj3 = -1;
        if (mask2 == 0u && l == 0) atomicOr(T.err, kErrEmpty);  // np.max of an empty selection (QLO:764)
    }
    // wait for the predecessors' values
    float m = m_static, pe = p;
    uint32_t spins = 0;
    const uint64_t want = (uint64_t)epoch;
    for (;;) {
        bool ok = true;
        float ml = m_static;
        uint64_t w;
        if (bj0 >= 0) { w = ld_relaxed_u64(T.slot + bj0); if ((w >> 32) != want) ok = false; else ml = fmax_plain(ml, __uint_as_float((uint32_t)w)); }
        if (bj1 >= 0) { w = ld_relaxed_u64(T.slot + bj1); if ((w >> 32) != want) ok = false; else ml = fmax_plain(ml, __uint_as_float((uint32_t)w)); }
        if (bj2 >= 0) { w = ld_relaxed_u64(T.slot + bj2); if ((w >> 32) != want) ok = false; else ml = fmax_plain(ml, __uint_as_float((uint32_t)w)); }
        if (bj3 >= 0) { w = ld_relaxed_u64(T.slot + bj3); if ((w >> 32) != want) ok = false; else ml = fmax_plain(ml, __uint_as_float((uint32_t)w)); }
        if (pj >= 0) { w = ld_relaxed_u64(T.slot + pj); if ((w >> 32) != want) ok = false; else pe = __uint_as_float((uint32_t)w); }
        if (group_all<LPA>(ok, gm)) { m = ml; break; }
        if (++spins > kSpinLimit) { if (l == 0) atomicOr(T.err, kErrTimeout); m = ml; break; }
    }
    m = group_max<LPA>(m, gm);
    const float v = td_value(pe, r, term ? 0.0f : m, lr, gamma);
    if (l == 0) {
        st_relaxed_u64(T.slot + i, ((uint64_t)epoch << 32) | (uint64_t)__float_as_uint(v));
        if (!later) T.q[(size_t)s * T.ld + a] = v;  // last writer of the cell commits
    }
}

// wipe the writer-list heads (every 255 steps, when the 8-bit tag wraps)
__device__ __forceinline__ void wipe_heads(const Table& T, int64_t S) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < (size_t)S; x += stride) T.head[x] = 0u;
}

// ------------------------------------------------------------------ unfused select
template <int LPA>
__global__ void __launch_bounds__(256) select_kernel(Table T, const int32_t* __restrict__ states,
                                                     const uint32_t* __restrict__ mask_bits, Uniforms U, uint64_t thresh,
                                                     int deterministic, int empty_all, int32_t* __restrict__ actions, int n) {
    const uint32_t gm = group_mask<LPA>();
    const int l = threadIdx.x & (LPA - 1);
    const int groups = gridDim.x * (blockDim.x / LPA);
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    for (int i = blockIdx.x * (blockDim.x / LPA) + threadIdx.x / LPA; i < n; i += groups) {
        const int s = states[i];
        const uint32_t valid = mask_bits ? (mask_bits[i] & full) : full;
        const float4 v = ld_row4(T.q + (size_t)s * T.ld + 4 * l);
        const bool explore = !deterministic && ((uint64_t)U.draw(i, 0) < thresh);
        float qsa;
        const int a = select_group<LPA>(v, valid, T.A, explore, empty_all != 0, U.draw(i, 1), gm, &qsa);
        if (l == 0) actions[i] = a;
    }
}

// general action counts (A > 32 or byte masks): one thread per agent, three passes over the row
__global__ void select_generic_kernel(Table T, const int32_t* __restrict__ states, const uint8_t* __restrict__ mask_bytes,
                                      Uniforms U, uint64_t thresh, int deterministic, int empty_all,
                                      int32_t* __restrict__ actions, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = T.q + (size_t)states[i] * T.ld;
    const uint8_t* mk = mask_bytes ? mask_bytes + (size_t)i * T.A : nullptr;
    const bool explore = !deterministic && ((uint64_t)U.draw(i, 0) < thresh);
    float best = -INFINITY;
    int nvalid = 0;
    for (int a = 0; a < T.A; ++a)
        if (!mk || mk[a]) { ++nvalid; best = fmax_plain(best, row[a]); }
    int cnt = 0;
    const bool all = (!explore && nvalid == 0 && empty_all);
    if (explore) cnt = nvalid;
    else if (all) cnt = T.A;
    else
        for (int a = 0; a < T.A; ++a) cnt += ((!mk || mk[a]) && row[a] == best);
    int res = -1;
    if (cnt > 0) {
        int idx = (int)pick(U.draw(i, 1), (uint32_t)cnt);
        for (int a = 0; a < T.A; ++a) {
            const bool c = all || ((!mk || mk[a]) && (explore || row[a] == best));
            if (c) { if (idx == 0) { res = a; break; } --idx; }
        }
    }
    actions[i] = res;
}

// ------------------------------------------------------------------ unfused exact learn (cooperative)
template <int LPA>
__global__ void __launch_bounds__(256) learn_exact_kernel(Table T, int64_t S, const int32_t* __restrict__ states,
                                                          const int32_t* __restrict__ actions, const float* __restrict__ rewards,
                                                          const int32_t* __restrict__ next_states,
                                                          const uint8_t* __restrict__ terminated,
                                                          const uint32_t* __restrict__ next_mask_bits, float lr, float gamma,
                                                          uint32_t tag, uint32_t epoch, int wipe, int n) {
    cg::grid_group grid = cg::this_grid();
    if (wipe) { wipe_heads(T, S); grid.sync(); }
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    for (int i = tid; i < n; i += nthreads) {
        const int s = states[i], a = actions[i];
        list_insert(T, i, s, a, __ldcg(T.q + (size_t)s * T.ld + a), tag);
    }
    grid.sync();
    const uint32_t gm = group_mask<LPA>();
    const int groups = nthreads / LPA;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int gl = (threadIdx.x & 31) / LPA;
    for (int ib = (tid / 32) * (32 / LPA); ib < n; ib += groups) {
        const int i = ib + gl;
        if (i < n) {
            const uint32_t m2 = next_mask_bits ? (next_mask_bits[i] & full) : full;
            learn_resolve<LPA>(T, i, states[i], actions[i], rewards[i], T.tr_p[i], next_states[i], terminated[i] != 0, m2, lr,
                               gamma, tag, epoch, gm);
        }
        __syncwarp();
    }
}

// general fallback (A > 32 / byte masks): the reference loop itself, one warp walks the agents in order
__global__ void learn_sequential_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                                        const float* __restrict__ rewards, const int32_t* __restrict__ next_states,
                                        const uint8_t* __restrict__ terminated, const uint8_t* __restrict__ next_mask_bytes,
                                        const uint32_t* __restrict__ next_mask_bits, float lr, float gamma, int n) {
    const int lane = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        float m = 0.0f;
        if (!terminated[i]) {
            const volatile float* row = T.q + (size_t)next_states[i] * T.ld;
            m = -INFINITY;
            for (int a = lane; a < T.A; a += 32) {
                const bool ok = next_mask_bytes ? next_mask_bytes[(size_t)i * T.A + a] != 0
                                                : (next_mask_bits ? ((next_mask_bits[i] >> a) & 1u) != 0 : true);
                if (ok) m = fmax_plain(m, row[a]);
            }
            for (int d = 16; d > 0; d >>= 1) m = fmax_plain(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
            if (m == -INFINITY && lane == 0) atomicOr(T.err, kErrEmpty);
        }
        if (lane == 0) {
            volatile float* cell = T.q + (size_t)states[i] * T.ld + actions[i];
            *cell = td_value(*cell, rewards[i], m, lr, gamma);
        }
        __syncwarp();
    }
}

// accumulate mode (learn_vec, QLO:853-891): deltas from the snapshot, then atomic scatter-add
__global__ void learn_delta_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                                   const float* __restrict__ rewards, const int32_t* __restrict__ next_states,
                                   const uint8_t* __restrict__ terminated, const uint8_t* __restrict__ next_mask_bytes,
                                   const uint32_t* __restrict__ next_mask_bits, float lr, float gamma, float* __restrict__ delta,
                                   int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = T.q + (size_t)next_states[i] * T.ld;
    float m = -INFINITY;
    for (int a = 0; a < T.A; ++a) {
        const bool ok = next_mask_bytes ? next_mask_bytes[(size_t)i * T.A + a] != 0
                                        : (next_mask_bits ? ((next_mask_bits[i] >> a) & 1u) != 0 : true);
        if (ok) m = fmax_plain(m, row[a]);
    }
    // targets = r + gamma * max * (1 - terminated)
    const float target = rewards[i] + gamma * m * (terminated[i] ? 0.0f : 1.0f);
    delta[i] = lr * (target - T.q[(size_t)states[i] * T.ld + actions[i]]);
}
__global__ void learn_scatter_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                                     const float* __restrict__ delta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(T.q + (size_t)states[i] * T.ld + actions[i], delta[i]);
}

__global__ void gather_kernel(Table T, const int32_t* __restrict__ states, const int32_t* __restrict__ actions,
                              float* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = T.q[(size_t)states[i] * T.ld + actions[i]];
}

__global__ void gather_rows_kernel(Table T, const int32_t* __restrict__ states, float* __restrict__ out, int n) {
    const size_t total = (size_t)n * T.A;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t i = x / T.A;
        out[x] = T.q[(size_t)states[i] * T.ld + (x - i * T.A)];
    }
}

__global__ void table_fill_kernel(Table T, int64_t S, float value, uint32_t seed, int random) {
    const size_t total = (size_t)S * T.ld;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const int a = (int)(x % T.ld);
        const size_t s = x / T.ld;
        float v = 0.0f;
        if (a < T.A) v = random ? (float)(fmix32((uint32_t)(s * (size_t)T.A + a) ^ (seed * kGold)) >> 8) * 5.9604644775390625e-08f : value;
        T.q[x] = v;
    }
}

// ------------------------------------------------------------------ unfused environments (one thread per agent)
__global__ void ttt_reset_kernel(uint32_t* __restrict__ boards, int32_t* __restrict__ states, uint32_t* __restrict__ masks,
                                 Uniforms U, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = ttt_reset(U.draw(i, 3), U.draw(i, 4));
    boards[i] = b;
    states[i] = ttt_state(b & 0x3FFFFu);
    masks[i] = ttt_empties(b);
}
__global__ void ttt_step_kernel(uint32_t* __restrict__ boards, const int32_t* __restrict__ actions, Uniforms U,
                                int32_t* __restrict__ next_states, uint32_t* __restrict__ next_masks, float* __restrict__ rewards,
                                uint8_t* __restrict__ terminated, int* err, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t b = boards[i];
    float r = 0.0f;
    bool term = false;
    if (!ttt_step(b, actions[i], U.draw(i, 2), U.draw(i, 3), U.draw(i, 4), r, term)) { atomicOr(err, kErrInvalidMove); return; }
    boards[i] = b;
    next_states[i] = ttt_state(b & 0x3FFFFu);
    next_masks[i] = ttt_empties(b);
    rewards[i] = r;
    terminated[i] = term;
}
__global__ void mdp_reset_kernel(int32_t* __restrict__ states, uint32_t* __restrict__ masks, uint32_t S, int A, uint32_t env_seed,
                                 Uniforms U, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = (int32_t)pick(U.draw(i, 3), S);
    states[i] = s;
    if (masks) masks[i] = mdp_mask((uint32_t)s, A, env_seed);
}
__global__ void mdp_step_kernel(int32_t* __restrict__ states, const int32_t* __restrict__ actions, uint32_t S, int A,
                                uint32_t env_seed, uint64_t term_thresh, Uniforms U, uint32_t* __restrict__ next_masks,
                                float* __restrict__ rewards, uint8_t* __restrict__ terminated, int* err, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = actions[i];
    if (a < 0 || a >= A) { atomicOr(err, kErrInvalidMove); return; }
    int32_t s = states[i];
    float r;
    bool term;
    mdp_step(s, a, S, A, env_seed, term_thresh, U.draw(i, 2), U.draw(i, 3), r, term);
    states[i] = s;
    if (next_masks) next_masks[i] = mdp_mask((uint32_t)s, A, env_seed);
    rewards[i] = r;
    terminated[i] = term;
}

// ------------------------------------------------------------------ fused persistent loop
struct FusedArgs {
    int env_kind, n, steps;
    int32_t* st_a;   // states (current at even local steps)
    int32_t* st_b;   // scratch
    uint32_t* envw;
    float* ep_ret;
    uint8_t* tr_a;   // [cap] action | term << 7
    float* tr_r;     // [cap]
    uint32_t env_seed, episode_len;
    uint64_t term_thresh;
    int64_t S;
    const uint64_t* eps_thresh;  // device [steps]
    const float* lr;             // device [steps]
    const uint32_t* uniforms;
    int slots;
    uint32_t stream_seed, t0, agent0, env_stream_seed, env_t0;
    int empty_all, use_masks;
    float gamma;
    uint32_t step0;              // engine-global step counter at launch (epoch/tag source)
    int32_t* trace_actions;
    float* trace_rewards;
    uint8_t* trace_term;
    int32_t* trace_next;
    float* trace_epret;
    double* ep_sum;
    unsigned long long* ep_count;
};

template <int ENV>
__device__ __forceinline__ uint32_t env_mask(int s, uint32_t envw, int A, uint32_t env_seed) {
    if (ENV == 0) return mdp_mask((uint32_t)s, A, env_seed);
    if (ENV == 1) return ttt_empties(envw);
    return A >= 32 ? 0xFFFFFFFFu : ((1u << A) - 1u);
}

template <int ENV, int LPA>
__global__ void __launch_bounds__(256) fused_kernel(Table T, FusedArgs F) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_sum[8];
    __shared__ unsigned int s_cnt[8];
    const uint32_t gm = group_mask<LPA>();
    const int l = threadIdx.x & (LPA - 1);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int groups = (gridDim.x * blockDim.x) / LPA;
    const int gl = (threadIdx.x & 31) / LPA;      // group within the warp
    const int gw0 = (tid / 32) * (32 / LPA);       // first group of this warp
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;

    for (int k = 0; k < F.steps; ++k) {
        const uint32_t gstep = F.step0 + (uint32_t)k;
        const uint32_t tag = gstep % 255u + 1u;
        const uint32_t epoch = gstep + 1u;
        if (tag == 1u && gstep != 0u) { wipe_heads(T, F.S); grid.sync(); }
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + env step + writer-list insert
        for (int ib = gw0; ib < n; ib += groups) {
            const int i = ib + gl;
            if (i < n) {
            const int s = cur[i];
            uint32_t ew = (ENV == 0) ? 0u : F.envw[i];
            const uint32_t valid = F.use_masks ? env_mask<ENV>(s, ew, T.A, F.env_seed) : full;
            const float4 v = ld_row4(T.q + (size_t)s * T.ld + 4 * l);
            const bool explore = (uint64_t)U.draw(i, 0) < thresh;
            float p;
            const int a = select_group<LPA>(v, valid, T.A, explore, F.empty_all != 0, U.draw(i, 1), gm, &p);
            int32_t s2 = s;
            float r = 0.0f;
            bool term = false;
            bool ok = a >= 0;
            if (!ok) { if (l == 0) atomicOr(T.err, kErrEmpty); }
            else if (ENV == 0) mdp_step(s2, a, (uint32_t)F.S, T.A, F.env_seed, F.term_thresh, U.draw(i, 2), U.draw(i, 3), r, term);
            else if (ENV == 1) {
                ok = ttt_step(ew, a, U.draw(i, 2), U.draw(i, 3), U.draw(i, 4), r, term);
                if (!ok) { if (l == 0) atomicOr(T.err, kErrInvalidMove); }
                s2 = ttt_state(ew & 0x3FFFFu);
            } else {  // bandit: reward = action, terminate every episode_len steps (rigged_two_armed_bandit.py:71-80)
                r = (float)a;
                ew += 1u;
                term = ew >= F.episode_len;
                if (term) ew = 0u;
                s2 = 0;
            }
            if (l == 0) {
                nxt[i] = s2;
                if (ENV != 0) F.envw[i] = ew;
                const int aa = ok ? a : 0;
                F.tr_a[i] = (uint8_t)(aa | (term ? 0x80 : 0) | (ok ? 0 : 0x40));
                F.tr_r[i] = r;
                float acc = F.ep_ret[i] + r;  // BRT:212
                float fin = __int_as_float(0x7FC00000);
                if (term) { fin = acc; loc_sum += (double)acc; ++loc_cnt; acc = 0.0f; }  // BRT:218-221
                F.ep_ret[i] = acc;
                const size_t o = (size_t)k * n + i;
                if (F.trace_actions) F.trace_actions[o] = a;
                if (F.trace_rewards) F.trace_rewards[o] = r;
                if (F.trace_term) F.trace_term[o] = term;
                if (F.trace_next) F.trace_next[o] = s2;
                if (F.trace_epret) F.trace_epret[o] = fin;
                if (ok) list_insert(T, i, s, a, p, tag);
            }
            }
            __syncwarp();  // reconverge the lane groups every iteration
        }
        if (F.ep_count) {  // block-level reduction of the episode statistics, one atomic per block
            for (int d = 16; d > 0; d >>= 1) {
                loc_sum += __shfl_xor_sync(0xFFFFFFFFu, loc_sum, d);
                loc_cnt += __shfl_xor_sync(0xFFFFFFFFu, loc_cnt, d);
            }
            if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = loc_sum; s_cnt[threadIdx.x >> 5] = loc_cnt; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double bs = 0.0;
                unsigned int bc = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { bs += s_sum[w]; bc += s_cnt[w]; }
                if (bc) { atomicAdd(F.ep_sum, bs); atomicAdd(F.ep_count, (unsigned long long)bc); }
            }
        }
        grid.sync();

        // ---------------- phase B: exact sequential TD update (resolve + commit)
        for (int ib = gw0; ib < n; ib += groups) {
            const int i = ib + gl;
            if (i < n) {
                const uint8_t at = F.tr_a[i];
                if (!(at & 0x40)) {  // else: agent had no legal action (error already flagged)
                    const int s2 = nxt[i];
                    const uint32_t ew = (ENV == 1) ? F.envw[i] : 0u;
                    const uint32_t m2 = F.use_masks ? env_mask<ENV>(s2, ew, T.A, F.env_seed) : full;
                    learn_resolve<LPA>(T, i, cur[i], at & 0x3F, F.tr_r[i], T.tr_p[i], s2, (at & 0x80) != 0, m2, lr, F.gamma, tag, epoch, gm);
                }
            }
            __syncwarp();
        }
        grid.sync();
    }
    // leave the current observation in F.st_a
    if (F.steps & 1) {
        const int nthreads = gridDim.x * blockDim.x;
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
