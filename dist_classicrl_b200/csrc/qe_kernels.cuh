// qe_kernels.cuh -- CUDA kernels of the engine (sm_100a): select, exact sequential TD update, env steps, fused loop.
//
// Table layout in HBM.  One *row block* per state: [Q row: 4*LPA floats][writer info: 8 words][spare inline entries],
// padded to a power of two (A=16: 128 B = one L2 line; A=8: 64 B).  A row gather therefore brings the writer info
// of that state along in the same line/DRAM page.
//
// Exact sequential TD update in parallel (DESIGN.md "TD update"):
//   The reference applies agents 0..N-1 one after the other (QLO:806-817), so agent i must see every write of
//   agents j<i -- to its own cell Q[s_i,a_i] AND to the row Q[s'_i,:] it bootstraps from.
//   Phase 1 (insert)  : every agent captures p_i = Q0[s_i,a_i] and registers itself as a writer of row s_i:
//                       count = atomicAdd on the row's epoch-stamped counter, entry {agent, action} stored inline
//                       in the row block (overflow: linked list through node[]).
//   -- grid-wide barrier --
//   Phase 2 (resolve) : agent i scans the writer entries of s_i and s'_i.  For every cell it needs it finds the
//                       latest writer j<i; the value "just before i" is slot[j] if such a writer exists, a writer's
//                       captured p (== Q0; the table may already hold a later commit) if the cell is written only
//                       by agents >= i, and the table itself if nobody writes the cell this step.
//                       v_i = p + lr*((r + gamma*m) - p) is published in slot[i] = {epoch, v_i} (one 64-bit store);
//                       waiting agents poll their predecessors' slots.  The dependency graph is a DAG in agent
//                       order and every warp processes its agents in increasing order, so the smallest unresolved
//                       agent can always proceed (cooperative launch => all CTAs are co-resident).
//                       The last writer of a cell commits v to the table.
//   Everything is stamped with a 32-bit epoch (global step counter), so nothing is cleared between steps.
//
// Thread mapping: scalar work (hashing, env step, writer scans, agent-array traffic) is one lane per agent, 32
// consecutive agents per warp iteration (coalesced); row gathers are transposed so that LPA lanes fetch one row with
// one 16-byte load each (one 128-byte line per row, not one per lane).
#pragma once
#include <cooperative_groups.h>

#include "qe_common.cuh"

namespace qe {
namespace cg = cooperative_groups;

constexpr uint32_t kNone = 0xFFFFFFu;       // list terminator (24-bit agent index)
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kErrInvalidMove = 1, kErrEmpty = 2, kErrTimeout = 4;
constexpr uint32_t kSpinLimit = 1u << 22;

struct Table {
    float* q;          // [S][ld] row blocks
    int ld;            // floats per row block (power of two)
    int A;             // actions
    int info_off;      // float offset of the writer info inside a row block
    int inline_cap;    // writer entries stored inline in the row block (>= 4)
    uint32_t* node;    // [cap] overflow list: (action << 24) | next24
    uint64_t* slot;    // [cap] (epoch << 32) | float bits of v_i
    float* tr_p;       // [cap] Q0[s_i, a_i] captured before any commit of this step
    int* err;          // device error flags
};
// writer info words: [0] count, [1] epoch (one u64, atomics), [2] overflow head idx, [3] its epoch (one u64),
//                    [4 .. 4+inline_cap) entries (agent24 | action << 24)
__device__ __forceinline__ uint32_t* row_info(const Table& T, int s) {
    return reinterpret_cast<uint32_t*>(T.q + (size_t)s * T.ld + T.info_off);
}

// ------------------------------------------------------------------ phase 1: register agent i as a writer of row s
__device__ __forceinline__ void row_insert(const Table& T, int i, int s, int a, float p, uint32_t epoch) {
    T.tr_p[i] = p;
    unsigned long long* info = reinterpret_cast<unsigned long long*>(row_info(T, s));
    const unsigned long long base = (unsigned long long)epoch << 32;
    atomicMax(info, base);  // a stale (older-epoch) counter restarts at {epoch, 0}
    const uint32_t c = (uint32_t)atomicAdd(info, 1ull);
    if (c < (uint32_t)T.inline_cap) {
        reinterpret_cast<uint32_t*>(info)[4 + c] = (uint32_t)i | ((uint32_t)a << 24);
    } else {
        const unsigned long long old = atomicExch(info + 1, base | (unsigned long long)i);
        T.node[i] = ((uint32_t)a << 24) | (((uint32_t)(old >> 32) == epoch) ? ((uint32_t)old & kNone) : kNone);
    }
}

struct RowWriters {  // snapshot of a row's writer info (taken after the grid barrier)
    const uint32_t* iw;
    uint4 e;          // first four inline entries
    uint32_t count;   // writers of this row in the current step (0 if the info is stale)
    uint32_t ovf;     // overflow list head or kNone
};
__device__ __forceinline__ RowWriters load_writers(const Table& T, int s, uint32_t epoch) {
    RowWriters w;
    w.iw = row_info(T, s);
    const uint4 h = __ldcg(reinterpret_cast<const uint4*>(w.iw));
    w.count = (h.y == epoch) ? h.x : 0u;
    w.ovf = (h.w == epoch) ? (h.z & kNone) : kNone;
    w.e = make_uint4(0, 0, 0, 0);
    if (w.count) w.e = __ldcg(reinterpret_cast<const uint4*>(w.iw) + 1);
    return w;
}
template <typename F>
__device__ __forceinline__ void for_each_writer(const Table& T, const RowWriters& w, F f) {
    if (w.count > 0) f(w.e.x);
    if (w.count > 1) f(w.e.y);
    if (w.count > 2) f(w.e.z);
    if (w.count > 3) f(w.e.w);
    if (w.count > 4) {
        const uint32_t ninl = min(w.count, (uint32_t)T.inline_cap);
        for (uint32_t k = 4; k < ninl; ++k) f(__ldcg(w.iw + 4 + k));
        if (w.count > (uint32_t)T.inline_cap) {
            uint32_t j = w.ovf;
            while (j != kNone) {
                const uint32_t nd = __ldcg(T.node + j);
                f(j | (nd & 0xFF000000u));
                j = nd & kNone;
            }
        }
    }
}

// ------------------------------------------------------------------ transposed row gathers (all 32 lanes participate)
// Each lane owns one agent; in sub-iteration q the LPA lanes of group g fetch the row of the agent owned by lane
// q*(32/LPA)+g with one float4 each and reduce; the result travels back to the owner lane.
template <int LPA>
__device__ __forceinline__ float coop_row_max(const Table& T, int s, uint32_t mask) {
    constexpr int G = 32 / LPA;
    const int lane = threadIdx.x & 31, l = lane & (LPA - 1), g = lane / LPA;
    float res = -INFINITY;
#pragma unroll
    for (int q = 0; q < LPA; ++q) {
        const int src = q * G + g;
        const int ss = __shfl_sync(kFull, s, src);
        const uint32_t my = (__shfl_sync(kFull, mask, src) >> (4 * l)) & 0xFu;
        float m = -INFINITY;
        if (my) {
            const float4 v = ld_row4(T.q + (size_t)ss * T.ld + 4 * l);
            if (my & 1u) m = fmax_plain(m, v.x);
            if (my & 2u) m = fmax_plain(m, v.y);
            if (my & 4u) m = fmax_plain(m, v.z);
            if (my & 8u) m = fmax_plain(m, v.w);
        }
#pragma unroll
        for (int d = LPA / 2; d > 0; d >>= 1) m = fmax_plain(m, __shfl_xor_sync(kFull, m, d));
        const float back = __shfl_sync(kFull, m, (lane & (G - 1)) * LPA);
        if (lane / G == q) res = back;
    }
    return res;
}
template <int LPA>
__device__ __forceinline__ int coop_select(const Table& T, int s, uint32_t valid, bool explore, bool empty_all,
                                           uint32_t bits_pick, float* q_sa) {
    constexpr int G = 32 / LPA;
    const int lane = threadIdx.x & 31, l = lane & (LPA - 1), g = lane / LPA;
    const uint32_t gm = group_mask<LPA>();  // `explore` differs between groups: group-scoped shuffles inside
    int res = -1;
    float qres = 0.0f;
#pragma unroll
    for (int q = 0; q < LPA; ++q) {
        const int src = q * G + g;
        const int ss = __shfl_sync(kFull, s, src);
        const uint32_t vv = __shfl_sync(kFull, valid, src);
        const bool ex = __shfl_sync(kFull, (int)explore, src) != 0;
        const uint32_t b1 = __shfl_sync(kFull, bits_pick, src);
        const float4 v = ld_row4(T.q + (size_t)ss * T.ld + 4 * l);
        float qsa;
        const int a = select_group<LPA>(v, vv, T.A, ex, empty_all, b1, gm, &qsa);
        __syncwarp();
        const int aback = __shfl_sync(kFull, a, (lane & (G - 1)) * LPA);
        const float qback = __shfl_sync(kFull, qsa, (lane & (G - 1)) * LPA);
        if (lane / G == q) { res = aback; qres = qback; }
    }
    *q_sa = qres;
    return res;
}

// ------------------------------------------------------------------ phase 2: resolve + commit (one lane per agent)
// `active` lanes own a real transition; every lane of the warp must call (transposed gathers inside).
// `best` points at this thread's column of a shared-memory table [4*LPA actions][256 threads]: for every contested
// action of the bootstrap row it holds the latest writer j<i (>= 0) or -(j+2) for "written only by agents >= i".
template <int LPA>
__device__ __forceinline__ void learn_resolve(const Table& T, int* best, bool active, int i, int s, int a, float r, float p,
                                              int s2, bool term, uint32_t mask2, float lr, float gamma, uint32_t epoch) {
    int pj = -1;
    bool later = false;
    uint32_t contested = 0;
    if (active) {
        const RowWriters w1 = load_writers(T, s, epoch);  // agent i itself is one of them
        for_each_writer(T, w1, [&](uint32_t w) {
            if ((int)(w >> 24) == a) {
                const int j = (int)(w & kNone);
                if (j < i) pj = max(pj, j);
                else if (j > i) later = true;
            }
        });
        if (!term) {
            if (mask2 == 0u) atomicOr(T.err, kErrEmpty);  // np.max of an empty selection (QLO:764)
            const RowWriters w2 = load_writers(T, s2, epoch);
            if (w2.count) {
                for (uint32_t b = mask2; b; b &= b - 1u) best[(__ffs(b) - 1) * 256] = -1;
                for_each_writer(T, w2, [&](uint32_t w) {  // one pass: latest earlier writer per action
                    const uint32_t aj = w >> 24;
                    if ((mask2 >> aj) & 1u) {
                        contested |= 1u << aj;
                        const int j = (int)(w & kNone);
                        int* b = best + aj * 256;
                        const int cur = *b;
                        if (j < i) *b = cur >= 0 ? max(cur, j) : j;
                        else if (cur == -1) *b = -(j + 2);
                    }
                });
            }
        }
    }
    const bool boot = active && !term;
    // cells of the bootstrap row nobody writes this step: straight from the table
    float m = coop_row_max<LPA>(T, boot ? s2 : 0, boot ? (mask2 & ~contested) : 0u);
    float pe = p;
    uint32_t dyn = contested;  // contested legal cells still to be folded into m
    bool need_pe = active && pj >= 0;
    bool done = !active;
    // Warp-level retry loop: no lane ever blocks while another lane of the same warp still has to publish (a lane
    // spinning inside a divergent region could otherwise wait for a lane parked at the reconvergence point).
    for (uint32_t spins = 0;; ++spins) {
        if (!done) {
            uint32_t left = 0;
            for (uint32_t b = dyn; b; b &= b - 1u) {  // value of each contested cell just before agent i
                const int a2 = __ffs(b) - 1;
                const int jb = best[a2 * 256];
                if (jb < 0) {  // written only by agents >= i: Q0 as captured by one of them
                    m = fmax_plain(m, __ldcg(T.tr_p + (-jb - 2)));
                } else {
                    const uint64_t w = ld_relaxed_u64(T.slot + jb);
                    if ((uint32_t)(w >> 32) == epoch) m = fmax_plain(m, __uint_as_float((uint32_t)w));
                    else left |= 1u << a2;
                }
            }
            dyn = left;
            if (need_pe) {
                const uint64_t w = ld_relaxed_u64(T.slot + pj);
                if ((uint32_t)(w >> 32) == epoch) { pe = __uint_as_float((uint32_t)w); need_pe = false; }
            }
            if (!dyn && !need_pe) {
                const float v = td_value(pe, r, term ? 0.0f : m, lr, gamma);
                st_relaxed_u64(T.slot + i, ((uint64_t)epoch << 32) | (uint64_t)__float_as_uint(v));
                if (!later) T.q[(size_t)s * T.ld + a] = v;  // last writer of the cell commits
                done = true;
            }
        }
        if (__all_sync(kFull, done)) break;
        if (spins > kSpinLimit) { if (!done) atomicOr(T.err, kErrTimeout); break; }
    }
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
__global__ void __launch_bounds__(256) learn_exact_kernel(Table T, const int32_t* __restrict__ states,
                                                          const int32_t* __restrict__ actions, const float* __restrict__ rewards,
                                                          const int32_t* __restrict__ next_states,
                                                          const uint8_t* __restrict__ terminated,
                                                          const uint32_t* __restrict__ next_mask_bits, float lr, float gamma,
                                                          uint32_t epoch, int n) {
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    for (int i = tid; i < n; i += nthreads) {
        const int s = states[i], a = actions[i];
        row_insert(T, i, s, a, __ldcg(T.q + (size_t)s * T.ld + a), epoch);
    }
    grid.sync();
    __shared__ int s_best[4 * LPA * 256];
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    for (int base = (tid & ~31); base < n; base += nthreads) {
        const int i = base + (threadIdx.x & 31);
        const bool active = i < n;
        const int ii = active ? i : 0;
        const uint32_t m2 = next_mask_bits ? (next_mask_bits[ii] & full) : full;
        learn_resolve<LPA>(T, s_best + threadIdx.x, active, i, states[ii], actions[ii], rewards[ii], T.tr_p[ii], next_states[ii],
                           terminated[ii] != 0, m2, lr, gamma, epoch);
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
    uint32_t step0;              // engine-global step counter at launch (epoch source)
    int* tile_counter;           // [2] phase-B tile cursors (double-buffered across steps), both 0 at launch
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
    __shared__ int s_best[4 * LPA * 256];
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    const uint32_t full = T.A >= 32 ? 0xFFFFFFFFu : ((1u << T.A) - 1u);
    const int n = F.n;

    for (int k = 0; k < F.steps; ++k) {
        const uint32_t epoch = F.step0 + (uint32_t)k + 1u;
        int32_t* cur = (k & 1) ? F.st_b : F.st_a;
        int32_t* nxt = (k & 1) ? F.st_a : F.st_b;
        Uniforms U{F.uniforms ? F.uniforms + (size_t)k * n * F.slots : nullptr, F.slots, F.stream_seed, F.t0 + (uint32_t)k, F.agent0,
                   F.env_stream_seed, F.env_t0 + (uint32_t)k};
        const uint64_t thresh = F.eps_thresh[k];
        const float lr = F.lr[k];
        double loc_sum = 0.0;
        unsigned int loc_cnt = 0;

        // ---------------- phase A: select + env step + writer registration (one lane per agent)
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
            float p;
            const int a = coop_select<LPA>(T, s, valid, explore, F.empty_all != 0, bits1, &p);
            if (active) {
                int32_t s2 = s;
                float r = 0.0f;
                bool term = false;
                bool ok = a >= 0;
                if (!ok) atomicOr(T.err, kErrEmpty);
                else if (ENV == 0) mdp_step(s2, a, (uint32_t)F.S, T.A, F.env_seed, F.term_thresh, U.draw(i, 2), U.draw(i, 3), r, term);
                else if (ENV == 1) {
                    ok = ttt_step(ew, a, U.draw(i, 2), U.draw(i, 3), U.draw(i, 4), r, term);
                    if (!ok) atomicOr(T.err, kErrInvalidMove);
                    s2 = ttt_state(ew & 0x3FFFFu);
                } else {  // bandit: reward = action, terminate every episode_len steps (rigged_two_armed_bandit.py:71-80)
                    r = (float)a;
                    ew += 1u;
                    term = ew >= F.episode_len;
                    if (term) ew = 0u;
                    s2 = 0;
                }
                nxt[i] = s2;
                if (ENV != 0) F.envw[i] = ew;
                F.tr_a[i] = (uint8_t)((ok ? a : 0) | (term ? 0x80 : 0) | (ok ? 0 : 0x40));
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
                if (ok) row_insert(T, i, s, a, p, epoch);
            }
            __syncwarp();
        }
        if (F.ep_count) {  // block-level reduction of the episode statistics, one atomic per block
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
        }
        grid.sync();

        // ---------------- phase B: exact sequential TD update (resolve + commit)
        // tiles of 32 agents are claimed dynamically in increasing order (balances the chain tails; a waiting agent's
        // predecessors always sit in tiles that were claimed earlier, i.e. by warps that are running)
        for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(F.tile_counter + (k & 1), 32);
            base = __shfl_sync(kFull, base, 0);
            if (base >= n) break;
            const int i = base + lane;
            bool active = i < n;
            const int ii = active ? i : 0;
            const uint8_t at = F.tr_a[ii];
            active = active && !(at & 0x40);  // 0x40: agent had no legal action (error already flagged)
            const int s2 = nxt[ii];
            const uint32_t ew = (ENV == 1) ? F.envw[ii] : 0u;
            const uint32_t m2 = F.use_masks ? env_mask<ENV>(s2, ew, T.A, F.env_seed) : full;
            learn_resolve<LPA>(T, s_best + threadIdx.x, active, i, cur[ii], at & 0x3F, F.tr_r[ii], T.tr_p[ii], s2, (at & 0x80) != 0,
                               m2, lr, F.gamma, epoch);
            __syncwarp();
        }
        if (tid == 0) F.tile_counter[(k + 1) & 1] = 0;  // the other counter is idle until the next step's phase B
        grid.sync();
    }
    // leave the current observation in F.st_a
    if (F.steps & 1) {
        for (int i = tid; i < n; i += nthreads) F.st_a[i] = F.st_b[i];
    }
}

}  // namespace qe
