// qe_common.cuh -- device-side building blocks of the B200 Q-learning engine (sm_100a).
//
// Semantics follow SURVEY.md Appendix B (one vector step of the reference loop):
//   select   : OptimalQLearningBase.choose_masked_action[_vec]   (QLO:304-348, 432-470)
//   TD update: OptimalQLearningBase.single_learn                 (QLO:728-768), fp32, no FMA
//   TicTacToe: TicTacToeEnv.step/reset (TTT:96-171, 183-237) + flatten radix (FLT:156-160, UTL:26-29)
//              + gymnasium SyncVectorEnv SAME_STEP autoreset
//   hash MDP : synthetic tabular MDP (new)
// The integer hash / uniform stream are the ones of oracle/rng.py (restated independently there).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qe {

// ------------------------------------------------------------------ hash + uniform stream
constexpr uint32_t kGold = 0x9E3779B9u;
constexpr uint32_t kStreamAdd = 0x7F4A7C15u;
constexpr uint32_t kSeedMix = 0x632BE5ABu;

__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x, uint32_t salt) { return fmix32(x + kGold * (salt + 1u)); }
// U[t, i, k]
__host__ __device__ __forceinline__ uint32_t stream_u32(uint32_t seed, uint32_t t, uint32_t i, uint32_t k) {
    return fmix32(fmix32((i * 8u + k) ^ (seed * kGold)) + t * kGold + kStreamAdd);
}
// (bits * n) >> 32
__host__ __device__ __forceinline__ uint32_t pick(uint32_t bits, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(bits, n);
#else
    return (uint32_t)(((uint64_t)bits * n) >> 32);
#endif
}

// Source of the per-step uniforms: a pre-drawn array U[N][K] (row of this vector step) or the counter stream.
// Slots 0,1 (select) come from the algorithm's stream (seed, t); slots >= 2 (environment) from the environment's
// stream (env_seed, env_t) -- the reference keeps two independent generators (QLO:97-98, TTT:88-94).
struct Uniforms {
    const uint32_t* pre;  // nullptr -> counter stream
    int slots;            // K
    uint32_t seed, t, agent0;
    uint32_t env_seed, env_t;
    const uint32_t* ids = nullptr;  // optional global agent ids (sharded table: local agent i is global agent ids[i])
    __device__ __forceinline__ uint32_t draw(int i, int k) const {
        if (pre) return __ldg(pre + (size_t)i * slots + k);
        const uint32_t g = ids ? __ldg(ids + i) : agent0 + (uint32_t)i;
        return k < 2 ? stream_u32(seed, t, g, (uint32_t)k) : stream_u32(env_seed, env_t, g, (uint32_t)k);
    }
};

// ------------------------------------------------------------------ cache-hinted memory ops
__device__ __forceinline__ float4 ld_row4(const float* p) {  // L2-only (random gather, no L1 reuse)
    return __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------ lane groups (LPA lanes per agent)
template <int LPA>
__device__ __forceinline__ uint32_t group_mask() {
    const uint32_t lane = threadIdx.x & 31u;
    return (LPA == 32) ? 0xFFFFFFFFu : (((1u << LPA) - 1u) << (lane & ~(uint32_t)(LPA - 1)));
}
template <int LPA>
__device__ __forceinline__ float group_max(float v, uint32_t gm) {
#pragma unroll
    for (int d = LPA / 2; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(gm, v, d));
    return v;
}
template <int LPA>
__device__ __forceinline__ uint32_t group_or(uint32_t v, uint32_t gm) {
#pragma unroll
    for (int d = LPA / 2; d > 0; d >>= 1) v |= __shfl_xor_sync(gm, v, d);
    return v;
}
template <int LPA>
__device__ __forceinline__ bool group_all(bool p, uint32_t gm) {
    if (LPA == 1) return p;
    return (__ballot_sync(gm, p) & gm) == gm;
}

// max that ignores nothing: numpy max / python '>' on finite values; NaN handling is not part of the contract.
__device__ __forceinline__ float fmax_plain(float a, float b) { return a > b ? a : b; }

// ------------------------------------------------------------------ select (one agent, group-cooperative)
// Lane l of the group holds row[4l .. 4l+3] in v.  Returns the chosen action (all lanes), -1 if no candidate,
// and the table value of the chosen action in *q_sa.
template <int LPA>
__device__ __forceinline__ int select_group(float4 v, uint32_t valid, int num_actions, bool explore, bool empty_all,
                                            uint32_t bits_pick, uint32_t gm, float* q_sa) {
    const int l = threadIdx.x & (LPA - 1);
    const uint32_t my = (valid >> (4 * l)) & 0xFu;
    uint32_t cand;
    if (explore) {
        cand = valid;  // QLO:336 / :465 -- every legal action, ascending
    } else {
        float m = -INFINITY;
        if (my & 1u) m = fmax_plain(m, v.x);
        if (my & 2u) m = fmax_plain(m, v.y);
        if (my & 4u) m = fmax_plain(m, v.z);
        if (my & 8u) m = fmax_plain(m, v.w);
        m = group_max<LPA>(m, gm);
        uint32_t tie = 0;
        if ((my & 1u) && v.x == m) tie |= 1u;  // exact == on fp32 (QLO:346 / :469)
        if ((my & 2u) && v.y == m) tie |= 2u;
        if ((my & 4u) && v.z == m) tie |= 4u;
        if ((my & 8u) && v.w == m) tie |= 8u;
        cand = group_or<LPA>(tie << (4 * l), gm);
        if (valid == 0u && empty_all)  // QLO:467-470: every masked value is -inf -> all actions tie
            cand = num_actions >= 32 ? 0xFFFFFFFFu : ((1u << num_actions) - 1u);
    }
    const int cnt = __popc(cand);
    int a = -1;
    if (cnt > 0) a = (int)__fns(cand, 0, (int)pick(bits_pick, (uint32_t)cnt) + 1);  // choice(cand) (QLO:348/:470)
    // value of Q[s, a]: lane a/4 holds it
    const int sel = a < 0 ? 0 : a;
    const int k = sel & 3;
    float mine = k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
    *q_sa = __shfl_sync(gm, mine, (threadIdx.x & 31 & ~(LPA - 1)) + (sel >> 2));
    return a;
}

// ------------------------------------------------------------------ fp32 TD update, one rounding per op (QLO:766-768)
__device__ __forceinline__ float td_value(float p, float r, float m, float lr, float gamma) {
    const float gm = __fmul_rn(gamma, m);
    const float target = __fadd_rn(r, gm);
    const float d = __fsub_rn(target, p);
    return __fadd_rn(p, __fmul_rn(lr, d));
}

// ------------------------------------------------------------------ TicTacToe (board: 2 bits / cell, bit 18 = agent_mark-1)
__host__ __device__ __forceinline__ uint32_t ttt_empties(uint32_t b) {
    const uint32_t occ = (b | (b >> 1)) & 0x15555u;  // bit 2c set if cell c occupied
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) m |= ((~occ >> (2 * c)) & 1u) << c;
    return m;
}
__host__ __device__ __forceinline__ bool ttt_line(uint32_t b, uint32_t mark) {
    // planes: bit 2c of `pl` set if cell c holds `mark`
    const uint32_t pl = (mark == 1u ? (b & ~(b >> 1)) : ((b >> 1) & ~b)) & 0x15555u;
    constexpr uint32_t L[8] = {0x15u, 0x540u, 0x15000u, 0x1041u, 0x4104u, 0x10410u, 0x10101u, 0x1110u};
    bool w = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) w |= (pl & L[i]) == L[i];
    return w;
}
__host__ __device__ __forceinline__ int32_t ttt_state(uint32_t b) {  // base 3, cell 0 most significant
    int32_t s = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) s = s * 3 + (int32_t)((b >> (2 * c)) & 3u);
    return s;
}
__host__ __device__ __forceinline__ uint32_t ttt_reset(uint32_t bits_coin, uint32_t bits_open) {  // TTT:96-108
    if (pick(bits_coin, 2) == 0) return 0u;                       // choice([True, False])[0] -> agent starts, mark 1
    return (1u << (2 * pick(bits_open, 9))) | (1u << 18);         // machine (mark 1) opens, agent is mark 2
}
__device__ __forceinline__ int kth_set(uint32_t m, int k) { return (int)__fns(m, 0, k + 1); }

// returns false on an illegal move (reference: AssertionError "Invalid move.", TTT:130)
__device__ __forceinline__ bool ttt_step(uint32_t& board, int action, uint32_t bm, uint32_t bcoin, uint32_t bopen, float& reward,
                                         bool& term) {
    uint32_t b = board;
    const uint32_t amark = ((b >> 18) & 1u) + 1u, mmark = 3u - amark;
    if (action < 0 || action > 8 || ((b >> (2 * action)) & 3u) != 0u) return false;
    b |= amark << (2 * action);
    reward = 0.0f;
    term = false;
    if (ttt_line(b, amark)) { reward = 1.0f; term = true; }
    else {
        const uint32_t e = ttt_empties(b);
        if (e == 0u) term = true;
        else {
            const int c = kth_set(e, (int)pick(bm, (uint32_t)__popc(e)));  // k-th empty cell ascending (TTT:183-197)
            b |= mmark << (2 * c);
            if (ttt_line(b, mmark)) { reward = -1.0f; term = true; }
            else if (ttt_empties(b) == 0u) term = true;
        }
    }
    if (term) b = ttt_reset(bcoin, bopen);  // SAME_STEP autoreset
    board = b;
    return true;
}

// ------------------------------------------------------------------ hash MDP
__host__ __device__ __forceinline__ uint32_t mdp_mask(uint32_t s, int A, uint32_t env_seed) {
    const uint32_t full = A >= 32 ? 0xFFFFFFFFu : ((1u << A) - 1u);
    return (mix32(s + env_seed * kSeedMix, 2u) & full) | 1u;
}
__device__ __forceinline__ void mdp_step(int32_t& state, int action, uint32_t S, int A, uint32_t env_seed, uint64_t term_thresh,
                                         uint32_t bterm, uint32_t breset, float& reward, bool& term) {
    const uint32_t h = mix32((uint32_t)state * (uint32_t)A + (uint32_t)action + env_seed * kSeedMix, 0u);
    const uint32_t h2 = mix32(h, 1u);
    reward = __fsub_rn(__fmul_rn(__fmul_rn((float)(h2 >> 8), 5.9604644775390625e-08f), 2.0f), 1.0f);
    term = (uint64_t)bterm < term_thresh;
    state = (int32_t)(term ? pick(breset, S) : pick(h, S));
}

}  // namespace qe
