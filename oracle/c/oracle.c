/*
 * oracle.c -- plain-C restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * NOT part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * `--impl reference` legs may load this library, and only as the checker or as the timed CPU arm.
 *
 * What it restates (file:line relative to /root/reference/src/dist_classicrl):
 *   orc_select        algorithms/base_algorithms/q_learning_optimal.py:304-348, 432-470 (masked
 *                     epsilon-greedy with uniform tie-break; all variants compute this function)
 *   orc_learn_seq     q_learning_optimal.py:728-768 driven by :770-817 / :893-934 (sequential
 *                     per-agent TD update, fp32, one rounding per operation -> -ffp-contract=off)
 *   orc_ttt_*         environments/tiktaktoe_mod.py:96-108 (reset), :110-171 (step),
 *                     :183-197 (machine move), :216-237 (winner); state id =
 *                     wrappers/flatten_multidiscrete_wrapper.py:156-160 + utils.py:26-29,48;
 *                     SAME_STEP autoreset = gymnasium SyncVectorEnv (third party, restated)
 *   orc_mdp_*         synthetic integer-hash MDP (new) -- same definition as oracle/envs.py
 *   orc_run_*         algorithms/runtime/base_runtime.py:184-222 loop (select, step, rewards +=,
 *                     learn, episode bookkeeping); schedules are evaluated by the caller.
 *
 * Pinned by tests/test_oracle_c.py against the NumPy oracle and the golden fixtures made from the
 * real reference (tests/golden/, oracle/make_golden.py).
 *
 * Build: make -C oracle/c   (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GOLD 0x9E3779B9u
#define STREAM_ADD 0x7F4A7C15u
#define SEEDMIX 0x632BE5ABu

static inline uint32_t fmix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
static inline uint32_t mix32(uint32_t x, uint32_t salt) { return fmix32(x + GOLD * (salt + 1u)); }
static inline uint32_t stream_u32(uint32_t seed, uint32_t t, uint32_t i, uint32_t k) {
    return fmix32(fmix32((i * 8u + k) ^ (seed * GOLD)) + t * GOLD + STREAM_ADD);
}
static inline uint32_t pick(uint32_t bits, uint32_t n) { return (uint32_t)(((uint64_t)bits * n) >> 32); }

uint32_t orc_stream_u32(uint32_t seed, uint32_t t, uint32_t i, uint32_t k) { return stream_u32(seed, t, i, k); }

/* uniform source: either a pre-drawn array U[N][K] for this step, or the counter stream */
typedef struct { const uint32_t* U; int K; uint32_t seed; uint32_t t; uint32_t agent0; } usrc_t;
static inline uint32_t udraw(const usrc_t* u, int i, int k) {
    return u->U ? u->U[(size_t)i * u->K + k] : stream_u32(u->seed, u->t, u->agent0 + (uint32_t)i, (uint32_t)k);
}

/* ---------------------------------------------------------------- select (one agent) */
static inline int select_one(const float* row, int A, uint32_t mask_bits, int has_mask, int explore, int empty_all,
                             uint32_t bits_pick) {
    uint32_t cand = 0;
    uint32_t valid = has_mask ? mask_bits : (A >= 32 ? 0xFFFFFFFFu : ((1u << A) - 1u));
    if (explore) {
        cand = valid;
    } else {
        float best = -INFINITY;
        for (int a = 0; a < A; ++a) if ((valid >> a) & 1u) { if (row[a] > best) best = row[a]; }
        for (int a = 0; a < A; ++a) if (((valid >> a) & 1u) && row[a] == best) cand |= 1u << a;
        if (valid == 0 && empty_all) cand = (A >= 32 ? 0xFFFFFFFFu : ((1u << A) - 1u));
    }
    int cnt = __builtin_popcount(cand);
    if (cnt == 0) return -1;
    int idx = (int)pick(bits_pick, (uint32_t)cnt);
    for (int a = 0; a < A; ++a) if ((cand >> a) & 1u) { if (idx == 0) return a; --idx; }
    return -1;
}

/* masks: either mask_bits[N] (A<=32) or NULL (no masks). U is [N][K] pre-drawn. */
void orc_select(const float* q, int A, const int32_t* states, const uint32_t* mask_bits, uint64_t explore_thresh,
                int deterministic, int empty_all, const uint32_t* U, int K, int N, int32_t* actions) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        int explore = !deterministic && ((uint64_t)U[(size_t)i * K + 0] < explore_thresh);
        actions[i] = select_one(q + (size_t)states[i] * A, A, mask_bits ? mask_bits[i] : 0, mask_bits != NULL, explore,
                                empty_all, U[(size_t)i * K + 1]);
    }
}

/* ---------------------------------------------------------------- sequential TD update */
static inline void learn_one(float* q, int A, int s, int a, float r, int s2, int term, uint32_t mask2, int has_mask,
                             float lr, float gamma) {
    float m = 0.0f;
    if (!term) {
        const float* row = q + (size_t)s2 * A;
        m = -INFINITY;
        for (int b = 0; b < A; ++b)
            if (!has_mask || ((mask2 >> b) & 1u)) { if (row[b] > m) m = row[b]; }
    }
    float gm = gamma * m;          /* each statement rounds once: no FMA (-ffp-contract=off) */
    float target = r + gm;
    float* cell = q + (size_t)s * A + a;
    float p = *cell;
    float d = target - p;
    float ld = lr * d;
    *cell = p + ld;
}

void orc_learn_seq(float* q, int A, const int32_t* s, const int32_t* a, const float* r, const int32_t* s2,
                   const uint8_t* term, const uint32_t* mask_bits2, float lr, float gamma, int N) {
    for (int i = 0; i < N; ++i)
        learn_one(q, A, s[i], a[i], r[i], s2[i], term[i], mask_bits2 ? mask_bits2[i] : 0, mask_bits2 != NULL, lr, gamma);
}

/* ---------------------------------------------------------------- TicTacToe */
/* board: 2 bits per cell, cell c at bits [2c, 2c+1]; bit 18 = agent_mark - 1 */
static const uint8_t LINES[8][3] = {{0,1,2},{3,4,5},{6,7,8},{0,3,6},{1,4,7},{2,5,8},{0,4,8},{2,4,6}};
static inline int cell(uint32_t b, int c) { return (b >> (2 * c)) & 3; }
static inline int has_line(uint32_t b, int mark) {
    for (int l = 0; l < 8; ++l)
        if (cell(b, LINES[l][0]) == mark && cell(b, LINES[l][1]) == mark && cell(b, LINES[l][2]) == mark) return 1;
    return 0;
}
static inline uint32_t empties(uint32_t b) { uint32_t m = 0; for (int c = 0; c < 9; ++c) if (cell(b, c) == 0) m |= 1u << c; return m; }
static inline int32_t ttt_state(uint32_t b) { int32_t s = 0; for (int c = 0; c < 9; ++c) s = s * 3 + cell(b, c); return s; }
static inline int kth_set(uint32_t m, int k) { for (int c = 0; c < 32; ++c) if ((m >> c) & 1u) { if (k == 0) return c; --k; } return -1; }

static inline uint32_t ttt_reset_one(uint32_t bits_coin, uint32_t bits_open) {
    uint32_t b = 0;
    int agent_starts = pick(bits_coin, 2) == 0;      /* choice([True, False]) */
    if (!agent_starts) { b |= 1u << (2 * pick(bits_open, 9)); b |= 1u << 18; }  /* machine = mark 1 opens */
    return b;
}

void orc_ttt_reset(uint32_t* boards, int32_t* states, uint32_t* masks, const uint32_t* U, int K, int N) {
    for (int i = 0; i < N; ++i) {
        boards[i] = ttt_reset_one(U[(size_t)i * K + 3], U[(size_t)i * K + 4]);
        states[i] = ttt_state(boards[i] & 0x3FFFF);
        masks[i] = empties(boards[i]);
    }
}

/* returns -1 on an illegal agent move (reference: AssertionError "Invalid move.") */
static inline int ttt_step_one(uint32_t* board, int action, uint32_t bm, uint32_t bcoin, uint32_t bopen, float* reward,
                               uint8_t* term) {
    uint32_t b = *board;
    int amark = ((b >> 18) & 1) + 1, mmark = 3 - amark;
    if (action < 0 || action > 8 || cell(b, action) != 0) return -1;
    b |= (uint32_t)amark << (2 * action);
    float r = 0.0f; int t = 0;
    if (has_line(b, amark)) { r = 1.0f; t = 1; }
    else if ((empties(b)) == 0) { t = 1; }
    else {
        uint32_t e = empties(b);
        int c = kth_set(e, (int)pick(bm, (uint32_t)__builtin_popcount(e)));
        b |= (uint32_t)mmark << (2 * c);
        if (has_line(b, mmark)) { r = -1.0f; t = 1; }
        else if (empties(b) == 0) { t = 1; }
    }
    if (t) b = ttt_reset_one(bcoin, bopen);          /* SAME_STEP autoreset */
    *board = b; *reward = r; *term = (uint8_t)t;
    return 0;
}

int orc_ttt_step(uint32_t* boards, const int32_t* actions, const uint32_t* U, int K, int N, int32_t* next_states,
                 uint32_t* next_masks, float* rewards, uint8_t* term) {
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(|:bad)
    for (int i = 0; i < N; ++i) {
        const uint32_t* u = U + (size_t)i * K;
        if (ttt_step_one(&boards[i], actions[i], u[2], u[3], u[4], &rewards[i], &term[i])) { bad |= 1; continue; }
        next_states[i] = ttt_state(boards[i] & 0x3FFFF);
        next_masks[i] = empties(boards[i]);
    }
    return bad ? -1 : 0;
}

/* ---------------------------------------------------------------- hash MDP */
static inline uint32_t mdp_mask(uint32_t s, int A, uint32_t env_seed) {
    uint32_t full = A >= 32 ? 0xFFFFFFFFu : ((1u << A) - 1u);
    return (mix32(s + env_seed * SEEDMIX, 2) & full) | 1u;
}
static inline void mdp_step_one(int32_t* state, int action, int S, int A, uint32_t env_seed, uint64_t term_thresh,
                                uint32_t bterm, uint32_t breset, float* reward, uint8_t* term) {
    uint32_t h = mix32((uint32_t)*state * (uint32_t)A + (uint32_t)action + env_seed * SEEDMIX, 0);
    uint32_t h2 = mix32(h, 1);
    *reward = (float)(h2 >> 8) * 5.9604644775390625e-08f * 2.0f - 1.0f;
    int t = (uint64_t)bterm < term_thresh;
    *state = (int32_t)(t ? pick(breset, (uint32_t)S) : pick(h, (uint32_t)S));
    *term = (uint8_t)t;
}

void orc_mdp_masks(const int32_t* states, int A, uint32_t env_seed, int N, uint32_t* masks) {
    for (int i = 0; i < N; ++i) masks[i] = mdp_mask((uint32_t)states[i], A, env_seed);
}
void orc_mdp_reset(int32_t* states, int S, const uint32_t* U, int K, int N) {
    for (int i = 0; i < N; ++i) states[i] = (int32_t)pick(U[(size_t)i * K + 3], (uint32_t)S);
}

/* ---------------------------------------------------------------- fused training loops */
/* One struct of optional per-step outputs (NULL = not recorded). Layout [steps][N]. */
typedef struct {
    int32_t* actions; float* rewards; uint8_t* term; int32_t* next_states;
    float* episode_returns;   /* NaN where no episode finished; else the finished episode's return */
} orc_trace_t;

/*
 * env_kind: 0 = hash MDP (env_state = int32 states), 1 = TicTacToe (env_state = uint32 boards).
 * U: pre-drawn [steps][N][K] or NULL -> counter stream (stream_seed, t0 + t, agent0 + i, k).
 * eps_thresh[steps]: explore thresholds (ceil(eps_t * 2^32)); lr[steps]: float32(lr_t).
 * Returns 0, or -(t+1) if step t hit an illegal move / empty candidate set.
 */
int orc_run(int env_kind, float* q, int S, int A, int N, void* env_state, int32_t* states, uint32_t* masks,
            uint32_t env_seed, uint64_t term_thresh, const uint32_t* U, int K, uint32_t stream_seed, uint32_t t0,
            uint32_t agent0, int steps, const uint64_t* eps_thresh, const float* lr, float gamma, int empty_all,
            float* agent_rewards, double* ep_sum, int64_t* ep_count, orc_trace_t* trace) {
    int32_t* actions = (int32_t*)malloc(sizeof(int32_t) * N);
    int32_t* s2 = (int32_t*)malloc(sizeof(int32_t) * N);
    uint32_t* m2 = (uint32_t*)malloc(sizeof(uint32_t) * N);
    float* r = (float*)malloc(sizeof(float) * N);
    uint8_t* term = (uint8_t*)malloc(N);
    int rc = 0;
    for (int t = 0; t < steps && rc == 0; ++t) {
        usrc_t us = {U ? U + (size_t)t * N * K : NULL, K, stream_seed, t0 + (uint32_t)t, agent0};
        int bad = 0;
#pragma omp parallel for schedule(static) reduction(|:bad)
        for (int i = 0; i < N; ++i) {
            int explore = (uint64_t)udraw(&us, i, 0) < eps_thresh[t];
            int a = select_one(q + (size_t)states[i] * A, A, masks[i], 1, explore, empty_all, udraw(&us, i, 1));
            actions[i] = a;
            if (a < 0) { bad |= 1; continue; }
            if (env_kind == 0) {
                int32_t st = states[i];
                mdp_step_one(&st, a, S, A, env_seed, term_thresh, udraw(&us, i, 2), udraw(&us, i, 3), &r[i], &term[i]);
                s2[i] = st; m2[i] = mdp_mask((uint32_t)st, A, env_seed);
            } else {
                uint32_t* boards = (uint32_t*)env_state;
                if (ttt_step_one(&boards[i], a, udraw(&us, i, 2), udraw(&us, i, 3), udraw(&us, i, 4), &r[i], &term[i])) {
                    bad |= 1; continue;
                }
                s2[i] = ttt_state(boards[i] & 0x3FFFF); m2[i] = empties(boards[i]);
            }
        }
        if (bad) { rc = -(t + 1); break; }
        for (int i = 0; i < N; ++i) agent_rewards[i] += r[i];                       /* BRT:212 */
        orc_learn_seq(q, A, states, actions, r, s2, term, m2, lr[t], gamma, N);     /* BRT:214 */
        for (int i = 0; i < N; ++i) {                                               /* BRT:218-221 */
            float er = NAN;
            if (term[i]) { er = agent_rewards[i]; *ep_sum += (double)er; *ep_count += 1; agent_rewards[i] = 0.0f; }
            if (trace && trace->episode_returns) trace->episode_returns[(size_t)t * N + i] = er;
        }
        if (trace) {
            if (trace->actions) memcpy(trace->actions + (size_t)t * N, actions, sizeof(int32_t) * N);
            if (trace->rewards) memcpy(trace->rewards + (size_t)t * N, r, sizeof(float) * N);
            if (trace->term) memcpy(trace->term + (size_t)t * N, term, N);
            if (trace->next_states) memcpy(trace->next_states + (size_t)t * N, s2, sizeof(int32_t) * N);
        }
        memcpy(states, s2, sizeof(int32_t) * N);
        memcpy(masks, m2, sizeof(uint32_t) * N);
    }
    free(actions); free(s2); free(m2); free(r); free(term);
    return rc;
}
