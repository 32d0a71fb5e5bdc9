/*
 * qe_engine.h -- C ABI of the B200-native tabular Q-learning engine (libqe_b200.so).
 *
 * Drop-in boundary for ONE hot path of j-moralejo-pinas/dist_classicrl: masked epsilon-greedy
 * select -> vector env step -> sequential TD update.  Each entry point names the reference
 * interface it replaces (paths relative to /root/reference/src/dist_classicrl):
 *
 *   QLO = algorithms/base_algorithms/q_learning_optimal.py      BRT = algorithms/runtime/base_runtime.py
 *   STR = algorithms/runtime/single_thread_runtime.py           TTT = environments/tiktaktoe_mod.py
 *   FLT = wrappers/flatten_multidiscrete_wrapper.py             UTL = utils.py
 *
 * Conventions: plain pointers and sizes only (no torch types).  Pointers are DEVICE pointers unless the
 * function name ends in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 * are asynchronous on it unless stated.  Every function returns 0 on success or a negative QE_ERR_* code;
 * qe_last_error() returns a thread-local message.  The engine is not re-entrant per handle: calls on one
 * handle are serialised by an internal mutex (the reference's MPI trainer calls choose_actions and learn
 * from two host threads, q_learning_async_dist.py:434-440).
 *
 * Random numbers: every stochastic decision consumes one uint32 U[t][i][k] (vector step t, agent i, slot k;
 * slot 0 explore test, 1 pick, 2.. environment; see oracle/rng.py).  `uniforms` is either a pre-drawn device
 * array for this call or NULL, in which case U is the counter hash qe_stream_u32(seed, t, agent0 + i, k).
 */
#ifndef QE_ENGINE_H
#define QE_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QE_OK 0
#define QE_ERR_ARG (-1)          /* bad argument (reference: AssertionError / ValueError) */
#define QE_ERR_CUDA (-2)         /* CUDA runtime error, see qe_last_error() */
#define QE_ERR_INVALID_MOVE (-3) /* env got an illegal action (reference: AssertionError "Invalid move.", TTT:130) */
#define QE_ERR_EMPTY (-4)        /* empty candidate / bootstrap set (reference: IndexError / np.max of empty, QLO:470,764) */
#define QE_ERR_TIMEOUT (-5)      /* dependency resolution did not converge (internal) */

#define QE_ENV_MDP 0    /* synthetic integer-hash MDP */
#define QE_ENV_TTT 1    /* TicTacToeEnv + FlattenMultiDiscreteObservationsWrapper + SyncVectorEnv(SAME_STEP) */
#define QE_ENV_BANDIT 2 /* RiggedTwoArmedBanditEnv in DummyVecWrapper (rigged_two_armed_bandit.py:55-80) */

#define QE_LEARN_SEQUENTIAL 0 /* learn / learn_iter: exact per-agent order (QLO:770-817, 893-934) */
#define QE_LEARN_ACCUMULATE 1 /* learn_vec / _learn_vec: snapshot bootstrap + np.add.at (QLO:853-891) */

#define QE_ABI_VERSION 2 /* bumped whenever a struct of this header or an entry point's signature changes */
typedef struct qe_engine qe_engine_t;
int qe_abi_version(void); /* the QE_ABI_VERSION the loaded library was built with (bindings compare it at load) */

/* ---- lifetime / table: replaces OptimalQLearningBase.__init__ and the q_table attribute (QLO:84-98, :80) ---- */
int qe_create(int64_t num_states, int32_t num_actions, float discount_factor, int32_t device, qe_engine_t** out);
int qe_destroy(qe_engine_t* e);
const char* qe_last_error(void);
int qe_set_last_error(int code, const char* msg); /* internal: sets the thread-local message, returns code */
int qe_set_discount(qe_engine_t* e, float discount_factor);
/* fp32 table in HBM, row-major with a padded row stride (in floats); rows are 16-byte aligned */
float* qe_table_ptr(qe_engine_t* e);
int32_t qe_table_stride(qe_engine_t* e);
int qe_table_upload_host(qe_engine_t* e, const float* dense_host /* [S][A] */);   /* q_table setter; synchronous */
int qe_table_download_host(qe_engine_t* e, float* dense_host /* [S][A] */);       /* q_table getter / save(); synchronous */
int qe_table_fill(qe_engine_t* e, float value, void* stream);
/* throughput runs: table[s][a] = (stream hash >> 8) * 2^-24, uniform in [0,1) */
int qe_table_fill_random(qe_engine_t* e, uint32_t seed, void* stream);
int qe_sync(qe_engine_t* e, void* stream); /* cudaStreamSynchronize + raise deferred device errors */
/* ---- flatten wrappers: batched mixed-radix encode / decode of MultiDiscrete vectors ------------------------------
 * Replaces utils.encode_multi_discretes (utils.py:51-69: out[i] = sum_d vectors[i][d] * radix[d]) and
 * utils.decode_to_multi_discretes (utils.py:95-115: out[i][d] = (indices[i] // radix[d]) % nvec[d]) behind
 * FlattenMultiDiscreteObservationsWrapper.observation / FlattenMultiDiscreteActionsWrapper.action
 * (wrappers/flatten_multidiscrete_wrapper.py:139-161, 61-76).  vectors / indices / out are DEVICE arrays
 * ([n][dims] int32 row-major, [n] int64); radix_host = compute_radix(nvec) and nvec_host are HOST arrays of `dims`
 * (<= 32) entries, passed to the kernel by value.  Asynchronous on `stream`. */
int qe_radix_encode(const int32_t* vectors, const int64_t* radix_host, int32_t dims, int64_t* out, int64_t n, void* stream);
int qe_radix_decode(const int64_t* indices, const int64_t* nvec_host, const int64_t* radix_host, int32_t dims, int32_t* out, int64_t n,
                    void* stream);
/* Page-lock a caller-owned host array in place (the reference's state dictionaries hold plain NumPy arrays the trainer
 * updates in place, STR:57, 70-75; page-locked, their per-step copies are asynchronous DMA).  Returns 1 if the range
 * was registered by this call (pair with qe_host_unregister), 0 if it already was page-locked, < 0 on error. */
int qe_host_register(void* host, uint64_t bytes);
int qe_host_unregister(void* host);

/* ---- select: replaces choose_actions and its 8 variants (QLO:263-726) ------------------------------------
 * mask_bits[N]: bit a set = action a legal (A <= 32), NULL = no masks.  mask_bytes[N][A] (uint8 truthy) is the
 * general form, required when A > 32.  explore_threshold = ceil(eps * 2^32) clamped to [0, 2^32].
 * empty_all: 0 -> all-zero mask yields -1 (QLO:348); 1 -> every action ties (QLO:467-470).  actions_out int32[N]. */
int qe_select(qe_engine_t* e, const int32_t* states, const uint32_t* mask_bits, const uint8_t* mask_bytes,
              const uint32_t* uniforms, int32_t slots, uint32_t stream_seed, uint32_t t, uint32_t agent0,
              uint64_t explore_threshold, int32_t deterministic, int32_t empty_all, int32_t* actions_out, int32_t n,
              void* stream);
/* same with HOST buffers (what BaseRuntime._choose_actions hands over, BRT:265-291); synchronous */
int qe_select_host(qe_engine_t* e, const int32_t* states, const uint32_t* mask_bits, const uint8_t* mask_bytes,
                   const uint32_t* uniforms, int32_t slots, uint32_t stream_seed, uint32_t t, uint32_t agent0,
                   uint64_t explore_threshold, int32_t deterministic, int32_t empty_all, int32_t* actions_out, int32_t n);

/* ---- learn: replaces learn / learn_iter / single_learn and learn_vec (QLO:728-934) -------------------------
 * terminated: uint8[N]; next_mask_bits / next_mask_bytes as in qe_select (NULL, NULL = unmasked max). */
int qe_learn(qe_engine_t* e, const int32_t* states, const int32_t* actions, const float* rewards,
             const int32_t* next_states, const uint8_t* terminated, const uint32_t* next_mask_bits,
             const uint8_t* next_mask_bytes, float lr, int32_t n, int32_t mode, void* stream);
int qe_learn_host(qe_engine_t* e, const int32_t* states, const int32_t* actions, const float* rewards,
                  const int32_t* next_states, const uint8_t* terminated, const uint32_t* next_mask_bits,
                  const uint8_t* next_mask_bytes, float lr, int32_t n, int32_t mode);

/* ---- accessors: get_q_values / add_q_values ... (QLO:100-250); device index arrays ---- */
int qe_gather(qe_engine_t* e, const int32_t* states, const int32_t* actions, float* out, int32_t n, void* stream);
/* get_states_q_values (QLO:154-171) with HOST buffers: out_host[n][A]; synchronous */
int qe_gather_rows_host(qe_engine_t* e, const int32_t* states_host, float* out_host, int32_t n);

/* ---- environments (device-resident state): replace env.reset / env.step of the bundled envs ---------------- */
/* TicTacToe boards: 2 bits per cell (cell c at bits 2c..2c+1), bit 18 = agent_mark - 1.  uniforms as above
 * (slots 2: machine move, 3: who starts, 4: machine opening). */
int qe_ttt_reset(uint32_t* boards, int32_t* states_out, uint32_t* mask_bits_out, const uint32_t* uniforms, int32_t slots,
                 uint32_t stream_seed, uint32_t t, uint32_t agent0, int32_t n, void* stream);
int qe_ttt_step(qe_engine_t* e, uint32_t* boards, const int32_t* actions, const uint32_t* uniforms, int32_t slots,
                uint32_t stream_seed, uint32_t t, uint32_t agent0, int32_t* next_states, uint32_t* next_mask_bits,
                float* rewards, uint8_t* terminated, int32_t n, void* stream);
/* hash MDP: slots 2: termination draw, 3: (re)start state */
int qe_mdp_reset(int32_t* states, uint32_t* mask_bits_out, int64_t num_states, int32_t num_actions, uint32_t env_seed,
                 const uint32_t* uniforms, int32_t slots, uint32_t stream_seed, uint32_t t, uint32_t agent0, int32_t n,
                 void* stream);
/* legal-action bits of the given states (a pure function of the state) */
int qe_mdp_masks(const int32_t* states, uint32_t* mask_bits_out, int32_t num_actions, uint32_t env_seed, int32_t n, void* stream);
int qe_mdp_step(qe_engine_t* e, int32_t* states, const int32_t* actions, int64_t num_states, int32_t num_actions,
                uint32_t env_seed, uint64_t term_threshold, const uint32_t* uniforms, int32_t slots, uint32_t stream_seed,
                uint32_t t, uint32_t agent0, uint32_t* next_mask_bits, float* rewards, uint8_t* terminated, int32_t n,
                void* stream);

/* ---- fused loop: replaces the body of SingleThreadQLearning.run_steps (STR:63-64) =
 *      BaseRuntime.run_single_step (BRT:184-222) x steps, one persistent cooperative kernel ------------------- */
typedef struct {
    int32_t env_kind;          /* QE_ENV_* */
    int32_t num_agents;        /* N <= 2^24 - 1 */
    int32_t* states;           /* [N] current observation; updated in place */
    int32_t* states_scratch;   /* [N] */
    uint32_t* env_words;       /* [N] TicTacToe board / bandit step counter; NULL for the MDP */
    float* episode_returns;    /* [N] agent_rewards accumulator (STR:57, BRT:212); updated in place */
    uint32_t env_seed;         /* MDP */
    uint32_t episode_len;      /* bandit */
    uint64_t term_threshold;   /* MDP: ceil(p_term * 2^32) */
} qe_agents_t;

typedef struct {
    int32_t steps;                    /* K vector steps in this launch */
    const uint64_t* explore_thresholds_host; /* [K] ceil(eps_t * 2^32)   (schedules evaluated by the caller, BRT:262-263) */
    const float* learning_rates_host;        /* [K] float32(lr_t) */
    const uint32_t* uniforms;         /* device [K][N][slots] or NULL -> counter stream */
    int32_t slots;
    uint32_t stream_seed, t0, agent0; /* select stream (slots 0,1): U[t0 + k][agent0 + i] */
    uint32_t env_stream_seed, env_t0; /* environment stream (slots >= 2) */
    int32_t empty_all;
    int32_t use_masks;                /* 0: every action legal (plain-array observations, BRT:255-259) */
    /* optional per-step traces, device [K][N]; NULL = not recorded */
    int32_t* trace_actions;
    float* trace_rewards;
    uint8_t* trace_terminated;
    int32_t* trace_next_states;
    float* trace_episode_returns;     /* NaN where no episode finished, else the finished episode's return (BRT:218-221) */
    /* optional statistics (device scalars, accumulated): sum / count of finished-episode returns */
    double* episode_sum;
    unsigned long long* episode_count;
    /* 1: evaluation loop (BaseRuntime.evaluate_steps / evaluate_episodes, BRT:293-384): select + env step only, the
     * table is not updated (pass explore thresholds of 0 for the reference's deterministic=True, exploration_rate=0) */
    int32_t evaluate;
    /* TD update of the loop: QE_LEARN_SEQUENTIAL (0; learn -> learn_iter, QLO:893-934, what the reference's trainers
     * call) or QE_LEARN_ACCUMULATE (1; learn_vec, QLO:819-891: snapshot bootstrap + accumulating scatter with plain
     * atomics -- equal to the reference up to the fp32 rounding of the order in which increments of one cell are summed) */
    int32_t learn_mode;
} qe_run_t;

int qe_fused_steps(qe_engine_t* e, const qe_agents_t* agents, const qe_run_t* run, void* stream);

/* ---- multi-GPU support (new relative to the reference, whose MPI trainer keeps ONE table on rank 0,
 *      q_learning_async_dist.py:59-281) ------------------------------------------------------------------------
 * Sharded table: a handle created with num_states = S_local owns the global states [first_state, first_state +
 * S_local); every state id passed to qe_select / qe_learn / qe_gather / qe_serve_bootstrap is then a GLOBAL id. */
int qe_set_state_base(qe_engine_t* e, int64_t first_state);
/* global ids of the local agents for the counter stream U[t][id][k] (device uint32[N], NULL = agent0 + i); used by
 * qe_select and qe_mdp_step */
int qe_set_agent_ids(qe_engine_t* e, const uint32_t* ids);
/* hold = 1: qe_learn (sequential mode) publishes every agent's new value but leaves the table untouched, so the
 * update can be repeated with better bootstrap values of remote states; qe_learn_commit applies the last one. */
int qe_set_hold(qe_engine_t* e, int32_t hold);
int qe_learn_commit(qe_engine_t* e, const int32_t* states, const int32_t* actions, int32_t n, void* stream);
/* bootstrap requests of agents living on another shard: out[r] = max over mask_bits[r] of the value of
 * (rows[r], a') just before an agent that sorts after exactly the first before[r] local agents of the last held
 * qe_learn (use_versions = 0: plain table max) */
int qe_serve_bootstrap(qe_engine_t* e, const int32_t* rows, const int32_t* before, const uint32_t* mask_bits, float* out,
                       int32_t n, int32_t use_versions, void* stream);
/* Replicated table: dense [S][A] device buffers; delta = Q - base, then Q = base = base + sum_of_deltas */
int qe_table_export_dense(qe_engine_t* e, float* dense, void* stream);
int qe_table_import_dense(qe_engine_t* e, const float* dense, void* stream);
int qe_table_delta_dense(qe_engine_t* e, const float* base, float* delta_out, void* stream);
int qe_table_merge_dense(qe_engine_t* e, float* base_inout, const float* delta_sum, void* stream);

/* ---- sharded table over peer memory (BASELINE config 4; round 2) -------------------------------------------------
 * Replaces what the reference's MPI trainer does with one table on rank 0 and transitions shipped to it
 * (q_learning_async_dist.py:59-281) by a state-range-sharded table: rank g owns the states [g * rows, (g + 1) * rows),
 * rows = ceil(S / G); agent i lives on rank i / ceil(N / G) for good.  Every rank runs ONE persistent kernel; rows,
 * writer records and targets travel as peer loads / stores over NVLink, phases are separated by flag barriers in peer
 * memory (no NCCL call on the data path; see csrc/qe_shard.cuh).  Hash MDP only.  Results (table, states, returns) are
 * identical to the single-GPU engine and to the reference's sequential loop in global agent order.
 * One process per GPU: create, exchange qe_shard_ipc_handle() with the peers (64 bytes), qe_shard_connect_ipc() each,
 * then qe_shard_steps(&mine, 1, ...) collectively.  All ranks in one process on one GPU (tests): qe_shard_connect_local()
 * and ONE call qe_shard_steps(all, G, ...). */
typedef struct qe_shard qe_shard_t;
/* external_slab: NULL (the library allocates the shared slab with cudaMalloc; peers map it with CUDA IPC) or
 * qe_shard_slab_bytes() of device memory the caller has made visible to the peers itself -- e.g. torch symmetric memory
 * (CUDA VMM, 2 MB pages): random peer accesses through a CUDA IPC mapping of a cudaMalloc block were measured at
 * 0.23 G/s on B200 NVLink against 6.8 G/s through a large-page mapping (profiles/r2_p2p_microbench.md). */
int64_t qe_shard_slab_bytes(int64_t num_states, int32_t num_actions, int32_t world, int32_t num_agents);
int qe_shard_create(int64_t num_states, int32_t num_actions, float discount_factor, int32_t device, int32_t rank, int32_t world,
                    int32_t num_agents, uint32_t env_seed, void* external_slab, qe_shard_t** out);
int qe_shard_connect_ptr(qe_shard_t* s, int32_t peer_rank, void* peer_slab);
int qe_shard_destroy(qe_shard_t* s);
int qe_shard_ipc_handle(qe_shard_t* s, void* out64);
int qe_shard_connect_ipc(qe_shard_t* s, int32_t peer_rank, const void* handle64);
int qe_shard_connect_local(qe_shard_t* s, int32_t peer_rank, qe_shard_t* peer);
int qe_shard_fill_random(qe_shard_t* s, uint32_t seed, void* stream);            /* this rank's slice of qe_table_fill_random */
int qe_shard_reset(qe_shard_t* s, uint32_t stream_seed, uint32_t t_init, void* stream); /* initial states of its agents (qe_mdp_reset by global id) */
int qe_shard_steps(qe_shard_t* const* ranks, int32_t nlocal, int32_t steps, const uint64_t* explore_thresholds_host,
                   const float* learning_rates_host, uint32_t stream_seed, uint32_t env_stream_seed, int32_t empty_all, int32_t use_masks,
                   uint64_t term_threshold, void* stream);
int qe_shard_sync(qe_shard_t* s, void* stream);
int qe_shard_download(qe_shard_t* s, float* table_host, int32_t* states_host, float* returns_host, double* episode_sum, uint64_t* episode_count);
int qe_shard_rows_host(qe_shard_t* s, const int64_t* states_host, float* out_host, int32_t n);
int qe_shard_phase_ns(qe_shard_t* s, uint64_t* out_host128); /* phase clock of the last launch, see csrc/qe_shard.cu */
double qe_shard_probe(qe_shard_t* s, int32_t peer_rank, int32_t mode); /* development aid: G random 32-byte loads (0) / 8-byte stores (1) per second over a rank's shard as mapped here */
int32_t qe_shard_info(qe_shard_t* s, int32_t what); /* 0: rows per shard, 1: agents per rank (ceil), 2: agents of this rank, 3: kernels launched */

/* introspection for benchmarks / tests */
uint32_t qe_stream_u32(uint32_t seed, uint32_t t, uint32_t i, uint32_t k);
int64_t qe_kernel_launches(qe_engine_t* e);     /* kernels launched by this handle so far */
int32_t qe_fused_grid_blocks(qe_engine_t* e);   /* grid of the last fused launch */
/* form of the TD update the last fused launch used: 0 = writer lists, 1 = per-step sort, 3 = target pipeline, 4 = one-CTA
 * loop for batches of at most 256 agents (csrc/qe_small.cuh; picked automatically under forms 3 and 5), 5 = one-pass
 * pipeline (csrc/qe_flow.cuh).  All exact. */
int32_t qe_fused_form(qe_engine_t* e);
/* Which exact form of the TD update the fused loop uses: 0 = writer lists, 1 = per-step sort, 2 = keep timing those two
 * and use the faster one, 3 = target pipeline (csrc/qe_pipe.cuh), 5 = one-pass pipeline (csrc/qe_flow.cuh; default;
 * QE_FORM in the environment sets the initial value).  All forms give identical results. */
int qe_set_fused_form(qe_engine_t* e, int32_t form);
/* phase clock of the last fused launch (synchronous): out_host[0] = %globaltimer (ns) at kernel start, then for each
 * of the first 10 vector steps the time after phase A (select + env step + writer registration), after phase B1 (TD
 * update, first pass) and after phase B2 (TD update, deferred agents).  One-pass form (5): after the in-order pass, after
 * the commit, after the bucket sorts; out_host[32 + k] = after the scatter of step k.  Returns the number of values written. */
int32_t qe_fused_phase_ns(qe_engine_t* e, uint64_t* out_host, int32_t cap);
/* GB/s of dependency-free random whole-row gathers over this table (the ceiling of the engine's dominant access pattern;
 * bench.py: roofline.gather_peak) */
double qe_debug_gather_gbs(qe_engine_t* e);
/* development aid: microseconds per grid-wide barrier at the fused loop's launch shape (cooperative launch, 4 CTAs/SM) */
double qe_debug_gridsync_us(qe_engine_t* e, int32_t iters);
/* development aid: counters of the sorted loop's phase Q accumulated since the last reset (see qe_sorted.cuh) */
int qe_debug_counters(qe_engine_t* e, uint64_t* out8_host, int32_t reset);
const char* qe_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* QE_ENGINE_H */
