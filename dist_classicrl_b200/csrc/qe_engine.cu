// qe_engine.cu -- C ABI (include/qe_engine.h) over the kernels in qe_kernels.cuh.  Build: see build.py
// (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cstdlib>
#include <mutex>
#include <new>

#include "../../include/qe_engine.h"
#include "qe_kernels.cuh"
#include "qe_sorted.cuh"
#include "qe_pipe.cuh"
#include "qe_flow.cuh"
#include "qe_small.cuh"
#include "qe_radix.cuh"

using namespace qe;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" int qe_set_last_error(int code, const char* msg) {  // for the library's other translation units
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess) return fail(QE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// the engine-less entry points launch on the device that owns their buffers (a process that drives several GPUs may
// have another device current)
static int use_device_of(const void* p) {
    cudaPointerAttributes at;
    if (p && cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice) CK(cudaSetDevice(at.device));
    else (void)cudaGetLastError();
    return QE_OK;
}

struct qe_engine {
    int64_t S = 0;
    int A = 0, ld = 0, lpa = 0, lpr = 0, device = 0, sms = 0;  // lpa: 16-byte lanes per row (select_kernel); lpr: 32-byte sectors per row
    float gamma = 0.0f;
    Table T{};               // T.q is biased by -state_base rows: kernels index it with GLOBAL state ids
    float* q_real = nullptr; // the allocation (row 0 = state `state_base`)
    uint32_t* info_real = nullptr;  // writer info of the list form (ensure_info)
    int64_t state_base = 0;
    const uint32_t* agent_ids = nullptr;  // optional global ids of the local agents (uniform stream), device
    int cap = 0;             // per-agent scratch capacity
    uint8_t* tr_a = nullptr;
    float* tr_r = nullptr;
    float* delta = nullptr;
    // staging for *_host entry points and per-step schedules
    void* stage = nullptr;
    size_t stage_bytes = 0;
    int* tile_counter = nullptr;  // [2]
    uint64_t* phase_ns = nullptr; // [31+] fused-loop phase clock (see qe_fused_phase_ns)
    int phase_steps = 0;
    uint64_t* d_thresh = nullptr;
    float* d_lr = nullptr;
    int sched_cap = 0;
    uint32_t step = 0;       // global step counter (epoch / tag source)
    SortedScratch X{};       // scratch of the sorted fused loop (qe_sorted.cuh)
    PipeScratch P{};         // scratch of the pipelined fused loop (qe_pipe.cuh)
    int pipe_sorters = 0;    // P.ghist was sized for this many blocks
    bool pipe_valid = false; // P.pos / P.seg / P.kv hold the sort of pipe_states[0 .. pipe_n) as the last pipelined launch left them
    const int32_t* pipe_states = nullptr;
    int pipe_n = 0;
    int pipe_sorted_n = 0;   // agents of the order P.kv / P.seg still describe (0: none; seg[] is all-empty then)
    int pipe_kind = 0;       // the form that left that order (3: target pipeline, 5: one-pass form; their records and buffers differ)
    int sorted_grid = 0;     // ghist was sized for this many blocks
    // The fused loop has two exact forms of the TD update: writer lists (qe_kernels.cuh; best while few agents share
    // a row) and the per-step sort (qe_sorted.cuh; best once agents herd).  Both give identical results, so the engine
    // times its launches and keeps using the faster form, trying the other one every kProbeEvery launches.
    int strategy = 5;        // 0 = writer lists, 1 = per-step sort, 2 = time both and keep the faster, 3 = target pipeline, 5 (default) = one-pass pipeline (QE_FORM / QE_SORTED env)
    int current = 0, since_probe = 0, timed_kind = -1, auto_launches = 0, auto_form = 0;  // state of pick_form (strategy 2)
    double timed_work = 0.0, rate[2] = {0.0, 0.0};  // agent-steps per millisecond of the last timed launch of each form
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    int last_grid = 0;
    std::mutex mu;
};

static Table local_table(const qe_engine* e) {  // rows indexed from 0 (whole-table kernels)
    Table T = e->T;
    T.q = e->q_real;
    return T;
}
static int lanes_per_agent(int A) { return A <= 4 ? 1 : (A <= 8 ? 2 : (A <= 16 ? 4 : 8)); }
static int sectors_per_row(int A) { return A <= 8 ? 1 : (A <= 16 ? 2 : 4); }

#define QE_FREE(p) do { cudaFree(p); (p) = nullptr; } while (0)
static int ensure_agents(qe_engine* e, int n) {
    if (n <= e->cap) return QE_OK;
    int cap = 1024;
    while (cap < n) cap <<= 1;
    CK(cudaDeviceSynchronize());
    // everything is released and forgotten first: if an allocation below fails, no pointer dangles, the capacity is 0 and
    // the next call starts over (qe_destroy frees whatever did get allocated)
    e->cap = 0;
    QE_FREE(e->T.node); QE_FREE(e->T.slot); QE_FREE(e->T.tr_p); QE_FREE(e->tr_a); QE_FREE(e->tr_r); QE_FREE(e->delta);
    QE_FREE(e->T.later_buf); QE_FREE(e->T.rec); QE_FREE(e->T.dmask); QE_FREE(e->T.smask);
    CK(cudaMalloc(&e->T.later_buf, cap));
    CK(cudaMalloc(&e->T.smask, sizeof(uint32_t) * (cap / 32)));
    CK(cudaMalloc(&e->T.rec, sizeof(uint32_t) * 8 * (size_t)cap));
    CK(cudaMalloc(&e->T.dmask, sizeof(uint32_t) * (cap / 32)));
    CK(cudaMalloc(&e->T.node, sizeof(uint32_t) * cap));
    CK(cudaMalloc(&e->T.slot, sizeof(uint64_t) * cap));
    CK(cudaMalloc(&e->T.tr_p, sizeof(float) * cap));
    CK(cudaMalloc(&e->tr_a, cap));
    CK(cudaMalloc(&e->tr_r, sizeof(float) * cap));
    CK(cudaMalloc(&e->delta, sizeof(float) * cap));
    CK(cudaMemset(e->T.slot, 0, sizeof(uint64_t) * cap));
    {
        SortedScratch& X = e->X;
        for (int b = 0; b < 2; ++b) { QE_FREE(X.key[b]); QE_FREE(X.val[b]); }
        QE_FREE(X.rank); QE_FREE(X.targ); QE_FREE(X.mhist); QE_FREE(X.rrec); QE_FREE(X.rmask); QE_FREE(X.hmask); QE_FREE(X.hrec);
        CK(cudaMalloc(&X.hrec, sizeof(uint32_t) * 4 * (size_t)cap));
        for (int b = 0; b < 2; ++b) {
            CK(cudaMalloc(&X.key[b], sizeof(int32_t) * cap));
            CK(cudaMalloc(&X.val[b], sizeof(int32_t) * cap));
        }
        CK(cudaMalloc(&X.rank, sizeof(int32_t) * cap));
        CK(cudaMalloc(&X.targ, sizeof(uint64_t) * cap));
        CK(cudaMalloc(&X.mhist, sizeof(uint64_t) * cap));
        CK(cudaMalloc(&X.rrec, sizeof(uint32_t) * 4 * (size_t)cap));
        CK(cudaMalloc(&X.rmask, sizeof(uint32_t) * (cap / 32)));
        CK(cudaMalloc(&X.hmask, sizeof(uint32_t) * (cap / 32)));
        CK(cudaMemset(X.targ, 0, sizeof(uint64_t) * cap));
        CK(cudaMemset(X.mhist, 0, sizeof(uint64_t) * cap));
    }
    {
        PipeScratch& P = e->P;
        for (int b = 0; b < 2; ++b) QE_FREE(P.kv[b]);
        QE_FREE(P.tw); QE_FREE(P.rec); QE_FREE(P.pos);
        e->pipe_valid = false;
        e->pipe_sorted_n = 0;  // the order kv[] described is gone with the buffers: seg[] starts from all-empty again
        if (P.seg) CK(cudaMemset(P.seg, 0, sizeof(uint2) * (size_t)e->S));
        CK(cudaMalloc(&P.tw, sizeof(uint4) * (size_t)cap));
        CK(cudaMalloc(&P.rec, sizeof(uint2) * ((size_t)cap + 8)));  // phase T reads whole 32-byte groups
        CK(cudaMalloc(&P.pos, sizeof(int32_t) * cap));
        CK(cudaMemset(P.rec, 0xFF, sizeof(uint2) * ((size_t)cap + 8)));
        for (int b = 0; b < 2; ++b) CK(cudaMalloc(&P.kv[b], sizeof(int2) * (size_t)cap));
    }
    e->cap = cap;
    return QE_OK;
}
// state-indexed scratch of the pipelined form (segment bounds, two parities), allocated at its first launch
static int ensure_pipe(qe_engine* e, int sorters) {
    PipeScratch& P = e->P;
    if (!P.seg) {
        CK(cudaMalloc(&P.seg, sizeof(uint2) * (size_t)e->S));
        CK(cudaMemset(P.seg, 0, sizeof(uint2) * (size_t)e->S));
        CK(cudaMalloc(&P.rowtot, sizeof(int) * kRadix));
        CK(cudaMalloc(&P.ctr, 64 * sizeof(unsigned int)));
        int bits = 1;
        while (bits < 31 && (1ll << bits) < e->S) ++bits;
        P.passes = (bits + kRadixBits - 1) / kRadixBits;
    }
    if (sorters > e->pipe_sorters) {
        CK(cudaDeviceSynchronize());
        cudaFree(P.ghist);
        P.ghist = nullptr;
        CK(cudaMalloc(&P.ghist, sizeof(int) * kRadix * (size_t)sorters));
        e->pipe_sorters = sorters;
    }
    return QE_OK;
}
// writer info of the list form: allocated (zeroed: epoch 0 never matches) the first time a writer-list kernel runs
static int ensure_info(qe_engine* e) {
    if (e->info_real || e->T.info_ld == 0) return QE_OK;
    const size_t bytes = sizeof(uint32_t) * (size_t)e->S * e->T.info_ld;
    CK(cudaMalloc(&e->info_real, bytes));
    CK(cudaMemset(e->info_real, 0, bytes));
    e->T.info = e->info_real - (size_t)e->state_base * (size_t)e->T.info_ld;
    return QE_OK;
}
static int ensure_stage(qe_engine* e, size_t bytes) {
    if (bytes <= e->stage_bytes) return QE_OK;
    CK(cudaDeviceSynchronize());
    cudaFree(e->stage);
    e->stage = nullptr;
    e->stage_bytes = 0;
    CK(cudaMalloc(&e->stage, bytes));
    e->stage_bytes = bytes;
    return QE_OK;
}
static int check_device_errors(qe_engine* e, cudaStream_t st) {
    int h = 0;
    CK(cudaMemcpyAsync(&h, e->T.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h) {
        CK(cudaMemsetAsync(e->T.err, 0, sizeof(int), st));
        // the pipelined form's persistent order may be half-written: start it over
        e->pipe_valid = false;
        e->pipe_sorted_n = 0;
        if (e->P.seg) CK(cudaMemsetAsync(e->P.seg, 0, sizeof(uint2) * (size_t)e->S, st));
        if (h & kErrInvalidMove) return fail(QE_ERR_INVALID_MOVE, "Invalid move.");
        if (h & kErrEmpty) return fail(QE_ERR_EMPTY, "empty candidate or bootstrap action set");
        return fail(QE_ERR_TIMEOUT, "TD-update dependency resolution timed out");
    }
    return QE_OK;
}
template <typename K>
static int coop_blocks(qe_engine* e, K kernel, long long work_threads, int* out) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0));
    if (per_sm < 1) return fail(QE_ERR_CUDA, "kernel cannot be made resident");
    long long want = (work_threads + 255) / 256;
    long long maxb = (long long)per_sm * e->sms;
    *out = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
    return QE_OK;
}

extern "C" {

const char* qe_last_error(void) { return g_err; }
const char* qe_build_info(void) { return "libqe_b200 sm_100a (" __DATE__ " " __TIME__ ")"; }
uint32_t qe_stream_u32(uint32_t seed, uint32_t t, uint32_t i, uint32_t k) { return stream_u32(seed, t, i, k); }

static int create_impl(qe_engine* e, int64_t num_states, int32_t num_actions, float discount_factor, int32_t device);
int qe_abi_version(void) { return QE_ABI_VERSION; }
int qe_create(int64_t num_states, int32_t num_actions, float discount_factor, int32_t device, qe_engine_t** out) {
    if (!out || num_states <= 0 || num_actions <= 0) return fail(QE_ERR_ARG, "state_size and action_size must be positive");
    if (num_states >= (1ll << 31)) return fail(QE_ERR_ARG, "state_size must be < 2^31");
    CK(cudaSetDevice(device));
    qe_engine* e = new (std::nothrow) qe_engine();
    if (!e) return fail(QE_ERR_ARG, "out of host memory");
    e->device = device;
    const int rc = create_impl(e, num_states, num_actions, discount_factor, device);
    if (rc) {  // whatever was allocated so far goes back (cudaFree(nullptr) is a no-op); the error message stays
        char keep[512];
        snprintf(keep, sizeof(keep), "%s", g_err);
        qe_destroy(e);
        snprintf(g_err, sizeof(g_err), "%s", keep);
        return rc;
    }
    *out = e;
    return QE_OK;
}
static int create_impl(qe_engine* e, int64_t num_states, int32_t num_actions, float discount_factor, int32_t device) {
    e->S = num_states;
    e->A = num_actions;
    e->gamma = discount_factor;
    e->device = device;
    e->lpa = lanes_per_agent(num_actions);
    e->lpr = sectors_per_row(num_actions);
    if (num_actions <= 32) {
        // dense Q rows of whole 32-byte sectors; the writer lists of the list form (three times the row size per state:
        // agents cluster on the deterministic environments, rows with dozens of writers are common) live in info[]
        e->ld = 8 * e->lpr;
        e->T.info_ld = 3 * e->ld;
        e->T.inline_cap = e->T.info_ld - 4 - 2;  // two words behind the inline entries hold the spill descriptor
    } else {  // generic path (sequential learn kernel): plain padded rows, no writer info
        e->ld = ((num_actions + 3) / 4) * 4;
        e->T.info_ld = 0;
        e->T.inline_cap = 0;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    e->sms = prop.multiProcessorCount;
    e->T.ld = e->ld;
    e->T.A = e->A;
    CK(cudaMalloc(&e->q_real, sizeof(float) * (size_t)e->S * e->ld));
    CK(cudaMemset(e->q_real, 0, sizeof(float) * (size_t)e->S * e->ld));
    e->T.q = e->q_real;
    CK(cudaMalloc(&e->T.err, sizeof(int)));
    CK(cudaMemset(e->T.err, 0, sizeof(int)));
    CK(cudaMalloc(&e->tile_counter, 8 * sizeof(int)));
    CK(cudaMemset(e->tile_counter, 0, 8 * sizeof(int)));
    if (num_actions <= 32) {
        CK(cudaMalloc(&e->X.seg, sizeof(uint32_t) * 4 * (size_t)e->S));
        CK(cudaMemset(e->X.seg, 0, sizeof(uint32_t) * 4 * (size_t)e->S));
        CK(cudaMalloc(&e->X.rowtot, sizeof(int) * kRadix));
        CK(cudaMalloc(&e->X.dbg, sizeof(unsigned long long) * 8));
        CK(cudaMemset(e->X.dbg, 0, sizeof(unsigned long long) * 8));
        int bits = 1;
        while (bits < 31 && (1ll << bits) < e->S) ++bits;
        e->X.passes = (bits + kRadixBits - 1) / kRadixBits;
    }
    e->strategy = getenv("QE_FORM") ? atoi(getenv("QE_FORM")) : (getenv("QE_SORTED") ? atoi(getenv("QE_SORTED")) : 5);
    if (e->strategy < 0 || e->strategy == 4 || e->strategy > 5) e->strategy = 5;
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
    e->T.spill_slots = 1024;
    CK(cudaMalloc(&e->T.spill, sizeof(uint32_t) * (size_t)e->T.spill_slots * kSpillCap));
    CK(cudaMalloc(&e->T.spill_next, 2 * sizeof(int)));
    CK(cudaMemset(e->T.spill_next, 0, 2 * sizeof(int)));
    CK(cudaMalloc(&e->phase_ns, 48 * sizeof(uint64_t)));
    CK(cudaMemset(e->phase_ns, 0, 48 * sizeof(uint64_t)));
    return ensure_agents(e, 1024);
}

int qe_destroy(qe_engine_t* e) {
    if (!e) return QE_OK;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (int b = 0; b < 2; ++b) { cudaFree(e->X.key[b]); cudaFree(e->X.val[b]); }
    cudaFree(e->X.rank); cudaFree(e->X.targ); cudaFree(e->X.mhist); cudaFree(e->X.rrec); cudaFree(e->X.rmask); cudaFree(e->X.hmask); cudaFree(e->X.hrec);
    cudaFree(e->X.seg); cudaFree(e->X.rowtot); cudaFree(e->X.dbg);
    for (int b = 0; b < 2; ++b) cudaFree(e->P.kv[b]);
    cudaFree(e->P.rec); cudaFree(e->P.pos); cudaFree(e->P.seg);
    cudaFree(e->P.ghist); cudaFree(e->P.rowtot); cudaFree(e->P.ctr); cudaFree(e->P.tw);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1); cudaFree(e->X.ghist);
    cudaFree(e->q_real); cudaFree(e->info_real); cudaFree(e->T.later_buf); cudaFree(e->T.spill); cudaFree(e->T.spill_next); cudaFree(e->T.err); cudaFree(e->tile_counter); cudaFree(e->phase_ns); cudaFree(e->T.rec); cudaFree(e->T.dmask); cudaFree(e->T.smask); cudaFree(e->T.node); cudaFree(e->T.slot); cudaFree(e->T.tr_p);
    cudaFree(e->tr_a); cudaFree(e->tr_r); cudaFree(e->delta); cudaFree(e->stage); cudaFree(e->d_thresh); cudaFree(e->d_lr);
    delete e;
    return QE_OK;
}

int qe_set_discount(qe_engine_t* e, float g) { e->gamma = g; return QE_OK; }
float* qe_table_ptr(qe_engine_t* e) { return e->q_real; }
int32_t qe_table_stride(qe_engine_t* e) { return e->ld; }
int64_t qe_kernel_launches(qe_engine_t* e) { return e->launches; }
int32_t qe_fused_grid_blocks(qe_engine_t* e) { return e->last_grid; }
int32_t qe_fused_form(qe_engine_t* e) { return e->current; }
int qe_set_fused_form(qe_engine_t* e, int32_t form) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (form < 0 || form == 4 || form > 5) return fail(QE_ERR_ARG, "form must be 0 (writer lists), 1 (per-step sort), 2 (pick between those two by measurement), 3 (target pipeline) or 5 (one-pass form)");
    e->strategy = form;
    return QE_OK;
}
int qe_debug_counters(qe_engine_t* e, uint64_t* out8_host, int32_t reset) {
    std::lock_guard<std::mutex> lk(e->mu);
    if ((e->current == 3 || e->current == 5) && e->P.ctr) {  // the pipelined forms: ctr[8..15] of the last launch
        unsigned int h[56];
        CK(cudaMemcpy(h, e->P.ctr + 8, sizeof(h), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 8; ++i) out8_host[i] = h[i];
        if (reset == 2) for (int i = 8; i < 56; ++i) out8_host[i] = h[i];  // development: sort stage laps (56 values)
        return QE_OK;
    }
    if (!e->X.dbg) return fail(QE_ERR_ARG, "no counters");
    CK(cudaMemcpy(out8_host, e->X.dbg, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(e->X.dbg, 0, 8 * sizeof(uint64_t)));
    return QE_OK;
}
double qe_debug_gridsync_us(qe_engine_t* e, int32_t iters) {
    std::lock_guard<std::mutex> lk(e->mu);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gridsync_probe_kernel, 256, 0) != cudaSuccess) return -1.0;
    per_sm = per_sm > 4 ? 4 : per_sm;
    int blocks = per_sm * e->sms;
    uint64_t* d_out = e->phase_ns;
    void* args[] = {&iters, &d_out};
    if (cudaLaunchCooperativeKernel((void*)gridsync_probe_kernel, dim3(blocks), dim3(256), args, 0, 0) != cudaSuccess) return -1.0;
    uint64_t ns = 0;
    if (cudaMemcpy(&ns, d_out, sizeof(ns), cudaMemcpyDeviceToHost) != cudaSuccess) return -1.0;
    return (double)ns / 1e3 / iters;
}
}  // extern "C"
// random row gather over the engine's own table, no dependencies: LPR lanes fetch one row (one 32-byte sector each),
// `inflight` rows per lane group in flight
template <int LPR>
static __global__ void __launch_bounds__(256) gather_probe_kernel(const float* __restrict__ q, uint32_t rows, int ld, int per_thread, uint32_t seed, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = tid / LPR, l = tid % LPR;
    float acc = 0.f;
    for (int it = 0; it < per_thread; it += 4) {
        F8 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_row8(q + (size_t)__umulhi(fmix32((grp * 7919u + (uint32_t)(it + u)) ^ seed), rows) * ld + 8 * l);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].v[0] + v[u].v[7];
    }
    if (acc == 12345.678f) out[0] = acc;
}
extern "C" {
/* What the memory system allows for this engine's dominant access pattern: GB/s of random whole-row gathers over THIS
 * table (no dependencies, four rows in flight per lane group).  bench.py reports it as roofline.gather_peak beside the
 * copy peak.  Synchronous; ~1 ms. */
double qe_debug_gather_gbs(qe_engine_t* e) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (cudaSetDevice(e->device) != cudaSuccess || e->A > 32) return -1.0;
    const int blocks = e->sms * 8, per = 64;
    const uint32_t rows = (uint32_t)std::min<int64_t>(e->S, 0x7FFFFFFF);
    float* out = (float*)e->phase_ns;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1.0;
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(a);
        switch (e->lpr) {
            case 1: gather_probe_kernel<1><<<blocks, 256>>>(e->q_real, rows, e->ld, per, 11u + r, out); break;
            case 2: gather_probe_kernel<2><<<blocks, 256>>>(e->q_real, rows, e->ld, per, 11u + r, out); break;
            default: gather_probe_kernel<4><<<blocks, 256>>>(e->q_real, rows, e->ld, per, 11u + r, out); break;
        }
        cudaEventRecord(b);
        if (cudaEventSynchronize(b) != cudaSuccess) { best = -1.0f; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms > 0.f && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    e->launches += 4;
    if (best <= 0.f) return -1.0;
    const double rows_read = (double)blocks * 256 / e->lpr * per;
    return rows_read * e->lpr * 32 / best / 1e6;
}
int32_t qe_fused_phase_ns(qe_engine_t* e, uint64_t* out_host, int32_t cap) {
    std::lock_guard<std::mutex> lk(e->mu);
    uint64_t h[48];
    if (cudaSetDevice(e->device) != cudaSuccess || cudaMemcpy(h, e->phase_ns, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(QE_ERR_CUDA, "cannot read the phase clock");
    const int m = 1 + 3 * e->phase_steps;
    for (int i = 0; i < m && i < cap; ++i) out_host[i] = h[i];
    if (cap >= 48) for (int i = 32; i < 48; ++i) out_host[i] = h[i];  // pipelined form: end of the sort inside phase "B1" of step i - 32
    return m < cap ? m : cap;
}

int qe_table_upload_host(qe_engine_t* e, const float* dense) {
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpy2D(e->q_real, sizeof(float) * e->ld, dense, sizeof(float) * e->A, sizeof(float) * e->A, (size_t)e->S, cudaMemcpyHostToDevice));
    return QE_OK;
}
int qe_table_download_host(qe_engine_t* e, float* dense) {
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpy2D(dense, sizeof(float) * e->A, e->q_real, sizeof(float) * e->ld, sizeof(float) * e->A, (size_t)e->S, cudaMemcpyDeviceToHost));
    return QE_OK;
}
int qe_table_fill(qe_engine_t* e, float value, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    table_fill_kernel<<<e->sms * 8, 256, 0, (cudaStream_t)stream>>>(local_table(e), e->S, value, 0u, 0, (uint64_t)e->state_base);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_table_fill_random(qe_engine_t* e, uint32_t seed, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    table_fill_kernel<<<e->sms * 8, 256, 0, (cudaStream_t)stream>>>(local_table(e), e->S, 0.0f, seed, 1, (uint64_t)e->state_base);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
// ---------------------------------------------------------------------------------------------- flatten wrappers
static int radix_spec(RadixSpec& R, const int64_t* nvec_host, const int64_t* radix_host, int dims) {
    if (dims <= 0 || dims > kRadixMaxDims) return fail(QE_ERR_ARG, "radix: 1 <= dims <= %d", kRadixMaxDims);
    if (radix_host == nullptr) return fail(QE_ERR_ARG, "radix: radix_host is NULL");
    R.dims = dims;
    for (int d = 0; d < kRadixMaxDims; ++d) {
        R.radix[d] = d < dims ? (long long)radix_host[d] : 1;
        R.nvec[d] = (d < dims && nvec_host) ? (long long)nvec_host[d] : 1;
    }
    return QE_OK;
}
static int radix_grid(int64_t n) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (n + 255) / 256;
    return (int)std::min<int64_t>(tiles, (int64_t)sms * 8);
}
int qe_radix_encode(const int32_t* vectors, const int64_t* radix_host, int32_t dims, int64_t* out, int64_t n, void* stream) {
    RadixSpec R;
    int rc = radix_spec(R, nullptr, radix_host, dims);
    if (rc) return rc;
    if (n <= 0) return QE_OK;
    if ((rc = use_device_of(vectors)) != QE_OK) return rc;
    radix_encode_kernel<<<radix_grid(n), 256, 256 * (dims | 1) * sizeof(int32_t), (cudaStream_t)stream>>>(vectors, R, (long long*)out, (long long)n);
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_radix_decode(const int64_t* indices, const int64_t* nvec_host, const int64_t* radix_host, int32_t dims, int32_t* out, int64_t n,
                    void* stream) {
    RadixSpec R;
    if (nvec_host == nullptr) return fail(QE_ERR_ARG, "radix: nvec_host is NULL");
    int rc = radix_spec(R, nvec_host, radix_host, dims);
    if (rc) return rc;
    for (int d = 0; d < dims; ++d)
        if (R.radix[d] == 0 || R.nvec[d] == 0) return fail(QE_ERR_ARG, "radix: zero radix / nvec entry");
    if (n <= 0) return QE_OK;
    if ((rc = use_device_of(indices)) != QE_OK) return rc;
    radix_decode_kernel<<<radix_grid(n), 256, 256 * (dims | 1) * sizeof(int32_t), (cudaStream_t)stream>>>((const long long*)indices, R, out, (long long)n);
    CK(cudaGetLastError());
    return QE_OK;
}

// Page-lock a caller-owned host array in place, so that the *_host copies and the runtimes' per-step transfers of
// it are asynchronous DMA.  1 = newly registered (pair with qe_host_unregister), 0 = was page-locked already.
int qe_host_register(void* host, uint64_t bytes) {
    if (host == nullptr || bytes == 0) return fail(QE_ERR_ARG, "qe_host_register: empty range");
    cudaError_t rc = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterDefault);
    if (rc == cudaSuccess) return 1;
    cudaGetLastError();  // not sticky: leave no stale error behind for the next launch check
    if (rc == cudaErrorHostMemoryAlreadyRegistered) return 0;
    return fail(QE_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(rc));
}
int qe_host_unregister(void* host) {
    cudaError_t rc = cudaHostUnregister(host);
    if (rc != cudaSuccess) {
        cudaGetLastError();
        return fail(QE_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(rc));
    }
    return QE_OK;
}
int qe_sync(qe_engine_t* e, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    return check_device_errors(e, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------- select
static int select_impl(qe_engine* e, const int32_t* states, const uint32_t* mask_bits, const uint8_t* mask_bytes,
                       const uint32_t* uniforms, int slots, uint32_t seed, uint32_t t, uint32_t agent0, uint64_t thresh,
                       int det, int empty_all, int32_t* out, int n, cudaStream_t st) {
    if (n <= 0) return QE_OK;
    if (uniforms && slots < 2) return fail(QE_ERR_ARG, "uniforms need at least 2 slots");
    Uniforms U{uniforms, slots, seed, t, agent0, seed, t};
    U.ids = e->agent_ids;
    if (e->A > 32 || mask_bytes) {
        select_generic_kernel<<<(n + 127) / 128, 128, 0, st>>>(e->T, states, mask_bytes, U, thresh, det, empty_all, out, n);
    } else {
        const long long threads = (long long)n * e->lpa;
        const int blocks = (int)((threads + 255) / 256 < e->sms * 16 ? (threads + 255) / 256 : e->sms * 16);
        switch (e->lpa) {
            case 1: select_kernel<1><<<blocks, 256, 0, st>>>(e->T, states, mask_bits, U, thresh, det, empty_all, out, n); break;
            case 2: select_kernel<2><<<blocks, 256, 0, st>>>(e->T, states, mask_bits, U, thresh, det, empty_all, out, n); break;
            case 4: select_kernel<4><<<blocks, 256, 0, st>>>(e->T, states, mask_bits, U, thresh, det, empty_all, out, n); break;
            default: select_kernel<8><<<blocks, 256, 0, st>>>(e->T, states, mask_bits, U, thresh, det, empty_all, out, n); break;
        }
    }
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}

int qe_select(qe_engine_t* e, const int32_t* states, const uint32_t* mask_bits, const uint8_t* mask_bytes,
              const uint32_t* uniforms, int32_t slots, uint32_t stream_seed, uint32_t t, uint32_t agent0,
              uint64_t explore_threshold, int32_t deterministic, int32_t empty_all, int32_t* actions_out, int32_t n,
              void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->device));
    return select_impl(e, states, mask_bits, mask_bytes, uniforms, slots, stream_seed, t, agent0, explore_threshold,
                       deterministic, empty_all, actions_out, n, (cudaStream_t)stream);
}

static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

int qe_select_host(qe_engine_t* e, const int32_t* states, const uint32_t* mask_bits, const uint8_t* mask_bytes,
                   const uint32_t* uniforms, int32_t slots, uint32_t stream_seed, uint32_t t, uint32_t agent0,
                   uint64_t explore_threshold, int32_t deterministic, int32_t empty_all, int32_t* actions_out, int32_t n) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    CK(cudaSetDevice(e->device));
    const size_t b_s = al(sizeof(int32_t) * n), b_m = mask_bits ? al(sizeof(uint32_t) * n) : 0,
                 b_mb = mask_bytes ? al((size_t)n * e->A) : 0, b_u = uniforms ? al(sizeof(uint32_t) * (size_t)n * slots) : 0,
                 b_o = al(sizeof(int32_t) * n);
    int rc = ensure_stage(e, b_s + b_m + b_mb + b_u + b_o);
    if (rc) return rc;
    char* p = (char*)e->stage;
    int32_t* d_s = (int32_t*)p; p += b_s;
    uint32_t* d_m = mask_bits ? (uint32_t*)p : nullptr; p += b_m;
    uint8_t* d_mb = mask_bytes ? (uint8_t*)p : nullptr; p += b_mb;
    uint32_t* d_u = uniforms ? (uint32_t*)p : nullptr; p += b_u;
    int32_t* d_o = (int32_t*)p;
    cudaStream_t st = 0;
    CK(cudaMemcpyAsync(d_s, states, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    if (d_m) CK(cudaMemcpyAsync(d_m, mask_bits, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    if (d_mb) CK(cudaMemcpyAsync(d_mb, mask_bytes, (size_t)n * e->A, cudaMemcpyHostToDevice, st));
    if (d_u) CK(cudaMemcpyAsync(d_u, uniforms, sizeof(uint32_t) * (size_t)n * slots, cudaMemcpyHostToDevice, st));
    rc = select_impl(e, d_s, d_m, d_mb, d_u, slots, stream_seed, t, agent0, explore_threshold, deterministic, empty_all, d_o, n, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(actions_out, d_o, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return QE_OK;
}

// ---------------------------------------------------------------------------------------------- learn
}  // extern "C"
template <int LPR>
static int launch_learn_exact(qe_engine* e, const int32_t* s, const int32_t* a, const float* r, const int32_t* s2,
                              const uint8_t* term, const uint32_t* m2, float lr, int n, cudaStream_t st) {
    int blocks = 0;
    int rc = coop_blocks(e, learn_exact_kernel<LPR>, (long long)n, &blocks);
    if (rc) return rc;
    if ((rc = ensure_info(e)) != QE_OK) return rc;
    uint32_t epoch = ++e->step;
    Table T = e->T;
    float gamma = e->gamma;
    int* cursor = e->tile_counter;
    CK(cudaMemsetAsync(cursor, 0, 8 * sizeof(int), st));
    CK(cudaMemsetAsync(e->T.spill_next, 0, 2 * sizeof(int), st));
    void* args[] = {&T, &s, &a, &r, &s2, &term, &m2, &lr, &gamma, &epoch, &cursor, &n};
    CK(cudaLaunchCooperativeKernel((void*)learn_exact_kernel<LPR>, dim3(blocks), dim3(256), args, 0, st));
    e->launches++;
    return QE_OK;
}

extern "C" {
static int learn_impl(qe_engine* e, const int32_t* s, const int32_t* a, const float* r, const int32_t* s2, const uint8_t* term,
                      const uint32_t* m2, const uint8_t* m2b, float lr, int n, int mode, cudaStream_t st) {
    if (n <= 0) return QE_OK;
    if (n >= (1 << 24)) return fail(QE_ERR_ARG, "at most 2^24-1 agents per call");
    int rc = ensure_agents(e, n);
    if (rc) return rc;
    if (mode == QE_LEARN_ACCUMULATE) {
        learn_delta_kernel<<<(n + 255) / 256, 256, 0, st>>>(e->T, s, a, r, s2, term, m2b, m2, lr, e->gamma, e->delta, n);
        learn_scatter_kernel<<<(n + 255) / 256, 256, 0, st>>>(e->T, s, a, e->delta, n);
        e->launches += 2;
        CK(cudaGetLastError());
        return QE_OK;
    }
    if (mode != QE_LEARN_SEQUENTIAL) return fail(QE_ERR_ARG, "unknown learn mode %d", mode);
    if (e->A > 32 || m2b) {
        learn_sequential_kernel<<<1, 32, 0, st>>>(e->T, s, a, r, s2, term, m2b, m2, lr, e->gamma, n);
        e->launches++;
        CK(cudaGetLastError());
        return QE_OK;
    }
    switch (e->lpr) {
        case 1: return launch_learn_exact<1>(e, s, a, r, s2, term, m2, lr, n, st);
        case 2: return launch_learn_exact<2>(e, s, a, r, s2, term, m2, lr, n, st);
        default: return launch_learn_exact<4>(e, s, a, r, s2, term, m2, lr, n, st);
    }
}

int qe_learn(qe_engine_t* e, const int32_t* states, const int32_t* actions, const float* rewards, const int32_t* next_states,
             const uint8_t* terminated, const uint32_t* next_mask_bits, const uint8_t* next_mask_bytes, float lr, int32_t n,
             int32_t mode, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    CK(cudaSetDevice(e->device));
    return learn_impl(e, states, actions, rewards, next_states, terminated, next_mask_bits, next_mask_bytes, lr, n, mode,
                      (cudaStream_t)stream);
}

int qe_learn_host(qe_engine_t* e, const int32_t* states, const int32_t* actions, const float* rewards,
                  const int32_t* next_states, const uint8_t* terminated, const uint32_t* next_mask_bits,
                  const uint8_t* next_mask_bytes, float lr, int32_t n, int32_t mode) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    CK(cudaSetDevice(e->device));
    const size_t b4 = al(sizeof(int32_t) * n), b1 = al(n), b_m = next_mask_bits ? b4 : 0,
                 b_mb = next_mask_bytes ? al((size_t)n * e->A) : 0;
    int rc = ensure_stage(e, 4 * b4 + b1 + b_m + b_mb);
    if (rc) return rc;
    char* p = (char*)e->stage;
    int32_t* d_s = (int32_t*)p; p += b4;
    int32_t* d_a = (int32_t*)p; p += b4;
    float* d_r = (float*)p; p += b4;
    int32_t* d_s2 = (int32_t*)p; p += b4;
    uint8_t* d_t = (uint8_t*)p; p += b1;
    uint32_t* d_m = next_mask_bits ? (uint32_t*)p : nullptr; p += b_m;
    uint8_t* d_mb = next_mask_bytes ? (uint8_t*)p : nullptr;
    cudaStream_t st = 0;
    CK(cudaMemcpyAsync(d_s, states, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_a, actions, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_r, rewards, sizeof(float) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_s2, next_states, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_t, terminated, n, cudaMemcpyHostToDevice, st));
    if (d_m) CK(cudaMemcpyAsync(d_m, next_mask_bits, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    if (d_mb) CK(cudaMemcpyAsync(d_mb, next_mask_bytes, (size_t)n * e->A, cudaMemcpyHostToDevice, st));
    rc = learn_impl(e, d_s, d_a, d_r, d_s2, d_t, d_m, d_mb, lr, n, mode, st);
    if (rc) return rc;
    return check_device_errors(e, st);
}

int qe_gather(qe_engine_t* e, const int32_t* states, const int32_t* actions, float* out, int32_t n, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    gather_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->T, states, actions, out, n);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}

int qe_gather_rows_host(qe_engine_t* e, const int32_t* states_host, float* out_host, int32_t n) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    CK(cudaSetDevice(e->device));
    const size_t b_s = al(sizeof(int32_t) * n), b_o = al(sizeof(float) * (size_t)n * e->A);
    int rc = ensure_stage(e, b_s + b_o);
    if (rc) return rc;
    int32_t* d_s = (int32_t*)e->stage;
    float* d_o = (float*)((char*)e->stage + b_s);
    CK(cudaMemcpyAsync(d_s, states_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice, 0));
    gather_rows_kernel<<<e->sms * 4, 256>>>(e->T, d_s, d_o, n);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, d_o, sizeof(float) * (size_t)n * e->A, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return QE_OK;
}


// ---------------------------------------------------------------------------------------------- multi-GPU support
int qe_set_state_base(qe_engine_t* e, int64_t first_state) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (first_state < 0) return fail(QE_ERR_ARG, "first_state must be >= 0");
    e->state_base = first_state;
    e->T.q = e->q_real - (size_t)first_state * (size_t)e->ld;  // never dereferenced outside [first_state, first_state + S)
    if (e->info_real) e->T.info = e->info_real - (size_t)first_state * (size_t)e->T.info_ld;
    return QE_OK;
}
int qe_set_agent_ids(qe_engine_t* e, const uint32_t* ids) {
    std::lock_guard<std::mutex> lk(e->mu);
    e->agent_ids = ids;
    return QE_OK;
}
int qe_set_hold(qe_engine_t* e, int32_t hold) {
    std::lock_guard<std::mutex> lk(e->mu);
    e->T.hold = hold != 0;
    return QE_OK;
}
int qe_learn_commit(qe_engine_t* e, const int32_t* states, const int32_t* actions, int32_t n, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    if (n > e->cap) return fail(QE_ERR_ARG, "qe_learn_commit: no held update of that size");
    learn_commit_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(e->T, states, actions, n);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_serve_bootstrap(qe_engine_t* e, const int32_t* rows, const int32_t* before, const uint32_t* mask_bits, float* out,
                       int32_t n, int32_t use_versions, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    if (e->A > 32) return fail(QE_ERR_ARG, "qe_serve_bootstrap supports at most 32 actions");
    { int rc = ensure_info(e); if (rc) return rc; }
    serve_bootstrap_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->T, rows, before, mask_bits, out, n, e->step, use_versions);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_table_export_dense(qe_engine_t* e, float* dense, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    table_export_kernel<<<e->sms * 8, 256, 0, (cudaStream_t)stream>>>(local_table(e), e->S, dense);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_table_import_dense(qe_engine_t* e, const float* dense, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    table_import_kernel<<<e->sms * 8, 256, 0, (cudaStream_t)stream>>>(local_table(e), e->S, dense);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_table_delta_dense(qe_engine_t* e, const float* base, float* delta_out, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    table_delta_kernel<<<e->sms * 8, 256, 0, (cudaStream_t)stream>>>(local_table(e), e->S, base, delta_out);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_table_merge_dense(qe_engine_t* e, float* base_inout, const float* delta_sum, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    table_merge_kernel<<<e->sms * 8, 256, 0, (cudaStream_t)stream>>>(local_table(e), e->S, base_inout, delta_sum);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}

// ---------------------------------------------------------------------------------------------- environments
int qe_ttt_reset(uint32_t* boards, int32_t* states_out, uint32_t* mask_bits_out, const uint32_t* uniforms, int32_t slots,
                 uint32_t stream_seed, uint32_t t, uint32_t agent0, int32_t n, void* stream) {
    if (n <= 0) return QE_OK;
    { int rc = use_device_of(boards); if (rc) return rc; }
    if (uniforms && slots < 5) return fail(QE_ERR_ARG, "TicTacToe needs 5 uniform slots");
    Uniforms U{uniforms, slots, stream_seed, t, agent0, stream_seed, t};
    ttt_reset_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(boards, states_out, mask_bits_out, U, n);
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_ttt_step(qe_engine_t* e, uint32_t* boards, const int32_t* actions, const uint32_t* uniforms, int32_t slots,
                uint32_t stream_seed, uint32_t t, uint32_t agent0, int32_t* next_states, uint32_t* next_mask_bits,
                float* rewards, uint8_t* terminated, int32_t n, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    if (uniforms && slots < 5) return fail(QE_ERR_ARG, "TicTacToe needs 5 uniform slots");
    Uniforms U{uniforms, slots, stream_seed, t, agent0, stream_seed, t};
    ttt_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(boards, actions, U, next_states, next_mask_bits, rewards,
                                                                    terminated, e->T.err, n);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_mdp_reset(int32_t* states, uint32_t* mask_bits_out, int64_t num_states, int32_t num_actions, uint32_t env_seed,
                 const uint32_t* uniforms, int32_t slots, uint32_t stream_seed, uint32_t t, uint32_t agent0, int32_t n,
                 void* stream) {
    if (n <= 0) return QE_OK;
    if (uniforms && slots < 4) return fail(QE_ERR_ARG, "the hash MDP needs 4 uniform slots");
    { int rc = use_device_of(states); if (rc) return rc; }
    Uniforms U{uniforms, slots, stream_seed, t, agent0, stream_seed, t};
    mdp_reset_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(states, mask_bits_out, (uint32_t)num_states, num_actions,
                                                                     env_seed, U, n);
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_mdp_masks(const int32_t* states, uint32_t* mask_bits_out, int32_t num_actions, uint32_t env_seed, int32_t n, void* stream) {
    if (n <= 0) return QE_OK;
    { int rc = use_device_of(states); if (rc) return rc; }
    mdp_masks_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(states, mask_bits_out, num_actions, env_seed, n);
    CK(cudaGetLastError());
    return QE_OK;
}
int qe_mdp_step(qe_engine_t* e, int32_t* states, const int32_t* actions, int64_t num_states, int32_t num_actions,
                uint32_t env_seed, uint64_t term_threshold, const uint32_t* uniforms, int32_t slots, uint32_t stream_seed,
                uint32_t t, uint32_t agent0, uint32_t* next_mask_bits, float* rewards, uint8_t* terminated, int32_t n,
                void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (n <= 0) return QE_OK;
    if (uniforms && slots < 4) return fail(QE_ERR_ARG, "the hash MDP needs 4 uniform slots");
    Uniforms U{uniforms, slots, stream_seed, t, agent0, stream_seed, t};
    U.ids = e->agent_ids;
    mdp_step_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(states, actions, (uint32_t)num_states, num_actions, env_seed,
                                                                    term_threshold, U, next_mask_bits, rewards, terminated,
                                                                    e->T.err, n);
    e->launches++;
    CK(cudaGetLastError());
    return QE_OK;
}

// ---------------------------------------------------------------------------------------------- fused loop
}  // extern "C"
constexpr int kProbeEvery = 12;
// Which exact form the next fused launch uses (strategy 2).  Agents herd as the table is learned, so the faster form
// changes during a run and a comparison is only worth something if both forms were timed at (nearly) the same training
// progress: the first four launches are {lists cold, lists timed, sort cold, sort timed} (a form's first launch pays for
// module loading and scratch allocation), and every kProbeEvery launches after that the other form and then the current
// one are timed in two ADJACENT launches.  A timed launch is collected by the next call (it waits for the event: two
// launches out of kProbeEvery); all other launches carry no events at all.  *time_it: record events around this launch.
static int pick_form(qe_engine* e, const FusedArgs& F, bool* time_it) {
    *time_it = false;
    // small batches: one CTA, no grid barrier (qe_small.cuh); QE_FORM / qe_set_fused_form(0..2) still pin the grid-wide forms
    if (F.n <= kSmallMaxAgents && e->A <= 32 && e->S < (1ll << 26) && !F.accumulate && (e->strategy == 3 || e->strategy == 5) && !getenv("QE_NO_SMALL")) return 4;
    if (e->state_base != 0 || e->A > 32 || !e->X.seg) return 0;
    if (F.accumulate || F.evaluate) return 0;  // the plain-atomics update and the evaluation loop live in fused_kernel
    if (e->strategy == 5) return (e->S < (1ll << 30) && F.n <= (1 << 24)) ? 5 : 1;  // writer records keep the agent in 24 bits
    if (e->strategy == 3) return e->S < (1ll << 30) ? 3 : 1;  // the pipeline's records keep the state in 30 bits
    if (e->strategy == 0 || e->strategy == 1) return e->strategy;
    if (e->timed_kind >= 0) {  // collect the launch timed last
        float ms = 0.0f;
        if (cudaEventSynchronize(e->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess && ms > 0.0f)
            e->rate[e->timed_kind] = e->timed_work / ms;
        e->timed_kind = -1;
        (void)cudaGetLastError();
    }
    const int seq = e->auto_launches < 1000000 ? e->auto_launches++ : e->auto_launches;  // launches of strategy 2 so far
    if (seq < 4) {                    // lists cold, lists timed, sort cold, sort timed
        *time_it = (seq & 1) != 0;
        return seq >> 1;
    }
    if (seq == 4) e->auto_form = e->rate[1] > e->rate[0] ? 1 : 0;
    int best = e->auto_form;
    if (e->since_probe == -1) {  // second half of a probe: the other form was timed by the last launch, now the current one
        e->since_probe = -2;
        *time_it = true;
        return best;
    }
    if (e->since_probe == -2) {  // both are in: decide
        best = e->auto_form = e->rate[1] > e->rate[0] ? 1 : 0;
        e->since_probe = 0;
    }
    if (++e->since_probe >= kProbeEvery) {
        e->since_probe = -1;
        *time_it = true;
        return best ^ 1;
    }
    return best;
}

template <int ENV, int LPR>
static int launch_fused(qe_engine* e, FusedArgs& F, cudaStream_t st) {
    int blocks = 0;
    Table T = e->T;
    bool timed = false;
    const int form = pick_form(e, F, &timed);
    if (timed) CK(cudaEventRecord(e->ev0, st));
    if (form == 4) {
        e->pipe_valid = false;  // (the agents move without the pipelined form's order being kept)
        blocks = 1;
        void* args[] = {&T, &F};
        CK(cudaLaunchKernel((void*)fused_small_kernel<ENV, LPR>, dim3(1), dim3(kSmallMaxAgents), args, 0, st));
    } else if (form == 5) {
        const size_t smem = flow_smem_bytes(LPR);
        CK(cudaFuncSetAttribute(fused_flow_kernel<ENV, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_flow_kernel<ENV, LPR>, 256, smem));
        if (per_sm < 1) return fail(QE_ERR_CUDA, "the one-pass kernel cannot be made resident");
        {
            const long long want = ((long long)F.n + 255) / 256, maxb = (long long)per_sm * e->sms;
            blocks = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
        }
        if (blocks > 32 * kScanPerLane) blocks = 32 * kScanPerLane;
        int rc = ensure_pipe(e, blocks);
        if (rc) return rc;
        if (e->pipe_kind != 5 && (e->pipe_sorted_n > 0 || e->pipe_valid)) {  // bounds left by the other pipelined form: start from all-empty
            CK(cudaMemsetAsync(e->P.seg, 0, sizeof(uint2) * (size_t)e->S, st));
            e->pipe_valid = false;
            e->pipe_sorted_n = 0;
        }
        FlowScratch X{};
        X.rec = reinterpret_cast<uint64_t*>(e->P.rec); X.seg = e->P.seg; X.pos = e->P.pos; X.kv[0] = e->P.kv[0]; X.kv[1] = e->P.kv[1];
        X.ghist = e->P.ghist; X.rowtot = e->P.rowtot; X.ctr = e->P.ctr;
        {
            int bits = 1;
            while (bits < 31 && (1ll << bits) < e->S) ++bits;
            X.msd_shift = bits > kRadixBits ? bits - kRadixBits : 0;
            X.local_passes = (X.msd_shift + kRadixBits - 1) / kRadixBits;
        }
        X.sorted_valid = (e->pipe_valid && e->pipe_states == F.st_a && e->pipe_n == F.n && !getenv("QE_PIPE_RESORT")) ? 1 : 0;
        X.old_n = e->pipe_sorted_n;
        X.flags = getenv("QE_FLOW_FLAGS") ? atoi(getenv("QE_FLOW_FLAGS")) : 0;
        CK(cudaMemsetAsync(e->P.ctr, 0, 64 * sizeof(unsigned int), st));
        void* args[] = {&T, &F, &X};
        CK(cudaLaunchCooperativeKernel((void*)fused_flow_kernel<ENV, LPR>, dim3(blocks), dim3(256), args, smem, st));
        e->pipe_valid = true;
        e->pipe_states = F.st_a;
        e->pipe_n = F.n;
        e->pipe_sorted_n = F.n;
        e->pipe_kind = 5;
    } else if (form == 3) {
        if (e->pipe_kind != 3 && (e->pipe_sorted_n > 0 || e->pipe_valid)) {  // bounds left by the one-pass form
            CK(cudaMemsetAsync(e->P.seg, 0, sizeof(uint2) * (size_t)e->S, st));
            e->pipe_valid = false;
            e->pipe_sorted_n = 0;
        }
        const size_t smem = pipe_smem_bytes(LPR);
        CK(cudaFuncSetAttribute(fused_pipe_kernel<ENV, LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_pipe_kernel<ENV, LPR>, 256, smem));
        if (per_sm < 1) return fail(QE_ERR_CUDA, "the pipelined kernel cannot be made resident");
        {
            const long long want = ((long long)F.n + 255) / 256, maxb = (long long)per_sm * e->sms;
            blocks = (int)(want < 1 ? 1 : (want > maxb ? maxb : want));
        }
        int rc = ensure_pipe(e, blocks);
        if (rc) return rc;
        if (blocks > 32 * kScanPerLane) return fail(QE_ERR_CUDA, "the pipelined kernel's sort supports at most %d blocks", 32 * kScanPerLane);
        e->P.sorted_valid = (e->pipe_valid && e->pipe_states == F.st_a && e->pipe_n == F.n && !getenv("QE_PIPE_RESORT")) ? 1 : 0;
        e->P.state_base = e->state_base;
        e->P.old_n = e->pipe_sorted_n;
        CK(cudaMemsetAsync(e->P.ctr, 0, 64 * sizeof(unsigned int), st));
        PipeScratch P = e->P;
        void* args[] = {&T, &F, &P};
        CK(cudaLaunchCooperativeKernel((void*)fused_pipe_kernel<ENV, LPR>, dim3(blocks), dim3(256), args, smem, st));
        e->pipe_valid = true;
        e->pipe_states = F.st_a;
        e->pipe_n = F.n;
        e->pipe_sorted_n = F.n;
        e->pipe_kind = 3;
    } else if (form == 1) {
        if (!F.evaluate) e->pipe_valid = false;
        int rc = coop_blocks(e, fused_sorted_kernel<ENV, LPR>, (long long)F.n, &blocks);
        if (rc) return rc;
        if (blocks > kSortMaxBlocks * kSortStride) blocks = kSortMaxBlocks * kSortStride;
        if (blocks > e->sorted_grid) {
            CK(cudaStreamSynchronize(st));
            cudaFree(e->X.ghist);
            CK(cudaMalloc(&e->X.ghist, sizeof(int) * kRadix * (size_t)blocks));
            e->sorted_grid = blocks;
        }
        SortedScratch X = e->X;
        void* args[] = {&T, &F, &X};
        CK(cudaLaunchCooperativeKernel((void*)fused_sorted_kernel<ENV, LPR>, dim3(blocks), dim3(256), args, 0, st));
    } else if (F.accumulate) {
        e->pipe_valid = false;
        int rc = coop_blocks(e, fused_kernel<ENV, LPR, true>, (long long)F.n, &blocks);
        if (rc) return rc;
        void* args[] = {&T, &F};
        CK(cudaLaunchCooperativeKernel((void*)fused_kernel<ENV, LPR, true>, dim3(blocks), dim3(256), args, 0, st));
    } else {
        e->pipe_valid = false;  // (an evaluation run moves the agents too)
        int rc = coop_blocks(e, fused_kernel<ENV, LPR>, (long long)F.n, &blocks);
        if (rc) return rc;
        if (!F.evaluate && (rc = ensure_info(e)) != QE_OK) return rc;
        T = e->T;
        void* args[] = {&T, &F};
        CK(cudaLaunchCooperativeKernel((void*)fused_kernel<ENV, LPR>, dim3(blocks), dim3(256), args, 0, st));
    }
    if (timed) {
        CK(cudaEventRecord(e->ev1, st));
        e->timed_kind = form;
        e->timed_work = (double)F.n * (double)F.steps;
    }
    if (!F.evaluate) e->current = form;
    e->launches++;
    e->last_grid = blocks;
    return QE_OK;
}
template <int ENV>
static int launch_fused_env(qe_engine* e, FusedArgs& F, cudaStream_t st) {
    switch (e->lpr) {
        case 1: return launch_fused<ENV, 1>(e, F, st);
        case 2: return launch_fused<ENV, 2>(e, F, st);
        default: return launch_fused<ENV, 4>(e, F, st);
    }
}

extern "C" {
int qe_fused_steps(qe_engine_t* e, const qe_agents_t* ag, const qe_run_t* run, void* stream) {
    std::lock_guard<std::mutex> lk(e->mu);
    if (!ag || !run) return fail(QE_ERR_ARG, "null argument");
    if (run->steps <= 0) return QE_OK;
    if (e->A > 32) return fail(QE_ERR_ARG, "the fused loop supports at most 32 actions");
    const int n = ag->num_agents;
    if (n <= 0 || n >= (1 << 24)) return fail(QE_ERR_ARG, "num_agents must be in [1, 2^24)");
    if (ag->env_kind == QE_ENV_TTT && (e->A != 9 || e->S != 19683)) return fail(QE_ERR_ARG, "TicTacToe needs a 19683 x 9 table");
    if (ag->env_kind == QE_ENV_MDP && (uint64_t)e->S * (uint64_t)e->A >= (1ull << 32)) return fail(QE_ERR_ARG, "hash MDP needs S*A < 2^32");
    if (ag->env_kind != QE_ENV_MDP && !ag->env_words) return fail(QE_ERR_ARG, "env_words required");
    const int need_slots = ag->env_kind == QE_ENV_TTT ? 5 : (ag->env_kind == QE_ENV_MDP ? 4 : 2);
    if (run->uniforms && run->slots < need_slots) return fail(QE_ERR_ARG, "uniforms need %d slots", need_slots);
    CK(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_agents(e, n);
    if (rc) return rc;
    if (run->steps > e->sched_cap) {
        CK(cudaDeviceSynchronize());
        cudaFree(e->d_thresh); cudaFree(e->d_lr);
        int cap = 256;
        while (cap < run->steps) cap <<= 1;
        CK(cudaMalloc(&e->d_thresh, sizeof(uint64_t) * cap));
        CK(cudaMalloc(&e->d_lr, sizeof(float) * cap));
        e->sched_cap = cap;
    }
    CK(cudaMemcpyAsync(e->d_thresh, run->explore_thresholds_host, sizeof(uint64_t) * run->steps, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->d_lr, run->learning_rates_host, sizeof(float) * run->steps, cudaMemcpyHostToDevice, st));
    if (e->step > 0xFFFF0000u - (uint32_t)run->steps) {  // the 32-bit epoch that stamps the scratch of forms 0 and 1 is about to wrap: start it over
        CK(cudaStreamSynchronize(st));
        CK(cudaMemset(e->T.slot, 0, sizeof(uint64_t) * e->cap));
        CK(cudaMemset(e->X.targ, 0, sizeof(uint64_t) * e->cap));
        CK(cudaMemset(e->X.mhist, 0, sizeof(uint64_t) * e->cap));
        if (e->X.seg) CK(cudaMemset(e->X.seg, 0, sizeof(uint32_t) * 4 * (size_t)e->S));
        if (e->info_real) CK(cudaMemset(e->info_real, 0, sizeof(uint32_t) * (size_t)e->S * e->T.info_ld));
        e->step = 0;
    }
    FusedArgs F{};
    F.env_kind = ag->env_kind; F.n = n; F.steps = run->steps;
    F.st_a = ag->states; F.st_b = ag->states_scratch; F.envw = ag->env_words; F.ep_ret = ag->episode_returns;
    F.tr_a = e->tr_a; F.tr_r = e->tr_r;
    F.env_seed = ag->env_seed; F.episode_len = ag->episode_len; F.term_thresh = ag->term_threshold; F.S = e->S;
    F.eps_thresh = e->d_thresh; F.lr = e->d_lr;
    F.uniforms = run->uniforms; F.slots = run->slots; F.stream_seed = run->stream_seed; F.t0 = run->t0; F.agent0 = run->agent0;
    F.env_stream_seed = run->env_stream_seed; F.env_t0 = run->env_t0;
    F.empty_all = run->empty_all; F.use_masks = run->use_masks; F.gamma = e->gamma;
    F.step0 = e->step;
    F.tile_counter = e->tile_counter;
    CK(cudaMemsetAsync(e->tile_counter, 0, 8 * sizeof(int), st));
    CK(cudaMemsetAsync(e->T.spill_next, 0, 2 * sizeof(int), st));
    F.trace_actions = run->trace_actions; F.trace_rewards = run->trace_rewards; F.trace_term = run->trace_terminated;
    F.trace_next = run->trace_next_states; F.trace_epret = run->trace_episode_returns;
    F.ep_sum = run->episode_sum; F.ep_count = run->episode_count;
    F.phase_ns = e->phase_ns;
    F.evaluate = run->evaluate != 0;
    F.accumulate = run->learn_mode == QE_LEARN_ACCUMULATE;
    if (run->learn_mode != QE_LEARN_SEQUENTIAL && run->learn_mode != QE_LEARN_ACCUMULATE) return fail(QE_ERR_ARG, "unknown learn_mode %d", run->learn_mode);
    e->phase_steps = run->steps < 10 ? run->steps : 10;
    if ((F.ep_sum == nullptr) != (F.ep_count == nullptr)) return fail(QE_ERR_ARG, "episode_sum and episode_count go together");
    if (!F.evaluate) e->step += (uint32_t)run->steps;
    switch (ag->env_kind) {
        case QE_ENV_MDP: return launch_fused_env<0>(e, F, st);
        case QE_ENV_TTT: return launch_fused_env<1>(e, F, st);
        case QE_ENV_BANDIT: return launch_fused_env<2>(e, F, st);
        default: return fail(QE_ERR_ARG, "unknown env kind %d", ag->env_kind);
    }
}

}  // extern "C"
