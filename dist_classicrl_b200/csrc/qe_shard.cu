// qe_shard.cu -- C ABI of the peer-memory sharded table (include/qe_engine.h, "sharded table"); kernels in qe_shard.cuh.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>

#include "../../include/qe_engine.h"
#include "qe_shard.cuh"

using namespace qe;

extern "C" int qe_set_last_error(int code, const char* msg);  // qe_engine.cu (thread-local message shared by the library)

static int sfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return qe_set_last_error(code, buf);
}
#define SCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) return sfail(QE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct qe_shard {
    int rank = 0, world = 0, device = 0, sms = 0;
    int64_t S = 0, rows = 0;
    int A = 0, ld = 0, lpr = 0, passes = 0, msd_shift = 0;
    int n_total = 0, n_home = 0, nh = 0;
    float gamma = 0.0f;
    uint32_t env_seed = 0;
    char* slab = nullptr;       // the shared part (one allocation: one IPC handle, or the caller's symmetric memory)
    bool slab_external = false;
    size_t slab_bytes = 0;
    size_t off_q = 0, off_rec = 0, off_seg = 0, off_inbox = 0, off_pos = 0, off_tw = 0, off_cin = 0, off_flag = 0;
    char* peer_slab[kMaxRanks] = {};   // every rank's slab as mapped here (own: slab)
    bool peer_ipc[kMaxRanks] = {};
    ShardLocal L{};
    int ghist_blocks = 0;
    uint64_t* d_thresh = nullptr;
    float* d_lr = nullptr;
    int sched_cap = 0;
    uint32_t epoch = 0;         // barrier epoch (advances identically on every rank)
    bool sorted_valid = false;
    uint32_t t = 0;             // vector steps done (both uniform streams)
    int64_t launches = 0;
};

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static ShardPeer view_of(const qe_shard* s, char* base) {
    ShardPeer v;
    v.q = (float*)(base + s->off_q);
    v.rec = (uint2*)(base + s->off_rec);
    v.seg = (uint2*)(base + s->off_seg);
    v.inbox = (int2*)(base + s->off_inbox);
    v.pos = (int32_t*)(base + s->off_pos);
    v.tw = (uint4*)(base + s->off_tw);
    v.cin = (unsigned int*)(base + s->off_cin);
    v.flag = (unsigned int*)(base + s->off_flag);
    return v;
}

__global__ void shard_fill_kernel(float* q, int64_t rows_here, int ld, int A, uint32_t seed, int64_t first_state) {
    const size_t total = (size_t)rows_here * ld, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const int a = (int)(x % ld);
        const size_t s = x / ld + (size_t)first_state;
        q[x] = a < A ? (float)(fmix32((uint32_t)(s * (size_t)A + a) ^ (seed * kGold)) >> 8) * 5.9604644775390625e-08f : 0.0f;
    }
}
__global__ void shard_reset_kernel(int32_t* states, uint32_t S, uint32_t seed, uint32_t t, uint32_t agent0, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) states[i] = (int32_t)pick(stream_u32(seed, t, agent0 + (uint32_t)i, 3u), S);
}
__global__ void shard_rows_kernel(const float* q, int ld, int A, const int64_t* local_rows, float* out, int n) {
    const size_t total = (size_t)n * A, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
        const size_t i = x / A;
        out[x] = q[(size_t)local_rows[i] * ld + (x - i * A)];
    }
}

// development aid: random 32-byte loads (mode 0) or 8-byte stores (mode 1) over a rank's table shard as mapped here
__global__ void __launch_bounds__(256) shard_probe_kernel(float* buf, uint32_t rows, int per_thread, uint32_t seed, int mode, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < per_thread; ++it) {
        const uint32_t r = __umulhi(fmix32((tid * 7919u + (uint32_t)it) ^ seed), rows);
        float* p = buf + (size_t)r * 8;
        if (mode == 0) { const F8 v = ld_row8(p); acc += v.v[0] + v.v[7]; }
        else *reinterpret_cast<uint2*>(p) = make_uint2(tid, (uint32_t)it);
    }
    if (acc == 12345.678f) out[0] = acc;
}

static size_t shard_layout(qe_shard* s) {
    size_t o = 0;
    s->off_q = o; o += al256(sizeof(float) * (size_t)s->rows * s->ld);
    s->off_rec = o; o += al256(sizeof(uint2) * ((size_t)s->n_total + 8));
    s->off_seg = o; o += al256(sizeof(uint2) * (size_t)s->rows);
    s->off_inbox = o; o += al256(sizeof(int2) * (size_t)s->n_total);
    s->off_pos = o; o += al256(sizeof(int32_t) * (size_t)s->n_home);
    s->off_tw = o; o += al256(sizeof(uint4) * (size_t)s->n_home);
    s->off_cin = o; o += al256(sizeof(unsigned int) * kMaxRanks);
    s->off_flag = o; o += al256(sizeof(unsigned int) * kMaxRanks);
    return o;
}

extern "C" {

/* bytes of the slab every rank shares with its peers (the same on every rank) */
int64_t qe_shard_slab_bytes(int64_t num_states, int32_t num_actions, int32_t world, int32_t num_agents) {
    if (num_states <= 0 || num_actions <= 0 || num_actions > 32 || world < 1 || num_agents <= 0) return -1;
    qe_shard t;
    t.lpr = num_actions <= 8 ? 1 : (num_actions <= 16 ? 2 : 4);
    t.ld = 8 * t.lpr;
    t.rows = (num_states + world - 1) / world;
    t.n_total = num_agents;
    t.n_home = (num_agents + world - 1) / world;
    return (int64_t)shard_layout(&t);
}

/* development aid: G random accesses per second over the table shard of `peer_rank` (its first `rows` rows of 32 bytes) */
double qe_shard_probe(qe_shard_t* s, int32_t peer_rank, int32_t mode) {
    if (cudaSetDevice(s->device) != cudaSuccess || !s->peer_slab[peer_rank]) return -1.0;
    float* buf = (float*)(s->peer_slab[peer_rank] + s->off_q);
    const uint32_t rows = (uint32_t)std::min<int64_t>(s->rows * s->ld / 8, 1ll << 30);
    float* out = nullptr;
    if (cudaMalloc(&out, 256) != cudaSuccess) return -1.0;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = s->sms * 3, per = 32;
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        shard_probe_kernel<<<blocks, 256>>>(buf, rows, per, 17u + r, mode, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    return (double)blocks * 256 * per / best / 1e6;
}

int qe_shard_create(int64_t num_states, int32_t num_actions, float discount_factor, int32_t device, int32_t rank, int32_t world,
                    int32_t num_agents, uint32_t env_seed, void* external_slab, qe_shard_t** out) {
    if (!out || num_states <= 0 || num_actions <= 0 || num_actions > 32) return sfail(QE_ERR_ARG, "sharded table: 1 <= actions <= 32");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return sfail(QE_ERR_ARG, "sharded table: 1 <= world <= %d", kMaxRanks);
    if (num_agents <= 0 || num_agents >= (1 << 24)) return sfail(QE_ERR_ARG, "num_agents must be in [1, 2^24)");
    if ((uint64_t)num_states * (uint64_t)num_actions >= (1ull << 32) || num_states >= (1ll << 31)) return sfail(QE_ERR_ARG, "hash MDP needs S*A < 2^32");
    SCK(cudaSetDevice(device));
    qe_shard* s = new (std::nothrow) qe_shard();
    if (!s) return sfail(QE_ERR_ARG, "out of host memory");
    s->rank = rank; s->world = world; s->device = device;
    s->S = num_states; s->A = num_actions; s->gamma = discount_factor; s->env_seed = env_seed;
    s->lpr = num_actions <= 8 ? 1 : (num_actions <= 16 ? 2 : 4);
    s->ld = 8 * s->lpr;
    s->rows = (num_states + world - 1) / world;
    s->n_total = num_agents;
    s->n_home = (num_agents + world - 1) / world;
    s->nh = std::max(0, std::min(s->n_home, num_agents - rank * s->n_home));
    int bits = 1;
    while (bits < 31 && (1ll << bits) < s->rows) ++bits;
    s->passes = (bits + kRadixBits - 1) / kRadixBits;
    s->msd_shift = bits > kRadixBits ? bits - kRadixBits : 0;
    cudaDeviceProp prop;
    SCK(cudaGetDeviceProperties(&prop, device));
    s->sms = prop.multiProcessorCount;
    const size_t o = shard_layout(s);
    s->slab_bytes = o;
    if (external_slab) {  // device memory the caller shares with the peers itself (symmetric memory: large pages over NVLink)
        s->slab = (char*)external_slab;
        s->slab_external = true;
    } else {
        SCK(cudaMalloc(&s->slab, o));
    }
    SCK(cudaMemset(s->slab, 0, o));
    SCK(cudaMemset(s->slab + s->off_rec, 0xFF, sizeof(uint2) * ((size_t)num_agents + 8)));
    s->peer_slab[rank] = s->slab;
    ShardLocal& L = s->L;
    SCK(cudaMalloc(&L.st_a, sizeof(int32_t) * (size_t)s->n_home));
    SCK(cudaMalloc(&L.st_b, sizeof(int32_t) * (size_t)s->n_home));
    SCK(cudaMalloc(&L.ep_ret, sizeof(float) * (size_t)s->n_home));
    SCK(cudaMemset(L.st_a, 0, sizeof(int32_t) * (size_t)s->n_home));
    SCK(cudaMemset(L.ep_ret, 0, sizeof(float) * (size_t)s->n_home));
    for (int b = 0; b < 2; ++b) SCK(cudaMalloc(&L.kv[b], sizeof(int2) * (size_t)num_agents));
    SCK(cudaMalloc(&L.rowtot, sizeof(int) * kRadix));
    SCK(cudaMalloc(&L.ctr, 16 * sizeof(unsigned int)));
    SCK(cudaMemset(L.ctr, 0, 16 * sizeof(unsigned int)));
    SCK(cudaMalloc(&L.ep_sum, sizeof(double)));
    SCK(cudaMalloc(&L.ep_count, sizeof(unsigned long long)));
    SCK(cudaMemset(L.ep_sum, 0, sizeof(double)));
    SCK(cudaMemset(L.ep_count, 0, sizeof(unsigned long long)));
    SCK(cudaMalloc(&L.err, sizeof(int)));
    SCK(cudaMemset(L.err, 0, sizeof(int)));
    SCK(cudaMalloc(&L.phase_ns, 128 * sizeof(unsigned long long)));
    SCK(cudaMemset(L.phase_ns, 0, 128 * sizeof(unsigned long long)));
    *out = s;
    return QE_OK;
}

int qe_shard_destroy(qe_shard_t* s) {
    if (!s) return QE_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    for (int g = 0; g < kMaxRanks; ++g)
        if (s->peer_ipc[g] && s->peer_slab[g]) cudaIpcCloseMemHandle(s->peer_slab[g]);
    if (!s->slab_external) cudaFree(s->slab);
    ShardLocal& L = s->L;
    cudaFree(L.st_a); cudaFree(L.st_b); cudaFree(L.ep_ret); cudaFree(L.kv[0]); cudaFree(L.kv[1]); cudaFree(L.ghist); cudaFree(L.rowtot);
    cudaFree(L.wcnt); cudaFree(L.bcnt); cudaFree(L.ctr); cudaFree(L.phase_ns); cudaFree(L.ep_sum); cudaFree(L.ep_count); cudaFree(L.err); cudaFree(s->d_thresh); cudaFree(s->d_lr);
    delete s;
    return QE_OK;
}

/* the CUDA IPC handle (64 bytes) of this rank's slab, to be opened by the other processes */
int qe_shard_ipc_handle(qe_shard_t* s, void* out64) {
    if (s->slab_external) return sfail(QE_ERR_ARG, "the slab belongs to the caller: share it the way it was allocated");
    SCK(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    SCK(cudaIpcGetMemHandle(&h, s->slab));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &h, 64);
    return QE_OK;
}
int qe_shard_connect_ipc(qe_shard_t* s, int32_t peer_rank, const void* handle64) {
    if (peer_rank < 0 || peer_rank >= s->world || peer_rank == s->rank) return sfail(QE_ERR_ARG, "bad peer rank %d", peer_rank);
    SCK(cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    SCK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_slab[peer_rank] = (char*)p;
    s->peer_ipc[peer_rank] = true;
    return QE_OK;
}
/* the peer's slab as the caller mapped it (symmetric memory) */
int qe_shard_connect_ptr(qe_shard_t* s, int32_t peer_rank, void* peer_slab) {
    if (peer_rank < 0 || peer_rank >= s->world || !peer_slab) return sfail(QE_ERR_ARG, "bad peer rank %d", peer_rank);
    s->peer_slab[peer_rank] = (char*)peer_slab;
    s->peer_ipc[peer_rank] = false;
    return QE_OK;
}
/* ranks that share a process (and a device): plain pointers */
int qe_shard_connect_local(qe_shard_t* s, int32_t peer_rank, qe_shard_t* peer) {
    if (peer_rank < 0 || peer_rank >= s->world || !peer || peer->rank != peer_rank) return sfail(QE_ERR_ARG, "bad peer rank %d", peer_rank);
    if (peer->S != s->S || peer->A != s->A || peer->world != s->world || peer->n_total != s->n_total) return sfail(QE_ERR_ARG, "peer has another shape");
    s->peer_slab[peer_rank] = peer->slab;
    s->peer_ipc[peer_rank] = false;
    return QE_OK;
}

int qe_shard_fill_random(qe_shard_t* s, uint32_t seed, void* stream) {
    SCK(cudaSetDevice(s->device));
    const int64_t first = (int64_t)s->rank * s->rows;
    const int64_t here = std::max<int64_t>(0, std::min<int64_t>(s->rows, s->S - first));
    if (here > 0) shard_fill_kernel<<<s->sms * 8, 256, 0, (cudaStream_t)stream>>>((float*)(s->slab + s->off_q), here, s->ld, s->A, seed, first);
    SCK(cudaGetLastError());
    s->launches++;
    return QE_OK;
}
/* initial states of this rank's agents (global ids rank * n_home ...): the hash MDP's reset draw, slot 3 of U[t_init][id] */
int qe_shard_reset(qe_shard_t* s, uint32_t stream_seed, uint32_t t_init, void* stream) {
    SCK(cudaSetDevice(s->device));
    if (s->nh > 0) shard_reset_kernel<<<(s->nh + 255) / 256, 256, 0, (cudaStream_t)stream>>>(s->L.st_a, (uint32_t)s->S, stream_seed, t_init, (uint32_t)(s->rank * s->n_home), s->nh);
    SCK(cudaGetLastError());
    SCK(cudaMemsetAsync(s->L.ep_ret, 0, sizeof(float) * (size_t)s->n_home, (cudaStream_t)stream));
    s->sorted_valid = false;
    s->t = 0;
    s->launches++;
    return QE_OK;
}

/* K vector steps of ALL ranks in `ranks` (the ranks of this process; one per process when every rank has its own GPU,
 * all of them on one GPU in the one-GPU emulation: they then run side by side in one cooperative launch) */
int qe_shard_steps(qe_shard_t* const* ranks, int32_t nlocal, int32_t steps, const uint64_t* explore_thresholds_host,
                   const float* learning_rates_host, uint32_t stream_seed, uint32_t env_stream_seed, int32_t empty_all, int32_t use_masks,
                   uint64_t term_threshold, void* stream) {
    if (!ranks || nlocal < 1 || nlocal > kMaxRanks || steps <= 0) return sfail(QE_ERR_ARG, "qe_shard_steps: bad arguments");
    qe_shard* s0 = ranks[0];
    const int G = s0->world;
    if (nlocal != 1 && nlocal != G) return sfail(QE_ERR_ARG, "qe_shard_steps: pass this process's one rank, or all %d ranks on one GPU", G);
    for (int v = 0; v < nlocal; ++v) {
        if (ranks[v]->rank != s0->rank + v || ranks[v]->device != s0->device) return sfail(QE_ERR_ARG, "qe_shard_steps: ranks must be consecutive and on one device");
        for (int g = 0; g < G; ++g)
            if (!ranks[v]->peer_slab[g]) return sfail(QE_ERR_ARG, "qe_shard_steps: rank %d is not connected to rank %d", ranks[v]->rank, g);
    }
    SCK(cudaSetDevice(s0->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (steps > s0->sched_cap) {
        SCK(cudaDeviceSynchronize());
        cudaFree(s0->d_thresh); cudaFree(s0->d_lr);
        int cap = 256;
        while (cap < steps) cap <<= 1;
        SCK(cudaMalloc(&s0->d_thresh, sizeof(uint64_t) * cap));
        SCK(cudaMalloc(&s0->d_lr, sizeof(float) * cap));
        s0->sched_cap = cap;
    }
    SCK(cudaMemcpyAsync(s0->d_thresh, explore_thresholds_host, sizeof(uint64_t) * steps, cudaMemcpyHostToDevice, st));
    SCK(cudaMemcpyAsync(s0->d_lr, learning_rates_host, sizeof(float) * steps, cudaMemcpyHostToDevice, st));
    // grid: as many resident blocks as fit, split evenly over the local ranks
    const size_t smem = shard_smem_bytes(s0->lpr);
    const void* kern = s0->lpr == 1 ? (const void*)shard_kernel<1> : (s0->lpr == 2 ? (const void*)shard_kernel<2> : (const void*)shard_kernel<4>);
    SCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) return sfail(QE_ERR_CUDA, "the sharded kernel cannot be made resident");
    int bpr = (per_sm * s0->sms) / nlocal;
    const int want = (s0->n_home + 255) / 256;
    if (bpr > want) bpr = want;
    if (bpr > 32 * kScanPerLane) bpr = 32 * kScanPerLane;
    if (bpr < 1) return sfail(QE_ERR_CUDA, "not enough resident blocks for %d ranks", nlocal);
    ShardArgs H{};
    H.G = G; H.first_rank = s0->rank; H.nlocal = nlocal; H.blocks_per_rank = bpr; H.multi_device = nlocal == 1 && G > 1;
    H.A = s0->A; H.ld = s0->ld; H.passes = s0->passes; H.msd_shift = s0->msd_shift;
    H.n_total = s0->n_total; H.n_home = s0->n_home; H.S = s0->S; H.rows = s0->rows;
    H.steps = steps; H.eps_thresh = s0->d_thresh; H.lr = s0->d_lr;
    H.stream_seed = stream_seed; H.t0 = s0->t; H.env_stream_seed = env_stream_seed; H.env_t0 = s0->t; H.env_seed = s0->env_seed;
    H.term_thresh = term_threshold; H.empty_all = empty_all; H.use_masks = use_masks; H.gamma = s0->gamma;
    H.epoch0 = s0->epoch;
    H.sorted_valid = s0->sorted_valid ? 1 : 0;
    for (int g = 0; g < G; ++g) H.peer[g] = view_of(s0, s0->peer_slab[g]);
    const int warps_per_rank = bpr * 8;
    for (int v = 0; v < nlocal; ++v) {
        qe_shard* s = ranks[v];
        if (s->sorted_valid != s0->sorted_valid || s->t != s0->t || s->epoch != s0->epoch) return sfail(QE_ERR_ARG, "qe_shard_steps: the local ranks are out of step");
        if (bpr > s->ghist_blocks) {
            SCK(cudaDeviceSynchronize());
            cudaFree(s->L.ghist); cudaFree(s->L.wcnt); cudaFree(s->L.bcnt);
            s->L.ghist = nullptr; s->L.wcnt = nullptr; s->L.bcnt = nullptr;
            SCK(cudaMalloc(&s->L.ghist, sizeof(int) * kRadix * (size_t)bpr));
            SCK(cudaMalloc(&s->L.wcnt, sizeof(unsigned int) * (size_t)warps_per_rank * kMaxRanks));
            SCK(cudaMalloc(&s->L.bcnt, sizeof(unsigned int) * (size_t)bpr * kMaxRanks));
            s->ghist_blocks = bpr;
        }
        SCK(cudaMemsetAsync(s->L.ctr + 4, 0, sizeof(unsigned int), st));
        H.loc[v] = s->L;
    }
    void* args[] = {&H};
    SCK(cudaLaunchCooperativeKernel(kern, dim3(bpr * nlocal), dim3(256), args, smem, st));
    // barriers over all ranks in this launch: 3 for the initial order, 3 per step
    const uint32_t xs = (s0->sorted_valid ? 0u : 3u) + 3u * (uint32_t)steps;
    for (int v = 0; v < nlocal; ++v) {
        qe_shard* s = ranks[v];
        if (H.multi_device) s->epoch += xs;
        s->sorted_valid = true;
        s->t += (uint32_t)steps;
        s->launches++;
    }
    return QE_OK;
}

/* cudaStreamSynchronize + deferred device errors of this rank */
int qe_shard_sync(qe_shard_t* s, void* stream) {
    SCK(cudaSetDevice(s->device));
    int h = 0;
    SCK(cudaMemcpyAsync(&h, s->L.err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SCK(cudaStreamSynchronize((cudaStream_t)stream));
    if (h) {
        SCK(cudaMemsetAsync(s->L.err, 0, sizeof(int), (cudaStream_t)stream));
        s->sorted_valid = false;
        if (h & kErrEmpty) return sfail(QE_ERR_EMPTY, "empty candidate or bootstrap action set");
        return sfail(QE_ERR_TIMEOUT, "sharded TD update: a wait timed out (rank %d)", s->rank);
    }
    return QE_OK;
}

/* this rank's part of the results (HOST buffers; any may be NULL): table rows [rows_here][A], states and running returns of
 * its home agents [n_home_here], episode statistics.  Synchronous. */
int qe_shard_download(qe_shard_t* s, float* table_host, int32_t* states_host, float* returns_host, double* episode_sum, uint64_t* episode_count) {
    SCK(cudaSetDevice(s->device));
    SCK(cudaDeviceSynchronize());
    const int64_t first = (int64_t)s->rank * s->rows;
    const int64_t here = std::max<int64_t>(0, std::min<int64_t>(s->rows, s->S - first));
    if (table_host && here > 0)
        SCK(cudaMemcpy2D(table_host, sizeof(float) * s->A, s->slab + s->off_q, sizeof(float) * s->ld, sizeof(float) * s->A, (size_t)here, cudaMemcpyDeviceToHost));
    if (states_host && s->nh > 0) SCK(cudaMemcpy(states_host, s->L.st_a, sizeof(int32_t) * s->nh, cudaMemcpyDeviceToHost));
    if (returns_host && s->nh > 0) SCK(cudaMemcpy(returns_host, s->L.ep_ret, sizeof(float) * s->nh, cudaMemcpyDeviceToHost));
    if (episode_sum) SCK(cudaMemcpy(episode_sum, s->L.ep_sum, sizeof(double), cudaMemcpyDeviceToHost));
    if (episode_count) SCK(cudaMemcpy(episode_count, s->L.ep_count, sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return QE_OK;
}
/* rows of this shard by GLOBAL state id (all ids must be owned by this rank): out_host[n][A].  Synchronous. */
int qe_shard_rows_host(qe_shard_t* s, const int64_t* states_host, float* out_host, int32_t n) {
    if (n <= 0) return QE_OK;
    SCK(cudaSetDevice(s->device));
    const int64_t first = (int64_t)s->rank * s->rows;
    int64_t* d_rows = nullptr;
    float* d_out = nullptr;
    int64_t* h = (int64_t*)malloc(sizeof(int64_t) * n);
    if (!h) return sfail(QE_ERR_ARG, "out of host memory");
    for (int i = 0; i < n; ++i) {
        h[i] = states_host[i] - first;
        if (h[i] < 0 || h[i] >= s->rows) { free(h); return sfail(QE_ERR_ARG, "state %lld is not owned by rank %d", (long long)states_host[i], s->rank); }
    }
    SCK(cudaMalloc(&d_rows, sizeof(int64_t) * n));
    SCK(cudaMalloc(&d_out, sizeof(float) * (size_t)n * s->A));
    SCK(cudaMemcpy(d_rows, h, sizeof(int64_t) * n, cudaMemcpyHostToDevice));
    free(h);
    shard_rows_kernel<<<s->sms * 4, 256>>>((const float*)(s->slab + s->off_q), s->ld, s->A, d_rows, d_out, n);
    SCK(cudaGetLastError());
    SCK(cudaMemcpy(out_host, d_out, sizeof(float) * (size_t)n * s->A, cudaMemcpyDeviceToHost));
    cudaFree(d_rows); cudaFree(d_out);
    return QE_OK;
}
/* phase clock of the last launch (synchronous): out_host[8 * k + j], j = 0: start of vector step k, 7: end of this rank's phase A
 * work, 1: after the barrier over all ranks that follows it, 4: after the scatter of the next order's pairs into the owners'
 * inboxes, 2: after phase T and its barrier, 3: after phase C, 6: after the local sort and its barrier; k < 16; %globaltimer ns */
int qe_shard_phase_ns(qe_shard_t* s, uint64_t* out_host128) {
    SCK(cudaSetDevice(s->device));
    SCK(cudaMemcpy(out_host128, s->L.phase_ns, 128 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return QE_OK;
}
int32_t qe_shard_info(qe_shard_t* s, int32_t what) {  /* 0: rows per shard, 1: agents per rank (ceil), 2: agents of this rank, 3: kernels launched */
    switch (what) {
        case 0: return (int32_t)s->rows;
        case 1: return s->n_home;
        case 2: return s->nh;
        case 3: return (int32_t)s->launches;
        default: return -1;
    }
}

}  // extern "C"
