// microbench.cu -- what the memory system of the box allows for this engine's access patterns (development aid and the
// source of bench.py's `roofline.gather_peak`): random row gathers without dependencies, L2 latencies of the load flavours
// the dependency resolution uses, and the cost of one publish -> poll hop between two SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench scripts/micro/microbench.cu ; build/microbench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t fmix32(uint32_t x) { x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16; return x; }
__device__ __forceinline__ uint64_t gns() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ld8(const float* p) {
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}

// (A) random gather of `sectors` x 32 B per row, one lane per sector, `unroll` independent rows in flight per lane group
template <int SECT, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ tab, uint32_t rows, int row_floats, int per_thread, uint32_t seed, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t grp = tid / SECT, l = tid % SECT;
    float acc = 0.f;
    for (int it = 0; it < per_thread; it += UNROLL) {
        F8 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t r = __umulhi(fmix32((grp * 7919u + (uint32_t)(it + u)) ^ seed), rows);
            v[u] = ld8(tab + (size_t)r * row_floats + 8 * l);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[u].v[j];
    }
    if (acc == 12345.678f) out[0] = acc;
}

// (B) pointer chase, one thread; flavour 0: ld.global.cg, 1: ld.relaxed.gpu, 2: ld.global (L1), 3: ld.acquire.gpu
__global__ void chase_kernel(const uint32_t* next, int hops, int flavour, uint64_t* out_ns, uint32_t* sink) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t p = 0;
    const uint64_t t0 = gns();
    for (int i = 0; i < hops; ++i) {
        const uint32_t* a = next + (size_t)p * 8;  // 32-byte stride
        if (flavour == 0) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(p) : "l"(a) : "memory");
        else if (flavour == 1) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(p) : "l"(a) : "memory");
        else if (flavour == 2) asm volatile("ld.global.u32 %0, [%1];" : "=r"(p) : "l"(a) : "memory");
        else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(p) : "l"(a) : "memory");
    }
    out_ns[0] = gns() - t0;
    sink[0] = p;
}
// background load for (B'): random gathers until stop flag
__global__ void __launch_bounds__(256) noise_kernel(const float* __restrict__ tab, uint32_t rows, int row_floats, volatile int* stop, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (uint32_t it = 0; !*stop && it < 40000000u; ++it) {
        const uint32_t r = __umulhi(fmix32((tid * 7919u + it) ^ 99u), rows);
        const F8 v = ld8(tab + (size_t)r * row_floats + 8 * (tid & 1));
        acc += v.v[0];
    }
    if (acc == 12345.678f) out[0] = acc;
}

// (C) ping-pong between block 0 and block `peer`: one publish -> poll hop = half a round trip
__global__ void pingpong_kernel(uint32_t* flags, int iters, int peer, uint64_t* out_ns) {
    if (threadIdx.x != 0) return;
    uint32_t* a = flags;        // written by block 0
    uint32_t* b = flags + 64;   // written by the peer (another 128-byte line)
    if (blockIdx.x == 0) {
        const uint64_t t0 = gns();
        for (int i = 1; i <= iters; ++i) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a), "r"(i) : "memory");
            uint32_t v;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(b) : "memory"); } while (v != (uint32_t)i);
        }
        out_ns[0] = gns() - t0;
    } else if ((int)blockIdx.x == peer) {
        for (int i = 1; i <= iters; ++i) {
            uint32_t v;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory"); } while (v != (uint32_t)i);
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(b), "r"(i) : "memory");
        }
    }
}

// (D) grid barrier
__global__ void __launch_bounds__(256, 4) gridsync_kernel(int iters, uint64_t* out) {
    cg::grid_group grid = cg::this_grid();
    const uint64_t t0 = gns();
    for (int i = 0; i < iters; ++i) grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gns() - t0;
}

// (E) random 4-byte atomics / stores into an array (writer registration, histograms)
__global__ void __launch_bounds__(256) atomic_kernel(unsigned int* cnt, uint32_t rows, int per_thread, int mode) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < per_thread; ++it) {
        const uint32_t r = __umulhi(fmix32((tid * 7919u + (uint32_t)it) ^ 5u), rows);
        if (mode == 0) atomicAdd(cnt + r, 1u);
        else cnt[r] = tid;
    }
}

template <typename F> static float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d", prop.name, sms);
    float* out; CK(cudaMalloc(&out, 256));
    uint64_t* d_ns; CK(cudaMalloc(&d_ns, 64)); uint64_t h_ns = 0;
    // ---- (A) gathers: 2^20 rows of 64 B (64 MB, config 3 dense), 2^20 rows with 256-byte stride (round 1's row blocks), 100M rows of 32 B (config 4)
    struct Cfg { const char* name; size_t rows; int row_floats; int sect; } cfgs[] = {
        {"c3_dense_64B_rows_64MB", 1000000, 16, 2}, {"c3_rowblock_256B_stride_256MB", 1000000, 64, 2}, {"c4_dense_32B_rows_3200MB", 100000000, 8, 1}};
    for (auto& c : cfgs) {
        float* tab; CK(cudaMalloc(&tab, c.rows * c.row_floats * sizeof(float))); CK(cudaMemset(tab, 0, c.rows * c.row_floats * sizeof(float)));
        const int blocks = sms * 8, per_thread = 64;
        const double rows_read = (double)blocks * 256 / c.sect * per_thread;
        float ms1, ms4;
        if (c.sect == 2) {
            ms1 = time_ms([&] { gather_kernel<2, 1><<<blocks, 256>>>(tab, (uint32_t)c.rows, c.row_floats, per_thread, 1u, out); });
            ms4 = time_ms([&] { gather_kernel<2, 4><<<blocks, 256>>>(tab, (uint32_t)c.rows, c.row_floats, per_thread, 2u, out); });
        } else {
            ms1 = time_ms([&] { gather_kernel<1, 1><<<blocks, 256>>>(tab, (uint32_t)c.rows, c.row_floats, per_thread, 1u, out); });
            ms4 = time_ms([&] { gather_kernel<1, 4><<<blocks, 256>>>(tab, (uint32_t)c.rows, c.row_floats, per_thread, 2u, out); });
        }
        const double bytes = rows_read * c.sect * 32;
        printf(", \"gather_%s\": {\"GBs_1_in_flight\": %.1f, \"GBs_4_in_flight\": %.1f, \"Grows_per_s\": %.2f}", c.name, bytes / ms1 / 1e6, bytes / ms4 / 1e6, rows_read / ms4 / 1e6);
        CK(cudaFree(tab));
    }
    // ---- (B) latencies
    {
        const size_t nodes = 1 << 19;  // 16 MB at 32-byte stride: L2 resident
        uint32_t* h = (uint32_t*)malloc(nodes * 32); uint32_t* perm = (uint32_t*)malloc(nodes * 4);
        for (size_t i = 0; i < nodes; ++i) perm[i] = (uint32_t)i;
        srand(1);
        for (size_t i = nodes - 1; i > 0; --i) { size_t j = ((size_t)rand() * RAND_MAX + rand()) % (i + 1); uint32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
        for (size_t i = 0; i < nodes; ++i) h[(size_t)perm[i] * 8] = perm[(i + 1) % nodes];
        uint32_t* d; CK(cudaMalloc(&d, nodes * 32)); CK(cudaMemcpy(d, h, nodes * 32, cudaMemcpyHostToDevice));
        uint32_t* sink; CK(cudaMalloc(&sink, 4));
        float* tab; const size_t rows = 1000000; CK(cudaMalloc(&tab, rows * 64)); CK(cudaMemset(tab, 0, rows * 64));
        int* stop; CK(cudaMallocHost(&stop, 4));
        const char* names[4] = {"ld_cg", "ld_relaxed_gpu", "ld_l1", "ld_acquire_gpu"};
        cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        for (int loaded = 0; loaded < 2; ++loaded) {
            for (int f = 0; f < 4; ++f) {
                const int hops = 20000;
                chase_kernel<<<1, 32, 0, s1>>>(d, hops, f, d_ns, sink);  // warm the L2
                CK(cudaStreamSynchronize(s1));
                *stop = 0;
                if (loaded) noise_kernel<<<sms * 8 - 1, 256, 0, s2>>>(tab, (uint32_t)rows, 16, stop, out);
                chase_kernel<<<1, 32, 0, s1>>>(d, hops, f, d_ns, sink);
                CK(cudaStreamSynchronize(s1));
                *stop = 1;
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(&h_ns, d_ns, 8, cudaMemcpyDeviceToHost));
                printf(", \"latency_ns_%s_%s\": %.0f", names[f], loaded ? "under_gather_load" : "idle", (double)h_ns / hops);
            }
        }
        CK(cudaFree(d)); CK(cudaFree(tab));
    }
    // ---- (C) publish -> poll hop
    {
        uint32_t* flags; CK(cudaMalloc(&flags, 1024)); 
        for (int peer : {1, 2, 75, 147}) {
            CK(cudaMemset(flags, 0, 1024));
            const int iters = 20000;
            void* args[] = {&flags, (void*)&iters, &peer, &d_ns};
            CK(cudaLaunchCooperativeKernel((void*)pingpong_kernel, dim3(sms), dim3(32), args, 0, 0));
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h_ns, d_ns, 8, cudaMemcpyDeviceToHost));
            printf(", \"hop_ns_block0_block%d\": %.0f", peer, (double)h_ns / iters / 2);
        }
    }
    // ---- (D) grid barrier, 4 CTAs of 256 threads per SM
    {
        int iters = 200;
        void* args[] = {&iters, &d_ns};
        CK(cudaLaunchCooperativeKernel((void*)gridsync_kernel, dim3(sms * 4), dim3(256), args, 0, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h_ns, d_ns, 8, cudaMemcpyDeviceToHost));
        printf(", \"gridsync_us_592x256\": %.2f", (double)h_ns / iters / 1e3);
    }
    // ---- (E) random atomics / stores into 1M counters (4 MB) and 100M (400 MB)
    for (size_t rows : {(size_t)1000000, (size_t)100000000}) {
        unsigned int* cnt; CK(cudaMalloc(&cnt, rows * 4)); CK(cudaMemset(cnt, 0, rows * 4));
        const int blocks = sms * 8, per = 16;
        for (int mode = 0; mode < 2; ++mode) {
            const float ms = time_ms([&] { atomic_kernel<<<blocks, 256>>>(cnt, (uint32_t)rows, per, mode); });
            printf(", \"%s_G_per_s_%zuM_slots\": %.2f", mode ? "random_store4" : "random_atomic_add", rows / 1000000, (double)blocks * 256 * per / ms / 1e6);
        }
        CK(cudaFree(cnt));
    }
    printf("}\n");
    return 0;
}
