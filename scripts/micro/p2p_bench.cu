// p2p_bench.cu -- what NVLink peer access gives for THIS engine's access patterns (random 32-byte row reads, 8-byte
// record stores, polling loads), one process, two GPUs.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/p2p_bench scripts/micro/p2p_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t fmix32(uint32_t x) { x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16; return x; }
struct __align__(32) F8 { float v[8]; };
template <int MODE>  // 0: ld.global.cg.v8 (32 B), 1: ld.relaxed.sys.v8, 2: ld.global.cg.u32 (4 B), 3: st 8 B, 4: st.relaxed.sys 4 B
__global__ void __launch_bounds__(256) k(float* buf, uint32_t rows, int per_thread, uint32_t seed, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < per_thread; ++it) {
        const uint32_t r = __umulhi(fmix32((tid * 7919u + (uint32_t)it) ^ seed), rows);
        float* p = buf + (size_t)r * 8;
        if (MODE == 0) { F8 v; asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v.v[0]), "=f"(v.v[1]), "=f"(v.v[2]), "=f"(v.v[3]), "=f"(v.v[4]), "=f"(v.v[5]), "=f"(v.v[6]), "=f"(v.v[7]) : "l"(p)); acc += v.v[0] + v.v[7]; }
        else if (MODE == 1) { uint32_t w[8]; asm volatile("ld.relaxed.sys.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p) : "memory"); acc += (float)(w[0] ^ w[7]); }
        else if (MODE == 2) { uint32_t w; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(w) : "l"(p)); acc += (float)w; }
        else if (MODE == 3) { *reinterpret_cast<uint2*>(p) = make_uint2(tid, it); }
        else { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(tid) : "memory"); }
    }
    if (acc == 12345.678f) out[0] = acc;
}
template <int MODE> static float run(float* buf, uint32_t rows, int blocks, int per, float* out) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a)); k<MODE><<<blocks, 256>>>(buf, rows, per, 17u + r, out); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("{\"error\": \"needs 2 GPUs\"}\n"); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    const size_t rows = 32u << 20;  // 1 GiB of 32-byte rows
    float *remote, *local, *out;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&remote, rows * 32)); CK(cudaMemset(remote, 0, rows * 32));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, rows * 32)); CK(cudaMemset(local, 0, rows * 32)); CK(cudaMalloc(&out, 256));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"can_access_peer\": %d", can);
    const char* names[5] = {"ld_cg_32B", "ld_relaxed_sys_32B", "ld_cg_4B", "st_8B", "st_relaxed_sys_4B"};
    for (int where = 0; where < 2; ++where) {
        float* buf = where ? remote : local;
        for (int blocks : {sms * 3, sms * 8}) {
            const int per = 32;
            const double ops = (double)blocks * 256 * per;
            float ms[5];
            ms[0] = run<0>(buf, (uint32_t)rows, blocks, per, out);
            ms[1] = run<1>(buf, (uint32_t)rows, blocks, per, out);
            ms[2] = run<2>(buf, (uint32_t)rows, blocks, per, out);
            ms[3] = run<3>(buf, (uint32_t)rows, blocks, per, out);
            ms[4] = run<4>(buf, (uint32_t)rows, blocks, per, out);
            for (int m = 0; m < 5; ++m) printf(", \"%s_%s_%dblocks_Gops\": %.3f", where ? "peer" : "local", names[m], blocks, ops / ms[m] / 1e6);
        }
    }
    printf("}\n");
    return 0;
}
