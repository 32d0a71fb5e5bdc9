// Batched mixed-radix encode / decode of MultiDiscrete vectors (reference: utils.py:12-115, used by the flatten
// wrappers, wrappers/flatten_multidiscrete_wrapper.py:61-76, 139-161).  Integer-only, HBM-bound: a block stages a
// tile of 256 vectors through shared memory so that both the [N][D] side and the [N] side move as coalesced rows.
#pragma once
#include <cstdint>

namespace qe {

constexpr int kRadixMaxDims = 32;
struct RadixSpec {
    int dims;
    long long radix[kRadixMaxDims];  // compute_radix(nvec): radix[d] = prod(nvec[d+1:])
    long long nvec[kRadixMaxDims];
};

// out[i] = sum_d vectors[i][d] * radix[d]          (encode_multi_discretes, utils.py:51-69)
static __global__ void __launch_bounds__(256) radix_encode_kernel(const int32_t* __restrict__ vectors, RadixSpec R, long long* __restrict__ out, long long n) {
    extern __shared__ int32_t s_tile[];  // [256][dims], padded to an odd stride
    const int D = R.dims, ld = D | 1;
    for (long long base = (long long)blockIdx.x * 256; base < n; base += (long long)gridDim.x * 256) {
        const int rows = (int)min(256ll, n - base);
        const int32_t* src = vectors + base * D;
        for (int j = threadIdx.x; j < rows * D; j += 256) s_tile[(j / D) * ld + (j % D)] = src[j];
        __syncthreads();
        if ((int)threadIdx.x < rows) {
            long long acc = 0;
            for (int d = 0; d < D; ++d) acc += (long long)s_tile[threadIdx.x * ld + d] * R.radix[d];
            out[base + threadIdx.x] = acc;
        }
        __syncthreads();
    }
}

// out[i][d] = (indices[i] // radix[d]) % nvec[d]    (decode_to_multi_discretes, utils.py:95-115; floor semantics)
static __global__ void __launch_bounds__(256) radix_decode_kernel(const long long* __restrict__ indices, RadixSpec R, int32_t* __restrict__ out, long long n) {
    extern __shared__ int32_t s_tile[];
    const int D = R.dims, ld = D | 1;
    for (long long base = (long long)blockIdx.x * 256; base < n; base += (long long)gridDim.x * 256) {
        const int rows = (int)min(256ll, n - base);
        if ((int)threadIdx.x < rows) {
            const long long idx = indices[base + threadIdx.x];
            for (int d = 0; d < D; ++d) {
                const long long r = R.radix[d], m = R.nvec[d];
                long long q = idx / r;
                if ((idx % r != 0) && ((idx < 0) != (r < 0))) --q;  // floor division, like NumPy
                long long v = q % m;
                if (v != 0 && ((v < 0) != (m < 0))) v += m;         // sign of the divisor, like NumPy
                s_tile[threadIdx.x * ld + d] = (int32_t)v;
            }
        }
        __syncthreads();
        int32_t* dst = out + base * D;
        for (int j = threadIdx.x; j < rows * D; j += 256) dst[j] = s_tile[(j / D) * ld + (j % D)];
        __syncthreads();
    }
}

}  // namespace qe
