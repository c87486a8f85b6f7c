// Microbenchmark: how many shared-memory wavefronts does a 64-/128-bit LDS cost when lanes of a warp read DUPLICATE
// addresses?  The frontend's two half-warps (one frame each) read the same twiddles: 16 distinct 16-byte entries per
// LDS.128.  Patterns (address of lane i, in units of the access size):
//   0  i            all distinct                       (512 B per LDS.128: 4 wavefronts by data volume)
//   1  i & 15       lanes i and i + 16 share           (what ww_mfcc.cuh does today)
//   2  i >> 1       lanes 2j and 2j + 1 share          (half-warps interleaved)
//   3  0            one address for the whole warp
//   4  i & 7        four lanes share, spread over the quarters
//   5  i >> 2       four adjacent lanes share
// Prints SM clocks per warp-level LDS at saturation (16 warps / SM, 8 independent loads in flight per warp).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_dup lds_dup.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int BYTES>
__device__ __forceinline__ float lds(unsigned addr) {
    if constexpr (BYTES == 16) {
        float a, b, c, d;
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
        return a + b + c + d;
    } else if constexpr (BYTES == 8) {
        float a, b;
        asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr) : "memory");
        return a + b;
    } else {
        float a;
        asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(a) : "r"(addr) : "memory");
        return a;
    }
}

template <int BYTES>
__global__ void __launch_bounds__(256) kern(float* out, long long* cyc, int iters, int pattern) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 0.5f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = lane;
    if (pattern == 1) idx = lane & 15;
    if (pattern == 2) idx = lane >> 1;
    if (pattern == 3) idx = 0;
    if (pattern == 4) idx = lane & 7;
    if (pattern == 5) idx = lane >> 2;
    unsigned base = (unsigned)__cvta_generic_to_shared(sm) + idx * BYTES;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += lds<BYTES>(base + ((it & 3) << 12) + u * 512);   // the address varies, or ptxas hoists the loads
    }
    const long long t1 = clock64();
    if (acc == 1234.5f) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int BYTES>
void run(float* d, long long* dc, int sms) {
    const int iters = 4096, blocks = sms * 2;
    for (int pattern = 0; pattern < 6; ++pattern) {
        kern<BYTES><<<blocks, 256, 20480>>>(d, dc, 64, pattern);
        cudaDeviceSynchronize();
        kern<BYTES><<<blocks, 256, 20480>>>(d, dc, iters, pattern);
        cudaDeviceSynchronize();
        long long h[4096];
        cudaMemcpy(h, dc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < blocks; ++i) mean += (double)h[i];
        mean /= blocks;
        // 16 warps per SM each issue iters * 8 loads in `mean` clocks
        printf("LDS.%-3d pattern %d: %.2f clocks per warp-level load (per SM)\n", BYTES * 8, pattern, mean / (16.0 * iters * 8));
    }
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* d;
    long long* dc;
    cudaMalloc(&d, 64);
    cudaMalloc(&dc, sizeof(long long) * 4096);
    run<4>(d, dc, p.multiProcessorCount);
    run<8>(d, dc, p.multiProcessorCount);
    run<16>(d, dc, p.multiProcessorCount);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
