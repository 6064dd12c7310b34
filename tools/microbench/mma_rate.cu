// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, for SS (A from smem) and TS (A from TMEM).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../vit_plasticity_b200/csrc -I../../include mma_rate.cu -o mma_rate -lcuda
#include <cstdio>
#include "ptx.cuh"
using namespace vb;
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int ts, int b_mn, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        if (elect_one()) { mbar_init(&bar, 1); fence_barrier_init(); }
        __syncwarp();
        tmem_alloc(&tptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tptr;
    if (warp == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, 0, b_mn);
        const uint32_t alo = (smem_u32(smem) >> 4) | (1u << 16);
        const uint32_t blo = ((smem_u32(smem) + 32768) >> 4) | (b_mn ? ((8192u >> 4) << 16) : (1u << 16));
        long long t0 = 0, t1 = 0, t2 = 0;
        for (int rep = 0; rep < 3; ++rep) {
            t0 = clock64();
            if (elect_one()) {
                for (int r = 0; r < reps; ++r) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t bk = b_mn ? k * 128 : 2 * k;
                        if (ts) umma_bf16_ts(tb, tb + 256 + k * 8, make_desc(blo + bk, DESC_HI), idesc, 1);
                        else umma_bf16_ss(tb, make_desc(alo + 2 * k, DESC_HI), make_desc(blo + bk, DESC_HI), idesc, 1);
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
            t1 = clock64();
            mbar_wait(&bar, rep & 1, 1);
            t2 = clock64();
        }
        if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 64;
    for (int ts = 0; ts < 2; ++ts)
        for (int b_mn = 0; b_mn < 2; ++b_mn)
            for (int N : {16, 32, 64, 96, 128, 208, 256}) {
                rate_kernel<<<1, 128, 100 * 1024>>>(N, ts, b_mn, reps, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                printf("%s B_%s N=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal %.1f)\n", ts ? "TS" : "SS", b_mn ? "MN" : "K ", N,
                       out[0] / (4.0 * reps), out[1] / (4.0 * reps), 128.0 * N / 256);
            }
    return 0;
}
