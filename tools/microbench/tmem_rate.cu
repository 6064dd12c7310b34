// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld (32x32b.x32 = 4 KB per warp-instruction) per SM, as a function of
// the number of warps reading (each warp reads its own lane quarter, all 512 columns round-robin), and of tcgen05.st.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../vit_plasticity_b200/csrc -I../../include tmem_rate.cu -o tmem_rate
#include <cstdio>
#define VB_MBAR_TRAP_PRINTF 0
#include "ptx.cuh"
using namespace vb;

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

template <int MODE>  // 0: ld + wait each; 1: two lds in flight per wait; 2: st
__global__ void __launch_bounds__(512, 1) k(int reps, long long* cyc, uint32_t* sink) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        tmem_alloc(&tptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t a[32], b[32], acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = b[i] = threadIdx.x + i;
    tmem_st_x32(base, a);
    tmem_st_wait();
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const uint32_t col = ((r * 2 + (warp >> 2)) * 32) & 511;
        if (MODE == 0) {
            tmem_ld_32x32b_x32(base + col, a);
            tmem_ld_wait_x32(a);
            acc += a[0] ^ a[31];
        } else if (MODE == 1) {
            tmem_ld_32x32b_x32(base + col, a);
            tmem_ld_32x32b_x32(base + ((col + 256) & 511), b);
            tmem_ld_wait_x32(a);
            tmem_ld_wait_x32(b);
            acc += a[0] ^ a[31] ^ b[0] ^ b[31];
        } else {
            a[0] += r;
            tmem_st_x32(base + col, a);
            tmem_st_wait();
        }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tptr, 512);
    }
}

int main() {
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&sink, 148 * 512 * 4);
    const int reps = 4000;
    long long c[148];
    for (int mode = 0; mode < 3; ++mode)
        for (int nw = 4; nw <= 16; nw *= 2) {
            if (mode == 0) k<0><<<148, nw * 32>>>(reps, cyc, sink);
            if (mode == 1) k<1><<<148, nw * 32>>>(reps, cyc, sink);
            if (mode == 2) k<2><<<148, nw * 32>>>(reps, cyc, sink);
            cudaDeviceSynchronize();
            cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
            const double bytes = (double)reps * nw * 4096.0 * (mode == 1 ? 2 : 1);
            printf("%-34s %2d warps/SM: %7.1f B/clk/SM  (%6.1f clk per 4 KB warp-instruction)  [%s]\n",
                   mode == 0 ? "tcgen05.ld x32, wait each" : mode == 1 ? "tcgen05.ld x32, two in flight" : "tcgen05.st x32, wait each", nw,
                   bytes / mx, (double)mx / reps / (mode == 1 ? 2 : 1), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
