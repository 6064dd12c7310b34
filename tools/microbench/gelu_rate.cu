// Microbenchmark: cost of the fc1 epilogue math (bias + exact-erf GELU + GELU' + bf16 packing) per 32-element chunk per
// warp, at the GEMM epilogue's occupancy (8 math warps per SM), scalar fp32 against packed f32x2 (FFMA2) arithmetic, plus
// raw FFMA / FFMA2 issue rates. Prints cycles per chunk per warp and the implied time for one fc1 tile (128 x 256).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../vit_plasticity_b200/csrc -I../../include gelu_rate.cu -o gelu_rate
#include <cstdio>
#include "ptx.cuh"
using namespace vb;

template <int MODE>
__global__ void __launch_bounds__(512, 1) gelu_kernel(const float* __restrict__ in, uint32_t* out, int reps, long long* cyc) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = in[(threadIdx.x & 255) * 32 + j];
    uint32_t acc = 0, acc2 = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t o[16], o2[16];
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float g0, d0, g1, d1;
                gelu_and_grad_erf(f[2 * j], g0, d0);
                gelu_and_grad_erf(f[2 * j + 1], g1, d1);
                o[j] = pack_bf16x2(g0, g1);
                o2[j] = pack_bf16x2(d0, d1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) gelu_and_grad_erf_x2(f[2 * j], f[2 * j + 1], o[j], o2[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            acc ^= o[j];
            acc2 += o2[j];
            f[2 * j] += __uint_as_float((acc & 0x3ff) | 0x30000000);  // keep the loop from being hoisted
            f[2 * j + 1] -= __uint_as_float((acc2 & 0x3ff) | 0x30000000);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ acc2;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) fma_kernel(float* out, int reps, long long* cyc) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = threadIdx.x * 0.001f + j;
    const float b = out[0], c = out[1];
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], b, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = fmaf(a[j], 1.0001f, 0.5f);
        } else {
            const uint64_t bb = pack_f32x2(b, b), cc = pack_f32x2(c, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint64_t v = pack_f32x2(a[2 * j], a[2 * j + 1]);
                v = fma_f32x2(v, bb, cc);
                unpack_f32x2(v, a[2 * j], a[2 * j + 1]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
    out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* in;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&in, 256 * 32 * 4);
    cudaMalloc(&out, (148 * 512 + 2) * 4);
    cudaMalloc(&cyc, 148 * 8);
    float h[256 * 32];
    for (int i = 0; i < 256 * 32; ++i) h[i] = ((i * 2654435761u) % 8000) / 1000.f - 4.f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    float bc[2] = {0.999f, 0.001f};
    cudaMemcpy(out, bc, 8, cudaMemcpyHostToDevice);
    const int reps = 2000;
    long long c[148];
    auto report = [&](const char* name, double per) {
        cudaDeviceSynchronize();
        cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
        printf("%-44s %8.1f cycles per %s (err %s)\n", name, (double)mx / reps, per > 0 ? "chunk of 32 elements per warp, 8 warps/SM" : "16 FMA per thread, 8 warps/SM",
               cudaGetErrorString(cudaGetLastError()));
        return (double)mx / reps;
    };
    for (int it = 0; it < 2; ++it) {
        for (int nt = 128; nt <= 512; nt *= 2) {
            gelu_kernel<0><<<148, nt>>>(in, out, reps, cyc);
            double a = report("gelu+gelu' scalar fp32", 1);
            gelu_kernel<1><<<148, nt>>>(in, out, reps, cyc);
            double b = report("gelu+gelu' packed f32x2", 1);
            printf("  -> %d warps/SM: math of one 128x256 tile (32 chunks): scalar %.0f clk, packed %.0f clk\n", nt / 32, a * 32 / (nt / 32), b * 32 / (nt / 32));
        }
        gelu_kernel<0><<<148, 256>>>(in, out, reps, cyc);
        double a = report("gelu+gelu' scalar fp32", 1);
        gelu_kernel<1><<<148, 256>>>(in, out, reps, cyc);
        double b = report("gelu+gelu' packed f32x2", 1);
        // one 128 x 256 tile = 4 chunks per warp on 8 warps
        printf("  -> epilogue math per fc1 tile: scalar %.0f clk, packed %.0f clk (main loop of K=768: 48 MMAs x 128..172 clk = 6144..8256 clk)\n", 4 * a, 4 * b);
        fma_kernel<0><<<148, 256>>>((float*)out, reps, cyc);
        report("FFMA reg,reg,reg x16", 0);
        fma_kernel<1><<<148, 256>>>((float*)out, reps, cyc);
        report("FFMA reg,imm,imm x16", 0);
        fma_kernel<2><<<148, 256>>>((float*)out, reps, cyc);
        report("FFMA2 x8 (same 16 FMAs)", 0);
    }
    return 0;
}
