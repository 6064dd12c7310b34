// Issue rate of the instructions the attention / GELU epilogue math is made of, per SM: FFMA, MUFU.EX2, F2FP (cvt.rn.bf16x2.f32),
// and a hand-rolled integer round-to-nearest-even pack (LOP3 / IADD3 / PRMT). One CTA per SM, W warps, a dependent-free stream
// of the instruction in registers. Build: make -C tools/microbench pipe_rate ; run on a B200: tools/microbench/pipe_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    float a[8], b = threadIdx.x * 1e-3f;
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = b + i, u[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], 1.0001f, 0.5f);
            if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 2) {  // the packed result is fed back as the next input: a dependent chain per i, 8 chains per thread
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b));
                a[i] = __uint_as_float(u[i]);
            }
            if (MODE == 3) {  // integer RNE pack of (a[i], a[i+1]) -> bf16x2
                uint32_t x = __float_as_uint(a[i]) + u[i], y = __float_as_uint(a[(i + 1) & 7]);
                x += 0x7fffu + ((x >> 16) & 1u);
                y += 0x7fffu + ((y >> 16) & 1u);
                u[i] = __byte_perm(x, y, 0x7632);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps, double per_iter) {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 20000;
    cudaEvent_t s, e;
    cudaEventCreate(&s);
    cudaEventCreate(&e);
    k<MODE><<<148, warps * 32>>>(out, 100);
    cudaEventRecord(s);
    k<MODE><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e);
    cudaEventSynchronize(e);
    float ms;
    cudaEventElapsedTime(&ms, s, e);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ops = (double)iters * 8 * warps * 32 * per_iter;  // per SM
    printf("%-28s %2d warps: %7.3f ms  -> %6.1f thread-ops / ns / SM (at %.2f GHz nominal: %5.1f per clk)\n", name, warps, ms, ops / (ms * 1e6),
           clk / 1e6, ops / (ms * 1e6) / (clk / 1e6));
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("FFMA", w, 1);
        run<1>("MUFU.EX2", w, 1);
        run<2>("F2FP.BF16 pack (per pair)", w, 1);
        run<3>("int RNE pack (per pair)", w, 1);
    }
    return 0;
}
