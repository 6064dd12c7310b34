// Microbenchmark: variants of the fc1 epilogue math (exact-erf GELU + GELU' of fp32 pre-activations, packed to bf16), at the
// GEMM epilogue's occupancy (8 math warps per SM). The MUFU pipe (16 lanes/clk/SM) is the tightest pipe of the shipped
// version (one rcp + one ex2 per element = 1024 of the 1658 cycles a 32 x 32 chunk takes per scheduler); the variants share
// ONE reciprocal between 2 / 4 elements (1 / (u0 u1) times the other factor), and optionally use the 3-term form of
// Abramowitz-Stegun's erfc (7.1.25, |err| <= 2.5e-5) instead of the 5-term one (7.1.26, 1.5e-7).
// Prints cycles per chunk per warp and the worst absolute error of gelu / gelu' against double precision on [-9, 9].
// build: nvcc -cudart shared -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../vit_plasticity_b200/csrc -I../../include gelu_variants.cu -o gelu_variants
#include <cmath>
#include <cstdio>
#include <vector>

#include "ptx.cuh"
using namespace vb;

// ---- the variants under test (not part of the library: none of them was adopted, see the results at the end of this file) ----
// gelu / gelu' of FOUR pre-activations with ONE MUFU.RCP: t_i = 1 / u_i, u_i = 1 + p |z_i| / sqrt 2 in [1, 3.5], is formed as
// (1 / (u0 u1 u2 u3)) times the product of the other three factors (9 FMA-pipe multiplies instead of 3 reciprocals; the
// product stays below 150 and every t_i keeps ~2 ulp). The MUFU pipe (16 lanes / clk / SM) is the tightest pipe of this
// epilogue: 2 -> 1.25 MUFU operations per element. THREE = true swaps the 5-term erfc polynomial (A&S 7.1.26, 1.5e-7) for
// the 3-term one (7.1.25, |err| <= 2.5e-5, still two orders below the bf16 rounding of the outputs); measurement only.
template <bool THREE>
__device__ __forceinline__ void gelu_and_grad_erf_x4(const float (&zz)[4], uint32_t (&g_bf16x2)[2], uint32_t (&dg_bf16x2)[2]) {
    constexpr float P = THREE ? 0.47047f * 0.70710678118654752f : 0.3275911f * 0.70710678118654752f;
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = fmaf(fabsf(zz[i]), P, 1.0f);
    const float p01 = u[0] * u[1], p23 = u[2] * u[3];
    const float r = fast_rcp(p01 * p23);
    const float r01 = r * p23, r23 = r * p01;  // 1 / (u0 u1), 1 / (u2 u3)
    const uint64_t tt[2] = {mul_f32x2(splat_f32x2(r01), pack_f32x2(u[1], u[0])), mul_f32x2(splat_f32x2(r23), pack_f32x2(u[3], u[2]))};
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        const float z0 = zz[2 * h2], z1 = zz[2 * h2 + 1];
        const uint64_t z = pack_f32x2(z0, z1);
        const uint64_t naz = pack_f32x2(__uint_as_float(__float_as_uint(z0) | 0x80000000u), __uint_as_float(__float_as_uint(z1) | 0x80000000u));
        const uint64_t t = tt[h2];
        const uint64_t w = mul_f32x2(mul_f32x2(z, splat_f32x2(-0.72134752044448170f)), z);
        float w0, w1;
        unpack_f32x2(w, w0, w1);
        const uint64_t e = pack_f32x2(fast_ex2(w0), fast_ex2(w1));  // exp(-z^2 / 2)
        uint64_t poly;
        if (THREE) {
            poly = fma_f32x2(splat_f32x2(0.5f * 0.7478556f), t, splat_f32x2(0.5f * -0.0958798f));
            poly = fma_f32x2(poly, t, splat_f32x2(0.5f * 0.3480242f));
        } else {
            poly = fma_f32x2(splat_f32x2(0.5f * 1.061405429f), t, splat_f32x2(0.5f * -1.453152027f));
            poly = fma_f32x2(poly, t, splat_f32x2(0.5f * 1.421413741f));
            poly = fma_f32x2(poly, t, splat_f32x2(0.5f * -0.284496736f));
            poly = fma_f32x2(poly, t, splat_f32x2(0.5f * 0.254829592f));
        }
        const uint64_t h = mul_f32x2(mul_f32x2(poly, t), e);
        const uint64_t relu = pack_f32x2(fmaxf(z0, 0.f), fmaxf(z1, 0.f));
        const uint64_t g = fma_f32x2(naz, h, relu);
        const uint64_t q = fma_f32x2(h, splat_f32x2(-1.0f), splat_f32x2(0.5f));
        const uint64_t sgn = pack_f32x2(__uint_as_float((__float_as_uint(z0) & 0x80000000u) | 0x3f800000u),
                                        __uint_as_float((__float_as_uint(z1) & 0x80000000u) | 0x3f800000u));
        const uint64_t cdf = fma_f32x2(sgn, q, splat_f32x2(0.5f));
        const uint64_t dg = fma_f32x2(mul_f32x2(z, splat_f32x2(0.3989422804014327f)), e, cdf);
        float a0, a1;
        unpack_f32x2(g, a0, a1);
        g_bf16x2[h2] = pack_bf16x2(a0, a1);
        unpack_f32x2(dg, a0, a1);
        dg_bf16x2[h2] = pack_bf16x2(a0, a1);
    }
}

// Pair version with one reciprocal: t0 = u1 / (u0 u1), t1 = u0 / (u0 u1) (microbenchmark comparison)
__device__ __forceinline__ void gelu_and_grad_erf_x2_rcp2(float z0, float z1, uint32_t& g_bf16x2, uint32_t& dg_bf16x2) {
    const uint64_t z = pack_f32x2(z0, z1);
    const uint64_t naz = pack_f32x2(__uint_as_float(__float_as_uint(z0) | 0x80000000u), __uint_as_float(__float_as_uint(z1) | 0x80000000u));
    const float u0 = fmaf(fabsf(z0), 0.3275911f * 0.70710678118654752f, 1.0f), u1 = fmaf(fabsf(z1), 0.3275911f * 0.70710678118654752f, 1.0f);
    const float r = fast_rcp(u0 * u1);
    const uint64_t t = mul_f32x2(splat_f32x2(r), pack_f32x2(u1, u0));
    const uint64_t w = mul_f32x2(mul_f32x2(z, splat_f32x2(-0.72134752044448170f)), z);
    float w0, w1;
    unpack_f32x2(w, w0, w1);
    const uint64_t e = pack_f32x2(fast_ex2(w0), fast_ex2(w1));
    uint64_t poly = fma_f32x2(splat_f32x2(0.5f * 1.061405429f), t, splat_f32x2(0.5f * -1.453152027f));
    poly = fma_f32x2(poly, t, splat_f32x2(0.5f * 1.421413741f));
    poly = fma_f32x2(poly, t, splat_f32x2(0.5f * -0.284496736f));
    poly = fma_f32x2(poly, t, splat_f32x2(0.5f * 0.254829592f));
    const uint64_t h = mul_f32x2(mul_f32x2(poly, t), e);
    const uint64_t relu = pack_f32x2(fmaxf(z0, 0.f), fmaxf(z1, 0.f));
    const uint64_t g = fma_f32x2(naz, h, relu);
    const uint64_t q = fma_f32x2(h, splat_f32x2(-1.0f), splat_f32x2(0.5f));
    const uint64_t sgn = pack_f32x2(__uint_as_float((__float_as_uint(z0) & 0x80000000u) | 0x3f800000u),
                                    __uint_as_float((__float_as_uint(z1) & 0x80000000u) | 0x3f800000u));
    const uint64_t cdf = fma_f32x2(sgn, q, splat_f32x2(0.5f));
    const uint64_t dg = fma_f32x2(mul_f32x2(z, splat_f32x2(0.3989422804014327f)), e, cdf);
    float a0, a1;
    unpack_f32x2(g, a0, a1);
    g_bf16x2 = pack_bf16x2(a0, a1);
    unpack_f32x2(dg, a0, a1);
    dg_bf16x2 = pack_bf16x2(a0, a1);
}

// ---- end of variants ----

// V: 0 = shipped pair version; 1 = reciprocal shared by a pair; 2 = shared by four; 3 = four + 3-term polynomial
template <int V>
__device__ __forceinline__ void gelu4(const float (&z)[4], uint32_t (&g)[2], uint32_t (&dg)[2]) {
    if (V == 0) {
        gelu_and_grad_erf_x2(z[0], z[1], g[0], dg[0]);
        gelu_and_grad_erf_x2(z[2], z[3], g[1], dg[1]);
        return;
    }
    if (V == 1) {
        gelu_and_grad_erf_x2_rcp2(z[0], z[1], g[0], dg[0]);
        gelu_and_grad_erf_x2_rcp2(z[2], z[3], g[1], dg[1]);
        return;
    }
    gelu_and_grad_erf_x4<V == 3>(z, g, dg);
}

template <int V>
__global__ void __launch_bounds__(512, 1) gelu_kernel(const float* __restrict__ in, uint32_t* out, int reps, long long* cyc) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = in[(threadIdx.x & 255) * 32 + j];
    uint32_t acc = 0, acc2 = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        uint32_t o[16], o2[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z[4] = {f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]};
            uint32_t g[2], d[2];
            gelu4<V>(z, g, d);
            o[2 * j] = g[0], o[2 * j + 1] = g[1], o2[2 * j] = d[0], o2[2 * j + 1] = d[1];
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

// accuracy: fp32 results before the bf16 rounding are not exposed, so compare the bf16 outputs with the bf16 rounding of the
// double-precision values and report the worst error in units of the bf16 spacing at that value, plus plain absolute error
template <int V>
__global__ void acc_kernel(const float* __restrict__ zs, uint32_t* g, uint32_t* dg, int n4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float z[4] = {zs[4 * i], zs[4 * i + 1], zs[4 * i + 2], zs[4 * i + 3]};
    uint32_t a[2], b[2];
    gelu4<V>(z, a, b);
    g[2 * i] = a[0], g[2 * i + 1] = a[1], dg[2 * i] = b[0], dg[2 * i + 1] = b[1];
}

static float bf16_to_float(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

template <int V>
static void run(const char* name, const float* in, uint32_t* out, long long* cyc, const float* dz, uint32_t* dgv, uint32_t* ddv, const std::vector<float>& hz) {
    const int reps = 2000;
    long long c[148];
    gelu_kernel<V><<<148, 256>>>(in, out, reps, cyc);
    cudaDeviceSynchronize();
    gelu_kernel<V><<<148, 256>>>(in, out, reps, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = c[i] > mx ? c[i] : mx;
    const int n = (int)hz.size();
    acc_kernel<V><<<(n / 4 + 255) / 256, 256>>>(dz, dgv, ddv, n / 4);
    std::vector<uint32_t> hg(n / 2), hd(n / 2);
    cudaMemcpy(hg.data(), dgv, n * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(hd.data(), ddv, n * 2, cudaMemcpyDeviceToHost);
    double eg = 0, ed = 0, ug = 0, ud = 0;
    for (int i = 0; i < n; ++i) {
        const double z = hz[i], cdf = 0.5 * erfc(-z / sqrt(2.0)), pdf = exp(-0.5 * z * z) / sqrt(2.0 * M_PI);
        const double rg = z * cdf, rd = cdf + z * pdf;
        const float gg = bf16_to_float((uint16_t)(hg[i / 2] >> (16 * (i & 1)))), dd = bf16_to_float((uint16_t)(hd[i / 2] >> (16 * (i & 1))));
        const double sg = fmax(fabs(rg), 1e-30) * 0.0078125, sd = fmax(fabs(rd), 1e-30) * 0.0078125;  // ~ one bf16 ulp (2^-7 relative)
        eg = fmax(eg, fabs(gg - rg)), ed = fmax(ed, fabs(dd - rd));
        if (fabs(rg) > 1e-20) ug = fmax(ug, fabs(gg - rg) / sg);
        if (fabs(rd) > 1e-20) ud = fmax(ud, fabs(dd - rd) / sd);
    }
    printf("%-38s %8.1f cycles per 32x32 chunk per warp (tile: %6.0f)   worst |err| gelu %.2e gelu' %.2e   worst err in bf16 ulps: gelu %.2f gelu' %.2f  (%s)\n",
           name, (double)mx / reps, 4.0 * mx / reps, eg, ed, ug, ud, cudaGetErrorString(cudaGetLastError()));
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
    std::vector<float> hz;
    for (int i = 0; i < 72000; ++i) hz.push_back(-9.f + i * 0.00025f);  // [-9, 9) in steps of 2.5e-4
    for (int e = -60; e < 0; ++e) {                                       // tiny magnitudes of both signs
        hz.push_back(ldexpf(1.37f, e));
        hz.push_back(-ldexpf(1.37f, e));
        hz.push_back(0.f);
        hz.push_back(-0.f);
    }
    float* dz;
    uint32_t *dgv, *ddv;
    cudaMalloc(&dz, hz.size() * 4);
    cudaMalloc(&dgv, hz.size() * 2);
    cudaMalloc(&ddv, hz.size() * 2);
    cudaMemcpy(dz, hz.data(), hz.size() * 4, cudaMemcpyHostToDevice);
    for (int it = 0; it < 2; ++it) {
        run<0>("shipped: rcp + ex2 per element", in, out, cyc, dz, dgv, ddv, hz);
        run<1>("one rcp per pair", in, out, cyc, dz, dgv, ddv, hz);
        run<2>("one rcp per four", in, out, cyc, dz, dgv, ddv, hz);
        run<3>("one rcp per four, 3-term erfc", in, out, cyc, dz, dgv, ddv, hz);
    }
    return 0;
}

// Measured on B200 (profiles/r04_c_gelu_variants.log), cycles per 32 x 32 chunk per warp at 8 warps / SM:
//   shipped 1833 | one rcp per pair 1799 | one rcp per four 1827 | one rcp per four + 3-term erfc 1684
// Removing 3 of every 8 MUFU operations changes nothing: the epilogue math is not MUFU-bound. It is bound by issue slots
// (a packed FFMA2 takes two): ~40 slots per pair of elements, 2 warps per scheduler. The 3-term polynomial (-4 slots per
// pair) is 8 % faster but its absolute error of 2.5e-5 is up to 29 bf16 ulps of gelu' in the negative tail: rejected.
