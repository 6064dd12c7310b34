// HBM-bound data-movement kernels either side of the GEMMs: casts, im2col for the patch embedding,
// token assembly (cls + patches + positional embedding), column sums for bias gradients, and the
// per-sample squared-distance reduction of the plasticity estimator.
#include "host_utils.h"
#include "ptx.cuh"

namespace vb {

constexpr int EW_THREADS = 256;

__global__ void __launch_bounds__(EW_THREADS)
cast_f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t n) {
    const int64_t n8 = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
        uint4 u;
        u.x = pack_bf16x2(a.x, a.y);
        u.y = pack_bf16x2(a.z, a.w);
        u.z = pack_bf16x2(b.x, b.y);
        u.w = pack_bf16x2(b.z, b.w);
        reinterpret_cast<uint4*>(dst)[i] = u;
    }
    // tail
    const int64_t t0 = n8 << 3;
    for (int64_t i = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void __launch_bounds__(EW_THREADS)
cast_bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n) {
    const int64_t n8 = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
        reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
    const int64_t t0 = n8 << 3;
    for (int64_t i = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __bfloat162float(src[i]);
}

__global__ void __launch_bounds__(EW_THREADS)
add_bf16_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ y, int64_t n) {
    const int64_t n8 = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a) + i);
        const uint4 ub = __ldg(reinterpret_cast<const uint4*>(b) + i);
        const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack_bf16x2(wa[j]), fb = unpack_bf16x2(wb[j]);
            o[j] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
        }
        reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    const int64_t t0 = n8 << 3;
    for (int64_t i = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

// One thread produces 8 consecutive patch columns (same c, py; px..px+7) = two float4 loads, one 16 B store.
// Requires P % 8 == 0.
__global__ void __launch_bounds__(EW_THREADS)
im2col_kernel(const float* __restrict__ img, const float* __restrict__ img2, bf16* __restrict__ patches, int n, int c,
              int h, int w, int p) {
    const int gh = h / p, gw = w / p;
    const int kdim = c * p * p;
    const int64_t total = (int64_t)n * gh * gw * (kdim >> 3);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int p8 = p >> 3;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int64_t t = idx;
        const int px8 = (int)(t % p8); t /= p8;
        const int py = (int)(t % p); t /= p;
        const int ch = (int)(t % c); t /= c;
        const int gx = (int)(t % gw); t /= gw;
        const int gy = (int)(t % gh); t /= gh;
        const int b = (int)t;
        const int64_t src = (((int64_t)b * c + ch) * h + (gy * p + py)) * w + gx * p + px8 * 8;
        float4 a0 = __ldg(reinterpret_cast<const float4*>(img + src));
        float4 a1 = __ldg(reinterpret_cast<const float4*>(img + src) + 1);
        if (img2 != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(img2 + src));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(img2 + src) + 1);
            a0.x -= b0.x; a0.y -= b0.y; a0.z -= b0.z; a0.w -= b0.w;
            a1.x -= b1.x; a1.y -= b1.y; a1.z -= b1.z; a1.w -= b1.w;
        }
        const int64_t rowi = ((int64_t)b * gh + gy) * gw + gx;
        const int col = (ch * p + py) * p + px8 * 8;
        uint4 u;
        u.x = pack_bf16x2(a0.x, a0.y);
        u.y = pack_bf16x2(a0.z, a0.w);
        u.z = pack_bf16x2(a1.x, a1.y);
        u.w = pack_bf16x2(a1.z, a1.w);
        *reinterpret_cast<uint4*>(patches + rowi * kdim + col) = u;
    }
}

// tokens[b, 0] = cls + pos[0]; tokens[b, 1 + i] = patch_out[b * np + i] + pos[1 + i]
__global__ void __launch_bounds__(EW_THREADS)
assemble_tokens_kernel(const bf16* __restrict__ patch_out, const float* __restrict__ patch_out_f32,
                       const float* __restrict__ cls, const float* __restrict__ pos, bf16* __restrict__ tokens,
                       float* __restrict__ tokens_f32, int batch, int np, int e) {
    const int e8 = e >> 3;
    const int64_t total = (int64_t)batch * (np + 1) * e8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int c8 = (int)(idx % e8);
        const int64_t r = idx / e8;
        const int l = (int)(r % (np + 1));
        const int b = (int)(r / (np + 1));
        float v[8];
        if (l == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(cls) + 2 * c8);
            const float4 bq = __ldg(reinterpret_cast<const float4*>(cls) + 2 * c8 + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
        } else {
            const int64_t prow = (int64_t)b * np + (l - 1);
            if (patch_out_f32 != nullptr) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(patch_out_f32 + prow * e) + 2 * c8);
                const float4 bq = __ldg(reinterpret_cast<const float4*>(patch_out_f32 + prow * e) + 2 * c8 + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = bq.x; v[5] = bq.y; v[6] = bq.z; v[7] = bq.w;
            } else {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(patch_out + prow * e) + c8);
                const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
            }
        }
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)l * e) + 2 * c8);
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)l * e) + 2 * c8 + 1);
        v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w; v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
        if (tokens != nullptr) {
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]);
            u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]);
            u.w = pack_bf16x2(v[6], v[7]);
            reinterpret_cast<uint4*>(tokens + r * e)[c8] = u;
        }
        if (tokens_f32 != nullptr) {
            reinterpret_cast<float4*>(tokens_f32 + r * e)[2 * c8] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(tokens_f32 + r * e)[2 * c8 + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

// grid: (ceil(e/8/32), chunks of batch). Each warp owns 256 consecutive columns? -> each thread 8 columns of one
// token position l, looping over a slice of the batch; dpos[l] += sum_b dtok[b,l]; dcls += sum_b dtok[b,0].
__global__ void __launch_bounds__(EW_THREADS)
assemble_tokens_bwd_kernel(const bf16* __restrict__ dtok, bf16* __restrict__ dpatch, float* __restrict__ dcls,
                           float* __restrict__ dpos, int batch, int np, int e, int b_per_block) {
    const int e8 = e >> 3;
    const int l = blockIdx.x;  // token position 0..np
    const int b0 = blockIdx.y * b_per_block;
    const int b1 = min(b0 + b_per_block, batch);
    for (int c8 = threadIdx.x; c8 < e8; c8 += blockDim.x) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int b = b0; b < b1; ++b) {
            const int64_t r = (int64_t)b * (np + 1) + l;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(dtok + r * e) + c8);
            if (l > 0) reinterpret_cast<uint4*>(dpatch + ((int64_t)b * np + (l - 1)) * e)[c8] = u;
            const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
            acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
        }
        if (dpos != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(dpos + (int64_t)l * e + c8 * 8 + j, acc[j]);
        }
        if (l == 0 && dcls != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(dcls + c8 * 8 + j, acc[j]);
        }
    }
}

// out[c] += sum_r x[r, c]; grid (ceil(cols/256), row_chunks); warp w of a block takes rows r0 + w, r0 + w + 8, ...
__global__ void __launch_bounds__(EW_THREADS)
colsum_bf16_kernel(const bf16* __restrict__ x, int64_t ldx, float* __restrict__ out, int rows, int cols,
                   int rows_per_block) {
    __shared__ float red[8][257];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 256 + lane * 8;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (col < cols) {
        auto add8 = [&](const uint4& u) {
            const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1.x; acc[3] += f1.y;
            acc[4] += f2.x; acc[5] += f2.y; acc[6] += f3.x; acc[7] += f3.y;
        };
        const bf16* base = x + col;
        int r = r0 + warp;
        for (; r + 24 < r1; r += 32) {  // four independent 128-bit loads in flight per thread
            const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)r * ldx));
            const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(r + 8) * ldx));
            const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(r + 16) * ldx));
            const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(r + 24) * ldx));
            add8(u0);
            add8(u1);
            add8(u2);
            add8(u3);
        }
        for (; r < r1; r += 8) add8(__ldg(reinterpret_cast<const uint4*>(base + (int64_t)r * ldx)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int t = threadIdx.x;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][t];
    const int c = blockIdx.x * 256 + t;
    if (c < cols) atomicAdd(out + c, s);
}

// dst = bf16(scale * src): the perturbation sweep re-uses the unit-noise projections W_qkv (dtok) for every magnitude
__global__ void __launch_bounds__(EW_THREADS)
scale_bf16_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int64_t n, float scale) {
    const int64_t n8 = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            o[j] = pack_bf16x2(f.x * scale, f.y * scale);
        }
        reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    const int64_t t0 = n8 << 3;
    for (int64_t i = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __float2bfloat16_rn(__bfloat162float(src[i]) * scale);
}

// Counter-based noise for the perturbation sweep: Philox4x32-10 (Salmon et al., SC'11) keyed by the seed, the IMAGE INDEX
// in the counter, so that the direction drawn for image i depends neither on the batching nor on the sharding over ranks.
// counter = (j, 0, image_lo, image_hi) -> r0..r3 -> Box-Muller -> elements 4j..4j+3 of the image (oracle/philox_oracle.py).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float philox_unit(uint32_t r) { return ((float)(r >> 9) + 0.5f) * 1.1920928955078125e-07f; }  // 2^-23
__global__ void __launch_bounds__(EW_THREADS)
philox_normal_kernel(float* __restrict__ out, int64_t quads_per_image, int64_t total_quads, uint64_t seed, uint64_t first_image) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += stride) {
        const uint64_t img = first_image + (uint64_t)(q / quads_per_image);
        const uint32_t j = (uint32_t)(q % quads_per_image);
        uint32_t c[4] = {j, 0u, (uint32_t)img, (uint32_t)(img >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float4 z;
        float sn, cs;
        float rad = sqrtf(-2.f * logf(philox_unit(c[0])));
        sincospif(2.f * philox_unit(c[1]), &sn, &cs);
        z.x = rad * cs; z.y = rad * sn;
        rad = sqrtf(-2.f * logf(philox_unit(c[2])));
        sincospif(2.f * philox_unit(c[3]), &sn, &cs);
        z.z = rad * cs; z.w = rad * sn;
        reinterpret_cast<float4*>(out)[q] = z;
    }
}

// out[s] += sum over sample s of (a - b)^2 ; block = (sample, slice)
__global__ void __launch_bounds__(EW_THREADS)
rowsumsq_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                     int64_t elems_per_sample) {
    __shared__ float red[EW_THREADS / 32];
    const int s = blockIdx.x;
    const int64_t n4 = elems_per_sample >> 2;
    const float4* ap = reinterpret_cast<const float4*>(a + (int64_t)s * elems_per_sample);
    const float4* bp = b ? reinterpret_cast<const float4*>(b + (int64_t)s * elems_per_sample) : nullptr;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.y * blockDim.x) {
        float4 va = __ldg(ap + i);
        if (bp) {
            const float4 vb4 = __ldg(bp + i);
            va.x -= vb4.x; va.y -= vb4.y; va.z -= vb4.z; va.w -= vb4.w;
        }
        acc += va.x * va.x + va.y * va.y + va.z * va.z + va.w * va.w;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < EW_THREADS / 32; ++w) t += red[w];
        atomicAdd(out + s, t);
    }
}

// pooled[n, :] = x[n, 0, :] (cls) or mean_l x[n, l, :], optionally divided by its L2 norm. One block per sample,
// thread t owns feature pairs t, t + 256, ...; the token loop reads 4 B per thread per token (coalesced rows).
__global__ void __launch_bounds__(EW_THREADS)
pool_tokens_kernel(const bf16* __restrict__ x, float* __restrict__ pooled, int seq, int dim, int cls_pooling, int normalize) {
    __shared__ float red[EW_THREADS / 32];
    const int n = blockIdx.x;
    const bf16* xs = x + (size_t)n * seq * dim;
    float* out = pooled + (size_t)n * dim;
    const int rows = cls_pooling ? 1 : seq;
    const float scale = 1.f / (float)rows;
    float ss = 0.f;
    for (int p = threadIdx.x; p < dim / 2; p += EW_THREADS) {
        float a0 = 0.f, a1 = 0.f;
        for (int l = 0; l < rows; ++l) {
            const float2 f = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(xs + (size_t)l * dim) + p));
            a0 += f.x;
            a1 += f.y;
        }
        a0 *= scale;
        a1 *= scale;
        out[2 * p] = a0;
        out[2 * p + 1] = a1;
        ss += a0 * a0 + a1 * a1;
    }
    if (!normalize) return;
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < EW_THREADS / 32; ++w) tot += red[w];
    const float inv = rsqrtf(tot);
    for (int p = threadIdx.x; p < dim / 2; p += EW_THREADS) {  // each thread rescales the values it wrote itself
        out[2 * p] *= inv;
        out[2 * p + 1] *= inv;
    }
}

static inline int ew_grid(int64_t work_items) {
    int64_t g = (work_items + EW_THREADS - 1) / EW_THREADS;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace vb

extern "C" int vb_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(src && dst && n >= 0, "vb_cast_f32_to_bf16: bad args");
    if (n == 0) return VB_OK;
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "vb_cast_f32_to_bf16: pointers must be 16B aligned");
    cast_f32_to_bf16_kernel<<<ew_grid(n / 8 + 1), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        src, static_cast<bf16*>(dst), n);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_cast_bf16_to_f32(const void* src, float* dst, int64_t n, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(src && dst && n >= 0, "vb_cast_bf16_to_f32: bad args");
    if (n == 0) return VB_OK;
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "vb_cast_bf16_to_f32: pointers must be 16B aligned");
    cast_bf16_to_f32_kernel<<<ew_grid(n / 8 + 1), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(src), dst, n);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_add_bf16(const void* a, const void* b, void* y, int64_t n, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(a && b && y && n >= 0, "vb_add_bf16: bad args");
    if (n == 0) return VB_OK;
    add_bf16_kernel<<<ew_grid(n / 8 + 1), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(y), n);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_im2col_patches(const float* img, const float* img2, void* patches, int32_t n, int32_t c, int32_t h,
                                 int32_t w, int32_t p, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(img && patches, "vb_im2col_patches: null pointer");
    VB_CHECK_ARG(n > 0 && c > 0 && p > 0 && p % 8 == 0 && h % p == 0 && w % p == 0,
                 "vb_im2col_patches: need P %% 8 == 0 and H, W divisible by P (n=%d c=%d h=%d w=%d p=%d)", n, c, h, w, p);
    const int64_t total = (int64_t)n * (h / p) * (w / p) * (c * p * p / 8);
    im2col_kernel<<<ew_grid(total), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        img, img2, static_cast<bf16*>(patches), n, c, h, w, p);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_assemble_tokens(const void* patch_out, const float* patch_out_f32, const float* cls, const float* pos,
                                  void* tokens, float* tokens_f32, int32_t batch, int32_t np, int32_t e,
                                  vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG((patch_out || patch_out_f32) && cls && pos && (tokens || tokens_f32), "vb_assemble_tokens: null pointer");
    VB_CHECK_ARG(batch > 0 && np > 0 && e % 8 == 0, "vb_assemble_tokens: e=%d must be a multiple of 8", e);
    const int64_t total = (int64_t)batch * (np + 1) * (e / 8);
    assemble_tokens_kernel<<<ew_grid(total), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(patch_out), patch_out_f32, cls, pos, static_cast<bf16*>(tokens), tokens_f32, batch, np, e);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_assemble_tokens_bwd(const void* dtokens, void* dpatch_out, float* dcls, float* dpos, int32_t batch,
                                      int32_t np, int32_t e, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(dtokens && dpatch_out, "vb_assemble_tokens_bwd: null pointer");
    VB_CHECK_ARG(batch > 0 && np > 0 && e % 8 == 0, "vb_assemble_tokens_bwd: e=%d must be a multiple of 8", e);
    const int b_per_block = 16;
    dim3 grid(np + 1, (batch + b_per_block - 1) / b_per_block);
    assemble_tokens_bwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(dtokens), static_cast<bf16*>(dpatch_out), dcls, dpos, batch, np, e, b_per_block);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

namespace vb {
int launch_colsum_bf16(const bf16* x, int64_t ldx, float* out, int rows, int cols, cudaStream_t stream) {
    return vb_colsum_bf16(x, ldx, out, rows, cols, static_cast<vb_stream_t>(stream));
}
}  // namespace vb

extern "C" int vb_colsum_bf16(const void* x, int64_t ldx, float* out, int32_t rows, int32_t cols, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(x && out, "vb_colsum_bf16: null pointer");
    VB_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0, "vb_colsum_bf16: cols and ldx must be multiples of 8");
    const int col_blocks = (cols + 255) / 256;
    int row_chunks = (num_sms() * 8 + col_blocks - 1) / col_blocks;
    if (row_chunks > (rows + 63) / 64) row_chunks = (rows + 63) / 64;
    if (row_chunks < 1) row_chunks = 1;
    const int rows_per_block = (rows + row_chunks - 1) / row_chunks;
    dim3 grid(col_blocks, (rows + rows_per_block - 1) / rows_per_block);
    colsum_bf16_kernel<<<grid, EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const bf16*>(x), ldx, out,
                                                                                   rows, cols, rows_per_block);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_pool_tokens(const void* x, float* pooled, int32_t n, int32_t seq, int32_t dim, int32_t cls_pooling,
                              int32_t normalize, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(x && pooled, "vb_pool_tokens: null pointer");
    VB_CHECK_ARG(n > 0 && seq > 0 && dim > 0 && dim % 2 == 0, "vb_pool_tokens: need n, seq > 0 and an even dim (n=%d seq=%d dim=%d)", n, seq, dim);
    pool_tokens_kernel<<<n, EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const bf16*>(x), pooled, seq, dim,
                                                                               cls_pooling, normalize);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_rowsumsq_diff_f32(const float* a, const float* b, float* out, int32_t n_samples,
                                    int32_t rows_per_sample, int32_t cols, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(a && out, "vb_rowsumsq_diff_f32: null pointer");
    const int64_t eps = (int64_t)rows_per_sample * cols;
    VB_CHECK_ARG(n_samples > 0 && eps > 0 && eps % 4 == 0, "vb_rowsumsq_diff_f32: rows_per_sample*cols must be a multiple of 4");
    int slices = (int)((eps / 4 + EW_THREADS * 4 - 1) / (EW_THREADS * 4));
    if (slices > 64) slices = 64;
    if (slices < 1) slices = 1;
    dim3 grid(n_samples, slices);
    rowsumsq_diff_kernel<<<grid, EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(a, b, out, eps);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_scale_bf16(const void* src, void* dst, int64_t n, float scale, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(src && dst && n >= 0, "vb_scale_bf16: bad args");
    if (n == 0) return VB_OK;
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "vb_scale_bf16: pointers must be 16B aligned");
    scale_bf16_kernel<<<ew_grid(n / 8 + 1), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const bf16*>(src), static_cast<bf16*>(dst), n, scale);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_philox_normal_f32(float* out, int64_t n_images, int64_t elems_per_image, uint64_t seed, uint64_t first_image,
                                    vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(out && n_images >= 0 && elems_per_image > 0, "vb_philox_normal_f32: bad args");
    VB_CHECK_ARG(elems_per_image % 4 == 0 && elems_per_image / 4 <= 0xFFFFFFFFll, "vb_philox_normal_f32: elems_per_image=%lld must be a multiple of 4 (< 2^34)", (long long)elems_per_image);
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "vb_philox_normal_f32: out must be 16B aligned");
    if (n_images == 0) return VB_OK;
    const int64_t quads = elems_per_image / 4;
    philox_normal_kernel<<<ew_grid(n_images * quads), EW_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(out, quads, n_images * quads, seed, first_image);
    VB_CHECK_LAUNCH();
    return VB_OK;
}
