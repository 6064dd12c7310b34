// Warp-per-row LayerNorm forward / backward (HBM-bound).
//   - 128-bit loads/stores (8 bf16 per access), the whole row lives in registers, fp32 statistics
//   - forward saves mean / rstd; backward fuses the residual-gradient add (dx = dres + LN'(dy)) and accumulates
//     dgamma / dbeta in registers across a persistent row loop, one atomicAdd per column per block at the end.
// Replaces nn.LayerNorm (reference: src/vitef/models/transformer/utils.py:293; call sites
// architecture.py:347,349 and transformer/utils.py:396) and native_layer_norm_backward under autograd.
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace vb {

constexpr int LN_WARPS = 8;

// Forward: persistent warps (grid = 6 blocks per SM), gamma / beta staged once per block in shared memory, the
// next row's 128-bit loads are issued before the current row is reduced (two rows in flight per warp).
template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     bf16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int cols,
                     float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = cols >> 3;
    const int stride = gridDim.x * LN_WARPS;
    __shared__ float sgamma[CHUNKS * 256], sbeta[CHUNKS * 256];
    for (int i = threadIdx.x; i < CHUNKS * 256; i += LN_WARPS * 32) {
        sgamma[i] = i < cols ? __ldg(gamma + i) : 0.f;
        sbeta[i] = i < cols ? __ldg(beta + i) : 0.f;
    }
    __syncthreads();
    int row = blockIdx.x * LN_WARPS + warp;
    if (row >= rows) return;
    uint4 cur[CHUNKS], nxt[CHUNKS];
    auto load_row = [&](uint4(&dst)[CHUNKS], int r) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)r * cols);
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            dst[i] = (c < nchunks) ? __ldg(xr + c) : make_uint4(0, 0, 0, 0);
        }
    };
    load_row(cur, row);
    const float inv_cols = 1.f / (float)cols;
    for (; row < rows; row += stride) {
        const bool has_next = row + stride < rows;
        if (has_next) load_row(nxt, row + stride);
        float v[CHUNKS][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(w[j]);
                v[i][2 * j] = f.x;
                v[i][2 * j + 1] = f.y;
                sum += f.x + f.y;
            }
        }
        const float mean = warp_sum(sum) * inv_cols;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            if (lane + i * 32 < nchunks) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[i][j] - mean;
                    sq += d * d;
                }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) * inv_cols + eps);  // biased variance, as torch
        if (lane == 0) {
            if (mean_out) mean_out[row] = mean;
            if (rstd_out) rstd_out[row] = rstd;
        }
        uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * cols);
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
                const float4 g0 = *reinterpret_cast<const float4*>(sgamma + c * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(sgamma + c * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(sbeta + c * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(sbeta + c * 8 + 4);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
                uint4 u;
                u.x = pack_bf16x2(o[0], o[1]);
                u.y = pack_bf16x2(o[2], o[3]);
                u.z = pack_bf16x2(o[4], o[5]);
                u.w = pack_bf16x2(o[6], o[7]);
                yr[c] = u;
            }
        }
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) cur[i] = nxt[i];
    }
}

// Column sums of the residual gradient (dres) on the way through the backward kernels: dres of LN2's backward is the block's
// incoming gradient (its column sums are fc2's bias gradient), dres of LN1's backward is the gradient of the attention
// branch's output (the output projection's bias gradient), so the 24 stand-alone column-sum passes per step (each a full
// read of a [tokens, E] tensor) disappear. Partials live in shared memory, not registers (the dgamma / dbeta partials already
// fill the register budget of two CTAs per SM): every warp owns a slice laid out [chunk half][lane] as float4, so its
// read-modify-write per row is conflict-free and needs no synchronisation; one block reduction + atomics at the end.
__device__ __forceinline__ void accumulate_dres(float4* my_dr, int i, int lane, const float (&r8)[8]) {
    float4 a0 = my_dr[(2 * i) * 32 + lane], a1 = my_dr[(2 * i + 1) * 32 + lane];
    a0.x += r8[0], a0.y += r8[1], a0.z += r8[2], a0.w += r8[3];
    a1.x += r8[4], a1.y += r8[5], a1.z += r8[6], a1.w += r8[7];
    my_dr[(2 * i) * 32 + lane] = a0;
    my_dr[(2 * i + 1) * 32 + lane] = a1;
}
template <int CHUNKS>
__device__ __forceinline__ void reduce_dres(const float4* sdr4, float* __restrict__ dres_colsum, int cols) {
    __syncthreads();
    const float* s = reinterpret_cast<const float*>(sdr4);
    for (int idx = threadIdx.x; idx < CHUNKS * 256; idx += LN_WARPS * 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) t += s[w * CHUNKS * 256 + idx];
        const int e = idx & 3, ln = (idx >> 2) & 31, ih = idx >> 7;  // ih = 2 * chunk slot + half
        const int col = ((ih >> 1) * 32 + ln) * 8 + (ih & 1) * 4 + e;
        if (col < cols) atomicAdd(dres_colsum + col, t);
    }
}

// Backward: persistent warps, gamma in shared memory, dgamma / dbeta partials in registers across the row loop.
// Two CTAs (16 warps) per SM, <= 128 registers: each warp has the 9 independent 128-bit loads of its row in flight and
// the other 15 warps cover their latency (the earlier one-CTA/SM version with a register-prefetched next row reached
// 62 % of the HBM peak). Normalised x and gamma*dy are recomputed in the second sweep instead of being kept.
template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
layernorm_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const bf16* __restrict__ dres,
                     bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dres_colsum, int rows, int cols) {
    __shared__ float red[LN_WARPS][32 * 8 + 1];
    __shared__ float sgamma[CHUNKS * 256];
    extern __shared__ float4 sdr4[];  // dres_colsum only: [LN_WARPS][CHUNKS * 2][32 lanes] float4 partial column sums of dres
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = cols >> 3;
    for (int i = threadIdx.x; i < CHUNKS * 256; i += LN_WARPS * 32) sgamma[i] = i < cols ? __ldg(gamma + i) : 0.f;
    float4* my_dr = sdr4 + warp * (CHUNKS * 64);  // this warp's slice: no synchronisation inside the row loop
    if (dres_colsum != nullptr)
        for (int k = lane; k < CHUNKS * 64; k += 32) my_dr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    float dg[CHUNKS][8], db[CHUNKS][8];
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dg[i][j] = db[i][j] = 0.f;
    const float inv_cols = 1.f / (float)cols;
    const int stride = gridDim.x * LN_WARPS;
    for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += stride) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * cols);
        const uint4* dyr = reinterpret_cast<const uint4*>(dy + (size_t)row * cols);
        const uint4* drr = dres ? reinterpret_cast<const uint4*>(dres + (size_t)row * cols) : nullptr;
        uint4 cx[CHUNKS], cd[CHUNKS], cr[CHUNKS];
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            const bool ok = c < nchunks;
            cx[i] = ok ? __ldg(xr + c) : make_uint4(0, 0, 0, 0);
            cd[i] = ok ? __ldg(dyr + c) : make_uint4(0, 0, 0, 0);
            cr[i] = (ok && drr) ? __ldg(drr + c) : make_uint4(0, 0, 0, 0);
        }
        const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            const uint32_t wx[4] = {cx[i].x, cx[i].y, cx[i].z, cx[i].w};
            const uint32_t wd[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w};
            const float4 g0 = *reinterpret_cast<const float4*>(sgamma + c * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(sgamma + c * 8 + 4);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const bool ok = c < nchunks;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fx = unpack_bf16x2(wx[j]);
                const float2 fd = unpack_bf16x2(wd[j]);
                const float h0 = ok ? (fx.x - mu) * rs : 0.f, h1 = ok ? (fx.y - mu) * rs : 0.f;
                dg[i][2 * j] += fd.x * h0;
                dg[i][2 * j + 1] += fd.y * h1;
                db[i][2 * j] += fd.x;
                db[i][2 * j + 1] += fd.y;
                const float y0 = fd.x * g[2 * j], y1 = fd.y * g[2 * j + 1];
                s1 += y0 + y1;
                s2 += y0 * h0 + y1 * h1;
            }
        }
        s1 = warp_sum(s1) * inv_cols;
        s2 = warp_sum(s2) * inv_cols;
        uint4* dxr = reinterpret_cast<uint4*>(dx + (size_t)row * cols);
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
                const uint32_t wx[4] = {cx[i].x, cx[i].y, cx[i].z, cx[i].w};
                const uint32_t wd[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w};
                const uint32_t wr[4] = {cr[i].x, cr[i].y, cr[i].z, cr[i].w};
                const float4 g0 = *reinterpret_cast<const float4*>(sgamma + c * 8);
                const float4 g1 = *reinterpret_cast<const float4*>(sgamma + c * 8 + 4);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                uint32_t o[4];
                float r8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fx = unpack_bf16x2(wx[j]);
                    const float2 fd = unpack_bf16x2(wd[j]);
                    const float2 fr = unpack_bf16x2(wr[j]);  // zeros when there is no residual gradient
                    const float h0 = (fx.x - mu) * rs, h1 = (fx.y - mu) * rs;
                    o[j] = pack_bf16x2(rs * (fd.x * g[2 * j] - s1 - h0 * s2) + fr.x, rs * (fd.y * g[2 * j + 1] - s1 - h1 * s2) + fr.y);
                    r8[2 * j] = fr.x;
                    r8[2 * j + 1] = fr.y;
                }
                dxr[c] = make_uint4(o[0], o[1], o[2], o[3]);
                if (dres_colsum != nullptr) accumulate_dres(my_dr, i, lane, r8);
            }
        }
    }
    if (dres_colsum != nullptr) reduce_dres<CHUNKS>(sdr4, dres_colsum, cols);
    if (dgamma == nullptr && dbeta == nullptr) return;  // frozen norm: parameter gradients not needed
    // block reduction of the per-warp partials, one chunk-slot at a time, then one atomic per column
    for (int pass = 0; pass < 2; ++pass) {
        float* out = pass == 0 ? dgamma : dbeta;
        if (out == nullptr) continue;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? dg[i][j] : db[i][j];
            __syncthreads();
            const int t = threadIdx.x;  // 256 threads <-> 256 columns of this slot
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) s += red[w][t];
            const int col = (i * 32 + (t >> 3)) * 8 + (t & 7);
            if (col < cols) atomicAdd(out + col, s);
        }
    }
}

// Backward for wide rows (cols > 768, e.g. ViT-L/H): one CTA per SM (the per-lane dgamma / dbeta partials alone take
// 16 registers per 256 columns), so x / dy / dres of the NEXT row are prefetched into registers while the current row is
// reduced.
template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_wide_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const bf16* __restrict__ dres,
                     bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dres_colsum, int rows, int cols) {
    __shared__ float red[LN_WARPS][32 * 8 + 1];
    __shared__ float sgamma[CHUNKS * 256];
    extern __shared__ float4 sdr4[];  // dres_colsum only: [LN_WARPS][CHUNKS * 2][32 lanes] float4 partial column sums of dres
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = cols >> 3;
    for (int i = threadIdx.x; i < CHUNKS * 256; i += LN_WARPS * 32) sgamma[i] = i < cols ? __ldg(gamma + i) : 0.f;
    float4* my_dr = sdr4 + warp * (CHUNKS * 64);  // this warp's slice: no synchronisation inside the row loop
    if (dres_colsum != nullptr)
        for (int k = lane; k < CHUNKS * 64; k += 32) my_dr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    float dg[CHUNKS][8], db[CHUNKS][8];
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dg[i][j] = db[i][j] = 0.f;
    const float inv_cols = 1.f / (float)cols;
    const int stride = gridDim.x * LN_WARPS;
    int row = blockIdx.x * LN_WARPS + warp;
    uint4 cx[CHUNKS], cd[CHUNKS], cr[CHUNKS], nx[CHUNKS], nd[CHUNKS], nr[CHUNKS];
    float cmu = 0.f, crs = 0.f, nmu = 0.f, nrs = 0.f;
    auto load_row = [&](uint4(&ox)[CHUNKS], uint4(&od)[CHUNKS], uint4(&orr)[CHUNKS], float& mu, float& rs, int r) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)r * cols);
        const uint4* dyr = reinterpret_cast<const uint4*>(dy + (size_t)r * cols);
        const uint4* drr = dres ? reinterpret_cast<const uint4*>(dres + (size_t)r * cols) : nullptr;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            const bool ok = c < nchunks;
            ox[i] = ok ? __ldg(xr + c) : make_uint4(0, 0, 0, 0);
            od[i] = ok ? __ldg(dyr + c) : make_uint4(0, 0, 0, 0);
            orr[i] = (ok && drr) ? __ldg(drr + c) : make_uint4(0, 0, 0, 0);
        }
        mu = __ldg(mean + r);
        rs = __ldg(rstd + r);
    };
    if (row < rows) load_row(cx, cd, cr, cmu, crs, row);
    for (; row < rows; row += stride) {
        if (row + stride < rows) load_row(nx, nd, nr, nmu, nrs, row + stride);
        const float mu = cmu, rs = crs;
        float xh[CHUNKS][8], gy[CHUNKS][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            const uint32_t wx[4] = {cx[i].x, cx[i].y, cx[i].z, cx[i].w};
            const uint32_t wd[4] = {cd[i].x, cd[i].y, cd[i].z, cd[i].w};
            const float4 g0 = *reinterpret_cast<const float4*>(sgamma + c * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(sgamma + c * 8 + 4);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const bool ok = c < nchunks;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fx = unpack_bf16x2(wx[j]);
                const float2 fd = unpack_bf16x2(wd[j]);
                const float h0 = ok ? (fx.x - mu) * rs : 0.f, h1 = ok ? (fx.y - mu) * rs : 0.f;
                xh[i][2 * j] = h0;
                xh[i][2 * j + 1] = h1;
                dg[i][2 * j] += fd.x * h0;
                dg[i][2 * j + 1] += fd.y * h1;
                db[i][2 * j] += fd.x;
                db[i][2 * j + 1] += fd.y;
                const float y0 = fd.x * g[2 * j], y1 = fd.y * g[2 * j + 1];
                gy[i][2 * j] = y0;
                gy[i][2 * j + 1] = y1;
                s1 += y0 + y1;
                s2 += y0 * h0 + y1 * h1;
            }
        }
        s1 = warp_sum(s1) * inv_cols;
        s2 = warp_sum(s2) * inv_cols;
        uint4* dxr = reinterpret_cast<uint4*>(dx + (size_t)row * cols);
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
                const uint32_t wr[4] = {cr[i].x, cr[i].y, cr[i].z, cr[i].w};
                uint32_t o[4];
                float r8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fr = unpack_bf16x2(wr[j]);  // zeros when there is no residual gradient
                    o[j] = pack_bf16x2(rs * (gy[i][2 * j] - s1 - xh[i][2 * j] * s2) + fr.x,
                                       rs * (gy[i][2 * j + 1] - s1 - xh[i][2 * j + 1] * s2) + fr.y);
                    r8[2 * j] = fr.x;
                    r8[2 * j + 1] = fr.y;
                }
                dxr[c] = make_uint4(o[0], o[1], o[2], o[3]);
                if (dres_colsum != nullptr) accumulate_dres(my_dr, i, lane, r8);
            }
        }
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            cx[i] = nx[i];
            cd[i] = nd[i];
            cr[i] = nr[i];
        }
        cmu = nmu;
        crs = nrs;
    }
    if (dres_colsum != nullptr) reduce_dres<CHUNKS>(sdr4, dres_colsum, cols);
    if (dgamma == nullptr && dbeta == nullptr) return;  // frozen norm: parameter gradients not needed
    // block reduction of the per-warp partials, one chunk-slot at a time, then one atomic per column
    for (int pass = 0; pass < 2; ++pass) {
        float* out = pass == 0 ? dgamma : dbeta;
        if (out == nullptr) continue;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? dg[i][j] : db[i][j];
            __syncthreads();
            const int t = threadIdx.x;  // 256 threads <-> 256 columns of this slot
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) s += red[w][t];
            const int col = (i * 32 + (t >> 3)) * 8 + (t & 7);
            if (col < cols) atomicAdd(out + col, s);
        }
    }
}

// u[s, d] = sum_l (zhat_a - zhat_b)^2 ; one warp per (sample,row) pair of rows, f32 inputs
template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_pair_sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ u,
                             int n_samples, int rows_per_sample, int cols, float eps) {
    // block = one sample; warps stride over the sample's rows; per-lane column accumulators
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    const int nchunks = cols >> 2;  // float4 chunks
    float acc[CHUNKS][4];
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int r = blockIdx.y * LN_WARPS + warp; r < rows_per_sample; r += gridDim.y * LN_WARPS) {
        const size_t row = (size_t)s * rows_per_sample + r;
        const float4* ar = reinterpret_cast<const float4*>(a + row * cols);
        const float4* br = reinterpret_cast<const float4*>(b + row * cols);
        float va[CHUNKS][4], vb_[CHUNKS][4];
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
                const float4 fa = __ldg(ar + c), fb = __ldg(br + c);
                va[i][0] = fa.x; va[i][1] = fa.y; va[i][2] = fa.z; va[i][3] = fa.w;
                vb_[i][0] = fb.x; vb_[i][1] = fb.y; vb_[i][2] = fb.z; vb_[i][3] = fb.w;
                sa += fa.x + fa.y + fa.z + fa.w;
                sb += fb.x + fb.y + fb.z + fb.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { va[i][j] = 0.f; vb_[i][j] = 0.f; }
            }
        }
        const float ma = warp_sum(sa) / (float)cols, mb = warp_sum(sb) / (float)cols;
        float qa = 0.f, qb = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float da = va[i][j] - ma, db = vb_[i][j] - mb;
                    qa += da * da;
                    qb += db * db;
                }
            }
        }
        const float ra = rsqrtf(warp_sum(qa) / (float)cols + eps), rb = rsqrtf(warp_sum(qb) / (float)cols + eps);
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // __fmul_rn: no FMA contraction, so identical inputs cancel exactly
                    const float d = __fmul_rn(va[i][j] - ma, ra) - __fmul_rn(vb_[i][j] - mb, rb);
                    acc[i][j] += d * d;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(u + (size_t)s * cols + c * 4 + j, acc[i][j]);
        }
    }
}

// Perturbation form of the kernel above: the second input is given as b = a + scale * d (fp32 token difference d), and the
// difference of the normalised rows is evaluated without cancellation,
//   zhat_b - zhat_a = cd * rstd_b + ca * (rstd_b - rstd_a),   ca = a - mean(a),  cd = scale * (d - mean(d)),
//   rstd_b - rstd_a = -(2 mean(ca cd) + mean(cd^2)) * rstd_a^2 rstd_b^2 / (rstd_a + rstd_b),
// so a perturbation of relative size 1e-3 (or 1e-6) keeps the full fp32 precision of d; d = 0 gives exactly 0.
template <int CHUNKS>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_delta_sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ d, float scale, float* __restrict__ u,
                              int n_samples, int rows_per_sample, int cols, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    const int nchunks = cols >> 2;  // float4 chunks
    float acc[CHUNKS][4];
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float inv_n = 1.f / (float)cols;
    for (int r = blockIdx.y * LN_WARPS + warp; r < rows_per_sample; r += gridDim.y * LN_WARPS) {
        const size_t row = (size_t)s * rows_per_sample + r;
        const float4* ar = reinterpret_cast<const float4*>(a + row * cols);
        const float4* dr = reinterpret_cast<const float4*>(d + row * cols);
        float va[CHUNKS][4], vd[CHUNKS][4];
        float sa = 0.f, sd = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
                const float4 fa = __ldg(ar + c), fd = __ldg(dr + c);
                va[i][0] = fa.x; va[i][1] = fa.y; va[i][2] = fa.z; va[i][3] = fa.w;
                vd[i][0] = fd.x * scale; vd[i][1] = fd.y * scale; vd[i][2] = fd.z * scale; vd[i][3] = fd.w * scale;
                sa += fa.x + fa.y + fa.z + fa.w;
                sd += (vd[i][0] + vd[i][1]) + (vd[i][2] + vd[i][3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { va[i][j] = 0.f; vd[i][j] = 0.f; }
            }
        }
        const float ma = warp_sum(sa) * inv_n, md = warp_sum(sd) * inv_n;
        float qa = 0.f, qb = 0.f, qx = 0.f, qd = 0.f;
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float ca = va[i][j] - ma, cd = vd[i][j] - md, cb = ca + cd;
                    va[i][j] = ca;
                    vd[i][j] = cd;
                    qa += ca * ca;
                    qb += cb * cb;
                    qx += ca * cd;
                    qd += cd * cd;
                }
            }
        }
        const float var_a = warp_sum(qa) * inv_n, var_b = warp_sum(qb) * inv_n;
        const float dvar = (2.f * warp_sum(qx) + warp_sum(qd)) * inv_n;  // = var_b - var_a, free of cancellation when small
        const float ra = rsqrtf(var_a + eps), rb = rsqrtf(var_b + eps);
        const float dr_ = -dvar * (ra * ra) * (rb * rb) / (ra + rb);      // = rb - ra
#pragma unroll
        for (int i = 0; i < CHUNKS; ++i) {
            const int c = lane + i * 32;
            if (c < nchunks) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dz = fmaf(vd[i][j], rb, va[i][j] * dr_);
                    acc[i][j] += dz * dz;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(u + (size_t)s * cols + c * 4 + j, acc[i][j]);
        }
    }
}

}  // namespace vb

#define VB_LN_DISPATCH(chunks, CALL)            \
    switch (chunks) {                           \
        case 1: { constexpr int C = 1; CALL; } break; \
        case 2: { constexpr int C = 2; CALL; } break; \
        case 3: { constexpr int C = 3; CALL; } break; \
        case 4: { constexpr int C = 4; CALL; } break; \
        case 5: { constexpr int C = 5; CALL; } break; \
        case 6: { constexpr int C = 6; CALL; } break; \
        case 7: { constexpr int C = 7; CALL; } break; \
        default: { constexpr int C = 8; CALL; } break; \
    }

extern "C" int vb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                int32_t rows, int32_t cols, float eps, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(x && gamma && beta && y, "vb_layernorm_fwd: null pointer");
    VB_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 2048, "vb_layernorm_fwd: cols=%d must be a multiple of 8, <= 2048", cols);
    const int chunks = (cols / 8 + 31) / 32;
    int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    static const int blocks_per_sm = []() {
        const char* e = getenv("VB_LN_FWD_BLOCKS_PER_SM");  // measurement knob
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : 8;
    }();
    if (grid > num_sms() * blocks_per_sm) grid = num_sms() * blocks_per_sm;
    VB_LN_DISPATCH(chunks, (layernorm_fwd_kernel<C><<<grid, LN_WARPS * 32, 0, stream>>>(
                               static_cast<const bf16*>(x), gamma, beta, static_cast<bf16*>(y), mean, rstd, rows, cols, eps)));
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int64_t vb_layernorm_bwd_workspace_bytes(int32_t cols) {
    (void)cols;
    return 0;  // parameter gradients are reduced with atomics; no workspace needed in this version
}

extern "C" int vb_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                const void* dres, void* dx, float* dgamma, float* dbeta, float* dres_colsum, int32_t rows,
                                int32_t cols, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(dres_colsum == nullptr || dres != nullptr, "vb_layernorm_bwd: dres_colsum needs dres");
    VB_CHECK_ARG(dy && x && gamma && mean && rstd && dx, "vb_layernorm_bwd: null pointer");
    VB_CHECK_ARG(rows > 0 && cols > 0 && cols % 8 == 0 && cols <= 2048, "vb_layernorm_bwd: cols=%d must be a multiple of 8, <= 2048", cols);
    const int chunks = (cols / 8 + 31) / 32;
    int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    const int max_grid = num_sms() * 2;
    if (grid > max_grid) grid = max_grid;
    if (dres_colsum != nullptr && chunks > 4) {
        // rows wider than 1024 columns: the shared-memory partials would need the > 48 KB opt-in; take the separate pass
        const int rc = vb_colsum_bf16(dres, cols, dres_colsum, rows, cols, stream_);
        if (rc != VB_OK) return rc;
        dres_colsum = nullptr;
    }
    const size_t dyn = dres_colsum != nullptr ? sizeof(float) * LN_WARPS * chunks * 256 : 0;
    if (chunks <= 3) {
        VB_LN_DISPATCH(chunks, (layernorm_bwd_kernel < C <= 3 ? C : 1 > <<<grid, LN_WARPS * 32, dyn, stream>>>(
                                   static_cast<const bf16*>(dy), static_cast<const bf16*>(x), gamma, mean, rstd,
                                   static_cast<const bf16*>(dres), static_cast<bf16*>(dx), dgamma, dbeta, dres_colsum, rows, cols)));
    } else {
        VB_LN_DISPATCH(chunks, (layernorm_bwd_wide_kernel<C><<<grid, LN_WARPS * 32, dyn, stream>>>(
                                   static_cast<const bf16*>(dy), static_cast<const bf16*>(x), gamma, mean, rstd,
                                   static_cast<const bf16*>(dres), static_cast<bf16*>(dx), dgamma, dbeta, dres_colsum, rows, cols)));
    }
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_layernorm_pair_sqdiff(const float* a, const float* b, float* u, int32_t n_samples,
                                        int32_t rows_per_sample, int32_t cols, float eps, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(a && b && u, "vb_layernorm_pair_sqdiff: null pointer");
    VB_CHECK_ARG(n_samples > 0 && rows_per_sample > 0 && cols % 4 == 0 && cols <= 1024 * 1, "vb_layernorm_pair_sqdiff: cols=%d must be a multiple of 4, <= 1024", cols);
    const int chunks = (cols / 4 + 31) / 32;
    dim3 grid(n_samples, (rows_per_sample + LN_WARPS * 4 - 1) / (LN_WARPS * 4));
    VB_LN_DISPATCH(chunks, (layernorm_pair_sqdiff_kernel<C><<<grid, LN_WARPS * 32, 0, stream>>>(a, b, u, n_samples,
                                                                                                  rows_per_sample, cols, eps)));
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_layernorm_delta_sqdiff(const float* a, const float* d, float scale, float* u, int32_t n_samples,
                                         int32_t rows_per_sample, int32_t cols, float eps, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(a && d && u, "vb_layernorm_delta_sqdiff: null pointer");
    VB_CHECK_ARG(n_samples > 0 && rows_per_sample > 0 && cols % 4 == 0 && cols <= 1024 * 1, "vb_layernorm_delta_sqdiff: cols=%d must be a multiple of 4, <= 1024", cols);
    const int chunks = (cols / 4 + 31) / 32;
    dim3 grid(n_samples, (rows_per_sample + LN_WARPS * 4 - 1) / (LN_WARPS * 4));
    VB_LN_DISPATCH(chunks, (layernorm_delta_sqdiff_kernel<C><<<grid, LN_WARPS * 32, 0, stream>>>(a, d, scale, u, n_samples,
                                                                                                   rows_per_sample, cols, eps)));
    VB_CHECK_LAUNCH();
    return VB_OK;
}
