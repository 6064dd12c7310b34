// Fused global-norm clip + SGD-momentum step over a flat gradient arena.
//
// Replaces, per optimisation step of apps/vit/train.py:277-283 with src/vitef/optim.py:76-82 (torch.optim.SGD, momentum,
// dampening 0, no nesterov): clip_grad_norm_ (one norm kernel per tensor + stack + norm + one multiply per tensor) and
// the foreach SGD update (three multi-tensor passes) by two launches: a sum of squares over the arena, then one pass that
// reads g and the momentum buffer and updates the (separately allocated, fp32) parameters through a chunk table.
// HBM-bound: 5 x 4 B per trainable element per step (g read twice, v read + write, p read + write = 2.06 GB for ViT-B).
#include "host_utils.h"
#include "ptx.cuh"

namespace vb {

constexpr int OPT_THREADS = 256;

// Fixed-order tree sum of a block's values (same result on every run and every rank: data-parallel replicas must compute
// bit-identical clip coefficients from their bit-identical all-reduced gradients).
__device__ __forceinline__ float block_sum_deterministic(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    return red[0];
}

// partials[blockIdx.x] = sum over this block's grid-stride share of x[i]^2 (no atomics: deterministic)
__global__ void __launch_bounds__(OPT_THREADS) sumsq_f32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    __shared__ float red[OPT_THREADS / 32];
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 128-bit loads in flight
        const float4 u0 = __ldg(x4 + i), u1 = __ldg(x4 + i + stride), u2 = __ldg(x4 + i + 2 * stride), u3 = __ldg(x4 + i + 3 * stride);
        a0 += u0.x * u0.x + u0.y * u0.y + u0.z * u0.z + u0.w * u0.w;
        a1 += u1.x * u1.x + u1.y * u1.y + u1.z * u1.z + u1.w * u1.w;
        a2 += u2.x * u2.x + u2.y * u2.y + u2.z * u2.z + u2.w * u2.w;
        a3 += u3.x * u3.x + u3.y * u3.y + u3.z * u3.z + u3.w * u3.w;
    }
    for (; i < n4; i += stride) {
        const float4 u = __ldg(x4 + i);
        a0 += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int64_t j = n4 << 2; j < n; ++j) a0 += x[j] * x[j];
    const float t = block_sum_deterministic((a0 + a1) + (a2 + a3), red);
    if (threadIdx.x == 0) out[blockIdx.x] = t;
}

struct OptChunk {
    float* param;        // first element of this chunk inside its parameter tensor
    long long arena_off; // same element inside the gradient / momentum arenas
    int count;
    int pad;
    bf16* shadow;        // optional: same element inside the parameter's bf16 shadow copy (the GEMM operand), refreshed here
};

// bf16 shadow of four / one freshly updated parameter values: saves the stand-alone cast pass (one launch per weight matrix
// and a second read of every parameter) that would otherwise follow the optimizer step
__device__ __forceinline__ void store_shadow4(bf16* shadow, int i, const float4& p4) {
    if (shadow != nullptr) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(p4.x, p4.y), pack_bf16x2(p4.z, p4.w));
}

// Hyper-parameters by value or, when `hyper` is given, from device memory (so a step captured in a CUDA graph follows an LR
// schedule / Adam's bias corrections without being re-captured).
//   SGD:   hyper = {lr, momentum, weight_decay, max_norm}
//   AdamW: hyper = {lr, weight_decay, max_norm, lr / bc1, 1 / sqrt(bc2)}

// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))  (torch.nn.utils.clip_grad_norm_);  g <- coef g;
// v <- momentum v + g  (first step: v = g);  p <- p - lr (v or g)
__global__ void __launch_bounds__(OPT_THREADS)
sgd_momentum_clip_kernel(const OptChunk* __restrict__ table, const float* __restrict__ grad, float* __restrict__ mom,
                         const float* __restrict__ sumsq_partials, int n_partials, float* __restrict__ norm_out, float max_norm,
                         float lr, float momentum, float weight_decay, int first_step, const float* __restrict__ hyper) {
    __shared__ float red[OPT_THREADS / 32];
    const OptChunk c = table[blockIdx.x];
    if (hyper != nullptr) {
        lr = __ldg(hyper);
        momentum = __ldg(hyper + 1);
        weight_decay = __ldg(hyper + 2);
        max_norm = __ldg(hyper + 3);
    }
    float part = 0.f;
    for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) part += __ldg(sumsq_partials + i);  // fixed order per thread
    const float norm = sqrtf(block_sum_deterministic(part, red));
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
    const float coef = fminf(1.f, max_norm / (norm + 1e-6f));
    const float* g = grad + c.arena_off;
    float* v = mom + c.arena_off;
    float* p = c.param;
    const int n4 = c.count >> 2;  // chunk starts are 16-byte aligned in all three arrays
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
        float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 p4 = reinterpret_cast<float4*>(p)[i];
        g4.x = fmaf(weight_decay, p4.x, g4.x * coef);
        g4.y = fmaf(weight_decay, p4.y, g4.y * coef);
        g4.z = fmaf(weight_decay, p4.z, g4.z * coef);
        g4.w = fmaf(weight_decay, p4.w, g4.w * coef);
        if (momentum != 0.f) {
            float4 v4 = reinterpret_cast<float4*>(v)[i];
            if (first_step) {
                v4 = g4;
            } else {
                v4.x = fmaf(momentum, v4.x, g4.x);
                v4.y = fmaf(momentum, v4.y, g4.y);
                v4.z = fmaf(momentum, v4.z, g4.z);
                v4.w = fmaf(momentum, v4.w, g4.w);
            }
            reinterpret_cast<float4*>(v)[i] = v4;
            g4 = v4;
        }
        p4.x = fmaf(-lr, g4.x, p4.x);
        p4.y = fmaf(-lr, g4.y, p4.y);
        p4.z = fmaf(-lr, g4.z, p4.z);
        p4.w = fmaf(-lr, g4.w, p4.w);
        reinterpret_cast<float4*>(p)[i] = p4;
        store_shadow4(c.shadow, i, p4);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < c.count; i += OPT_THREADS) {
        float gi = fmaf(weight_decay, p[i], g[i] * coef);
        if (momentum != 0.f) {
            const float vi = first_step ? gi : fmaf(momentum, v[i], gi);
            v[i] = vi;
            gi = vi;
        }
        const float pn = fmaf(-lr, gi, p[i]);
        p[i] = pn;
        if (c.shadow != nullptr) c.shadow[i] = __float2bfloat16_rn(pn);
    }
}

// AdamW (torch.optim.AdamW, amsgrad off; src/vitef/optim.py:83-88) with the same clip coefficient folded in:
//   g <- coef g;  p <- p (1 - lr wd);  m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2;
//   p <- p - (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)          (bc1 = 1 - b1^t, bc2 = 1 - b2^t, computed by the host)
__global__ void __launch_bounds__(OPT_THREADS)
adamw_clip_kernel(const OptChunk* __restrict__ table, const float* __restrict__ grad, float* __restrict__ exp_avg,
                  float* __restrict__ exp_avg_sq, const float* __restrict__ sumsq_partials, int n_partials,
                  float* __restrict__ norm_out, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                  float step_size, float inv_bc2_sqrt, const float* __restrict__ hyper) {
    __shared__ float red[OPT_THREADS / 32];
    const OptChunk c = table[blockIdx.x];
    if (hyper != nullptr) {
        lr = __ldg(hyper);
        weight_decay = __ldg(hyper + 1);
        max_norm = __ldg(hyper + 2);
        step_size = __ldg(hyper + 3);
        inv_bc2_sqrt = __ldg(hyper + 4);
    }
    float part = 0.f;
    for (int i = threadIdx.x; i < n_partials; i += OPT_THREADS) part += __ldg(sumsq_partials + i);
    const float norm = sqrtf(block_sum_deterministic(part, red));
    if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
    const float coef = fminf(1.f, max_norm / (norm + 1e-6f));
    const float decay = 1.f - lr * weight_decay, omb1 = 1.f - beta1, omb2 = 1.f - beta2;
    const float* g = grad + c.arena_off;
    float* m = exp_avg + c.arena_off;
    float* v = exp_avg_sq + c.arena_off;
    float* p = c.param;
    auto upd = [&](float gi, float& mi, float& vi, float& pi) {
        gi *= coef;
        pi *= decay;
        mi = fmaf(omb1, gi - mi, mi);
        vi = fmaf(omb2 * gi, gi, beta2 * vi);
        pi = fmaf(-step_size, mi / fmaf(sqrtf(vi), inv_bc2_sqrt, eps), pi);
    };
    const int n4 = c.count >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
        upd(g4.x, m4.x, v4.x, p4.x);
        upd(g4.y, m4.y, v4.y, p4.y);
        upd(g4.z, m4.z, v4.z, p4.z);
        upd(g4.w, m4.w, v4.w, p4.w);
        reinterpret_cast<float4*>(m)[i] = m4;
        reinterpret_cast<float4*>(v)[i] = v4;
        reinterpret_cast<float4*>(p)[i] = p4;
        store_shadow4(c.shadow, i, p4);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < c.count; i += OPT_THREADS) {
        upd(g[i], m[i], v[i], p[i]);
        if (c.shadow != nullptr) c.shadow[i] = __float2bfloat16_rn(p[i]);
    }
}

}  // namespace vb

extern "C" int vb_sumsq_partials_f32(const float* x, int64_t n, float* partials, int32_t n_partials, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(x && partials && n >= 0 && n_partials > 0, "vb_sumsq_partials_f32: bad args");
    VB_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "vb_sumsq_partials_f32: x must be 16-byte aligned");
    sumsq_f32_kernel<<<n_partials, OPT_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(x, n, partials);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_sgd_momentum_clip_step(const void* chunk_table, int32_t n_chunks, const float* grad_arena, float* momentum_arena,
                                         const float* sumsq_partials, int32_t n_partials, float* grad_norm_out, float max_norm,
                                         float lr, float momentum, float weight_decay, int32_t first_step, const float* hyper_dev,
                                         vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(chunk_table && grad_arena && sumsq_partials && n_chunks > 0 && n_partials > 0, "vb_sgd_momentum_clip_step: bad args");
    VB_CHECK_ARG(momentum == 0.f || momentum_arena != nullptr, "vb_sgd_momentum_clip_step: momentum needs a momentum arena");
    VB_CHECK_ARG(sizeof(OptChunk) == 32, "vb_sgd_momentum_clip_step: chunk layout");
    sgd_momentum_clip_kernel<<<n_chunks, OPT_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const OptChunk*>(chunk_table), grad_arena, momentum_arena, sumsq_partials, n_partials, grad_norm_out, max_norm, lr,
        momentum, weight_decay, first_step, hyper_dev);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

extern "C" int vb_adamw_clip_step(const void* chunk_table, int32_t n_chunks, const float* grad_arena, float* exp_avg_arena,
                                  float* exp_avg_sq_arena, const float* sumsq_partials, int32_t n_partials, float* grad_norm_out,
                                  float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  float bias_correction1, float bias_correction2, const float* hyper_dev, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(chunk_table && grad_arena && exp_avg_arena && exp_avg_sq_arena && sumsq_partials && n_chunks > 0 && n_partials > 0,
                 "vb_adamw_clip_step: bad args");
    VB_CHECK_ARG(bias_correction1 > 0.f && bias_correction2 > 0.f, "vb_adamw_clip_step: bias corrections must be > 0");
    adamw_clip_kernel<<<n_chunks, OPT_THREADS, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const OptChunk*>(chunk_table), grad_arena, exp_avg_arena, exp_avg_sq_arena, sumsq_partials, n_partials, grad_norm_out,
        max_norm, lr, beta1, beta2, eps, weight_decay, lr / bias_correction1, 1.f / sqrtf(bias_correction2), hyper_dev);
    VB_CHECK_LAUNCH();
    return VB_OK;
}
