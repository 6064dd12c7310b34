// Device-side input pipeline: the reference's PIL / torchvision preprocessing of uint8 images, bit for bit, on the GPU.
//   eval / analysis / probing:  Resize(224) -> CenterCrop(224) -> ToTensor -> Normalize     (data/images/utils.py:348-366)
//   train:                      RandomResizedCrop(224) -> RandomHorizontalFlip -> ToTensor -> Normalize   (:339-347)
// The resize is Pillow's ImagingResample (bilinear, 8-bit): separable, horizontal pass first, the intermediate rounded to
// 8 bits, taps in 22-bit fixed point. The host computes the tap tables once per (source size, 224) pair; this kernel applies
// them: one CTA per (image, strip of output rows) resamples the source rows the strip needs horizontally into shared
// memory (uint8, exactly Pillow's intermediate image), then vertically, maps the 8-bit result through the 3 x 256
// Normalize(ToTensor(.)) table and writes fp32 NCHW (the tensor the reference model is fed) and / or the bf16 patch rows
// the patch-embed GEMM consumes, so a training batch crosses PCIe as 3 KB per CIFAR image instead of 602 KB.
#include "host_utils.h"
#include "ptx.cuh"

namespace vb {

constexpr int PRE_THREADS = 256;
constexpr int PRECISION_BITS = 22;  // Pillow Resample.c, 8 bits per channel

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// params[i] = {top, left, crop_h, crop_w, flip, row_table, col_table, unused}
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_u8_kernel(const uint8_t* __restrict__ src, int src_h, int src_w, const int* __restrict__ params,
                     const int* __restrict__ tab_bounds, const int* __restrict__ tab_coef, int ksize,
                     const float* __restrict__ lut, int out, int strip, float* __restrict__ out_f32,
                     bf16* __restrict__ out_patches, int patch, int max_rows) {
    extern __shared__ uint8_t hbuf[];  // [rows of the strip's source window][out][3]
    __shared__ float slut[3 * 256];
    const int img = blockIdx.x, oy0 = blockIdx.y * strip;
    int top = 0, left = 0, flip = 0, trow = 0, tcol = 0;
    if (params != nullptr) {
        const int* pr = params + (size_t)img * 8;
        top = pr[0], left = pr[1], flip = pr[4], trow = pr[5], tcol = pr[6];
    }
    for (int i = threadIdx.x; i < 3 * 256; i += PRE_THREADS) slut[i] = lut[i];
    const int* brow = tab_bounds + (size_t)trow * out * 2;
    const int* bcol = tab_bounds + (size_t)tcol * out * 2;
    const int* krow = tab_coef + (size_t)trow * out * ksize;
    const int* kcol = tab_coef + (size_t)tcol * out * ksize;
    const int oy1 = min(oy0 + strip, out);
    const int y_first = brow[2 * oy0];
    const int y_end = brow[2 * (oy1 - 1)] + brow[2 * (oy1 - 1) + 1];  // bounds are monotone in the output index
    const int nrows = min(y_end - y_first, max_rows);
    const uint8_t* simg = src + (size_t)img * src_h * src_w * 3;

    // horizontal pass (ImagingResampleHorizontal_8bpc) of the source rows this strip needs
    for (int idx = threadIdx.x; idx < nrows * out * 3; idx += PRE_THREADS) {
        const int r = idx / (out * 3), rem = idx - r * out * 3, ox = rem / 3, c = rem - ox * 3;
        const int x0 = bcol[2 * ox], cnt = bcol[2 * ox + 1];
        const uint8_t* row = simg + ((size_t)(top + y_first + r) * src_w + left + x0) * 3 + c;
        int acc = 1 << (PRECISION_BITS - 1);
        for (int x = 0; x < cnt; ++x) acc += (int)row[3 * x] * kcol[ox * ksize + x];
        hbuf[idx] = clip8(acc);
    }
    __syncthreads();

    // vertical pass (ImagingResampleVertical_8bpc), Normalize(ToTensor(.)) through the table, flip on the way out
    const int np_side = patch > 0 ? out / patch : 0;
    for (int idx = threadIdx.x; idx < (oy1 - oy0) * out * 3; idx += PRE_THREADS) {
        const int c = idx / ((oy1 - oy0) * out), rem = idx - c * (oy1 - oy0) * out, oyl = rem / out, ox = rem - oyl * out;
        const int oy = oy0 + oyl;
        const int y0 = brow[2 * oy] - y_first, cnt = brow[2 * oy + 1];
        int acc = 1 << (PRECISION_BITS - 1);
        for (int y = 0; y < cnt; ++y) acc += (int)hbuf[((y0 + y) * out + ox) * 3 + c] * krow[oy * ksize + y];
        const float v = slut[c * 256 + clip8(acc)];
        const int oxo = flip ? out - 1 - ox : ox;
        if (out_f32 != nullptr) out_f32[(((size_t)img * 3 + c) * out + oy) * out + oxo] = v;
        if (out_patches != nullptr) {
            const size_t prow = (size_t)img * np_side * np_side + (size_t)(oy / patch) * np_side + oxo / patch;
            out_patches[prow * (3 * patch * patch) + c * patch * patch + (oy % patch) * patch + oxo % patch] = __float2bfloat16(v);
        }
    }
}

}  // namespace vb

extern "C" int vb_preprocess_u8(const uint8_t* src, int32_t n, int32_t src_h, int32_t src_w, const int32_t* params,
                                const int32_t* tab_bounds, const int32_t* tab_coef, int32_t n_tables, int32_t ksize,
                                int32_t max_src_rows_per_strip, const float* lut, int32_t out, float* out_f32,
                                void* out_patches, int32_t patch, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(src && tab_bounds && tab_coef && lut, "vb_preprocess_u8: null pointer");
    VB_CHECK_ARG(out_f32 || out_patches, "vb_preprocess_u8: no output requested");
    VB_CHECK_ARG(n > 0 && src_h > 0 && src_w > 0 && out > 0 && n_tables > 0 && ksize > 0, "vb_preprocess_u8: bad sizes");
    if (out_patches) VB_CHECK_ARG(patch > 0 && out % patch == 0, "vb_preprocess_u8: out=%d not divisible by patch=%d", out, patch);
    const int strip = (out_patches && patch > 0) ? patch : 16;
    VB_CHECK_ARG(max_src_rows_per_strip > 0, "vb_preprocess_u8: max_src_rows_per_strip must be > 0");
    const size_t smem = (size_t)max_src_rows_per_strip * out * 3;
    VB_CHECK_ARG(smem <= 200 * 1024, "vb_preprocess_u8: a strip of %d output rows needs %d source rows (%zu B of shared memory)",
                 strip, max_src_rows_per_strip, smem);
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(preprocess_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid(n, (out + strip - 1) / strip);
    preprocess_u8_kernel<<<grid, PRE_THREADS, smem, static_cast<cudaStream_t>(stream_)>>>(
        src, src_h, src_w, params, tab_bounds, tab_coef, ksize, lut, out, strip, out_f32, static_cast<bf16*>(out_patches),
        patch, max_src_rows_per_strip);
    VB_CHECK_LAUNCH();
    return VB_OK;
}
