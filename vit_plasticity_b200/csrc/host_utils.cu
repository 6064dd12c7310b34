#include "host_utils.h"

#include <atomic>
#include <mutex>
#include <string.h>

namespace vb {

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        // resolved through the runtime so that the library links against libcudart only (no -lcuda)
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    return fn;
}

int make_tensor_map_2d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* ptr, uint64_t inner,
                       uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                       CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VB_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (row_stride_bytes & 15) != 0) {
        set_error("TMA operand must be 16-byte aligned (ptr=%p, row stride=%llu bytes)", ptr,
                  (unsigned long long)row_stride_bytes);
        return VB_ERR_INVALID;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    (void)elem_bytes;
    CUresult r = fn(map, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu box=%ux%u)", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
                  box_outer);
        return VB_ERR_CUDA;
    }
    return VB_OK;
}

int make_tensor_map_3d(CUtensorMap* map, CUtensorMapDataType dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2,
                       CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return VB_ERR_CUDA;
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (stride1_bytes & 15) != 0 || (stride2_bytes & 15) != 0) {
        set_error("TMA operand must be 16-byte aligned (ptr=%p, strides=%llu,%llu bytes)", ptr,
                  (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
        return VB_ERR_INVALID;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, dtype, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed with CUresult %d (dims=%llu,%llu,%llu strides=%llu,%llu box=%u,%u,%u)", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)stride1_bytes,
                  (unsigned long long)stride2_bytes, box0, box1, box2);
        return VB_ERR_CUDA;
    }
    return VB_OK;
}

}  // namespace vb

extern "C" {

int vb_version(void) { return 100; }
const char* vb_last_error(void) { return vb::g_err; }
int64_t vb_launch_count(void) { return vb::g_launches.load(std::memory_order_relaxed); }
void vb_reset_launch_count(void) { vb::g_launches.store(0, std::memory_order_relaxed); }

}  // extern "C"
