// Host-side helpers shared by all translation units: error reporting, launch accounting, TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "vitb200.h"

namespace vb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

#define VB_CHECK_ARG(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            ::vb::set_error(__VA_ARGS__);  \
            return VB_ERR_INVALID;         \
        }                                  \
    } while (0)

#define VB_CHECK_CUDA(expr)                                                                          \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            ::vb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return VB_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define VB_CHECK_LAUNCH()                                                                            \
    do {                                                                                             \
        cudaError_t _e = cudaGetLastError();                                                         \
        if (_e != cudaSuccess) {                                                                     \
            ::vb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return VB_ERR_CUDA;                                                                      \
        }                                                                                            \
        ::vb::count_launch();                                                                        \
    } while (0)

// 2-D row-major tensor map: `inner` contiguous elements per row, `outer` rows, row stride in bytes.
// Returns 0 on success (error string set otherwise).
int make_tensor_map_2d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* ptr, uint64_t inner,
                       uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                       CUtensorMapSwizzle swizzle);

// 3-D tensor map (dims innermost first); strides[0] = bytes between rows of dim1, strides[1] = bytes between dim2 slabs.
int make_tensor_map_3d(CUtensorMap* map, CUtensorMapDataType dtype, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2,
                       CUtensorMapSwizzle swizzle);

}  // namespace vb
