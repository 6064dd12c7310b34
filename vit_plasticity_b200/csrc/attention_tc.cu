// Attention backward on tcgen05 for L <= 208 (ViT-B/L at 224x224: L = 197), head_dim = 64.
//
// One CTA per (image, head). Everything is computed in the TRANSPOSED domain, keys on the MMA M axis (two tiles of
// 128 keys, j = 0,1) and queries on the N / K axes (two halves of 112 + 96 queries, h = 0,1), because backward needs
// no row reduction (log-sum-exp and D = rowsum(dO*O) are per-query scalars read from smem):
//
//   MMA1(j,h):  S^T  = K_j  Q_h^T            [128 keys x Nh]   A = K_j  (K-major)   B = Q_h  (K-major)
//               dP^T = V_j dO_h^T            [128 keys x Nh]   A = V_j  (K-major)   B = dO_h (K-major)
//   elementwise (8 warps, thread = key row, tcgen05.ld):  P^T = exp(S^T/8 - lse[q]),  dS^T = P^T (dP^T - D[q]) / 8
//               both written as bf16 into 128B-swizzled smem tiles [128 keys][128 q]
//   MMA2(j,h):  dV_j  += P^T  dO_h           [128 keys x 64]   A = P^T  (K-major)   B = dO_h (MN-major)
//               dK_j  += dS^T Q_h            [128 keys x 64]   A = dS^T (K-major)   B = Q_h  (MN-major)
//               dQ_h  += dS   K_j            [128 q    x 64]   A = dS^T read as MN-major (same bytes!)  B = K_j (MN-major)
//
// Q, K, V, dO arrive by TMA through 3-D tensor maps (feature, token, image): rows >= L are zero-filled by the TMA
// unit, so no masking of loads is needed; P^T / dS^T rows of padded keys are forced to zero in the elementwise pass.
// TMEM (512 columns): S^T [0,112) | dP^T [128,240) | dV [256,320) | dK [320,384) | dQ_0 [384,448) | dQ_1 [448,512).
// Only descriptor forms already exercised by the GEMM kernel are used (SS MMAs, K-major / MN-major SWIZZLE_128B).
#include "host_utils.h"
#include "ptx.cuh"

namespace vb {

namespace attn_tc {

constexpr int HD = 64;
constexpr int TILE_BYTES = 128 * 128;       // 128 rows x 64 bf16
constexpr int OPER_BYTES = 2 * TILE_BYTES;  // 256 rows
constexpr int EW_WARPS = 8;
constexpr int THREADS = (EW_WARPS + 1) * 32;
constexpr int SMEM_BYTES = 4 * OPER_BYTES + 2 * OPER_BYTES + 2 * 256 * 4 + 128 + 1024;
constexpr uint32_t COL_ST = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;
constexpr int Q0[2] = {0, 112};
constexpr int NH[2] = {112, 96};

__device__ __forceinline__ int q0_of(int h) { return h == 0 ? 0 : 112; }
__device__ __forceinline__ int nh_of(int h) { return h == 0 ? 112 : 96; }

__device__ __forceinline__ void store_row64_bf16(bf16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(a[8 * i + 0]), __uint_as_float(a[8 * i + 1]));
        u.y = pack_bf16x2(__uint_as_float(a[8 * i + 2]), __uint_as_float(a[8 * i + 3]));
        u.z = pack_bf16x2(__uint_as_float(a[8 * i + 4]), __uint_as_float(a[8 * i + 5]));
        u.w = pack_bf16x2(__uint_as_float(a[8 * i + 6]), __uint_as_float(a[8 * i + 7]));
        d4[i] = u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(b[8 * i + 0]), __uint_as_float(b[8 * i + 1]));
        u.y = pack_bf16x2(__uint_as_float(b[8 * i + 2]), __uint_as_float(b[8 * i + 3]));
        u.z = pack_bf16x2(__uint_as_float(b[8 * i + 4]), __uint_as_float(b[8 * i + 5]));
        u.w = pack_bf16x2(__uint_as_float(b[8 * i + 6]), __uint_as_float(b[8 * i + 7]));
        d4[4 + i] = u;
    }
}

__global__ void __launch_bounds__(THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                        bf16* __restrict__ dqkv, int L, int H) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sDO = sQ + OPER_BYTES;
    uint8_t* sK = sDO + OPER_BYTES;
    uint8_t* sV = sK + OPER_BYTES;
    uint8_t* sP = sV + OPER_BYTES;   // P^T  tile: 2 chunks (64 q each) x [128 keys x 128 B]
    uint8_t* sS = sP + OPER_BYTES;   // dS^T tile, same layout
    float* sL = reinterpret_cast<float*>(sS + OPER_BYTES);  // [256] lse * log2(e), +inf for padded queries
    float* sD = sL + 256;                                    // [256] rowsum(dO * O)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
    uint64_t* bar_load = bars;
    uint64_t* bar_s = bars + 1;      // MMA1 done (also implies every earlier MMA is done)
    uint64_t* bar_p = bars + 2;      // elementwise done: smem tiles written, TMEM S^T/dP^T consumed
    uint64_t* bar_mma2 = bars + 3;   // MMA2 done
    uint64_t* bar_drain = bars + 4;  // dV_0 / dK_0 drained from TMEM
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / H, hd = blockIdx.x % H;
    const int E = H * HD;
    const int64_t ld3 = 3 * (int64_t)E;

    if (warp == EW_WARPS) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            mbar_init(bar_load, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, EW_WARPS);
            mbar_init(bar_mma2, 1);
            mbar_init(bar_drain, EW_WARPS);
            fence_barrier_init();
            // operands: two 128-row boxes each; rows >= L come back as zeros
            mbar_arrive_expect_tx(bar_load, 4 * OPER_BYTES);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                tma_load_3d(sQ + t * TILE_BYTES, &tmQKV, bar_load, hd * HD, t * 128, b);
                tma_load_3d(sK + t * TILE_BYTES, &tmQKV, bar_load, E + hd * HD, t * 128, b);
                tma_load_3d(sV + t * TILE_BYTES, &tmQKV, bar_load, 2 * E + hd * HD, t * 128, b);
                tma_load_3d(sDO + t * TILE_BYTES, &tmDO, bar_load, hd * HD, t * 128, b);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    } else {
        // zero the P^T / dS^T tiles once (padding columns are read by the dQ MMA as rows that are never stored)
        uint4* z = reinterpret_cast<uint4*>(sP);
        for (int i = tid; i < 2 * OPER_BYTES / 16; i += EW_WARPS * 32) z[i] = make_uint4(0, 0, 0, 0);
        // D = rowsum(dO * O) and lse (scaled to log2 units); 8 consecutive lanes share a row
        const bf16* obase = out + (int64_t)b * L * E + hd * HD;
        const bf16* dobase = dout + (int64_t)b * L * E + hd * HD;
        // all 16 global loads of a thread are issued before the first use (memory-level parallelism)
        uint4 uo[8], ud[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * EW_WARPS * 32;
            const int r = idx >> 3, c = idx & 7;
            uo[it] = make_uint4(0, 0, 0, 0);
            ud[it] = make_uint4(0, 0, 0, 0);
            if (r < L) {
                uo[it] = __ldg(reinterpret_cast<const uint4*>(obase + (int64_t)r * E + c * 8));
                ud[it] = __ldg(reinterpret_cast<const uint4*>(dobase + (int64_t)r * E + c * 8));
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = tid + it * EW_WARPS * 32;
            const int r = idx >> 3, c = idx & 7;
            const uint32_t wo[4] = {uo[it].x, uo[it].y, uo[it].z, uo[it].w}, wd[4] = {ud[it].x, ud[it].y, ud[it].z, ud[it].w};
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fo = unpack_bf16x2(wo[k]), fd = unpack_bf16x2(wd[k]);
                acc += fo.x * fd.x + fo.y * fd.y;
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (c == 0) {
                sD[r] = acc;
                sL[r] = (r < L) ? __ldg(lse + ((int64_t)b * H + hd) * L + r) * 1.4426950408889634f : INFINITY;
            }
        }
        fence_proxy_async_smem();  // the zero fill must be visible to the tensor core's (async proxy) smem reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == EW_WARPS) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO), aK = smem_u32(sK), aV = smem_u32(sV);
            const uint32_t aP = smem_u32(sP), aS = smem_u32(sS);
            auto mma1 = [&](int j, int h) {
                const int q0 = q0_of(h), nh = nh_of(h);
                const uint32_t idesc = make_idesc_bf16(128, nh, 0, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_bf16_ss(tmem_base + COL_ST, make_smem_desc_sw128(aK + j * TILE_BYTES + k * 32, 16, 1024),
                                 make_smem_desc_sw128(aQ + q0 * 128 + k * 32, 16, 1024), idesc, k > 0);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_bf16_ss(tmem_base + COL_DP, make_smem_desc_sw128(aV + j * TILE_BYTES + k * 32, 16, 1024),
                                 make_smem_desc_sw128(aDO + q0 * 128 + k * 32, 16, 1024), idesc, k > 0);
                }
                umma_commit(bar_s);
            };
            auto mma2 = [&](int j, int h) {
                const int q0 = q0_of(h), nh = nh_of(h);
                const uint32_t idesc_kmn = make_idesc_bf16(128, 64, 0, 1);
                const uint32_t idesc_mnmn = make_idesc_bf16(128, 64, 1, 1);
                const int ksteps = nh / 16;
                for (int k = 0; k < ksteps; ++k) {
                    // A: 16 queries of the [128 keys x 128 q] tile: 64-q chunks 16 KB apart, 32 B per step inside
                    const uint32_t aoff = (k >> 2) * TILE_BYTES + (k & 3) * 32;
                    // B: 16 query rows (2 KB) of dO / Q, used as [K = query][N = d] (MN-major)
                    const uint32_t boff = (q0 + k * 16) * 128;
                    umma_bf16_ss(tmem_base + COL_DV, make_smem_desc_sw128(aP + aoff, 16, 1024),
                                 make_smem_desc_sw128(aDO + boff, 8192, 1024), idesc_kmn, (h > 0 || k > 0));
                    umma_bf16_ss(tmem_base + COL_DK, make_smem_desc_sw128(aS + aoff, 16, 1024),
                                 make_smem_desc_sw128(aQ + boff, 8192, 1024), idesc_kmn, (h > 0 || k > 0));
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // dQ_h += dS K_j: A = dS^T bytes read MN-major (M = q, chunks 16 KB apart), 16 key rows per step
                    umma_bf16_ss(tmem_base + COL_DQ + h * 64, make_smem_desc_sw128(aS + k * 2048, TILE_BYTES, 1024),
                                 make_smem_desc_sw128(aK + j * TILE_BYTES + k * 2048, 8192, 1024), idesc_mnmn,
                                 (j > 0 || k > 0));
                }
                umma_commit(bar_mma2);
            };
            mbar_wait(bar_load, 0, 10);
            tc_fence_after();
            mma1(0, 0);
            for (int t = 0; t < 4; ++t) {
                const int j = t >> 1, h = t & 1;
                mbar_wait(bar_p, t & 1, 11);
                tc_fence_after();
                if (t == 2) {  // dV_0 / dK_0 must have left TMEM before the accumulators are re-used for j = 1
                    mbar_wait(bar_drain, 0, 12);
                    tc_fence_after();
                }
                mma2(j, h);
                if (t < 3) mma1((t + 1) >> 1, (t + 1) & 1);
            }
        }
    } else {
        // =========================== elementwise + epilogue warps ===========================
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32)
        const int hf = warp >> 2;      // which half of the query columns of this pass
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float scale = 0.125f, scale_log2 = 0.125f * 1.4426950408889634f;
        for (int t = 0; t < 4; ++t) {
            const int j = t >> 1, h = t & 1;
            const int q0 = q0_of(h), nh = nh_of(h);
            const int key = j * 128 + row;
            const bool key_ok = key < L;
            mbar_wait(bar_s, t & 1, 20);
            tc_fence_after();
            const int cbeg = hf * (nh >> 1), nchunks = nh >> 4;  // 8 columns per chunk, nh/2 columns per warp
            // one batch of TMEM loads for the warp's 56 (h = 0) or 48 (h = 1) columns, a single wait
            uint32_t sv[56], dv[56];
            {
                uint32_t (&s32)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
                uint32_t (&s16)[16] = *reinterpret_cast<uint32_t(*)[16]>(&sv[32]);
                uint32_t (&s8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&sv[48]);
                uint32_t (&d32)[32] = *reinterpret_cast<uint32_t(*)[32]>(&dv[0]);
                uint32_t (&d16)[16] = *reinterpret_cast<uint32_t(*)[16]>(&dv[32]);
                uint32_t (&d8)[8] = *reinterpret_cast<uint32_t(*)[8]>(&dv[48]);
                tmem_ld_32x32b_x32(lane_addr + COL_ST + cbeg, s32);
                tmem_ld_32x32b_x32(lane_addr + COL_DP + cbeg, d32);
                tmem_ld_32x32b_x16(lane_addr + COL_ST + cbeg + 32, s16);
                tmem_ld_32x32b_x16(lane_addr + COL_DP + cbeg + 32, d16);
                if (h == 0) {
                    tmem_ld_32x32b_x8(lane_addr + COL_ST + cbeg + 48, s8);
                    tmem_ld_32x32b_x8(lane_addr + COL_DP + cbeg + 48, d8);
                }
                tmem_ld_wait();
            }
#pragma unroll
            for (int cc = 0; cc < 7; ++cc) {
                if (cc < nchunks) {
                    const int c0 = cbeg + cc * 8;
                    const float4 l0 = *reinterpret_cast<const float4*>(sL + q0 + c0);
                    const float4 l1 = *reinterpret_cast<const float4*>(sL + q0 + c0 + 4);
                    const float4 d0 = *reinterpret_cast<const float4*>(sD + q0 + c0);
                    const float4 d1 = *reinterpret_cast<const float4*>(sD + q0 + c0 + 4);
                    const float lq[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
                    const float dq[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
                    float p[8], ds[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        p[i] = key_ok ? exp2f(__uint_as_float(sv[cc * 8 + i]) * scale_log2 - lq[i]) : 0.f;
                        ds[i] = p[i] * (__uint_as_float(dv[cc * 8 + i]) - dq[i]) * scale;
                    }
                    // 8 queries = one 16-byte unit of the 128B-swizzled [key row][64 q] chunk
                    const uint32_t off = (c0 >> 6) * TILE_BYTES + row * 128 + ((((c0 & 63) >> 3) ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(sP) + off),
                                 "r"(pack_bf16x2(p[0], p[1])), "r"(pack_bf16x2(p[2], p[3])), "r"(pack_bf16x2(p[4], p[5])),
                                 "r"(pack_bf16x2(p[6], p[7]))
                                 : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(sS) + off),
                                 "r"(pack_bf16x2(ds[0], ds[1])), "r"(pack_bf16x2(ds[2], ds[3])),
                                 "r"(pack_bf16x2(ds[4], ds[5])), "r"(pack_bf16x2(ds[6], ds[7]))
                                 : "memory");
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p);

            if (h == 1) {
                // all query halves of key tile j have been accumulated: drain dV_j (warps 0-3) / dK_j (warps 4-7)
                mbar_wait(bar_mma2, t & 1, 21);
                tc_fence_after();
                uint32_t a[32], c[32];
                const uint32_t col = hf == 0 ? COL_DV : COL_DK;
                tmem_ld_32x32b_x32(lane_addr + col, a);
                tmem_ld_32x32b_x32(lane_addr + col + 32, c);
                tmem_ld_wait();
                if (key_ok) {
                    bf16* dst = dqkv + ((int64_t)b * L + key) * ld3 + (hf == 0 ? 2 * E : E) + hd * HD;
                    store_row64_bf16(dst, a, c);
                }
                if (j == 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_drain);
                }
            }
        }
        // dQ_0 (warps 0-3: queries [0,112)) and dQ_1 (warps 4-7: queries [112,208)); MMA2(1,1) completion was
        // observed above (bar_mma2, t = 3)
        {
            uint32_t a[32], c[32];
            tmem_ld_32x32b_x32(lane_addr + COL_DQ + hf * 64, a);
            tmem_ld_32x32b_x32(lane_addr + COL_DQ + hf * 64 + 32, c);
            tmem_ld_wait();
            const int q = q0_of(hf) + row;
            if (row < nh_of(hf) && q < L) {
                bf16* dst = dqkv + ((int64_t)b * L + q) * ld3 + hd * HD;
                store_row64_bf16(dst, a, c);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EW_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace attn_tc

int launch_attention_bwd_tc(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int batch, int L,
                            int H, cudaStream_t stream) {
    using namespace attn_tc;
    const int64_t E = (int64_t)H * HD;
    CUtensorMap tmQKV, tmDO;
    int rc = make_tensor_map_3d(&tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * E, L, batch, 3 * E * 2, (uint64_t)L * 3 * E * 2, 64,
                                128, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map_3d(&tmDO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, E, L, batch, E * 2, (uint64_t)L * E * 2, 64, 128, 1,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set[dev] = true;
    }
    attention_bwd_tc_kernel<<<batch * H, THREADS, SMEM_BYTES, stream>>>(tmQKV, tmDO, out, dout, lse, dqkv, L, H);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

}  // namespace vb
