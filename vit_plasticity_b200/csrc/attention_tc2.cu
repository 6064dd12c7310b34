// Attention forward / backward on tcgen05 for L <= 208 (ViT-B/L at 224x224: L = 197), head_dim = 64.
//
// Three kernels, each sized so that TWO CTAs are resident per SM (<= 96 KB smem, 256 TMEM columns, <= 113 regs):
// the serial MMA -> tcgen05.ld -> math -> tcgen05.st -> MMA chain of one CTA overlaps with the other CTA's.
// Probabilities never touch shared memory: they are written back to TMEM as packed bf16 pairs (tcgen05.st) in place
// of the fp32 scores they were computed from, and consumed as the A operand of the next MMA (A-from-TMEM).
// Operands arrive by TMA through 3-D tensor maps (feature, token, image); rows >= L are zero-filled by the TMA unit.
//
//   fwd   CTA = (image, head, 128-query tile)   S = Q K^T (N = 208) -> exact softmax (two TMEM passes) -> O = P V
//   dQ    CTA = (image, head, 128-query tile)   4 passes over 64-key chunks: S, dP -> dS -> dQ += dS K ; also writes
//                                               delta = rowsum(dO * O) for the dK/dV kernel
//   dKdV  CTA = (image, head, 128-key tile)     transposed domain (keys on M; lse / delta are per-COLUMN scalars, so
//                                               no row reduction): 4 passes over 64-query chunks:
//                                               S^T, dP^T -> P^T, dS^T -> dV += P^T dO, dK += dS^T Q
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace vb {
namespace attn2 {

constexpr int HD = 64;
constexpr int TILE = 128 * 128;  // bytes: 128 rows x 64 bf16
constexpr int EW_WARPS = 8;
constexpr int THREADS = (EW_WARPS + 1) * 32;
constexpr float LOG2E = 1.4426950408889634f;
// smem descriptor halves: hi = SBO 1024 B | version 1 | SWIZZLE_128B ; lo = (address >> 4) | (LBO >> 4) << 16
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t LBO_K = 1u << 16;             // K-major operands: LBO field unused (16 B)
constexpr uint32_t LBO_MN = (8192u >> 4) << 16;  // MN-major, single 64-wide chunk: unused as well

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

__device__ __forceinline__ uint8_t* align_smem(uint8_t* raw) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}

// 32 fp32 accumulator columns -> 32 bf16 -> 64 contiguous bytes
__device__ __forceinline__ void store_32cols_bf16(bf16* dst, const uint32_t (&a)[32], float mul) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(a[8 * i + 0]) * mul, __uint_as_float(a[8 * i + 1]) * mul);
        u.y = pack_bf16x2(__uint_as_float(a[8 * i + 2]) * mul, __uint_as_float(a[8 * i + 3]) * mul);
        u.z = pack_bf16x2(__uint_as_float(a[8 * i + 4]) * mul, __uint_as_float(a[8 * i + 5]) * mul);
        u.w = pack_bf16x2(__uint_as_float(a[8 * i + 6]) * mul, __uint_as_float(a[8 * i + 7]) * mul);
        d4[i] = u;
    }
}

// =====================================================================================================
// forward
// =====================================================================================================
constexpr int FWD_SMEM = 5 * TILE + 3 * 128 * 4 * 2 + 128 + 1024;  // Q tile, K (2 tiles), V (2 tiles), max/sum exchange
constexpr uint32_t F_COL_S = 0, F_COL_O = 160;  // S [0,208); P packed in place: [0,56) and [112,160); O [160,224)

__global__ void __launch_bounds__(THREADS, 2)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out, int L,
                        int H, long long* __restrict__ dbg) {
    // dbg (development only, normally nullptr): per-CTA clock64 stamps of the pipeline phases
#define VB_STAMP(slot)                                                                       \
    do {                                                                                     \
        if (dbg != nullptr) dbg[(size_t)blockIdx.x * 16 + (slot)] = clock64();               \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    uint8_t* sQ = smem;             // 128 x 64
    uint8_t* sK = sQ + TILE;        // 256 x 64
    uint8_t* sV = sK + 2 * TILE;    // 256 x 64
    float* sMax = reinterpret_cast<float*>(sV + 2 * TILE);  // [2][128] partial row max per column half
    float* sSum = sMax + 256;                               // [2][128] partial row sums
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 256);
    uint64_t *bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_o = bars + 3;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
    const int qt = blockIdx.x & 1;  // query tile
    const int bh = blockIdx.x >> 1;
    const int b = bh / H, hd = bh % H;
    const int E = H * HD;
    if (tid == 0) VB_STAMP(0);

    if (warp == EW_WARPS) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQKV);
            mbar_init(bar_load, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, EW_WARPS);
            mbar_init(bar_o, 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(bar_load, 5 * TILE);
            tma_load_3d(sQ, &tmQKV, bar_load, hd * HD, qt * 128, b);
            for (int t = 0; t < 2; ++t) {
                tma_load_3d(sK + t * TILE, &tmQKV, bar_load, E + hd * HD, t * 128, b);
                tma_load_3d(sV + t * TILE, &tmQKV, bar_load, 2 * E + hd * HD, t * 128, b);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == EW_WARPS) {
        // all 32 lanes run this (uniform control flow); one elected lane issues the asynchronous instructions
        const uint32_t qlo = (smem_u32(sQ) >> 4) | LBO_K, klo = (smem_u32(sK) >> 4) | LBO_K, vlo = (smem_u32(sV) >> 4) | LBO_MN;
        if (lane == 0) VB_STAMP(1);
        mbar_wait(bar_load, 0, 30);
        tc_fence_after();
        if (lane == 0) VB_STAMP(2);
        if (elect_one()) {
            const uint32_t idesc_s = make_idesc_bf16(128, 208, 0, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_base + F_COL_S, make_desc(qlo + 2 * k, DESC_HI), make_desc(klo + 2 * k, DESC_HI), idesc_s, k > 0);
            umma_commit(bar_s);
        }
        __syncwarp();
        if (lane == 0) VB_STAMP(3);
        mbar_wait(bar_p, 0, 31);
        tc_fence_after();
        if (lane == 0) VB_STAMP(4);
        if (elect_one()) {
            // O = P V: A = packed P in TMEM (8 columns per 16 keys), B = V rows as [K = key][N = d] (MN-major)
            const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                const uint32_t acol = k < 7 ? k * 8 : 112 + (k - 7) * 8;  // keys [0,112) then [112,208)
                umma_bf16_ts(tmem_base + F_COL_O, tmem_base + acol, make_desc(vlo + k * 128, DESC_HI), idesc_o, k > 0);
            }
            umma_commit(bar_o);
        }
        __syncwarp();
        if (lane == 0) VB_STAMP(5);
    } else {
        const int quarter = warp & 3, hf = warp >> 2;
        const int row = quarter * 32 + lane;
        const int q = qt * 128 + row;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        const int cbeg = hf == 0 ? 0 : 112;
        const int nsub = hf == 0 ? 7 : 6;  // 16-column sub-chunks owned by this warp
        if (tid == 0) VB_STAMP(6);
        mbar_wait(bar_s, 0, 32);
        tc_fence_after();
        if (tid == 0) VB_STAMP(7);
        // pass 1: row max over this warp's valid columns
        float m = -INFINITY;
        for (int s = 0; s < nsub; ++s) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(lane_addr + F_COL_S + cbeg + s * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (cbeg + s * 16 + i < L) m = fmaxf(m, __uint_as_float(v[i]));
        }
        sMax[hf * 128 + row] = m;
        if (tid == 0) VB_STAMP(8);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // the two warps sharing this lane quarter
        m = fmaxf(m, sMax[(hf ^ 1) * 128 + row]);  // L >= 1: column 0 is valid, so m is finite
        const float mc = m * c;
        // pass 2: p = exp2(s c - m c), packed to bf16 pairs in place (own column range only)
        float sum = 0.f;
        for (int s = 0; s < nsub; ++s) {
            uint32_t v[16], pk[8];
            tmem_ld_32x32b_x16(lane_addr + F_COL_S + cbeg + s * 16, v);
            tmem_ld_wait();
            float p[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                p[i] = (cbeg + s * 16 + i < L) ? fast_ex2(__uint_as_float(v[i]) * c - mc) : 0.f;
                sum += p[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
            tmem_st_32x32b_x8(lane_addr + F_COL_S + cbeg + s * 8, pk);
        }
        sSum[hf * 128 + row] = sum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        if (tid == 0) VB_STAMP(9);
        mbar_wait(bar_o, 0, 33);
        tc_fence_after();
        if (tid == 0) VB_STAMP(10);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // partner's partial sum is visible
        const float tot = sum + sSum[(hf ^ 1) * 128 + row];
        uint32_t o[32];
        tmem_ld_32x32b_x32(lane_addr + F_COL_O + hf * 32, o);
        tmem_ld_wait();
        if (q < L) {
            store_32cols_bf16(out + ((int64_t)b * L + q) * E + hd * HD + hf * 32, o, 1.f / tot);
            if (hf == 0 && lse_out != nullptr) lse_out[((int64_t)b * H + hd) * L + q] = m * 0.125f + __logf(tot);
        }
        if (tid == 0) VB_STAMP(11);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EW_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
    if (tid == 0) VB_STAMP(12);
#undef VB_STAMP
}

// =====================================================================================================
// backward, dQ (+ delta)
// =====================================================================================================
constexpr int BWD_SMEM = 6 * TILE + 2 * 256 * 4 + 128 + 1024;
constexpr uint32_t B_COL_S = 0, B_COL_DP = 64, B_COL_ACC = 128;  // dQ kernel: dQ [128,192); dKdV: dV [128,192) dK [192,256)

__global__ void __launch_bounds__(THREADS, 2)
attention_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                           const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                           float* __restrict__ delta, bf16* __restrict__ dqkv, int L, int H) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    uint8_t* sQ = smem;            // 128 rows (this query tile)
    uint8_t* sDO = sQ + TILE;      // 128 rows
    uint8_t* sK = sDO + TILE;      // 256 rows
    uint8_t* sV = sK + 2 * TILE;   // 256 rows
    float* sD = reinterpret_cast<float*>(sV + 2 * TILE);  // [128] delta of this tile's rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 512);
    uint64_t *bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_done = bars + 3;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
    const int qt = blockIdx.x & 1;
    const int bh = blockIdx.x >> 1;
    const int b = bh / H, hd = bh % H;
    const int E = H * HD;
    const int64_t ld3 = 3 * (int64_t)E;

    if (warp == EW_WARPS) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            mbar_init(bar_load, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, EW_WARPS);
            mbar_init(bar_done, 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(bar_load, 6 * TILE);
            tma_load_3d(sQ, &tmQKV, bar_load, hd * HD, qt * 128, b);
            tma_load_3d(sDO, &tmDO, bar_load, hd * HD, qt * 128, b);
            for (int t = 0; t < 2; ++t) {
                tma_load_3d(sK + t * TILE, &tmQKV, bar_load, E + hd * HD, t * 128, b);
                tma_load_3d(sV + t * TILE, &tmQKV, bar_load, 2 * E + hd * HD, t * 128, b);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 256);
        tmem_relinquish();
    } else if (warp < 4) {
        // delta[q] = sum_d dO[q,d] O[q,d] for this tile's rows: one thread per row, 16 independent 16-byte loads
        const int q = qt * 128 + tid;
        float acc = 0.f;
        if (q < L) {
            const uint4* po = reinterpret_cast<const uint4*>(out + ((int64_t)b * L + q) * E + hd * HD);
            const uint4* pd = reinterpret_cast<const uint4*>(dout + ((int64_t)b * L + q) * E + hd * HD);
            uint4 uo[8], ud[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uo[i] = __ldg(po + i);
                ud[i] = __ldg(pd + i);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t wo[4] = {uo[i].x, uo[i].y, uo[i].z, uo[i].w}, wd[4] = {ud[i].x, ud[i].y, ud[i].z, ud[i].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 fo = unpack_bf16x2(wo[k]), fd = unpack_bf16x2(wd[k]);
                    acc += fo.x * fd.x + fo.y * fd.y;
                }
            }
            delta[((int64_t)b * H + hd) * L + q] = acc;
        }
        sD[tid] = acc;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == EW_WARPS) {
        const uint32_t qlo = (smem_u32(sQ) >> 4) | LBO_K, dolo = (smem_u32(sDO) >> 4) | LBO_K;
        const uint32_t klo = (smem_u32(sK) >> 4) | LBO_K, vlo = (smem_u32(sV) >> 4) | LBO_K;
        const uint32_t kmn = (smem_u32(sK) >> 4) | LBO_MN;
        const uint32_t idesc1 = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t idesc2 = make_idesc_bf16(128, 64, 0, 1);
        auto mma1 = [&](int kc) {  // S = Q K_kc^T, dP = dO V_kc^T  (64 keys = 8 KB of rows = 512 descriptor units)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_base + B_COL_S, make_desc(qlo + 2 * k, DESC_HI), make_desc(klo + kc * 512 + 2 * k, DESC_HI), idesc1, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_base + B_COL_DP, make_desc(dolo + 2 * k, DESC_HI), make_desc(vlo + kc * 512 + 2 * k, DESC_HI), idesc1, k > 0);
            umma_commit(bar_s);
        };
        mbar_wait(bar_load, 0, 40);
        tc_fence_after();
        if (elect_one()) mma1(0);
        __syncwarp();
        for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(bar_p, kc & 1, 41);
            tc_fence_after();
            if (elect_one()) {
                // dQ += dS_kc K_kc: A = packed dS in TMEM (keys [0,32) at dP cols [0,16), keys [32,64) at [32,48))
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ts(tmem_base + B_COL_ACC, tmem_base + B_COL_DP + (k >> 1) * 32 + (k & 1) * 8,
                                 make_desc(kmn + kc * 512 + k * 128, DESC_HI), idesc2, (kc > 0 || k > 0));
                if (kc < 3)
                    mma1(kc + 1);
                else
                    umma_commit(bar_done);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3, hf = warp >> 2;
        const int row = quarter * 32 + lane;
        const int q = qt * 128 + row;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        const float lse2 = q < L ? __ldg(lse + ((int64_t)b * H + hd) * L + q) * LOG2E : INFINITY;
        const float dlt = sD[row];
        for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(bar_s, kc & 1, 42);
            tc_fence_after();
            {
                const int c0 = hf * 32;  // this warp's 32 columns, one batch of TMEM loads, one wait
                uint32_t sv[32], dv[32], pk[16];
                tmem_ld_32x32b_x32(lane_addr + B_COL_S + c0, sv);
                tmem_ld_32x32b_x32(lane_addr + B_COL_DP + c0, dv);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float ds2[2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int key = kc * 64 + c0 + 2 * i + j;
                        const float p = key < L ? fast_ex2(__uint_as_float(sv[2 * i + j]) * c - lse2) : 0.f;
                        ds2[j] = p * (__uint_as_float(dv[2 * i + j]) - dlt) * 0.125f;
                    }
                    pk[i] = pack_bf16x2(ds2[0], ds2[1]);
                }
                tmem_st_32x32b_x16(lane_addr + B_COL_DP + c0, pk);  // in place, own column range
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p);
        }
        mbar_wait(bar_done, 0, 43);
        tc_fence_after();
        uint32_t a[32];
        tmem_ld_32x32b_x32(lane_addr + B_COL_ACC + hf * 32, a);
        tmem_ld_wait();
        if (q < L) store_32cols_bf16(dqkv + ((int64_t)b * L + q) * ld3 + hd * HD + hf * 32, a, 1.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EW_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// =====================================================================================================
// backward, dK / dV (transposed domain)
// =====================================================================================================
__global__ void __launch_bounds__(THREADS, 2)
attention_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                             const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int L,
                             int H) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    uint8_t* sK = smem;            // 128 rows (this key tile)
    uint8_t* sV = sK + TILE;       // 128 rows
    uint8_t* sQ = sV + TILE;       // 256 rows
    uint8_t* sDO = sQ + 2 * TILE;  // 256 rows
    float* sL = reinterpret_cast<float*>(sDO + 2 * TILE);  // [256] lse * log2(e), +inf for padded queries
    float* sD = sL + 256;                                   // [256] delta
    uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 256);
    uint64_t *bar_load = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_done = bars + 3;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 4);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
    const int kt = blockIdx.x & 1;  // key tile
    const int bh = blockIdx.x >> 1;
    const int b = bh / H, hd = bh % H;
    const int E = H * HD;
    const int64_t ld3 = 3 * (int64_t)E;

    if (warp == EW_WARPS) {
        if (elect_one()) {
            tma_prefetch_desc(&tmQKV);
            tma_prefetch_desc(&tmDO);
            mbar_init(bar_load, 1);
            mbar_init(bar_s, 1);
            mbar_init(bar_p, EW_WARPS);
            mbar_init(bar_done, 1);
            fence_barrier_init();
            mbar_arrive_expect_tx(bar_load, 6 * TILE);
            tma_load_3d(sK, &tmQKV, bar_load, E + hd * HD, kt * 128, b);
            tma_load_3d(sV, &tmQKV, bar_load, 2 * E + hd * HD, kt * 128, b);
            for (int t = 0; t < 2; ++t) {
                tma_load_3d(sQ + t * TILE, &tmQKV, bar_load, hd * HD, t * 128, b);
                tma_load_3d(sDO + t * TILE, &tmDO, bar_load, hd * HD, t * 128, b);
            }
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 256);
        tmem_relinquish();
    } else {
        const float* lp = lse + ((int64_t)b * H + hd) * L;
        const float* dp = delta + ((int64_t)b * H + hd) * L;
        if (tid < 256) {
            sL[tid] = tid < L ? __ldg(lp + tid) * LOG2E : INFINITY;
            sD[tid] = tid < L ? __ldg(dp + tid) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == EW_WARPS) {
        const uint32_t klo = (smem_u32(sK) >> 4) | LBO_K, vlo = (smem_u32(sV) >> 4) | LBO_K;
        const uint32_t qlo = (smem_u32(sQ) >> 4) | LBO_K, dolo = (smem_u32(sDO) >> 4) | LBO_K;
        const uint32_t qmn = (smem_u32(sQ) >> 4) | LBO_MN, domn = (smem_u32(sDO) >> 4) | LBO_MN;
        const uint32_t idesc1 = make_idesc_bf16(128, 64, 0, 0);
        const uint32_t idesc2 = make_idesc_bf16(128, 64, 0, 1);
        auto mma1 = [&](int qc) {  // S^T = K Q_qc^T, dP^T = V dO_qc^T  (64 queries = 8 KB of rows = 512 descriptor units)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_base + B_COL_S, make_desc(klo + 2 * k, DESC_HI), make_desc(qlo + qc * 512 + 2 * k, DESC_HI), idesc1, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_ss(tmem_base + B_COL_DP, make_desc(vlo + 2 * k, DESC_HI), make_desc(dolo + qc * 512 + 2 * k, DESC_HI), idesc1, k > 0);
            umma_commit(bar_s);
        };
        mbar_wait(bar_load, 0, 50);
        tc_fence_after();
        if (elect_one()) mma1(0);
        __syncwarp();
        for (int qc = 0; qc < 4; ++qc) {
            mbar_wait(bar_p, qc & 1, 51);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t acol = (k >> 1) * 32 + (k & 1) * 8;
                    const uint32_t boff = qc * 512 + k * 128;  // 16 query rows of dO / Q as [K = query][N = d]
                    umma_bf16_ts(tmem_base + B_COL_ACC, tmem_base + B_COL_S + acol, make_desc(domn + boff, DESC_HI), idesc2, (qc > 0 || k > 0));
                    umma_bf16_ts(tmem_base + B_COL_ACC + 64, tmem_base + B_COL_DP + acol, make_desc(qmn + boff, DESC_HI), idesc2, (qc > 0 || k > 0));
                }
                if (qc < 3)
                    mma1(qc + 1);
                else
                    umma_commit(bar_done);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3, hf = warp >> 2;
        const int row = quarter * 32 + lane;
        const int key = kt * 128 + row;
        const bool key_ok = key < L;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        for (int qc = 0; qc < 4; ++qc) {
            mbar_wait(bar_s, qc & 1, 52);
            tc_fence_after();
            {
                const int c0 = hf * 32;
                uint32_t sv[32], dv[32], pp[16], pd[16];
                tmem_ld_32x32b_x32(lane_addr + B_COL_S + c0, sv);
                tmem_ld_32x32b_x32(lane_addr + B_COL_DP + c0, dv);
                const float* lq = sL + qc * 64 + c0;  // same address for the whole warp: smem broadcast
                const float* dq = sD + qc * 64 + c0;
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 l4 = lds128(lq + 4 * g), d4 = lds128(dq + 4 * g);
                    const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
                    float p[4], ds[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        p[i] = key_ok ? fast_ex2(__uint_as_float(sv[4 * g + i]) * c - lv[i]) : 0.f;
                        ds[i] = p[i] * (__uint_as_float(dv[4 * g + i]) - dl[i]) * 0.125f;
                    }
                    pp[2 * g] = pack_bf16x2(p[0], p[1]);
                    pp[2 * g + 1] = pack_bf16x2(p[2], p[3]);
                    pd[2 * g] = pack_bf16x2(ds[0], ds[1]);
                    pd[2 * g + 1] = pack_bf16x2(ds[2], ds[3]);
                }
                tmem_st_32x32b_x16(lane_addr + B_COL_S + c0, pp);
                tmem_st_32x32b_x16(lane_addr + B_COL_DP + c0, pd);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p);
        }
        mbar_wait(bar_done, 0, 53);
        tc_fence_after();
        // warps 0-3 drain dV, warps 4-7 drain dK (64 columns each)
        uint32_t a[32], a2[32];
        tmem_ld_32x32b_x32(lane_addr + B_COL_ACC + hf * 64, a);
        tmem_ld_32x32b_x32(lane_addr + B_COL_ACC + hf * 64 + 32, a2);
        tmem_ld_wait();
        if (key_ok) {
            bf16* dst = dqkv + ((int64_t)b * L + key) * ld3 + (hf == 0 ? 2 * E : E) + hd * HD;
            store_32cols_bf16(dst, a, 1.f);
            store_32cols_bf16(dst + 32, a2, 1.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EW_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

template <typename K>
static int set_smem(K kern, int bytes, bool& done) {
    if (!done) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        done = true;
    }
    return VB_OK;
}

static int make_maps(CUtensorMap* tmQKV, const bf16* qkv, CUtensorMap* tmDO, const bf16* dout, int batch, int L, int H) {
    const int64_t E = (int64_t)H * HD;
    int rc = make_tensor_map_3d(tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * E, L, batch, 3 * E * 2, (uint64_t)L * 3 * E * 2, 64,
                                128, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc || tmDO == nullptr) return rc;
    return make_tensor_map_3d(tmDO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, E, L, batch, E * 2, (uint64_t)L * E * 2, 64, 128, 1,
                              CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace attn2

int launch_attention_fwd_tc2(const bf16* qkv, bf16* out, float* lse, int batch, int L, int H, cudaStream_t stream) {
    using namespace attn2;
    CUtensorMap tmQKV;
    int rc = make_maps(&tmQKV, qkv, nullptr, nullptr, batch, L, H);
    if (rc) return rc;
    static bool done = false;
    rc = set_smem(attention_fwd_tc_kernel, FWD_SMEM, done);
    if (rc) return rc;
    static const bool dbg_on = getenv("VITB200_DBG_TIMING") != nullptr;  // development only
    if (dbg_on) {
        const int nblk = batch * H * 2;
        long long* dbg = nullptr;
        VB_CHECK_CUDA(cudaMallocManaged(&dbg, (size_t)nblk * 16 * sizeof(long long)));
        VB_CHECK_CUDA(cudaMemset(dbg, 0, (size_t)nblk * 16 * sizeof(long long)));
        attention_fwd_tc_kernel<<<nblk, THREADS, FWD_SMEM, stream>>>(tmQKV, out, lse, L, H, dbg);
        VB_CHECK_CUDA(cudaStreamSynchronize(stream));
        const char* names[13] = {"start", "mma:pre-load-wait", "mma:loaded", "mma:S issued", "mma:P ready", "mma:O issued",
                                 "ew:pre S wait", "ew:S ready", "ew:max done", "ew:P stored", "ew:O ready", "ew:stored", "end"};
        for (int blk : {0, nblk / 2, nblk - 3}) {
            printf("[fwd timing] block %d:", blk);
            for (int i = 0; i < 13; ++i) printf(" %s=%lld", names[i], dbg[(size_t)blk * 16 + i] - dbg[(size_t)blk * 16]);
            printf("\n");
        }
        cudaFree(dbg);
        return VB_OK;
    }
    attention_fwd_tc_kernel<<<batch * H * 2, THREADS, FWD_SMEM, stream>>>(tmQKV, out, lse, L, H, nullptr);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// delta: caller workspace, f32 [batch, heads, L]
int launch_attention_bwd_tc2(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                             int batch, int L, int H, cudaStream_t stream) {
    using namespace attn2;
    CUtensorMap tmQKV, tmDO;
    int rc = make_maps(&tmQKV, qkv, &tmDO, dout, batch, L, H);
    if (rc) return rc;
    static bool done1 = false, done2 = false;
    rc = set_smem(attention_bwd_dq_tc_kernel, BWD_SMEM, done1);
    if (rc) return rc;
    rc = set_smem(attention_bwd_dkdv_tc_kernel, BWD_SMEM, done2);
    if (rc) return rc;
    attention_bwd_dq_tc_kernel<<<batch * H * 2, THREADS, BWD_SMEM, stream>>>(tmQKV, tmDO, out, dout, lse, delta, dqkv, L, H);
    VB_CHECK_LAUNCH();
    attention_bwd_dkdv_tc_kernel<<<batch * H * 2, THREADS, BWD_SMEM, stream>>>(tmQKV, tmDO, lse, delta, dqkv, L, H);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

}  // namespace vb
