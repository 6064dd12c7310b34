// Persistent, software-pipelined attention forward / backward on tcgen05 for L <= 208 tokens, head_dim = 64
// (ViT-B/L at 224x224: L = 197). One CTA per SM walks over (image, head) items:
//
//   warp 16      TMA producer: the item's Q/K/V(/dO) rows as 208-row SWIZZLE_128B operand tiles (two 104-row boxes
//                each; rows >= L are zero-filled by the TMA unit), double-buffered: item i+1 loads while item i computes
//   warp 17      MMA issuer (one elected lane): runs ahead of the math warps, so the tensor pipe works on the next
//                scores while the math warps exponentiate the current ones
//   warps 0..15  math: TMEM -> registers -> TMEM (4 warps per scheduler: the exp2 / FMA chains of one warp hide behind
//                the others). Probabilities never touch shared memory: they are packed to bf16 over the fp32 scores
//                and consumed as the A operand (from TMEM) of the next MMA.
//
// forward, per item two units (query tiles of 128 rows): S = Q_t K^T (N = 208) -> exact softmax, one TMEM read, the row
//   split over 4 warps (64+48+48+48 columns held in registers) -> O = P V.
//   TMEM: S/P buffers [0,208) and [208,416) (unit parity), O accumulator [416,480).
// backward: attention_bwd_kd_kernel (key-domain schedule, described at its definition below). The two-domain kernel of round 1
//   (S / dP computed in the query AND the key domain, 690 -> 525 us) was removed in round 2; git history has it (80de1b7).
//   Measured (tools/microbench/mma_rate.cu): a 128xNx16 MMA costs ~44 + N/2 cycles with A in smem, ~12 + N/2 with A in
//   TMEM, so N = 64 tiles run the tensor pipe at < 50 % of its rate: head_dim 64 bounds these kernels, not HBM.
#include <stdlib.h>

#include <type_traits>

#include "host_utils.h"
#define VB_MBAR_TRAP_PRINTF 0  // no CALL in these kernels: per-role register budgets (setmaxnreg), see ptx.cuh
#include "ptx.cuh"

namespace vb {
namespace attn3 {

constexpr int HD = 64;
constexpr int ROWS = 208;                    // operand rows staged per item (13 x 16)
constexpr int BOX_ROWS = 104;                // 104 x 128 B = 13 KB per TMA box, a multiple of the 1024-B swizzle atom
constexpr int OPER_BYTES = ROWS * 128;       // 26624
constexpr int TILE_BYTES = 128 * 128;        // second 128-row tile of an operand starts here
constexpr int MATH_WARPS = 16;
constexpr int WARP_TMA = 16, WARP_MMA = 17;
constexpr int THREADS = 18 * 32;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B | version 1 | SWIZZLE_128B
constexpr uint32_t LBO_K = 1u << 16;                                  // K-major operands: LBO unused (16 B)
constexpr uint32_t LBO_MN = (8192u >> 4) << 16;                       // MN-major, one 64-wide chunk: unused as well

__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
        "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// The registers of an asynchronous tcgen05.ld are defined only after tcgen05.wait::ld: pin every later use behind it.
template <int N>
__device__ __forceinline__ void reg_fence(uint32_t (&r)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ uint8_t* align_smem(uint8_t* raw) {
    return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// 32 fp32 accumulator columns -> 32 bf16 -> 64 contiguous bytes (dst 32-byte aligned), as two 256-bit stores. Every lane
// writes its own row, so one store instruction touches 32 different 128-byte lines whatever its width: 256-bit stores halve
// the LSU wavefronts of the accumulator drain, which share the SM's memory pipeline with the math warps' lse / delta loads
// and dS^T stores (a math phase that overlaps a drain takes 3 300 - 4 800 cycles instead of 900, profiles/r04_e_kd_timeline.log).
// c_attn_store128 (measurement only, env VITB200_ATTN_STG128=1): the 128-bit stores these kernels used before, for A/B runs
// of the same binary on the same box.
__constant__ int c_attn_store128 = 0;
__device__ __forceinline__ void store_32cols_bf16(bf16* dst, const uint32_t (&a)[32], float mul) {
    if (c_attn_store128) {
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
        return;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint32_t u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(__uint_as_float(a[16 * i + 2 * j]) * mul, __uint_as_float(a[16 * i + 2 * j + 1]) * mul);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 16 * i), "r"(u[0]), "r"(u[1]), "r"(u[2]),
                     "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                     : "memory");
    }
}

// 16 bf16 (32 contiguous bytes, dst 32-byte aligned) as ONE 256-bit store (see store_32cols_bf16)
__device__ __forceinline__ void store_16cols_packed(bf16* dst, const uint32_t (&u)[8]) {
    if (c_attn_store128) {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        d4[0] = make_uint4(u[0], u[1], u[2], u[3]);
        d4[1] = make_uint4(u[4], u[5], u[6], u[7]);
        return;
    }
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
                 "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                 : "memory");
}

// dst[c] += sum over the warp's 32 lanes (rows) of a[c], c = 0..31: butterfly transpose-reduce (31 shuffles), lane c ends
// with column c's total and issues one reduction. Rows that must not count are excluded with row_ok.
__device__ __forceinline__ void warp_colsum32_atomic(const uint32_t (&a)[32], bool row_ok, float* dst, int lane) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = row_ok ? __uint_as_float(a[i]) : 0.f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = hi ? v[i] : v[i + off];
            const float keep = hi ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    atomicAdd(dst + lane, v[0]);
}

// =====================================================================================================
// forward
// =====================================================================================================
constexpr int F_STAGE = 3 * OPER_BYTES;  // Q, K, V
constexpr int F_EXCH_FLOATS = 2 * 4 * 128;  // sMax / sSum: [unit parity][column part][row]
constexpr int F_SMEM = 2 * F_STAGE + 2 * F_EXCH_FLOATS * 4 + 256 + 1024;
constexpr uint32_t F_COL_O = 416;

// One row's share of the softmax: NC (64 or 48) scores at TMEM columns [cbeg, cbeg + NC) of the S buffer, held in
// registers across the row-max exchange, exponentiated and packed to bf16 pairs at columns [cbeg / 2, +NC / 2).
template <int NC>
__device__ __forceinline__ void softmax_part(uint32_t s_buf, int cbeg, int nvalid, float c, float* sMaxU, float* sSumU, int part,
                                             int row, int quarter, float& m_out, float& sum_out) {
    uint32_t v0[32], v1[NC - 32];
    tmem_ld_32x32b_x32(s_buf + cbeg, v0);
    if constexpr (NC == 64)
        tmem_ld_32x32b_x32(s_buf + cbeg + 32, v1);
    else
        tmem_ld_32x32b_x16(s_buf + cbeg + 32, v1);
    tmem_ld_wait();
    reg_fence(v0);
    reg_fence(v1);
    float m = -INFINITY;
    if (nvalid >= NC) {
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v0[i]));
#pragma unroll
        for (int i = 0; i < NC - 32; ++i) m = fmaxf(m, __uint_as_float(v1[i]));
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i < nvalid) m = fmaxf(m, __uint_as_float(v0[i]));
#pragma unroll
        for (int i = 0; i < NC - 32; ++i)
            if (32 + i < nvalid) m = fmaxf(m, __uint_as_float(v1[i]));
    }
    sMaxU[part * 128 + row] = m;
    named_bar_sync(1 + quarter, 128);  // the four warps sharing this lane quarter: all their S columns are in registers now
    m = fmaxf(fmaxf(sMaxU[row], sMaxU[128 + row]), fmaxf(sMaxU[256 + row], sMaxU[384 + row]));  // column 0 is valid: finite
    const float mc = m * c;
    uint32_t pk0[16], pk1[(NC - 32) / 2];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (nvalid >= NC) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float a = fast_ex2(fmaf(__uint_as_float(v0[2 * i]), c, -mc)), b = fast_ex2(fmaf(__uint_as_float(v0[2 * i + 1]), c, -mc));
            pk0[i] = pack_bf16x2(a, b);
            if (i & 1) s2 += a, s3 += b; else s0 += a, s1 += b;
        }
#pragma unroll
        for (int i = 0; i < (NC - 32) / 2; ++i) {
            const float a = fast_ex2(fmaf(__uint_as_float(v1[2 * i]), c, -mc)), b = fast_ex2(fmaf(__uint_as_float(v1[2 * i + 1]), c, -mc));
            pk1[i] = pack_bf16x2(a, b);
            if (i & 1) s2 += a, s3 += b; else s0 += a, s1 += b;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float a = (2 * i < nvalid) ? fast_ex2(fmaf(__uint_as_float(v0[2 * i]), c, -mc)) : 0.f;
            const float b = (2 * i + 1 < nvalid) ? fast_ex2(fmaf(__uint_as_float(v0[2 * i + 1]), c, -mc)) : 0.f;
            pk0[i] = pack_bf16x2(a, b);
            s0 += a, s1 += b;
        }
#pragma unroll
        for (int i = 0; i < (NC - 32) / 2; ++i) {
            const float a = (32 + 2 * i < nvalid) ? fast_ex2(fmaf(__uint_as_float(v1[2 * i]), c, -mc)) : 0.f;
            const float b = (32 + 2 * i + 1 < nvalid) ? fast_ex2(fmaf(__uint_as_float(v1[2 * i + 1]), c, -mc)) : 0.f;
            pk1[i] = pack_bf16x2(a, b);
            s2 += a, s3 += b;
        }
    }
    tmem_st_x16(s_buf + (cbeg >> 1), pk0);
    if constexpr (NC == 64)
        tmem_st_x16(s_buf + (cbeg >> 1) + 16, pk1);
    else
        tmem_st_x8(s_buf + (cbeg >> 1) + 16, pk1);
    const float sum = (s0 + s1) + (s2 + s3);
    sSumU[part * 128 + row] = sum;
    m_out = m;
    sum_out = sum;
}

// PAIR = false: out = attention(qkv), lse written (training forward).
// PAIR = true (plasticity estimator): every item is run on two inputs, tmQKV (a) then tmQKV2 (b), and
// out = attention(a) - attention(b), subtracted in fp32 (the normalised rows of a wait in registers) before the single
// bf16 rounding. Items then also range over `layers`: the qkv tensors hold all layers' projections side by side
// (feature = layer * 3E + which * E + head * 64) and out is [layers][batch * L][E].
template <bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQKV2,
                                bf16* __restrict__ out, float* __restrict__ lse_out, int L, int H, int batch, int n_items) {
    constexpr int SUBS = PAIR ? 2 : 1;  // smem stages / TMEM units are per (item, input)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    float* sMax = reinterpret_cast<float*>(smem + 2 * F_STAGE);  // [2][4][128]
    float* sSum = sMax + F_EXCH_FLOATS;                          // [2][4][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + F_EXCH_FLOATS);
    uint64_t *full = bars, *empty = bars + 2, *s_ready = bars + 4, *p_ready = bars + 6, *o_ready = bars + 8, *o_free = bars + 9;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
    const int E = H * HD;
    const int n_local = SUBS * ((n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);  // (item, input) stages
    const int U = 2 * n_local;  // units (query tiles) this CTA processes

    if (warp == WARP_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        if (PAIR) tma_prefetch_desc(&tmQKV2);
    }
    if (warp == WARP_MMA) {
        if (elect_one()) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
                mbar_init(&s_ready[i], 1);
                mbar_init(&p_ready[i], MATH_WARPS);
            }
            mbar_init(o_ready, 1);
            mbar_init(o_free, MATH_WARPS);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == WARP_TMA) {
        // =========================== TMA producer ===========================
        for (int n = 0; n < n_local; ++n) {
            const int it = blockIdx.x + (n / SUBS) * gridDim.x;
            const int hd = it % H, b = (it / H) % batch, layer = it / (H * batch);
            const CUtensorMap* tm = (PAIR && (n & 1)) ? &tmQKV2 : &tmQKV;
            const int s = n & 1;
            mbar_wait(&empty[s], ((n >> 1) & 1) ^ 1, 60);
            if (elect_one()) {
                uint8_t* st = smem + s * F_STAGE;
                mbar_arrive_expect_tx(&full[s], F_STAGE);
#pragma unroll
                for (int op = 0; op < 3; ++op)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2)
                        tma_load_3d(st + op * OPER_BYTES + h2 * BOX_ROWS * 128, tm, &full[s], (layer * 3 + op) * E + hd * HD, h2 * BOX_ROWS, b);
            }
            __syncwarp();
        }
    } else if (warp == WARP_MMA) {
        // =========================== MMA issuer ===========================
        // iteration u: O(u-1) = P(u-1) V, then S(u+1) = Q K^T into the buffer P(u-1) just vacated (in-order tensor pipe)
        const uint32_t idesc_s = make_idesc_bf16(128, ROWS, 0, 0);
        const uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
        const uint32_t smem_lo = smem_u32(smem) >> 4;
        for (int u = -1; u <= U; ++u) {
            const int v = u - 1;
            if (v >= 0 && v < U) {
                const int n = v >> 1, t = v & 1, s = n & 1;
                mbar_wait(&p_ready[t], (v >> 1) & 1, 61);
                mbar_wait(o_free, (v & 1) ^ 1, 62);  // O of unit v-1 has been read out (passes at once for v = 0)
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t vlo = (smem_lo + ((s * F_STAGE + 2 * OPER_BYTES) >> 4)) | LBO_MN;
#pragma unroll
                    for (int k = 0; k < 13; ++k)  // 16 keys per step = 8 packed columns
                        umma_bf16_ts(tmem_base + F_COL_O, tmem_base + t * ROWS + k * 8, make_desc(vlo + k * 128, DESC_HI), idesc_o, k > 0);
                    umma_commit(o_ready);
                    if (t == 1) umma_commit(&empty[s]);  // both tiles of the item are done with this smem stage
                }
                __syncwarp();
            }
            const int w = u + 1;
            if (w >= 0 && w < U) {
                const int n = w >> 1, t = w & 1, s = n & 1;
                if (t == 0) mbar_wait(&full[s], (n >> 1) & 1, 63);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t qlo = (smem_lo + ((s * F_STAGE + t * TILE_BYTES) >> 4)) | LBO_K;
                    const uint32_t klo = (smem_lo + ((s * F_STAGE + OPER_BYTES) >> 4)) | LBO_K;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + t * ROWS, make_desc(qlo + 2 * k, DESC_HI), make_desc(klo + 2 * k, DESC_HI), idesc_s, k > 0);
                    umma_commit(&s_ready[t]);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== softmax / epilogue warps ===========================
        const int quarter = warp & 3, part = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        const int cbeg = part == 0 ? 0 : 16 + 48 * part;  // 0, 64, 112, 160
        float m_prev = 0.f, sum_prev = 0.f;
        float okeep[PAIR ? 2 : 1][16];  // PAIR: normalised output rows of input a, per query tile, until input b's arrive
        for (int u = 0; u <= U; ++u) {
            float m_cur = 0.f, sum_cur = 0.f;
            if (u < U) {
                const int t = u & 1;
                mbar_wait(&s_ready[t], (u >> 1) & 1, 64);
                tc_fence_after();
                if (part == 0)
                    softmax_part<64>(lane_addr + t * ROWS, cbeg, L - cbeg, c, sMax + t * 512, sSum + t * 512, part, row, quarter, m_cur, sum_cur);
                else
                    softmax_part<48>(lane_addr + t * ROWS, cbeg, L - cbeg, c, sMax + t * 512, sSum + t * 512, part, row, quarter, m_cur, sum_cur);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[t]);
            }
            if (u >= 1) {
                // ---- read out O of unit u-1 (its P V ran while the softmax of unit u was computed) ----
                const int v = u - 1, t = v & 1;
                const int it = blockIdx.x + ((v >> 1) / SUBS) * gridDim.x;
                const int hd = it % H, b = (it / H) % batch, layer = it / (H * batch);
                mbar_wait(o_ready, v & 1, 65);
                tc_fence_after();
                uint32_t o[16];
                tmem_ld_32x32b_x16(lane_addr + F_COL_O + part * 16, o);
                // o_ready implies every warp arrived on p_ready(v), i.e. wrote its partial sum before (release / acquire chain)
                const float* sS = sSum + t * 512;
                const float tot = (sS[row] + sS[128 + row]) + (sS[256 + row] + sS[384 + row]);
                tmem_ld_wait();
                reg_fence(o);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                const int q = t * 128 + row;
                const float inv = 1.f / tot;
                float r[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(o[i]) * inv;
                bool store = q < L;
                if (PAIR) {
                    if (((v >> 1) & 1) == 0) {  // input a: keep, nothing to store yet
#pragma unroll
                        for (int i = 0; i < 16; ++i) okeep[t][i] = r[i];
                        store = false;
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = okeep[t][i] - r[i];
                    }
                }
                if (store) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(r[2 * i], r[2 * i + 1]);
                    store_16cols_packed(out + (((int64_t)layer * batch + b) * L + q) * E + hd * HD + part * 16, w);
                    if (!PAIR && part == 0 && lse_out != nullptr) lse_out[((int64_t)b * H + hd) * L + q] = m_prev * 0.125f + __logf(tot);
                }
            }
            m_prev = m_cur;
            sum_prev = sum_cur;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================
// plasticity estimator: attention of a perturbed input, attn(a + d) - attn(a), evaluated in PERTURBATION FORM
// =====================================================================================================
// Inputs per (layer, image, head): q, k, v of the base input a (qkv_a = W t_a + b) and dq, dk, dv = W d_tok of the token
// difference (no bias). Two independent bf16 forward passes lose the difference once |d| << |a| (34-53 % error at a
// relative perturbation of 1e-2); here every small quantity is its own bf16 tensor with its own exponent:
//   S   = q k^T                                   (SS MMA, N = 208)              -> buffer A
//   dS  = q dk^T + dq k^T + dq dk^T  (= S_b - S)  (3 x SS MMA into one accumulator) -> buffer B
//   p   = exp2((S - rowmax S) c),  l = sum p
//   g   = p * expm1((dS - rowmax dS) / 8)         (the softmax is shift invariant per row: the row max removes the
//                                                  common mode of dS; expm1 by a polynomial above -1/8, so p_b - p = g
//                                                  keeps full relative precision however small dS is). g is parked in
//                                                  fp32 over the dS columns it came from until the row sums are known.
//   dl  = sum g,  lb = sum (p + g)                (P_a = p / l,  P_b = (p + g) / lb)
//   H   = P_b - P_a = (g - (dl / l) p) / lb       fp32, THEN one bf16 rounding: relative precision at any perturbation size
//   out = O_b - O_a = H v + H dv + P_a dv         (TS MMAs: P_a and H packed to bf16 over buffer A, accumulator over B)
// Uniformly accurate (tools/emulate_delta_attention.py: <= 3e-3 of the fp64 difference per layer and <= 5e-3 per token row
// from |d| / |a| = 1e-4 to 40, where the perturbed softmax is one-hot on a key that had negligible weight before).
// One unit (128 query rows) at a time: TMEM holds S (208) + dS (208) columns, so units are not double-buffered; the
// Q/K and V operand groups have their own barriers, so the next item's Q/K load overlaps this item's softmax and P V.
constexpr int D_OFF_QA = 0, D_OFF_KA = OPER_BYTES, D_OFF_DQ = 2 * OPER_BYTES, D_OFF_DK = 3 * OPER_BYTES, D_OFF_VA = 4 * OPER_BYTES,
              D_OFF_DV = 5 * OPER_BYTES;
constexpr int D_EXCH_FLOATS = 4 * 128;  // one value per (column part, row)
constexpr int D_SMEM = 6 * OPER_BYTES + 5 * D_EXCH_FLOATS * 4 + 256 + 1024;
static_assert(D_SMEM <= 232448, "shared memory budget");
constexpr uint32_t D_COL_A = 0, D_COL_H = 104, D_COL_B = 208, D_COL_ACC = 224;

// exp(w) - 1 for w <= 0: 4th-order polynomial above -1/8 (truncation 3e-7 relative), exp2 - 1 below
__device__ __forceinline__ float expm1_neg(float w) {
    const float poly = w * fmaf(w, fmaf(w, fmaf(w, 1.f / 24.f, 1.f / 6.f), 0.5f), 1.f);
    const float big = fast_ex2(w * LOG2E) - 1.f;
    return w > -0.125f ? poly : big;
}

__device__ __forceinline__ void tmem_st_x4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}
template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[W]) {
    if constexpr (W == 16)
        tmem_ld_32x32b_x16(taddr, r);
    else
        tmem_ld_32x32b_x8(taddr, r);
}
template <int W>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&r)[W]) {
    if constexpr (W == 16)
        tmem_st_x16(taddr, r);
    else if constexpr (W == 8)
        tmem_st_x8(taddr, r);
    else
        tmem_st_x4(taddr, r);
}

// One row's share of a unit. The 208 score columns are 13 pieces of 16; piece q belongs to column part q & 3, and part 0
// also owns the tail piece 12 (columns 192..207, of which L - 192 hold keys: 5 for 197 tokens), processed 8 columns at a
// time and only as far as keys exist: every part has 48 columns of work, part 0 eight more (the contiguous 64/48/48/48 split
// made the other twelve warps wait for part 0). Leaves packed bf16 P_a at buffer-A columns [8 q, 8 q + 8) and packed H at
// [104 + 8 q, ...) for each of its pieces.
template <bool TAIL>
__device__ __forceinline__ void delta_softmax_part(uint32_t lane_addr, int L, float* sMaxA, float* sMaxW, float* sSumA, float* sSumG,
                                                   float* sSumB, int part, int row, int quarter) {
    const float c = 0.125f * LOG2E;
    const int nvt = TAIL ? L - 192 : 0;  // keys in the tail piece
    uint32_t v[3][16], vt[TAIL ? 16 : 1];
#pragma unroll
    for (int k = 0; k < 3; ++k) tmem_ld_32x32b_x16(lane_addr + D_COL_A + 16 * (part + 4 * k), v[k]);
    if constexpr (TAIL) tmem_ld_32x32b_x16(lane_addr + D_COL_A + 192, vt);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 3; ++k) reg_fence(v[k]);
    if constexpr (TAIL) reg_fence(vt);
    float m = -INFINITY, mw = -INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int nv = L - 16 * (part + 4 * k);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nv) m = fmaxf(m, __uint_as_float(v[k][i]));
    }
    if constexpr (TAIL) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nvt) m = fmaxf(m, __uint_as_float(vt[i]));
    }
    auto max_piece = [&](auto wtag, int col, int nv) {  // running max of the valid dS columns [col, col + W)
        constexpr int W = decltype(wtag)::value;
        uint32_t w[W];
        tmem_ld_cols<W>(lane_addr + D_COL_B + col, w);
        tmem_ld_wait();
        reg_fence(w);
#pragma unroll
        for (int i = 0; i < W; ++i)
            if (i < nv) mw = fmaxf(mw, __uint_as_float(w[i]));
    };
    using W16 = std::integral_constant<int, 16>;
    using W8 = std::integral_constant<int, 8>;
#pragma unroll
    for (int k = 0; k < 3; ++k) max_piece(W16{}, 16 * (part + 4 * k), L - 16 * (part + 4 * k));
    if constexpr (TAIL) {
        if (nvt > 0) max_piece(W8{}, 192, nvt);
        if (nvt > 8) max_piece(W8{}, 200, nvt - 8);
    }
    sMaxA[part * 128 + row] = m;
    sMaxW[part * 128 + row] = mw;
    named_bar_sync(1 + quarter, 128);  // the four warps sharing this lane quarter: all their S columns are in registers now
    m = fmaxf(fmaxf(sMaxA[row], sMaxA[128 + row]), fmaxf(sMaxA[256 + row], sMaxA[384 + row]));  // column 0 is valid: finite
    mw = fmaxf(fmaxf(sMaxW[row], sMaxW[128 + row]), fmaxf(sMaxW[256 + row], sMaxW[384 + row]));
    const float mc = m * c;
    // ---- p = exp2((S - m) c), kept in fp32 in the registers that held S ----
    float la = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int nv = L - 16 * (part + 4 * k);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float a = (i < nv) ? fast_ex2(fmaf(__uint_as_float(v[k][i]), c, -mc)) : 0.f;
            v[k][i] = __float_as_uint(a);
            la += a;
        }
    }
    if constexpr (TAIL) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float a = (i < nvt) ? fast_ex2(fmaf(__uint_as_float(vt[i]), c, -mc)) : 0.f;
            vt[i] = __float_as_uint(a);
            la += a;
        }
    }
    // ---- g = p * expm1((dS - mw) / 8), parked in fp32 over the dS columns it came from (this thread's own columns) ----
    const float mw8 = mw * 0.125f;
    float dl = 0.f, lb = 0.f;
    auto make_g = [&](auto wtag, auto& pv, int base, int col) {  // pv[base .. base + W) hold p of columns [col, col + W)
        constexpr int W = decltype(wtag)::value;
        uint32_t w[W];
        tmem_ld_cols<W>(lane_addr + D_COL_B + col, w);
        tmem_ld_wait();
        reg_fence(w);
        float wmin = 0.f;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            // the clamp keeps masked columns (p = 0, dS = 0 there) at expm1(0) = 0 instead of 0 * inf
            const float x = fminf(fmaf(__uint_as_float(w[i]), 0.125f, -mw8), 0.f);
            w[i] = __float_as_uint(x);
            wmin = fminf(wmin, x);
        }
        // small perturbations (the usual case of a sweep): the whole warp is in the polynomial range, no exp2 at all
        const bool poly_only = __all_sync(0xffffffffu, wmin > -0.125f);
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const float x = __uint_as_float(w[i]), p0 = __uint_as_float(pv[base + i]);
            const float poly = x * fmaf(x, fmaf(x, fmaf(x, 1.f / 24.f, 1.f / 6.f), 0.5f), 1.f);
            const float e = poly_only ? poly : (x > -0.125f ? poly : fast_ex2(x * LOG2E) - 1.f);
            const float g0 = p0 * e;
            w[i] = __float_as_uint(g0);
            dl += g0;
            lb += p0 + g0;
        }
        tmem_st_cols<W>(lane_addr + D_COL_B + col, w);
    };
#pragma unroll
    for (int k = 0; k < 3; ++k) make_g(W16{}, v[k], 0, 16 * (part + 4 * k));
    if constexpr (TAIL) {
        if (nvt > 0) make_g(W8{}, vt, 0, 192);
        if (nvt > 8) make_g(W8{}, vt, 8, 200);
    }
    sSumA[part * 128 + row] = la;
    sSumG[part * 128 + row] = dl;
    sSumB[part * 128 + row] = lb;
    tmem_st_wait();  // g is re-read below by this same thread
    named_bar_sync(5 + quarter, 128);
    la = (sSumA[row] + sSumA[128 + row]) + (sSumA[256 + row] + sSumA[384 + row]);
    dl = (sSumG[row] + sSumG[128 + row]) + (sSumG[256 + row] + sSumG[384 + row]);
    lb = (sSumB[row] + sSumB[128 + row]) + (sSumB[256 + row] + sSumB[384 + row]);
    const float inv_la = 1.f / la, inv_lb = 1.f / fmaxf(lb, 1e-37f);
    const float rho = dl * inv_la;
    // ---- P_a = p / la and H = (g - rho p) / lb, packed to bf16 over buffer A (every S column of this quarter is in
    //      registers since the first barrier) ----
    auto pack_h = [&](auto wtag, auto& pv, int base, int col) {
        constexpr int W = decltype(wtag)::value;
        uint32_t w[W], pa[W / 2], ph[W / 2];
        tmem_ld_cols<W>(lane_addr + D_COL_B + col, w);
        tmem_ld_wait();
        reg_fence(w);
#pragma unroll
        for (int i = 0; i < W / 2; ++i) {
            const float p0 = __uint_as_float(pv[base + 2 * i]), p1 = __uint_as_float(pv[base + 2 * i + 1]);
            pa[i] = pack_bf16x2(p0 * inv_la, p1 * inv_la);
            ph[i] = pack_bf16x2(fmaf(-rho, p0, __uint_as_float(w[2 * i])) * inv_lb, fmaf(-rho, p1, __uint_as_float(w[2 * i + 1])) * inv_lb);
        }
        tmem_st_cols<W / 2>(lane_addr + D_COL_A + (col >> 1), pa);
        tmem_st_cols<W / 2>(lane_addr + D_COL_H + (col >> 1), ph);
    };
#pragma unroll
    for (int k = 0; k < 3; ++k) pack_h(W16{}, v[k], 0, 16 * (part + 4 * k));
    if constexpr (TAIL) {
        // halves of the tail piece without keys: zero operands (the fp32 scores left there must not reach the MMA as bf16)
        const uint32_t zero[4] = {0u, 0u, 0u, 0u};
        if (nvt > 0) {
            pack_h(W8{}, vt, 0, 192);
        } else {
            tmem_st_x4(lane_addr + D_COL_A + 96, zero);
            tmem_st_x4(lane_addr + D_COL_H + 96, zero);
        }
        if (nvt > 8) {
            pack_h(W8{}, vt, 8, 200);
        } else {
            tmem_st_x4(lane_addr + D_COL_A + 100, zero);
            tmem_st_x4(lane_addr + D_COL_H + 100, zero);
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1)
attention_perturb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmD, bf16* __restrict__ out, int L,
                         int H, int batch, int n_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    float* sMaxA = reinterpret_cast<float*>(smem + 6 * OPER_BYTES);  // [4][128] each
    float* sMaxW = sMaxA + D_EXCH_FLOATS;
    float* sSumA = sMaxW + D_EXCH_FLOATS;
    float* sSumG = sSumA + D_EXCH_FLOATS;
    float* sSumB = sSumG + D_EXCH_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSumB + D_EXCH_FLOATS);
    uint64_t *full_qk = bars, *full_v = bars + 1, *empty_qk = bars + 2, *empty_v = bars + 3, *s_ready = bars + 4, *p_ready = bars + 5,
             *o_ready = bars + 6, *o_free = bars + 7;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
    const int E = H * HD;
    const int n_local = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int U = 2 * n_local;  // units (query tiles) this CTA processes

    if (warp == WARP_TMA && elect_one()) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmD);
    }
    if (warp == WARP_MMA) {
        if (elect_one()) {
            mbar_init(full_qk, 1);
            mbar_init(full_v, 1);
            mbar_init(empty_qk, 1);
            mbar_init(empty_v, 1);
            mbar_init(s_ready, 1);
            mbar_init(p_ready, MATH_WARPS);
            mbar_init(o_ready, 1);
            mbar_init(o_free, MATH_WARPS);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == WARP_TMA) {
        // =========================== TMA producer ===========================
        for (int n = 0; n < n_local; ++n) {
            const int it = blockIdx.x + n * gridDim.x;
            const int hd = it % H, b = (it / H) % batch, layer = it / (H * batch);
            const int f0 = layer * 3 * E + hd * HD;  // feature of q; k at + E, v at + 2E
            mbar_wait(empty_qk, (n & 1) ^ 1, 80);
            if (elect_one()) {
                mbar_arrive_expect_tx(full_qk, 4 * OPER_BYTES);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    tma_load_3d(smem + D_OFF_QA + h2 * BOX_ROWS * 128, &tmA, full_qk, f0, h2 * BOX_ROWS, b);
                    tma_load_3d(smem + D_OFF_KA + h2 * BOX_ROWS * 128, &tmA, full_qk, f0 + E, h2 * BOX_ROWS, b);
                    tma_load_3d(smem + D_OFF_DQ + h2 * BOX_ROWS * 128, &tmD, full_qk, f0, h2 * BOX_ROWS, b);
                    tma_load_3d(smem + D_OFF_DK + h2 * BOX_ROWS * 128, &tmD, full_qk, f0 + E, h2 * BOX_ROWS, b);
                }
            }
            __syncwarp();
            mbar_wait(empty_v, (n & 1) ^ 1, 81);
            if (elect_one()) {
                mbar_arrive_expect_tx(full_v, 2 * OPER_BYTES);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    tma_load_3d(smem + D_OFF_VA + h2 * BOX_ROWS * 128, &tmA, full_v, f0 + 2 * E, h2 * BOX_ROWS, b);
                    tma_load_3d(smem + D_OFF_DV + h2 * BOX_ROWS * 128, &tmD, full_v, f0 + 2 * E, h2 * BOX_ROWS, b);
                }
            }
            __syncwarp();
        }
    } else if (warp == WARP_MMA) {
        // =========================== MMA issuer ===========================
        const uint32_t idesc_s = make_idesc_bf16(128, ROWS, 0, 0);
        const uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
        const uint32_t smem_lo = smem_u32(smem) >> 4;
        const uint32_t ka = (smem_lo + (D_OFF_KA >> 4)) | LBO_K, dk = (smem_lo + (D_OFF_DK >> 4)) | LBO_K;
        const uint32_t va = (smem_lo + (D_OFF_VA >> 4)) | LBO_MN, dv = (smem_lo + (D_OFF_DV >> 4)) | LBO_MN;
        for (int u = 0; u < U; ++u) {
            const int n = u >> 1, t = u & 1;
            if (t == 0) mbar_wait(full_qk, n & 1, 82);
            mbar_wait(o_free, (u & 1) ^ 1, 83);  // the accumulator of unit u-1 (over buffer B) has been read out
            tc_fence_after();
            if (elect_one()) {
                const uint32_t qa = (smem_lo + ((D_OFF_QA + t * TILE_BYTES) >> 4)) | LBO_K;
                const uint32_t dq = (smem_lo + ((D_OFF_DQ + t * TILE_BYTES) >> 4)) | LBO_K;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base + D_COL_A, make_desc(qa + 2 * k, DESC_HI), make_desc(ka + 2 * k, DESC_HI), idesc_s, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base + D_COL_B, make_desc(qa + 2 * k, DESC_HI), make_desc(dk + 2 * k, DESC_HI), idesc_s, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base + D_COL_B, make_desc(dq + 2 * k, DESC_HI), make_desc(ka + 2 * k, DESC_HI), idesc_s, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base + D_COL_B, make_desc(dq + 2 * k, DESC_HI), make_desc(dk + 2 * k, DESC_HI), idesc_s, 1);
                umma_commit(s_ready);
                if (t == 1) umma_commit(empty_qk);  // both query tiles have read Q / K / dQ / dK: the next item may load
            }
            __syncwarp();
            mbar_wait(p_ready, u & 1, 84);
            if (t == 0) mbar_wait(full_v, n & 1, 85);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 13; ++k)  // 16 keys per step = 8 packed columns
                    umma_bf16_ts(tmem_base + D_COL_ACC, tmem_base + D_COL_H + k * 8, make_desc(va + k * 128, DESC_HI), idesc_o, k > 0);
#pragma unroll
                for (int k = 0; k < 13; ++k)
                    umma_bf16_ts(tmem_base + D_COL_ACC, tmem_base + D_COL_H + k * 8, make_desc(dv + k * 128, DESC_HI), idesc_o, 1);
#pragma unroll
                for (int k = 0; k < 13; ++k)
                    umma_bf16_ts(tmem_base + D_COL_ACC, tmem_base + D_COL_A + k * 8, make_desc(dv + k * 128, DESC_HI), idesc_o, 1);
                umma_commit(o_ready);
                if (t == 1) umma_commit(empty_v);
            }
            __syncwarp();
        }
    } else {
        // =========================== softmax / epilogue warps ===========================
        const int quarter = warp & 3, part = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        for (int u = 0; u < U; ++u) {
            const int t = u & 1;
            const int it = blockIdx.x + (u >> 1) * gridDim.x;
            const int hd = it % H, b = (it / H) % batch, layer = it / (H * batch);
            mbar_wait(s_ready, u & 1, 86);
            tc_fence_after();
            if (part == 0)
                delta_softmax_part<true>(lane_addr, L, sMaxA, sMaxW, sSumA, sSumG, sSumB, part, row, quarter);
            else
                delta_softmax_part<false>(lane_addr, L, sMaxA, sMaxW, sSumA, sSumG, sSumB, part, row, quarter);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            // ---- read the accumulator out (16 of the 64 output columns per warp): it already holds O_b - O_a ----
            mbar_wait(o_ready, u & 1, 87);
            tc_fence_after();
            uint32_t o[16];
            tmem_ld_32x32b_x16(lane_addr + D_COL_ACC + part * 16, o);
            tmem_ld_wait();
            reg_fence(o);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
            const int q = t * 128 + row;
            if (q < L) {
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1]));
                store_16cols_packed(out + (((int64_t)layer * batch + b) * L + q) * E + hd * HD + part * 16, w);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================
// backward
// =====================================================================================================
// delta[b, h, q] = sum_d dO[b, q, h, d] * O[b, q, h, d]   (one warp per token row; 8 lanes per head per 128-bit load)
__global__ void attention_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta, int rows,
                                       int L, int H) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int E = H * HD;
    const int b = warp / L, q = warp - b * L;
    const uint4* po = reinterpret_cast<const uint4*>(out + (int64_t)warp * E);
    const uint4* pd = reinterpret_cast<const uint4*>(dout + (int64_t)warp * E);
    for (int c = lane; c < E / 8; c += 32) {
        const uint4 uo = __ldg(po + c), ud = __ldg(pd + c);
        const uint32_t wo[4] = {uo.x, uo.y, uo.z, uo.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fo = unpack_bf16x2(wo[k]), fd = unpack_bf16x2(wd[k]);
            acc += fo.x * fd.x + fo.y * fd.y;
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if ((lane & 7) == 0) delta[((int64_t)b * H + (c >> 3)) * L + q] = acc;
    }
}

constexpr int GROUP_WARPS = 8;
constexpr int B_WARP_TMA = 16, B_WARP_MMA1 = 17, B_WARP_MMA2 = 18;  // (warp 19 idles: roles are assigned per warpgroup)
constexpr int B_WARP_DRAIN0 = 20, DRAIN_WARPS = 4;                   // one warp per TMEM lane quarter
constexpr int B_THREADS = 24 * 32;

// =====================================================================================================
// backward, key-domain schedule ("kd"): S and dP are computed ONCE per (key tile, query chunk), keys on the TMEM lanes
// =====================================================================================================
// Per item: key tiles j = 0, 1 (keys [128 j, 128 j + 128)), query chunks c = 0..3 (64, 64, 64, 16 queries). Chunk (j, c):
//   MMA1   S^T = K_j Q_c^T, dP^T = V_j dO_c^T                       (SS, N = 64 / 16)        -> chunk buffer (2 of them)
//   math   P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - delta_q) / 8, packed bf16 over the fp32 values in TMEM (A operands
//          of MMA2) and dS^T ALSO to shared memory as an MN-major A operand [key][query] (block = two query chunks)
//   MMA2   dV_j += P^T dO_c, dK_j += dS^T Q_c                       (TS)
//   MMA3   after the odd chunk of a block: dQ_block += dS_block K_j  (SS, M = 128 queries, K = the tile's 128 / 80 keys)
// Against the two-domain kernel above this halves the score-shaped MMAs, the exponentials and the TMEM traffic; dQ stays in
// TMEM across both key tiles, dV_j / dK_j are drained after each tile by the drain warpgroup.
// TMEM: chunk buffers [0,128) [128,256) (S^T at +0, dP^T at +64), dV [256,320), dK [320,384), dQ block 0 [384,448), 1 [448,512).
// Shared memory: K, V single-buffered but reloaded per key tile (tile 0 of the next item loads while tile 1 computes);
// Q, dO double-buffered per item; two 32 KB dS^T blocks; lse / delta per item stage.
constexpr int KD_KV_BYTES = 2 * OPER_BYTES;                 // K, V
constexpr int KD_QDO_BYTES = 2 * OPER_BYTES;                // Q, dO (per stage)
constexpr int KD_DS_BLOCK = 2 * 128 * 128;                  // two 64-query MN chunks of [128 keys][128 B]
constexpr int KD_OFF_QDO = KD_KV_BYTES;
constexpr int KD_OFF_DS = KD_OFF_QDO + 2 * KD_QDO_BYTES;
constexpr int KD_OFF_TAIL = KD_OFF_DS + 2 * KD_DS_BLOCK;
constexpr int KD_SMEM = KD_OFF_TAIL + 4096 + 512 + 1024;    // + lse / delta + barriers + alignment slack
static_assert(KD_SMEM <= 232448, "shared memory budget");
constexpr uint32_t KD_COL_DV = 256, KD_COL_DK = 320, KD_COL_DQ = 384;
constexpr int KD_THREADS = 24 * 32;

// EARLY = true (default): ONE fp32 chunk buffer [0,128) that the math warps hand back as soon as its S^T / dP^T columns are in
// their registers, and two packed buffers [128,192) [192,256) (P^T at +0, dS^T at +32, 16 columns per 32 queries) for the A
// operands of MMA2: the next chunk's MMA1 runs while this chunk is still being exponentiated. With EARLY = false the packed
// values overwrite the fp32 ones in place in two chunk buffers, so MMA1 of chunk g + 2 has to wait for MMA2 of chunk g, and
// the math warps spend a third of their time waiting for scores (ncu: 34 % of the stall samples on that one wait).
template <bool EARLY>
__global__ void __launch_bounds__(KD_THREADS, 1)
attention_bwd_kd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmKV0, const __grid_constant__ CUtensorMap tmKV1,
                        const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv,
                        float* __restrict__ dbias, int bias_q_only, int L, int H, int n_items, long long* __restrict__ dbg) {
#define KD_STAMP(g, slot)                                                                                   \
    do {                                                                                                    \
        if (dbg != nullptr && blockIdx.x == 0 && (g) >= 16 && (g) < 80) dbg[((g)-16) * 16 + (slot)] = clock64(); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    float* sL = reinterpret_cast<float*>(smem + KD_OFF_TAIL);  // [2][256] lse * log2(e); +inf for q >= L
    float* sD = sL + 512;                                      // [2][256] delta / 8;     0 for q >= L
    uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 512);
    uint64_t *qdo_full = bars, *qdo_empty = bars + 2, *kv_full = bars + 4, *kv_empty = bars + 6, *s_ready = bars + 8,
             *p_ready = bars + 10, *c_free = bars + 12, *ds_free = bars + 14, *kv_acc_ready = bars + 16, *kv_acc_free = bars + 17,
             *q_acc_ready = bars + 18, *q_acc_free = bars + 19, *sd_free = bars + 20;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 22);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int E = H * HD;
    const int64_t ld3 = 3 * (int64_t)E;
    const int n_local = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int G = 8 * n_local;  // chunks: 2 key tiles x 4 query chunks per item

    if (warp == B_WARP_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
        tma_prefetch_desc(&tmKV0);
        tma_prefetch_desc(&tmKV1);
    }
    if (warp == B_WARP_MMA1) {
        if (elect_one()) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&qdo_full[i], 2);  // TMA bytes + the producer warp's lse / delta staging
                mbar_init(&qdo_empty[i], 1);
                mbar_init(&kv_full[i], 1);   // [tile]
                mbar_init(&kv_empty[i], 1);
                mbar_init(&s_ready[i], 1);
                mbar_init(&p_ready[i], GROUP_WARPS);
                mbar_init(&c_free[i], 1);
                mbar_init(&ds_free[i], 1);   // [block]
            }
            mbar_init(kv_acc_ready, 1);
            mbar_init(kv_acc_free, DRAIN_WARPS);
            mbar_init(q_acc_ready, 1);
            mbar_init(q_acc_free, DRAIN_WARPS);
            mbar_init(sd_free, GROUP_WARPS);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t smem_lo = smem_u32(smem) >> 4;
    constexpr uint32_t A_K = 0, A_V = OPER_BYTES >> 4;  // 16-byte units
    constexpr uint32_t TILE16 = TILE_BYTES >> 4;

    if (warp >= B_WARP_DRAIN0) {
        // =========================== drain: dV_j, dK_j after every key tile, dQ after the item ===========================
        setmaxnreg_inc<96>();
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        uint32_t a[32], a2[32];
        auto load64 = [&](uint32_t col) {
            tmem_ld_32x32b_x32(lane_addr + col, a);
            tmem_ld_32x32b_x32(lane_addr + col + 32, a2);
            tmem_ld_wait();
            reg_fence(a);
            reg_fence(a2);
        };
        auto release = [&](uint64_t* bar) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        // bias_q_only: the key third of the qkv bias gradient is identically zero (sum over keys of dS = 0: the softmax does
        // not see a per-query shift of the scores) and the value third equals the column sums of dO (softmax rows sum to
        // one), which the GEMM that produced dO delivers from its epilogue: only dQ's columns are summed here, and the
        // per-key-tile drain of dV / dK, which the next tile's first MMA2 waits for, is stores only.
        auto put = [&](bool ok, bf16* dst, float* dcol, bool sum) {
            if (ok) {
                store_32cols_bf16(dst, a, 1.f);
                store_32cols_bf16(dst + 32, a2, 1.f);
            }
            if (dbias != nullptr && sum) {
                warp_colsum32_atomic(a, ok, dcol, lane);
                warp_colsum32_atomic(a2, ok, dcol + 32, lane);
            }
        };
        for (int n = 0; n < n_local; ++n) {
            const int it = blockIdx.x + n * gridDim.x;
            const int b = it / H, hd = it - b * H;
            bf16* dbase = dqkv + (int64_t)b * L * ld3 + hd * HD;
            for (int j = 0; j < 2; ++j) {
                const int t = 2 * n + j;
                mbar_wait(kv_acc_ready, t & 1, 74);
                tc_fence_after();
                const int r = j * 128 + row;  // key
                load64(KD_COL_DV);
                put(r < L, dbase + (int64_t)r * ld3 + 2 * E, dbias + 2 * E + hd * HD, !bias_q_only);
                load64(KD_COL_DK);
                release(kv_acc_free);
                put(r < L, dbase + (int64_t)r * ld3 + E, dbias + E + hd * HD, !bias_q_only);
            }
            mbar_wait(q_acc_ready, n & 1, 77);
            tc_fence_after();
            load64(KD_COL_DQ);
            put(row < L, dbase + (int64_t)row * ld3, dbias + hd * HD, true);  // queries 0..127
            load64(KD_COL_DQ + 64);
            release(q_acc_free);
            put(128 + row < L, dbase + (int64_t)(128 + row) * ld3, dbias + hd * HD, true);  // queries 128..255
        }
    } else if (warp >= B_WARP_TMA) {
    setmaxnreg_dec<32>();
    if (warp == B_WARP_TMA) {
        // =========================== TMA producer (+ lse / delta staging) ===========================
        for (int n = 0; n < n_local; ++n) {
            const int it = blockIdx.x + n * gridDim.x;
            const int b = it / H, hd = it - b * H;
            const int s = n & 1;
            // K, V of key tile 0 (free once tile 0 of the previous item is done: that is half an item ago)
            mbar_wait(&kv_empty[0], (n & 1) ^ 1, 70);
            if (elect_one()) {
                mbar_arrive_expect_tx(&kv_full[0], 2 * TILE_BYTES);
                tma_load_3d(smem, &tmKV0, &kv_full[0], E + hd * HD, 0, b);
                tma_load_3d(smem + OPER_BYTES, &tmKV0, &kv_full[0], 2 * E + hd * HD, 0, b);
            }
            __syncwarp();
            // Q, dO of the item (double-buffered)
            mbar_wait(&qdo_empty[s], ((n >> 1) & 1) ^ 1, 71);
            if (elect_one()) {
                uint8_t* st = smem + KD_OFF_QDO + s * KD_QDO_BYTES;
                mbar_arrive_expect_tx(&qdo_full[s], KD_QDO_BYTES);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r0 = h2 * BOX_ROWS;
                    tma_load_3d(st + r0 * 128, &tmQKV, &qdo_full[s], hd * HD, r0, b);
                    tma_load_3d(st + OPER_BYTES + r0 * 128, &tmDO, &qdo_full[s], hd * HD, r0, b);
                }
            }
            __syncwarp();
            float lv[8], dv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int q = lane + 32 * i;
                lv[i] = q < L ? __ldg(lse + (int64_t)it * L + q) : INFINITY;
                dv[i] = q < L ? __ldg(delta + (int64_t)it * L + q) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sL[s * 256 + lane + 32 * i] = lv[i] * LOG2E;
                sD[s * 256 + lane + 32 * i] = dv[i] * 0.125f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&qdo_full[s]);
            // K, V of key tile 1 (rows 128..207; free once tile 1 of the previous item is done)
            mbar_wait(&kv_empty[1], (n & 1) ^ 1, 72);
            if (elect_one()) {
                mbar_arrive_expect_tx(&kv_full[1], 2 * (ROWS - 128) * 128);
                tma_load_3d(smem + TILE_BYTES, &tmKV1, &kv_full[1], E + hd * HD, 128, b);
                tma_load_3d(smem + OPER_BYTES + TILE_BYTES, &tmKV1, &kv_full[1], 2 * E + hd * HD, 128, b);
            }
            __syncwarp();
        }
    } else if (warp == B_WARP_MMA1) {
        // =========================== MMA1 issuer: S^T, dP^T of chunk g into buffer g & 1 ===========================
        const uint32_t idesc64 = make_idesc_bf16(128, 64, 0, 0), idesc16 = make_idesc_bf16(128, 16, 0, 0);
        for (int g = 0; g < G; ++g) {
            const int n = g >> 3, j = (g >> 2) & 1, c = g & 3, s = n & 1, cb = g & 1;
            if ((g & 7) == 0) mbar_wait(&qdo_full[s], (n >> 1) & 1, 73);
            if (c == 0) mbar_wait(&kv_full[j], n & 1, 78);
            if (lane == 0) KD_STAMP(g, 3);
            if (EARLY)
                mbar_wait(sd_free, (g & 1) ^ 1, 79);  // the math group of chunk g - 1 holds its scores in registers
            else
                mbar_wait(&c_free[cb], ((g >> 1) & 1) ^ 1, 79);
            if (lane == 0) KD_STAMP(g, 4);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t qdo = smem_lo + ((KD_OFF_QDO + s * KD_QDO_BYTES) >> 4);
                const uint32_t crow = c * 64 * 8;  // chunk queries = operand rows [64 c, ...)
                const uint32_t a1 = smem_lo + A_K + j * TILE16, a2 = smem_lo + A_V + j * TILE16;
                const uint32_t b1 = qdo + crow, b2 = qdo + (OPER_BYTES >> 4) + crow;
                const uint32_t idesc = c == 3 ? idesc16 : idesc64;
                const uint32_t d = tmem_base + (EARLY ? 0 : cb * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d, make_desc((a1 | LBO_K) + 2 * k, DESC_HI), make_desc((b1 | LBO_K) + 2 * k, DESC_HI), idesc, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d + 64, make_desc((a2 | LBO_K) + 2 * k, DESC_HI), make_desc((b2 | LBO_K) + 2 * k, DESC_HI), idesc, k > 0);
                umma_commit(&s_ready[cb]);
            }
            __syncwarp();
        }
    } else if (warp == B_WARP_MMA2) {
        // =========================== MMA2 / MMA3 issuer ===========================
        const uint32_t idesc2 = make_idesc_bf16(128, HD, 0, 1);  // A from TMEM (K-major), B MN-major
        const uint32_t idesc3 = make_idesc_bf16(128, HD, 1, 1);  // A = dS^T block in smem, MN-major; B = K_j MN-major
        constexpr uint32_t LBO_DS = (16384u >> 4) << 16;          // the two 64-query chunks of a block are 16 KB apart
        for (int g = 0; g < G; ++g) {
            const int n = g >> 3, j = (g >> 2) & 1, c = g & 3, s = n & 1, cb = g & 1;
            const int t = 2 * n + j;
            if (lane == 0) KD_STAMP(g, 0);
            mbar_wait(&p_ready[cb], (g >> 1) & 1, 72);
            if (lane == 0) KD_STAMP(g, 1);
            if (c == 0) mbar_wait(kv_acc_free, (t & 1) ^ 1, 73);           // dV / dK of the previous key tile read out
            if (j == 0 && c == 1) mbar_wait(q_acc_free, (n & 1) ^ 1, 80);  // dQ of the previous item read out
            if (lane == 0) KD_STAMP(g, 2);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t qdo = smem_lo + ((KD_OFF_QDO + s * KD_QDO_BYTES) >> 4);
                const uint32_t crow = c * 64 * 8;
                const int ksteps = c == 3 ? 1 : 4;  // 16 queries per k-step
                const uint32_t pbuf = tmem_base + (EARLY ? 128 + cb * 64 : cb * 128);
                const uint32_t domn = (qdo + (OPER_BYTES >> 4) + crow) | LBO_MN, qmn = (qdo + crow) | LBO_MN;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < ksteps) {
                        const uint32_t acol = EARLY ? k * 8 : (k >> 1) * 32 + (k & 1) * 8;
                        const uint32_t dscol = (EARLY ? 32 : 64) + acol;
                        umma_bf16_ts(tmem_base + KD_COL_DV, pbuf + acol, make_desc(domn + k * 128, DESC_HI), idesc2, (c > 0 || k > 0));    // dV_j += P^T dO_c
                        umma_bf16_ts(tmem_base + KD_COL_DK, pbuf + dscol, make_desc(qmn + k * 128, DESC_HI), idesc2, (c > 0 || k > 0));  // dK_j += dS^T Q_c
                    }
                umma_commit(&c_free[cb]);
                if (c & 1) {
                    // dQ_block += dS_block K_j over the tile's keys (tile 1 holds 80 staged keys: rows 208.. are not K)
                    const int blk = c >> 1;
                    const int kk = j == 0 ? 8 : (ROWS - 128) / 16;
                    const uint32_t ds = (smem_lo + ((KD_OFF_DS + blk * KD_DS_BLOCK) >> 4)) | LBO_DS;
                    const uint32_t kmn = (smem_lo + A_K + j * TILE16) | LBO_MN;
                    for (int k = 0; k < kk; ++k)
                        umma_bf16_ss(tmem_base + KD_COL_DQ + blk * 64, make_desc(ds + k * 128, DESC_HI), make_desc(kmn + k * 128, DESC_HI), idesc3,
                                     (j > 0 || k > 0));
                    umma_commit(&ds_free[blk]);
                }
                if (c == 3) {
                    umma_commit(kv_acc_ready);
                    umma_commit(&kv_empty[j]);  // K_j, V_j may be overwritten by the next item's tile
                    if (j == 1) {
                        umma_commit(q_acc_ready);
                        umma_commit(&qdo_empty[s]);
                    }
                }
            }
            __syncwarp();
        }
    }
    } else {
        // =========================== math warps: group 0 takes even chunks, group 1 odd chunks ===========================
        setmaxnreg_inc<88>();
        const int grp = warp >> 3;
        const int quarter = warp & 3, hf = (warp >> 2) & 1;
        const int row = quarter * 32 + lane;  // key within the tile
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        const int c0 = hf * 32;
        const uint32_t ds_row = smem_u32(smem + KD_OFF_DS) + row * 128;  // this key's 128-byte row inside a 64-query chunk
        const uint32_t sw = row & 7;

        for (int g = grp; g < G; g += 2) {
            const int n = g >> 3, j = (g >> 2) & 1, cc = g & 3, s = n & 1, cb = g & 1;
            const float* sLs = sL + s * 256;
            const float* sDs = sD + s * 256;
            if ((g & 7) == grp) mbar_wait(&qdo_full[s], (n >> 1) & 1, 76);  // lse / delta of the item staged
            // the dS^T block this chunk writes into must have been consumed by MMA3 of the previous key tile
            const bool stp = (warp & 7) == 0 && lane == 0;
            if (stp) KD_STAMP(g, 5);
            mbar_wait(&ds_free[cc >> 1], ((2 * n + j) & 1) ^ 1, 81);
            if (stp) KD_STAMP(g, 6);
            // one barrier per math group in both modes (EARLY shares the fp32 buffer, not the barrier: a parity-1 wait on a
            // barrier shared by both groups would pass at once for group 1's first chunk)
            mbar_wait(&s_ready[cb], (g >> 1) & 1, 75);
            if (stp) KD_STAMP(g, 7);
            tc_fence_after();
            const uint32_t sbuf = lane_addr + (EARLY ? 0 : cb * 128), dbuf = sbuf + 64;
            // where the packed A operands of MMA2 go: over the fp32 values, or (EARLY) into this chunk's packed buffer
            const uint32_t pk_p = EARLY ? lane_addr + 128 + cb * 64 + hf * 16 : sbuf + c0;
            const uint32_t pk_ds = EARLY ? pk_p + 32 : dbuf + c0;
            auto release_scores = [&]() {
                if (EARLY) {
                    // the fp32 scores of this chunk are in registers (or not needed by this warp): MMA1 may refill the buffer;
                    // the packed buffer is free once the MMA2 that read it two chunks ago has completed
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sd_free);
                    mbar_wait(&c_free[cb], ((g >> 1) & 1) ^ 1, 82);
                    tc_fence_after();
                }
            };
            const uint32_t ds_chunk = ds_row + (cc >> 1) * KD_DS_BLOCK + (cc & 1) * 16384;
            if (cc < 3) {
                uint32_t sv[32], dv[32];
                tmem_ld_32x32b_x32(sbuf + c0, sv);
                tmem_ld_32x32b_x32(dbuf + c0, dv);
                tmem_ld_wait();
                reg_fence(sv);
                reg_fence(dv);
                release_scores();
                if (stp) KD_STAMP(g, 10);
                uint32_t pp[16], pd[16];
                const float* lq = sLs + cc * 64 + c0;  // same address for the whole warp: smem broadcast
                const float* dq = sDs + cc * 64 + c0;
#pragma unroll
                for (int gq = 0; gq < 8; ++gq) {
                    const float4 l4 = lds128(lq + 4 * gq), d4 = lds128(dq + 4 * gq);
                    const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
                    float p[4], ds[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        p[i] = fast_ex2(fmaf(__uint_as_float(sv[4 * gq + i]), c, -lv[i]));
                        ds[i] = p[i] * fmaf(__uint_as_float(dv[4 * gq + i]), 0.125f, -dl[i]);
                    }
                    pp[2 * gq] = pack_bf16x2(p[0], p[1]);
                    pp[2 * gq + 1] = pack_bf16x2(p[2], p[3]);
                    pd[2 * gq] = pack_bf16x2(ds[0], ds[1]);
                    pd[2 * gq + 1] = pack_bf16x2(ds[2], ds[3]);
                }
                if (stp) KD_STAMP(g, 11);
                tmem_st_x16(pk_p, pp);
                tmem_st_x16(pk_ds, pd);
                // dS^T also as the A operand of the dQ MMA: 32 queries = 64 bytes = pieces 4 hf .. 4 hf + 3 of this key's row
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_chunk + (((4 * hf + i) ^ sw) << 4)), "r"(pd[4 * i]),
                                 "r"(pd[4 * i + 1]), "r"(pd[4 * i + 2]), "r"(pd[4 * i + 3])
                                 : "memory");
            } else if (hf == 0) {
                // 16-query tail chunk (queries 192..207)
                uint32_t sv[16], dv[16];
                tmem_ld_32x32b_x16(sbuf, sv);
                tmem_ld_32x32b_x16(dbuf, dv);
                tmem_ld_wait();
                reg_fence(sv);
                reg_fence(dv);
                release_scores();
                uint32_t pp[8], pd[8];
                const float* lq = sLs + 192;
                const float* dq = sDs + 192;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const float4 l4 = lds128(lq + 4 * gq), d4 = lds128(dq + 4 * gq);
                    const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
                    float p[4], ds[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        p[i] = fast_ex2(fmaf(__uint_as_float(sv[4 * gq + i]), c, -lv[i]));
                        ds[i] = p[i] * fmaf(__uint_as_float(dv[4 * gq + i]), 0.125f, -dl[i]);
                    }
                    pp[2 * gq] = pack_bf16x2(p[0], p[1]);
                    pp[2 * gq + 1] = pack_bf16x2(p[2], p[3]);
                    pd[2 * gq] = pack_bf16x2(ds[0], ds[1]);
                    pd[2 * gq + 1] = pack_bf16x2(ds[2], ds[3]);
                }
                tmem_st_x8(pk_p, pp);
                tmem_st_x8(pk_ds, pd);
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_chunk + ((i ^ sw) << 4)), "r"(pd[4 * i]),
                                 "r"(pd[4 * i + 1]), "r"(pd[4 * i + 2]), "r"(pd[4 * i + 3])
                                 : "memory");
            } else {
                release_scores();  // the upper-half warps have nothing to do in the 16-query tail chunk
            }
            if (stp) KD_STAMP(g, 8);
            tmem_st_wait();
            fence_proxy_async_smem();  // the dS^T rows must be visible to the tensor core's shared-memory reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[cb]);
            if (stp) KD_STAMP(g, 9);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == B_WARP_MMA1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// VITB200_ATTN_STG128=1 (measurement only): copied into the kernels' constant once per device
static int ensure_store_mode() {
    static const int want = []() {
        const char* e = getenv("VITB200_ATTN_STG128");
        return (e != nullptr && e[0] == '1') ? 1 : 0;
    }();
    static bool done[64] = {false};
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    if (want && dev < 64 && !done[dev]) {
        VB_CHECK_CUDA(cudaMemcpyToSymbol(attn3::c_attn_store128, &want, sizeof(int)));
        done[dev] = true;
    }
    return VB_OK;
}

template <typename K>
static int set_smem(K kern, int bytes, bool& done) {
    if (!done) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done = true;
    }
    return VB_OK;
}

static int make_maps(CUtensorMap* tmQKV, const bf16* qkv, CUtensorMap* tmDO, const bf16* dout, int batch, int L, int H) {
    const int64_t E = (int64_t)H * HD;
    int rc = make_tensor_map_3d(tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * E, L, batch, 3 * E * 2, (uint64_t)L * 3 * E * 2, 64,
                                BOX_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc || tmDO == nullptr) return rc;
    return make_tensor_map_3d(tmDO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dout, E, L, batch, E * 2, (uint64_t)L * E * 2, 64, BOX_ROWS, 1,
                              CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace attn3

int launch_attention_fwd_tc3(const bf16* qkv, bf16* out, float* lse, int batch, int L, int H, cudaStream_t stream) {
    using namespace attn3;
    CUtensorMap tmQKV;
    int rc = make_maps(&tmQKV, qkv, nullptr, nullptr, batch, L, H);
    if (rc) return rc;
    static bool done = false;
    rc = ensure_store_mode();
    if (rc) return rc;
    rc = set_smem(attention_fwd_persistent_kernel<false>, F_SMEM, done);
    if (rc) return rc;
    const int n_items = batch * H;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_fwd_persistent_kernel<false><<<grid, THREADS, F_SMEM, stream>>>(tmQKV, tmQKV, out, lse, L, H, batch, n_items);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// delta[layer][batch * L][E] = attention(qkv_a) - attention(qkv_b); qkv_*: [batch * L, ld] with ld >= layers * 3E
int launch_attention_pair_tc3(const bf16* qkv_a, const bf16* qkv_b, int64_t ld, bf16* delta, int layers, int batch, int L, int H,
                              cudaStream_t stream) {
    using namespace attn3;
    CUtensorMap tmA, tmB;
    const int64_t E = (int64_t)H * HD;
    int rc = make_tensor_map_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv_a, (uint64_t)layers * 3 * E, L, batch, ld * 2, (uint64_t)L * ld * 2, 64,
                                BOX_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map_3d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv_b, (uint64_t)layers * 3 * E, L, batch, ld * 2, (uint64_t)L * ld * 2, 64,
                            BOX_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool done = false;
    rc = ensure_store_mode();
    if (rc) return rc;
    rc = set_smem(attention_fwd_persistent_kernel<true>, F_SMEM, done);
    if (rc) return rc;
    const int n_items = layers * batch * H;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_fwd_persistent_kernel<true><<<grid, THREADS, F_SMEM, stream>>>(tmA, tmB, delta, nullptr, L, H, batch, n_items);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// delta[layer][batch * L][E] = attention(a + d) - attention(a) in perturbation form; qkv_a: projections of a (with bias),
// dqkv: projections of the token difference (no bias); both [batch * L, ld] with ld >= layers * 3E
int launch_attention_delta_tc3(const bf16* qkv_a, const bf16* dqkv, int64_t ld, bf16* delta, int layers, int batch, int L, int H,
                               cudaStream_t stream) {
    using namespace attn3;
    CUtensorMap tmA, tmD;
    const int64_t E = (int64_t)H * HD;
    int rc = make_tensor_map_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv_a, (uint64_t)layers * 3 * E, L, batch, ld * 2, (uint64_t)L * ld * 2, 64,
                                BOX_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map_3d(&tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dqkv, (uint64_t)layers * 3 * E, L, batch, ld * 2, (uint64_t)L * ld * 2, 64,
                            BOX_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool done = false;
    rc = ensure_store_mode();
    if (rc) return rc;
    rc = set_smem(attention_perturb_kernel, D_SMEM, done);
    if (rc) return rc;
    const int n_items = layers * batch * H;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_perturb_kernel<<<grid, THREADS, D_SMEM, stream>>>(tmA, tmD, delta, L, H, batch, n_items);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// delta: caller workspace, f32 [batch, heads, L]
int launch_attention_bwd_tc3(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                             float* dbias, int bias_q_only, int batch, int L, int H, cudaStream_t stream) {
    using namespace attn3;
    CUtensorMap tmQKV, tmDO;
    int rc = make_maps(&tmQKV, qkv, &tmDO, dout, batch, L, H);
    if (rc) return rc;
    const int rows = batch * L;
    if (out != nullptr) {  // (nullptr: the caller already left delta in the workspace, e.g. from the proj-dgrad epilogue)
        attention_delta_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(out, dout, delta, rows, L, H);
        VB_CHECK_LAUNCH();
    }
    const int n_items = batch * H;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    static const bool dbg_on = getenv("VITB200_DBG_TIMING") != nullptr;  // development only
    // key-domain kernel: K / V come in per key tile (128 and 80 rows), so they get their own boxes
    CUtensorMap tmKV0, tmKV1;
    const int64_t E = (int64_t)H * HD;
    rc = make_tensor_map_3d(&tmKV0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * E, L, batch, 3 * E * 2, (uint64_t)L * 3 * E * 2, 64, 128, 1,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_tensor_map_3d(&tmKV1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * E, L, batch, 3 * E * 2, (uint64_t)L * 3 * E * 2, 64, ROWS - 128, 1,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool done_kd = false, done_kd0 = false;
    static const bool early = []() {
        const char* e = getenv("VITB200_ATTN_BWD_EARLY");  // 0: packed operands in place, MMA1 two chunks behind MMA2 (A/B runs)
        return !(e != nullptr && e[0] == '0');
    }();
    rc = ensure_store_mode();
    if (rc) return rc;
    rc = early ? set_smem(attention_bwd_kd_kernel<true>, KD_SMEM, done_kd) : set_smem(attention_bwd_kd_kernel<false>, KD_SMEM, done_kd0);
    if (rc) return rc;
    if (dbg_on) {
        long long* dbg = nullptr;
        VB_CHECK_CUDA(cudaMallocManaged(&dbg, 64 * 16 * sizeof(long long)));
        VB_CHECK_CUDA(cudaMemset(dbg, 0, 64 * 16 * sizeof(long long)));
        attention_bwd_kd_kernel<true><<<grid, KD_THREADS, KD_SMEM, stream>>>(tmQKV, tmDO, tmKV0, tmKV1, lse, delta, dqkv, dbias, bias_q_only, L, H, n_items, dbg);
        VB_CHECK_CUDA(cudaStreamSynchronize(stream));
        const long long t0 = dbg[0];
        printf("[kd timing] g: mma2{waitP< waitP> accfree>} mma1{cfree< cfree>} math{start dsfree> Sready> stores> arrive> ld> math>}\n");
        for (int g = 0; g < 64; ++g) {
            printf("[kd timing] %3d:", g + 16);
            for (int i = 0; i < 12; ++i) printf(" %7lld", dbg[g * 16 + i] ? dbg[g * 16 + i] - t0 : -1);
            printf("\n");
        }
        cudaFree(dbg);
        return VB_OK;
    }
    if (early)
        attention_bwd_kd_kernel<true><<<grid, KD_THREADS, KD_SMEM, stream>>>(tmQKV, tmDO, tmKV0, tmKV1, lse, delta, dqkv, dbias, bias_q_only, L, H, n_items, nullptr);
    else
        attention_bwd_kd_kernel<false><<<grid, KD_THREADS, KD_SMEM, stream>>>(tmQKV, tmDO, tmKV0, tmKV1, lse, delta, dqkv, dbias, bias_q_only, L, H, n_items, nullptr);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

}  // namespace vb
