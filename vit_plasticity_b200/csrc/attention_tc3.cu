// Persistent, software-pipelined attention forward / backward on tcgen05 for L <= 208 tokens, head_dim = 64
// (ViT-B/L at 224x224: L = 197). One CTA per SM walks over (image, head) items:
//
//   warp 8      TMA producer: the item's Q/K/V(/dO) rows as 208-row SWIZZLE_128B operand tiles (two 104-row boxes each;
//               rows >= L are zero-filled by the TMA unit), double-buffered, so item i+1 loads while item i computes
//   warp 9      MMA issuer (one elected lane): runs one unit ahead of the math warps, so the tensor pipe works on the
//               next scores while the math warps exponentiate the current ones
//   warps 0..7  math: TMEM -> registers -> TMEM. Probabilities never touch shared memory: they are packed to bf16 in
//               place of the fp32 scores and consumed as the A operand (from TMEM) of the next MMA.
//
// forward, per item two units (query tiles of 128 rows): S = Q_t K^T (N = 208) -> exact two-pass softmax -> O = P V.
//   TMEM: S/P buffers [0,208) and [208,416) (unit parity), O accumulator [416,480).
// backward, per item four units: dQ_t (t = 0,1; queries on the TMEM lanes) and dK_j/dV_j (j = 0,1; keys on the lanes,
//   "transposed domain", so lse / delta are per-COLUMN scalars there). Each unit walks 4 column chunks (64,64,64,16):
//   MMA1: S_c, dP_c -> math: P = exp2(S c - lse), dS = P (dP - delta) / 8 -> MMA2: accumulate.
//   TMEM: chunk buffers [0,128) / [128,256) (S at +0, dP at +64), accumulators [256,384) / [384,512) (unit parity).
//   No masking is needed: padded K/V/Q/dO rows are zero and padded lse entries are +inf (p = 0).
#include "host_utils.h"
#include "ptx.cuh"

namespace vb {
namespace attn3 {

constexpr int HD = 64;
constexpr int ROWS = 208;                    // operand rows staged per item (13 x 16)
constexpr int BOX_ROWS = 104;                // 104 x 128 B = 13 KB per TMA box, a multiple of the 1024-B swizzle atom
constexpr int OPER_BYTES = ROWS * 128;       // 26624
constexpr int TILE_BYTES = 128 * 128;        // second 128-row tile of an operand starts here
constexpr int MATH_WARPS = 8;
constexpr int WARP_TMA = 8, WARP_MMA = 9;
constexpr int THREADS = 10 * 32;
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
constexpr int F_STAGE = 3 * OPER_BYTES;  // Q, K, V
constexpr int F_EXCH_BYTES = 2 * 2 * 128 * 4 * 2;  // sMax / sSum: [unit parity][column half][row]
constexpr int F_SMEM = 2 * F_STAGE + F_EXCH_BYTES + 256 + 1024;
constexpr uint32_t F_COL_O = 416;

// 16 scores -> running max (nvalid = number of unmasked columns among these 16)
__device__ __forceinline__ float max16(const uint32_t (&v)[16], float m, int nvalid) {
    if (nvalid >= 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) m = fmaxf(m, __uint_as_float(v[i]));
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (i < nvalid) m = fmaxf(m, __uint_as_float(v[i]));
    }
    return m;
}
// 16 scores -> 16 probabilities packed as 8 bf16 pairs; returns their fp32 sum
__device__ __forceinline__ float exp16(const uint32_t (&v)[16], uint32_t (&pk)[8], float c, float mc, int nvalid) {
    float p[16];
    if (nvalid >= 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = fast_ex2(fmaf(__uint_as_float(v[i]), c, -mc));
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = (i < nvalid) ? fast_ex2(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        s0 += p[4 * i];
        s1 += p[4 * i + 1];
        s2 += p[4 * i + 2];
        s3 += p[4 * i + 3];
    }
    return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(THREADS, 1)
attention_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQKV, bf16* __restrict__ out, float* __restrict__ lse_out, int L,
                                int H, int n_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    float* sMax = reinterpret_cast<float*>(smem + 2 * F_STAGE);  // [2][2][128]
    float* sSum = sMax + 512;                                    // [2][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSum + 512);
    uint64_t *full = bars, *empty = bars + 2, *s_ready = bars + 4, *p_ready = bars + 6, *o_ready = bars + 8, *o_free = bars + 9;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 10);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
    const int E = H * HD;
    const int n_local = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int U = 2 * n_local;  // units (query tiles) this CTA processes

    if (warp == WARP_TMA && elect_one()) tma_prefetch_desc(&tmQKV);
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
            const int it = blockIdx.x + n * gridDim.x;
            const int b = it / H, hd = it - b * H;
            const int s = n & 1;
            mbar_wait(&empty[s], ((n >> 1) & 1) ^ 1, 60);
            if (elect_one()) {
                uint8_t* st = smem + s * F_STAGE;
                mbar_arrive_expect_tx(&full[s], F_STAGE);
#pragma unroll
                for (int op = 0; op < 3; ++op)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2)
                        tma_load_3d(st + op * OPER_BYTES + h2 * BOX_ROWS * 128, &tmQKV, &full[s], op * E + hd * HD, h2 * BOX_ROWS, b);
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
                    for (int k = 0; k < 13; ++k) {
                        const uint32_t acol = t * ROWS + (k < 7 ? k * 8 : 112 + (k - 7) * 8);  // keys [0,112) then [112,208)
                        umma_bf16_ts(tmem_base + F_COL_O, tmem_base + acol, make_desc(vlo + k * 128, DESC_HI), idesc_o, k > 0);
                    }
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
        const int quarter = warp & 3, hf = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        const int cbeg = hf == 0 ? 0 : 112;
        const int nsub = hf == 0 ? 7 : 6;  // 16-column sub-chunks owned by this warp
        float m_prev = 0.f, sum_prev = 0.f;
        for (int u = 0; u <= U; ++u) {
            float m_cur = 0.f, sum_cur = 0.f;
            if (u < U) {
                const int t = u & 1;
                const uint32_t s_addr = lane_addr + t * ROWS + cbeg;
                mbar_wait(&s_ready[t], (u >> 1) & 1, 64);
                tc_fence_after();
                uint32_t va[16], vb2[16];
                // ---- pass 1: row max over this warp's columns (next sub-chunk's TMEM load in flight) ----
                float m = -INFINITY;
                tmem_ld_32x32b_x16(s_addr, va);
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    if (s < nsub) {
                        tmem_ld_wait();
                        const int nvalid = L - (cbeg + s * 16);
                        if ((s & 1) == 0) {
                            reg_fence(va);
                            if (s + 1 < nsub) tmem_ld_32x32b_x16(s_addr + (s + 1) * 16, vb2);
                            m = max16(va, m, nvalid);
                        } else {
                            reg_fence(vb2);
                            if (s + 1 < nsub) tmem_ld_32x32b_x16(s_addr + (s + 1) * 16, va);
                            m = max16(vb2, m, nvalid);
                        }
                    }
                }
                sMax[(t * 2 + hf) * 128 + row] = m;
                tmem_ld_32x32b_x16(s_addr, va);  // first sub-chunk of pass 2, in flight across the exchange
                named_bar_sync(1 + quarter, 64);   // the two warps sharing this lane quarter
                m = fmaxf(m, sMax[(t * 2 + (hf ^ 1)) * 128 + row]);  // L >= 1: column 0 is valid, so m is finite
                const float mc = m * c;
                // ---- pass 2: p = exp2(s c - m c), packed to bf16 pairs in place ----
                float sum = 0.f;
#pragma unroll
                for (int s = 0; s < 7; ++s) {
                    if (s < nsub) {
                        tmem_ld_wait();
                        const int nvalid = L - (cbeg + s * 16);
                        uint32_t pk[8];
                        if ((s & 1) == 0) {
                            reg_fence(va);
                            if (s + 1 < nsub) tmem_ld_32x32b_x16(s_addr + (s + 1) * 16, vb2);
                            sum += exp16(va, pk, c, mc, nvalid);
                        } else {
                            reg_fence(vb2);
                            if (s + 1 < nsub) tmem_ld_32x32b_x16(s_addr + (s + 1) * 16, va);
                            sum += exp16(vb2, pk, c, mc, nvalid);
                        }
                        tmem_st_x8(s_addr + s * 8, pk);
                    }
                }
                sSum[(t * 2 + hf) * 128 + row] = sum;
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[t]);
                m_cur = m;
                sum_cur = sum;
            }
            if (u >= 1) {
                // ---- read out O of unit u-1 (its P V ran while this warp did the softmax of unit u) ----
                const int v = u - 1, t = v & 1;
                const int it = blockIdx.x + (v >> 1) * gridDim.x;
                const int b = it / H, hd = it - b * H;
                mbar_wait(o_ready, v & 1, 65);
                tc_fence_after();
                named_bar_sync(1 + quarter, 64);  // partner's partial sum is visible
                const float tot = sum_prev + sSum[(t * 2 + (hf ^ 1)) * 128 + row];
                uint32_t o[32];
                tmem_ld_32x32b_x32(lane_addr + F_COL_O + hf * 32, o);
                tmem_ld_wait();
                reg_fence(o);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(o_free);
                const int q = t * 128 + row;
                if (q < L) {
                    store_32cols_bf16(out + ((int64_t)b * L + q) * E + hd * HD + hf * 32, o, 1.f / tot);
                    if (hf == 0 && lse_out != nullptr) lse_out[((int64_t)b * H + hd) * L + q] = m_prev * 0.125f + __logf(tot);
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

constexpr int B_STAGE = 4 * OPER_BYTES;  // Q, K, V, dO
constexpr int B_TAIL = 6144;             // lse / delta staging + barriers; also absorbs the 48-row over-read of the last tile
constexpr int B_SMEM = 2 * B_STAGE + B_TAIL + 1024;
constexpr uint32_t OFF_Q = 0, OFF_K = OPER_BYTES >> 4, OFF_V = (2 * OPER_BYTES) >> 4, OFF_DO = (3 * OPER_BYTES) >> 4;  // 16-B units
constexpr uint32_t B_COL_ACC = 256;

__global__ void __launch_bounds__(THREADS, 1)
attention_bwd_persistent_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                                const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dqkv, int L, int H,
                                int n_items) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align_smem(smem_raw);
    float* sL = reinterpret_cast<float*>(smem + 2 * B_STAGE);  // [2][256] lse * log2(e); +inf for q >= L
    float* sD = sL + 512;                                      // [2][256] delta / 8;     0 for q >= L
    uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 512);
    uint64_t *full = bars, *empty = bars + 2, *s_ready = bars + 4, *p_ready = bars + 6, *acc_ready = bars + 8, *acc_free = bars + 10;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform
    const int E = H * HD;
    const int64_t ld3 = 3 * (int64_t)E;
    const int n_local = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int G = 16 * n_local;  // column chunks: 4 units x 4 chunks per item

    if (warp == WARP_TMA && elect_one()) {
        tma_prefetch_desc(&tmQKV);
        tma_prefetch_desc(&tmDO);
    }
    if (warp == WARP_MMA) {
        if (elect_one()) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
                mbar_init(&s_ready[i], 1);
                mbar_init(&p_ready[i], MATH_WARPS);
                mbar_init(&acc_ready[i], 1);
                mbar_init(&acc_free[i], MATH_WARPS);
            }
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
            const int b = it / H, hd = it - b * H;
            const int s = n & 1;
            mbar_wait(&empty[s], ((n >> 1) & 1) ^ 1, 70);
            if (elect_one()) {
                uint8_t* st = smem + s * B_STAGE;
                mbar_arrive_expect_tx(&full[s], B_STAGE);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r0 = h2 * BOX_ROWS;
                    tma_load_3d(st + 0 * OPER_BYTES + r0 * 128, &tmQKV, &full[s], hd * HD, r0, b);
                    tma_load_3d(st + 1 * OPER_BYTES + r0 * 128, &tmQKV, &full[s], E + hd * HD, r0, b);
                    tma_load_3d(st + 2 * OPER_BYTES + r0 * 128, &tmQKV, &full[s], 2 * E + hd * HD, r0, b);
                    tma_load_3d(st + 3 * OPER_BYTES + r0 * 128, &tmDO, &full[s], hd * HD, r0, b);
                }
            }
            __syncwarp();
        }
    } else if (warp == WARP_MMA) {
        // =========================== MMA issuer ===========================
        // chunk g = 16 n + 4 unit + c. Order: MMA1(0) MMA1(1) | MMA2(g) MMA1(g+2) ... : the tensor pipe executes in issue
        // order, so MMA1(g+2) may overwrite the chunk buffer as soon as MMA2(g) (which read the packed operands) is queued.
        const uint32_t idesc64 = make_idesc_bf16(128, 64, 0, 0), idesc16 = make_idesc_bf16(128, 16, 0, 0);
        const uint32_t idesc2 = make_idesc_bf16(128, HD, 0, 1);
        const uint32_t smem_lo = smem_u32(smem) >> 4;
        auto issue_mma1 = [&](int g) {
            const int n = g >> 4, un = (g >> 2) & 3, c = g & 3, s = n & 1, cb = g & 1;
            if ((g & 15) == 0) mbar_wait(&full[s], (n >> 1) & 1, 71);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t base = smem_lo + ((s * B_STAGE) >> 4);
                const uint32_t tile = (un & 1) * (TILE_BYTES >> 4);  // lanes = rows [128 (un & 1), +128) of the unit's operand
                const uint32_t crow = c * 64 * 8;                    // chunk columns = rows [64 c, ...) of the other operand
                uint32_t a1, b1, a2, b2;
                if (un < 2) {  // S = Q_t K_c^T, dP = dO_t V_c^T
                    a1 = base + OFF_Q + tile, b1 = base + OFF_K + crow, a2 = base + OFF_DO + tile, b2 = base + OFF_V + crow;
                } else {       // S^T = K_j Q_c^T, dP^T = V_j dO_c^T
                    a1 = base + OFF_K + tile, b1 = base + OFF_Q + crow, a2 = base + OFF_V + tile, b2 = base + OFF_DO + crow;
                }
                const uint32_t idesc = c == 3 ? idesc16 : idesc64;
                const uint32_t d = tmem_base + cb * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d, make_desc((a1 | LBO_K) + 2 * k, DESC_HI), make_desc((b1 | LBO_K) + 2 * k, DESC_HI), idesc, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d + 64, make_desc((a2 | LBO_K) + 2 * k, DESC_HI), make_desc((b2 | LBO_K) + 2 * k, DESC_HI), idesc, k > 0);
                umma_commit(&s_ready[cb]);
            }
            __syncwarp();
        };
        auto issue_mma2 = [&](int g) {
            const int n = g >> 4, un = (g >> 2) & 3, c = g & 3, s = n & 1, cb = g & 1;
            const int ug = g >> 2, ab = ug & 1;
            mbar_wait(&p_ready[cb], (g >> 1) & 1, 72);
            if (c == 0) mbar_wait(&acc_free[ab], ((ug >> 1) & 1) ^ 1, 73);  // accumulator of unit ug-2 has been read out
            tc_fence_after();
            if (elect_one()) {
                const uint32_t base = smem_lo + ((s * B_STAGE) >> 4);
                const uint32_t crow = c * 64 * 8;
                const int ksteps = c == 3 ? 1 : 4;  // 16 operand rows per k-step
                const uint32_t acc = tmem_base + B_COL_ACC + ab * 128;
                const uint32_t pbuf = tmem_base + cb * 128;
                if (un < 2) {
                    const uint32_t kmn = (base + OFF_K + crow) | LBO_MN;  // dQ_t += dS_c K_c
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < ksteps)
                            umma_bf16_ts(acc, pbuf + 64 + (k >> 1) * 32 + (k & 1) * 8, make_desc(kmn + k * 128, DESC_HI), idesc2, (c > 0 || k > 0));
                } else {
                    const uint32_t domn = (base + OFF_DO + crow) | LBO_MN, qmn = (base + OFF_Q + crow) | LBO_MN;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < ksteps) {
                            const uint32_t acol = (k >> 1) * 32 + (k & 1) * 8;
                            umma_bf16_ts(acc, pbuf + acol, make_desc(domn + k * 128, DESC_HI), idesc2, (c > 0 || k > 0));           // dV_j += P^T dO_c
                            umma_bf16_ts(acc + 64, pbuf + 64 + acol, make_desc(qmn + k * 128, DESC_HI), idesc2, (c > 0 || k > 0));  // dK_j += dS^T Q_c
                        }
                }
                if (c == 3) {
                    umma_commit(&acc_ready[ab]);
                    if (un == 3) umma_commit(&empty[s]);  // last MMA reading this item's smem stage
                }
            }
            __syncwarp();
        };
        if (G > 0) {
            issue_mma1(0);
            issue_mma1(1);
            for (int g = 0; g < G; ++g) {
                issue_mma2(g);
                if (g + 2 < G) issue_mma1(g + 2);
            }
        }
    } else {
        // =========================== math warps ===========================
        const int quarter = warp & 3, hf = warp >> 2;
        const int row = quarter * 32 + lane;
        const int tid = threadIdx.x;  // 0..255: token index for the lse / delta staging
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = 0.125f * LOG2E;
        const int c0 = hf * 32;

        auto fetch_stats = [&](int n, float& l, float& d) {  // global loads for item n, staged into smem one item later
            l = INFINITY;
            d = 0.f;
            if (n < n_local && tid < L) {
                const int it = blockIdx.x + n * gridDim.x;
                l = __ldg(lse + (int64_t)it * L + tid) * LOG2E;
                d = __ldg(delta + (int64_t)it * L + tid) * 0.125f;
            }
        };
        auto readout = [&](int ug) {
            const int ab = ug & 1, un = ug & 3;
            const int it = blockIdx.x + (ug >> 2) * gridDim.x;
            const int b = it / H, hd = it - b * H;
            mbar_wait(&acc_ready[ab], (ug >> 1) & 1, 74);
            tc_fence_after();
            const int r = (un & 1) * 128 + row;  // query (dQ units) or key (dK/dV units)
            uint32_t a[32], a2[32];
            if (un < 2) {
                tmem_ld_32x32b_x32(lane_addr + B_COL_ACC + ab * 128 + hf * 32, a);
                tmem_ld_wait();
                reg_fence(a);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_free[ab]);
                if (r < L) store_32cols_bf16(dqkv + ((int64_t)b * L + r) * ld3 + hd * HD + hf * 32, a, 1.f);
            } else {
                // warps 0-3 drain dV, warps 4-7 drain dK (64 columns each)
                tmem_ld_32x32b_x32(lane_addr + B_COL_ACC + ab * 128 + hf * 64, a);
                tmem_ld_32x32b_x32(lane_addr + B_COL_ACC + ab * 128 + hf * 64 + 32, a2);
                tmem_ld_wait();
                reg_fence(a);
                reg_fence(a2);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_free[ab]);
                if (r < L) {
                    bf16* dst = dqkv + ((int64_t)b * L + r) * ld3 + (hf == 0 ? 2 * E : E) + hd * HD;
                    store_32cols_bf16(dst, a, 1.f);
                    store_32cols_bf16(dst + 32, a2, 1.f);
                }
            }
        };

        float l_next, d_next;
        fetch_stats(0, l_next, d_next);
        sL[tid] = l_next;
        sD[tid] = d_next;
        fetch_stats(1, l_next, d_next);
        named_bar_sync(5, 256);

        float lse_row = 0.f, dlt_row = 0.f;
        for (int g = 0; g < G; ++g) {
            const int n = g >> 4, un = (g >> 2) & 3, cc = g & 3, s = n & 1, cb = g & 1;
            const float* sLs = sL + s * 256;
            const float* sDs = sD + s * 256;
            if (cc == 0 && un < 2) {
                lse_row = sLs[(un & 1) * 128 + row];
                dlt_row = sDs[(un & 1) * 128 + row];
            }
            mbar_wait(&s_ready[cb], (g >> 1) & 1, 75);
            tc_fence_after();
            const uint32_t sbuf = lane_addr + cb * 128, dbuf = sbuf + 64;
            if (cc < 3) {
                uint32_t sv[32], dv[32];
                tmem_ld_32x32b_x32(sbuf + c0, sv);
                tmem_ld_32x32b_x32(dbuf + c0, dv);
                tmem_ld_wait();
                reg_fence(sv);
                reg_fence(dv);
                if (un < 2) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = fast_ex2(fmaf(__uint_as_float(sv[2 * i]), c, -lse_row));
                        const float p1 = fast_ex2(fmaf(__uint_as_float(sv[2 * i + 1]), c, -lse_row));
                        pk[i] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv[2 * i]), 0.125f, -dlt_row),
                                            p1 * fmaf(__uint_as_float(dv[2 * i + 1]), 0.125f, -dlt_row));
                    }
                    tmem_st_x16(dbuf + c0, pk);  // dS packed in place of dP (own column range)
                } else {
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
                    tmem_st_x16(sbuf + c0, pp);
                    tmem_st_x16(dbuf + c0, pd);
                }
            } else if (hf == 0) {
                // 16-column tail chunk (operand rows 192..207): handled by the first column-half warps only
                uint32_t sv[16], dv[16];
                tmem_ld_32x32b_x16(sbuf, sv);
                tmem_ld_32x32b_x16(dbuf, dv);
                tmem_ld_wait();
                reg_fence(sv);
                reg_fence(dv);
                uint32_t pp[8], pd[8];
                if (un < 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float p0 = fast_ex2(fmaf(__uint_as_float(sv[2 * i]), c, -lse_row));
                        const float p1 = fast_ex2(fmaf(__uint_as_float(sv[2 * i + 1]), c, -lse_row));
                        pd[i] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv[2 * i]), 0.125f, -dlt_row),
                                            p1 * fmaf(__uint_as_float(dv[2 * i + 1]), 0.125f, -dlt_row));
                    }
                    tmem_st_x8(dbuf, pd);
                } else {
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
                    tmem_st_x8(sbuf, pp);
                    tmem_st_x8(dbuf, pd);
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[cb]);

            if (cc == 0 && g >= 4) readout((g >> 2) - 1);  // previous unit's accumulator: its last MMA2 ran during this chunk
            if ((g & 15) == 15) {
                // item n is done with its lse / delta slot; stage item n+1's (fetched one item ago) and fetch item n+2's
                sL[((n + 1) & 1) * 256 + tid] = l_next;
                sD[((n + 1) & 1) * 256 + tid] = d_next;
                fetch_stats(n + 2, l_next, d_next);
                named_bar_sync(5, 256);
            }
        }
        if (G > 0) readout((G >> 2) - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
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
    rc = set_smem(attention_fwd_persistent_kernel, F_SMEM, done);
    if (rc) return rc;
    const int n_items = batch * H;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_fwd_persistent_kernel<<<grid, THREADS, F_SMEM, stream>>>(tmQKV, out, lse, L, H, n_items);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// delta: caller workspace, f32 [batch, heads, L]
int launch_attention_bwd_tc3(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                             int batch, int L, int H, cudaStream_t stream) {
    using namespace attn3;
    CUtensorMap tmQKV, tmDO;
    int rc = make_maps(&tmQKV, qkv, &tmDO, dout, batch, L, H);
    if (rc) return rc;
    static bool done = false;
    rc = set_smem(attention_bwd_persistent_kernel, B_SMEM, done);
    if (rc) return rc;
    const int rows = batch * L;
    attention_delta_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(out, dout, delta, rows, L, H);
    VB_CHECK_LAUNCH();
    const int n_items = batch * H;
    const int grid = n_items < num_sms() ? n_items : num_sms();
    attention_bwd_persistent_kernel<<<grid, THREADS, B_SMEM, stream>>>(tmQKV, tmDO, lse, delta, dqkv, L, H, n_items);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

}  // namespace vb
