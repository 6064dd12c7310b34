// Fused attention core for ViT sequence lengths (L = 197 / 257): the whole key range of one (image, head) fits in
// shared memory, so softmax is exact in one pass (no online rescaling) and neither the score matrix nor the
// probabilities ever reach HBM. Replaces q@k^T / sqrt(d) -> softmax -> @v (reference architecture.py:212-233,
// which materialises two (N,h,L,L) fp32 tensors) and its autograd.
//
// Round-1 implementation: one CTA per (image, head), bf16 mma.sync.m16n8k16 with fp32 accumulation, operands staged
// with cp.async into padded (conflict-free ldmatrix) shared memory.
//   forward : warp owns 16-query row blocks; S (16 x L) lives in registers; P re-used as the A operand of P.V
//   backward: warp owns 16-key blocks; computes S^T, dP^T in [key][query] layout so that P^T / dS^T feed the
//             dV / dK MMAs directly; dS is transposed in registers (movmatrix) for dQ, accumulated in smem (fp32)
//   pair    : forward on two inputs, writes attn(a) - attn(b) subtracted in fp32 (plasticity estimator)
#include <stdlib.h>

#include "host_utils.h"
#include "ptx.cuh"

namespace vb {

constexpr int HD = 64;         // head dim
constexpr int SROW = HD + 8;   // padded smem row (elements): 144 B stride -> conflict-free ldmatrix

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
// D(16x8, f32) += A(16x16, bf16 row) * B(16x8, bf16 col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// Stage `L` rows of 64 bf16 (row stride ld elements in global) into padded smem, zero the rows up to `rows_pad`.
__device__ __forceinline__ void stage_rows(bf16* s, const bf16* g, int64_t ld, int L, int rows_pad, int tid, int nthreads) {
    for (int idx = tid; idx < rows_pad * 8; idx += nthreads) {
        const int r = idx >> 3, c = idx & 7;
        bf16* dst = s + r * SROW + c * 8;
        if (r < L)
            cp_async16(smem_u32(dst), g + (int64_t)r * ld + c * 8);
        else
            *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
}

// One 16-query row block of softmax(Q K^T * scale) V. Returns normalised O in o[8][4] (C-fragment layout) and the
// row log-sum-exp (natural log, of the scaled scores) for rows g and g+8.
template <int NB16>
__device__ __forceinline__ void attn_row_block(const bf16* sQ, const bf16* sK, const bf16* sV, int rb, int L, int lane,
                                               float (&o)[8][4], float& lse0, float& lse1) {
    constexpr int NT = NB16 * 2;
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(aq[ks], smem_u32(sQ + (rb * 16 + (lane & 15)) * SROW + ks * 16 + (lane >> 4) * 8));
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            uint32_t bk[4];
            ldmatrix_x4(bk, smem_u32(sK + (nt * 8 + (lane & 7)) * SROW + kp * 32 + (lane >> 3) * 8));
            mma_bf16(s[nt], aq[2 * kp], bk[0], bk[1]);
            mma_bf16(s[nt], aq[2 * kp + 1], bk[2], bk[3]);
        }
    }
    // mask padded keys, row max
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + (lane & 3) * 2;
        if (c >= L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (c + 1 >= L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    const float mb0 = m0 * scale_log2, mb1 = m1 * scale_log2;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = exp2f(s[nt][0] * scale_log2 - mb0);
        s[nt][1] = exp2f(s[nt][1] * scale_log2 - mb0);
        s[nt][2] = exp2f(s[nt][2] * scale_log2 - mb1);
        s[nt][3] = exp2f(s[nt][3] * scale_log2 - mb1);
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    lse0 = m0 * 0.125f + __logf(sum0);
    lse1 = m1 * 0.125f + __logf(sum1);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NB16; ++ks) {
        uint32_t ap[4];
        ap[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
        ap[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
        ap[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
        ap[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t bv[4];
            ldmatrix_x4_trans(bv, smem_u32(sV + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SROW + np * 16 + (lane >> 4) * 8));
            mma_bf16(o[2 * np], ap, bv[0], bv[1]);
            mma_bf16(o[2 * np + 1], ap, bv[2], bv[3]);
        }
    }
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        // __fmul_rn: keeps the product a rounded value so that the paired variant's o_a - o_b cannot be contracted
        // into an FMA (identical inputs must cancel exactly)
        o[nt][0] = __fmul_rn(o[nt][0], inv0);
        o[nt][1] = __fmul_rn(o[nt][1], inv0);
        o[nt][2] = __fmul_rn(o[nt][2], inv1);
        o[nt][3] = __fmul_rn(o[nt][3], inv1);
    }
}

// Write a 16 x 64 C-fragment tile as bf16 through the warp's own 16 smem rows, then coalesced 16-byte stores.
__device__ __forceinline__ void store_tile_via_smem(bf16* srows /* 16 rows, SROW stride, owned by this warp */,
                                                    const float (&o)[8][4], bf16* gdst, int64_t ld, int row_base, int L,
                                                    int lane) {
    __syncwarp();
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(srows + g * SROW + nt * 8 + t * 2) = pack_bf16x2(o[nt][0], o[nt][1]);
        *reinterpret_cast<uint32_t*>(srows + (g + 8) * SROW + nt * 8 + t * 2) = pack_bf16x2(o[nt][2], o[nt][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + i * 32;
        const int r = idx >> 3, c = idx & 7;
        if (row_base + r < L)
            *reinterpret_cast<uint4*>(gdst + (int64_t)(row_base + r) * ld + c * 8) =
                *reinterpret_cast<const uint4*>(srows + r * SROW + c * 8);
    }
    __syncwarp();
}

constexpr int FWD_WARPS = 4;

template <int NB16, bool PAIR>
__global__ void __launch_bounds__(FWD_WARPS * 32)
attention_fwd_kernel(const bf16* __restrict__ qkv_a, const bf16* __restrict__ qkv_b, int64_t ld_qkv,
                     bf16* __restrict__ out, int64_t ld_out, float* __restrict__ lse, int L, int H) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    constexpr int ROWS = NB16 * 16;
    bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
    bf16* sK = sQ + ROWS * SROW;
    bf16* sV = sK + ROWS * SROW;
    bf16* sQ2 = sV + ROWS * SROW;  // PAIR only
    bf16* sK2 = sQ2 + ROWS * SROW;
    bf16* sV2 = sK2 + ROWS * SROW;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int E = H * HD;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bf16* base_a = qkv_a + (int64_t)b * L * ld_qkv + h * HD;
    stage_rows(sQ, base_a, ld_qkv, L, ROWS, threadIdx.x, FWD_WARPS * 32);
    stage_rows(sK, base_a + E, ld_qkv, L, ROWS, threadIdx.x, FWD_WARPS * 32);
    stage_rows(sV, base_a + 2 * E, ld_qkv, L, ROWS, threadIdx.x, FWD_WARPS * 32);
    if (PAIR) {
        const bf16* base_b = qkv_b + (int64_t)b * L * ld_qkv + h * HD;
        stage_rows(sQ2, base_b, ld_qkv, L, ROWS, threadIdx.x, FWD_WARPS * 32);
        stage_rows(sK2, base_b + E, ld_qkv, L, ROWS, threadIdx.x, FWD_WARPS * 32);
        stage_rows(sV2, base_b + 2 * E, ld_qkv, L, ROWS, threadIdx.x, FWD_WARPS * 32);
    }
    cp_async_wait_all();
    __syncthreads();

    const int nblocks = (L + 15) / 16;
    bf16* gout = out + (int64_t)b * L * ld_out + h * HD;
    for (int rb = warp; rb < nblocks; rb += FWD_WARPS) {
        float o[8][4];
        float l0, l1;
        attn_row_block<NB16>(sQ, sK, sV, rb, L, lane, o, l0, l1);
        if (PAIR) {
            float o2[8][4];
            float k0, k1;
            attn_row_block<NB16>(sQ2, sK2, sV2, rb, L, lane, o2, k0, k1);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[nt][j] -= o2[nt][j];
        } else if (lse != nullptr) {
            const int r0 = rb * 16 + (lane >> 2);
            if ((lane & 3) == 0) {
                float* lp = lse + ((int64_t)b * H + h) * L;
                if (r0 < L) lp[r0] = l0;
                if (r0 + 8 < L) lp[r0 + 8] = l1;
            }
        }
        // this warp's 16 Q rows are no longer needed: reuse them as the output staging buffer
        store_tile_via_smem(sQ + rb * 16 * SROW, o, gout, ld_out, rb * 16, L, lane);
    }
}

constexpr int BWD_WARPS = 8;
// fp32 elements per dQ accumulator row (padded when it fits: 2-way bank conflicts at worst)
template <int NB16> struct DqStride { static constexpr int value = NB16 <= 13 ? HD + 8 : HD; };

template <int NB16>
__global__ void __launch_bounds__(BWD_WARPS * 32, 1)
attention_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                     const float* __restrict__ lse, bf16* __restrict__ dqkv, int L, int H) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    constexpr int ROWS = NB16 * 16;
    constexpr int DQ_STRIDE = DqStride<NB16>::value;
    bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
    bf16* sK = sQ + ROWS * SROW;
    bf16* sV = sK + ROWS * SROW;
    bf16* sdO = sV + ROWS * SROW;
    float* sdQ = reinterpret_cast<float*>(sdO + ROWS * SROW);  // [ROWS][DQ_STRIDE]
    float* sD = sdQ + ROWS * DQ_STRIDE;                        // [ROWS]   rowsum(dO * O)
    float* sL = sD + ROWS;                                     // [ROWS]   lse * log2(e)
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int E = H * HD;
    const int64_t ld3 = 3 * (int64_t)E;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bf16* base = qkv + (int64_t)b * L * ld3 + h * HD;
    const bf16* obase = out + (int64_t)b * L * E + h * HD;
    const bf16* dobase = dout + (int64_t)b * L * E + h * HD;
    stage_rows(sQ, base, ld3, L, ROWS, tid, BWD_WARPS * 32);
    stage_rows(sK, base + E, ld3, L, ROWS, tid, BWD_WARPS * 32);
    stage_rows(sV, base + 2 * E, ld3, L, ROWS, tid, BWD_WARPS * 32);
    stage_rows(sdO, dobase, E, L, ROWS, tid, BWD_WARPS * 32);
    for (int idx = tid; idx < ROWS * DQ_STRIDE; idx += BWD_WARPS * 32) sdQ[idx] = 0.f;
    // D = rowsum(dO * O); 8 consecutive lanes share one row
    for (int idx = tid; idx < ROWS * 8; idx += BWD_WARPS * 32) {
        const int r = idx >> 3, c = idx & 7;
        float acc = 0.f;
        if (r < L) {
            const uint4 uo = __ldg(reinterpret_cast<const uint4*>(obase + (int64_t)r * E + c * 8));
            const uint4 ud = __ldg(reinterpret_cast<const uint4*>(dobase + (int64_t)r * E + c * 8));
            const uint32_t wo[4] = {uo.x, uo.y, uo.z, uo.w}, wd[4] = {ud.x, ud.y, ud.z, ud.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 fo = unpack_bf16x2(wo[j]), fd = unpack_bf16x2(wd[j]);
                acc += fo.x * fd.x + fo.y * fd.y;
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (c == 0) {
            sD[r] = acc;
            sL[r] = (r < L) ? __ldg(lse + ((int64_t)b * H + h) * L + r) * 1.4426950408889634f : 0.f;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const float scale = 0.125f;
    const float scale_log2 = 0.125f * 1.4426950408889634f;
    const int nblocks = (L + 15) / 16;
    const int g = lane >> 2, t = lane & 3;
    bf16* gdk = dqkv + (int64_t)b * L * ld3 + E + h * HD;
    bf16* gdv = dqkv + (int64_t)b * L * ld3 + 2 * E + h * HD;

    for (int jb = warp; jb < nblocks; jb += BWD_WARPS) {
        // operands of this key block that stay in registers for the whole query loop
        uint32_t aK[4][4], aV[4][4], bKt[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ldmatrix_x4(aK[ks], smem_u32(sK + (jb * 16 + (lane & 15)) * SROW + ks * 16 + (lane >> 4) * 8));
            ldmatrix_x4(aV[ks], smem_u32(sV + (jb * 16 + (lane & 15)) * SROW + ks * 16 + (lane >> 4) * 8));
        }
#pragma unroll
        for (int np = 0; np < 4; ++np)  // B[k = key][n = d] for dQ = dS K_j
            ldmatrix_x4_trans(bKt[np], smem_u32(sK + (jb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SROW + np * 16 + (lane >> 4) * 8));
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
            dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
        }
        const int key0 = jb * 16 + g, key1 = key0 + 8;
#pragma unroll 1
        for (int ib = 0; ib < nblocks; ++ib) {
            // S^T = K_j Q_i^T and dP^T = V_j dO_i^T : [16 keys] x [16 queries]
            float st[2][4], dpt[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
                dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t bq[4], bd[4];
                const int off = (ib * 16 + (lane >> 4) * 8 + (lane & 7)) * SROW + ks * 16 + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(bq, smem_u32(sQ + off));
                ldmatrix_x4(bd, smem_u32(sdO + off));
                mma_bf16(st[0], aK[ks], bq[0], bq[1]);
                mma_bf16(st[1], aK[ks], bq[2], bq[3]);
                mma_bf16(dpt[0], aV[ks], bd[0], bd[1]);
                mma_bf16(dpt[1], aV[ks], bd[2], bd[3]);
            }
            // P^T = exp(S^T * scale - lse[q]); dS^T = P^T * (dP^T - D[q]) * scale
            float pt[2][4], dst[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int q0 = ib * 16 + nt * 8 + t * 2;
                const float l0 = sL[q0], l1 = sL[q0 + 1];
                const float d0 = sD[q0], d1 = sD[q0 + 1];
                const bool qv0 = q0 < L, qv1 = q0 + 1 < L;
                pt[nt][0] = (qv0 && key0 < L) ? exp2f(st[nt][0] * scale_log2 - l0) : 0.f;
                pt[nt][1] = (qv1 && key0 < L) ? exp2f(st[nt][1] * scale_log2 - l1) : 0.f;
                pt[nt][2] = (qv0 && key1 < L) ? exp2f(st[nt][2] * scale_log2 - l0) : 0.f;
                pt[nt][3] = (qv1 && key1 < L) ? exp2f(st[nt][3] * scale_log2 - l1) : 0.f;
                dst[nt][0] = pt[nt][0] * (dpt[nt][0] - d0) * scale;
                dst[nt][1] = pt[nt][1] * (dpt[nt][1] - d1) * scale;
                dst[nt][2] = pt[nt][2] * (dpt[nt][2] - d0) * scale;
                dst[nt][3] = pt[nt][3] * (dpt[nt][3] - d1) * scale;
            }
            uint32_t aP[4], aS[4];
            aP[0] = pack_bf16x2(pt[0][0], pt[0][1]);
            aP[1] = pack_bf16x2(pt[0][2], pt[0][3]);
            aP[2] = pack_bf16x2(pt[1][0], pt[1][1]);
            aP[3] = pack_bf16x2(pt[1][2], pt[1][3]);
            aS[0] = pack_bf16x2(dst[0][0], dst[0][1]);
            aS[1] = pack_bf16x2(dst[0][2], dst[0][3]);
            aS[2] = pack_bf16x2(dst[1][0], dst[1][1]);
            aS[3] = pack_bf16x2(dst[1][2], dst[1][3]);
            // dV_j += P^T dO_i ; dK_j += dS^T Q_i   (B[k = query][n = d] via ldmatrix.trans)
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                uint32_t bo[4], bq[4];
                const int off = (ib * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SROW + np * 16 + (lane >> 4) * 8;
                ldmatrix_x4_trans(bo, smem_u32(sdO + off));
                ldmatrix_x4_trans(bq, smem_u32(sQ + off));
                mma_bf16(dv[2 * np], aP, bo[0], bo[1]);
                mma_bf16(dv[2 * np + 1], aP, bo[2], bo[3]);
                mma_bf16(dk[2 * np], aS, bq[0], bq[1]);
                mma_bf16(dk[2 * np + 1], aS, bq[2], bq[3]);
            }
            // dQ_i += dS K_j : A = dS = (dS^T)^T, 8x8 blocks transposed in registers, off-diagonal blocks swapped
            uint32_t aT[4];
            aT[0] = movmatrix_trans(aS[0]);
            aT[1] = movmatrix_trans(aS[2]);
            aT[2] = movmatrix_trans(aS[1]);
            aT[3] = movmatrix_trans(aS[3]);
            float dq[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                mma_bf16(dq[2 * np], aT, bKt[np][0], bKt[np][1]);
                mma_bf16(dq[2 * np + 1], aT, bKt[np][2], bKt[np][3]);
            }
            float* r0p = sdQ + (ib * 16 + g) * DQ_STRIDE + t * 2;
            float* r1p = r0p + 8 * DQ_STRIDE;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                atomicAdd(r0p + nt * 8, dq[nt][0]);
                atomicAdd(r0p + nt * 8 + 1, dq[nt][1]);
                atomicAdd(r1p + nt * 8, dq[nt][2]);
                atomicAdd(r1p + nt * 8 + 1, dq[nt][3]);
            }
        }
        // K_j / V_j rows are only ever read by this warp: reuse them to stage dK_j / dV_j for coalesced stores
        store_tile_via_smem(sK + jb * 16 * SROW, dk, gdk, ld3, jb * 16, L, lane);
        store_tile_via_smem(sV + jb * 16 * SROW, dv, gdv, ld3, jb * 16, L, lane);
    }
    __syncthreads();
    // dQ: fp32 smem -> bf16 global
    bf16* gdq = dqkv + (int64_t)b * L * ld3 + h * HD;
    for (int idx = tid; idx < L * 8; idx += BWD_WARPS * 32) {
        const int r = idx >> 3, c = idx & 7;
        const float* p = sdQ + r * DQ_STRIDE + c * 8;
        uint4 u;
        u.x = pack_bf16x2(p[0], p[1]);
        u.y = pack_bf16x2(p[2], p[3]);
        u.z = pack_bf16x2(p[4], p[5]);
        u.w = pack_bf16x2(p[6], p[7]);
        *reinterpret_cast<uint4*>(gdq + (int64_t)r * ld3 + c * 8) = u;
    }
}

template <typename K>
static int set_smem_attr(K kern, int bytes) {
    VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return VB_OK;
}

template <int NB16, bool PAIR>
static int launch_fwd(const bf16* qa, const bf16* qb, int64_t ld_qkv, bf16* out, int64_t ld_out, float* lse, int batch,
                      int L, int H, cudaStream_t stream) {
    const int smem = (PAIR ? 6 : 3) * NB16 * 16 * SROW * 2;
    auto kern = attention_fwd_kernel<NB16, PAIR>;
    int rc = set_smem_attr(kern, smem);
    if (rc) return rc;
    kern<<<batch * H, FWD_WARPS * 32, smem, stream>>>(qa, qb, ld_qkv, out, ld_out, lse, L, H);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

template <int NB16>
static int launch_bwd(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, bf16* dqkv, int batch, int L,
                      int H, cudaStream_t stream) {
    const int rows = NB16 * 16;
    const int smem = 4 * rows * SROW * 2 + rows * DqStride<NB16>::value * 4 + 2 * rows * 4;
    auto kern = attention_bwd_kernel<NB16>;
    int rc = set_smem_attr(kern, smem);
    if (rc) return rc;
    kern<<<batch * H, BWD_WARPS * 32, smem, stream>>>(qkv, out, dout, lse, dqkv, L, H);
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// tcgen05 implementations, used for seq <= 208 (ViT-B/L at 224x224)
int launch_attention_fwd_tc3(const bf16* qkv, bf16* out, float* lse, int batch, int L, int H, cudaStream_t stream);
int launch_attention_bwd_tc3(const bf16* qkv, const bf16* out, const bf16* dout, const float* lse, float* delta, bf16* dqkv,
                             float* dbias, int bias_q_only, int batch, int L, int H, cudaStream_t stream);
int launch_colsum_bf16(const bf16* x, int64_t ldx, float* out, int rows, int cols, cudaStream_t stream);
int launch_attention_pair_tc3(const bf16* qkv_a, const bf16* qkv_b, int64_t ld, bf16* delta, int layers, int batch, int L, int H,
                              cudaStream_t stream);

// Development aid only (never set by the package): VITB200_ATTN=mma forces the mma.sync kernels of this file (the
// production path for sequences of 209..272 tokens) for every length; default = the persistent tcgen05 kernels of
// attention_tc3.cu for seq <= 208.
int launch_attention_delta_tc3(const bf16* qkv_a, const bf16* dqkv, int64_t ld, bf16* delta, int layers, int batch, int L, int H,
                               cudaStream_t stream);
static int attn_impl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VITB200_ATTN");
        v = (e != nullptr && e[0] == 'm') ? 0 : 3;
    }
    return v;
}

}  // namespace vb

extern "C" int vb_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t seq, int32_t heads,
                                int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(qkv && out, "vb_attention_fwd: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_fwd: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(batch > 0 && heads > 0 && seq > 0 && seq <= 272, "vb_attention_fwd: seq=%d must be in [1, 272]", seq);
    const int64_t E = (int64_t)heads * HD;
    if (seq <= 208 && attn_impl() == 3)
        return launch_attention_fwd_tc3(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, batch, seq, heads, stream);
    if (seq <= 208)
        return launch_fwd<13, false>(static_cast<const bf16*>(qkv), nullptr, 3 * E, static_cast<bf16*>(out), E, lse, batch,
                                     seq, heads, stream);
    return launch_fwd<17, false>(static_cast<const bf16*>(qkv), nullptr, 3 * E, static_cast<bf16*>(out), E, lse, batch, seq,
                                 heads, stream);
}

extern "C" int vb_attention_pair_delta(const void* qkv_a, const void* qkv_b, int64_t ld_qkv, void* delta, int64_t ld_delta,
                                       int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(qkv_a && qkv_b && delta, "vb_attention_pair_delta: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_pair_delta: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(batch > 0 && heads > 0 && seq > 0 && seq <= 208, "vb_attention_pair_delta: seq=%d must be in [1, 208]", seq);
    VB_CHECK_ARG(ld_qkv % 8 == 0 && ld_delta % 8 == 0, "vb_attention_pair_delta: leading dims must be multiples of 8");
    if (attn_impl() == 3 && ld_delta == (int64_t)heads * HD)
        return launch_attention_pair_tc3(static_cast<const bf16*>(qkv_a), static_cast<const bf16*>(qkv_b), ld_qkv,
                                         static_cast<bf16*>(delta), 1, batch, seq, heads, stream);
    return launch_fwd<13, true>(static_cast<const bf16*>(qkv_a), static_cast<const bf16*>(qkv_b), ld_qkv,
                                static_cast<bf16*>(delta), ld_delta, nullptr, batch, seq, heads, stream);
}

extern "C" int vb_attention_pair_delta_layers(const void* qkv_a, const void* qkv_b, int64_t ld_qkv, void* delta, int32_t layers,
                                              int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(qkv_a && qkv_b && delta, "vb_attention_pair_delta_layers: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_pair_delta_layers: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(layers > 0 && batch > 0 && heads > 0 && seq > 0 && seq <= 208, "vb_attention_pair_delta_layers: seq=%d must be in [1, 208]", seq);
    const int64_t E = (int64_t)heads * HD;
    VB_CHECK_ARG(ld_qkv % 8 == 0 && ld_qkv >= layers * 3 * E, "vb_attention_pair_delta_layers: ld_qkv=%lld must be a multiple of 8, >= layers * 3E",
                 (long long)ld_qkv);
    if (attn_impl() == 3)
        return launch_attention_pair_tc3(static_cast<const bf16*>(qkv_a), static_cast<const bf16*>(qkv_b), ld_qkv,
                                         static_cast<bf16*>(delta), layers, batch, seq, heads, stream);
    for (int i = 0; i < layers; ++i) {  // development fallback (VITB200_ATTN=...): one legacy launch per layer
        int rc = launch_fwd<13, true>(static_cast<const bf16*>(qkv_a) + i * 3 * E, static_cast<const bf16*>(qkv_b) + i * 3 * E, ld_qkv,
                                      static_cast<bf16*>(delta) + (int64_t)i * batch * seq * E, E, nullptr, batch, seq, heads, stream);
        if (rc) return rc;
    }
    return VB_OK;
}

extern "C" int vb_attention_perturb_delta_layers(const void* qkv_a, const void* dqkv, int64_t ld_qkv, void* delta, int32_t layers,
                                                 int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(qkv_a && dqkv && delta, "vb_attention_perturb_delta_layers: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_perturb_delta_layers: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(layers > 0 && batch > 0 && heads > 0 && seq > 0 && seq <= 208, "vb_attention_perturb_delta_layers: seq=%d must be in [1, 208]", seq);
    const int64_t E = (int64_t)heads * HD;
    VB_CHECK_ARG(ld_qkv % 8 == 0 && ld_qkv >= layers * 3 * E, "vb_attention_perturb_delta_layers: ld_qkv=%lld must be a multiple of 8, >= layers * 3E",
                 (long long)ld_qkv);
    return launch_attention_delta_tc3(static_cast<const bf16*>(qkv_a), static_cast<const bf16*>(dqkv), ld_qkv, static_cast<bf16*>(delta), layers,
                                      batch, seq, heads, stream);
}

extern "C" int64_t vb_attention_bwd_workspace_bytes(int32_t batch, int32_t seq, int32_t heads) {
    return (int64_t)batch * heads * seq * 4;
}

static int attention_bwd_dispatch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* workspace,
                                  int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, float* dbias, bool* dbias_done,
                                  cudaStream_t stream);

extern "C" int vb_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                void* workspace, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                                vb_stream_t stream_) {
    bool done = false;
    return attention_bwd_dispatch(qkv, out, dout, lse, dqkv, workspace, batch, seq, heads, head_dim, nullptr, &done,
                                  static_cast<cudaStream_t>(stream_));
}

extern "C" int vb_attention_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                     float* dbias, void* workspace, int32_t batch, int32_t seq, int32_t heads,
                                     int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(dbias != nullptr, "vb_attention_bwd_bias: null dbias");
    bool done = false;
    int rc = attention_bwd_dispatch(qkv, out, dout, lse, dqkv, workspace, batch, seq, heads, head_dim, dbias, &done, stream);
    if (rc != VB_OK || done) return rc;
    // kernels without the fused reduction (legacy / long-sequence paths): one column-sum pass over dqkv
    return launch_colsum_bf16(static_cast<const bf16*>(dqkv), 3LL * heads * head_dim, dbias, batch * seq, 3 * heads * head_dim, stream);
}

extern "C" int vb_attention_bwd_with_delta(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias,
                                           const void* workspace, int32_t batch, int32_t seq, int32_t heads,
                                           int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(qkv && dout && lse && dqkv && workspace, "vb_attention_bwd_with_delta: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_bwd_with_delta: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(batch > 0 && heads > 0 && seq > 0 && seq <= 208, "vb_attention_bwd_with_delta: seq=%d must be in [1, 208]", seq);
    return launch_attention_bwd_tc3(static_cast<const bf16*>(qkv), nullptr, static_cast<const bf16*>(dout), lse,
                                    static_cast<float*>(const_cast<void*>(workspace)), static_cast<bf16*>(dqkv), dbias, 0, batch, seq, heads,
                                    static_cast<cudaStream_t>(stream_));
}

extern "C" int vb_attention_bwd_with_delta_qbias(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias_q,
                                                 const void* workspace, int32_t batch, int32_t seq, int32_t heads,
                                                 int32_t head_dim, vb_stream_t stream_) {
    using namespace vb;
    VB_CHECK_ARG(qkv && dout && lse && dqkv && workspace && dbias_q, "vb_attention_bwd_with_delta_qbias: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_bwd_with_delta_qbias: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(batch > 0 && heads > 0 && seq > 0 && seq <= 208, "vb_attention_bwd_with_delta_qbias: seq=%d must be in [1, 208]", seq);
    return launch_attention_bwd_tc3(static_cast<const bf16*>(qkv), nullptr, static_cast<const bf16*>(dout), lse,
                                    static_cast<float*>(const_cast<void*>(workspace)), static_cast<bf16*>(dqkv), dbias_q, 1, batch, seq,
                                    heads, static_cast<cudaStream_t>(stream_));
}

static int attention_bwd_dispatch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* workspace,
                                  int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, float* dbias, bool* dbias_done,
                                  cudaStream_t stream) {
    using namespace vb;
    VB_CHECK_ARG(qkv && out && dout && lse && dqkv, "vb_attention_bwd: null pointer");
    VB_CHECK_ARG(head_dim == HD, "vb_attention_bwd: head_dim must be 64 (got %d)", head_dim);
    VB_CHECK_ARG(batch > 0 && heads > 0 && seq > 0 && seq <= 272, "vb_attention_bwd: seq=%d must be in [1, 272]", seq);
    if (seq <= 208 && attn_impl() == 3) {
        VB_CHECK_ARG(workspace != nullptr, "vb_attention_bwd: workspace of vb_attention_bwd_workspace_bytes() bytes required");
        *dbias_done = true;
        return launch_attention_bwd_tc3(static_cast<const bf16*>(qkv), static_cast<const bf16*>(out),
                                        static_cast<const bf16*>(dout), lse, static_cast<float*>(workspace),
                                        static_cast<bf16*>(dqkv), dbias, 0, batch, seq, heads, stream);
    }
    if (seq <= 208)
        return launch_bwd<13>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(out), static_cast<const bf16*>(dout),
                              lse, static_cast<bf16*>(dqkv), batch, seq, heads, stream);
    return launch_bwd<17>(static_cast<const bf16*>(qkv), static_cast<const bf16*>(out), static_cast<const bf16*>(dout), lse,
                          static_cast<bf16*>(dqkv), batch, seq, heads, stream);
}
