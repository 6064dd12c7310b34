// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring -> tcgen05.mma -> double-buffered fp32 accumulators in
//   TMEM (2 x 256 columns) -> tcgen05.ld epilogue on 8 warps -> swizzled smem staging -> TMA store / TMA reduce-add.
// Tile mapping G: 2 (default) = a CTA pair (cluster of 2) per 256 x 256 tile, ONE tcgen05.mma.cta_group::2 256x256x16 per
//   k-step, each CTA stages its own 128 rows of A and half of the B tile (6 x 32 KB ring, 5 for the two-output / aux
//   epilogues); 1 = one CTA per 128 x 256 tile (cta_group::1, 4 x 48 KB ring).
//
// One kernel serves every dense contraction of the ViT hot path:
//   forward  y  = x W^T (+bias, +GELU(+GELU'), +residual)          A K-major,  B K-major
//   dgrad    dx = dy W  (x gelu'(z), + row dots with a 2nd operand)  A K-major,  B MN-major (weights used as stored)
//   wgrad    dW += dy^T x  (split-K, TMA reduce-add)                 A MN-major, B MN-major (activations used as stored)
//   plasticity: sum of squares of the accumulator per sample, nothing written back
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA of a pair), warp 2 = TMEM allocator,
// warp 3 = tile fetcher of the optional work-stealing scheduler (else idle); warps 4..11 = epilogue (warp % 4 selects the
// TMEM lane quarter, (warp - 4) / 4 the 128-column half of the tile). Registers: 56 for warps 0..3, 224 for the epilogue
// (setmaxnreg; the kernel must stay free of CALLs for that, see ptx.cuh).
#include "host_utils.h"
#define VB_MBAR_TRAP_PRINTF 0  // no CALL in this kernel: see ptx.cuh (per-role register budgets via setmaxnreg)
#include "ptx.cuh"

namespace vb {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
// G = CTAs per tile: 1 = one CTA computes 128 x 256 (cta_group::1); 2 = a CTA pair computes 256 x 256 with ONE
// tcgen05.mma.cta_group::2 per k-step: each CTA stages its own 128 rows of A and only HALF of the B tile (128 columns),
// so the smem operand traffic per SM per MMA drops from 12 KB to 8 KB and the ring holds 6 stages instead of 4.
// DEEP (pairs only) trades one ring stage for twice the epilogue staging: 5 x 32 KB + 8 x 8 KB instead of 6 x 32 KB +
// 8 x 4 KB, so a two-output (GELU + GELU') chunk or an fp32 chunk is double-buffered against its TMA store.
template <int G, int DEEP>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = (BN / G) * BK * 2;  // 32 KB / 16 KB
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = G == 1 ? 4 : (DEEP ? 5 : 6);
    // per epilogue warp: 2 KB buffers (32 rows x 64 B bf16; an fp32 chunk or a two-output chunk takes two)
    static constexpr int STAGING_BYTES = (G == 2 && DEEP) ? 8192 : 4096;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 8 * STAGING_BYTES + 512 + 1024;  // + barriers + alignment slack
};
constexpr int MAX_STAGES = 6;
constexpr int SCHED_SLOTS = 4;  // tile indices in flight between the fetcher and the slowest role (the epilogue)

constexpr int EPI_WARPS = 8;
constexpr int EPI_FIRST_WARP = 4;
constexpr int GEMM_THREADS = (EPI_FIRST_WARP + EPI_WARPS) * 32;
constexpr int TMEM_COLS = 512;
static_assert(GemmCfg<1, 0>::SMEM_BYTES <= 232448 && GemmCfg<2, 0>::SMEM_BYTES <= 232448 && GemmCfg<2, 1>::SMEM_BYTES <= 232448,
              "dynamic smem limit of sm_100");

struct GemmKernelParams {
    int M, N, K;
    int num_m_blocks, num_n_blocks, num_k_blocks;
    int split_k, kb_per_split;
    int epi;
    const float* bias;
    const bf16* aux;
    long long ld_aux;
    float* sumsq;
    int rows_per_sample, cols_per_group, n_groups;
    int has_out2;
    float* out_colsum;  // f32 [N] or null: += column sums of the bf16 output (bias gradient of the layer that produced A's grad)
    unsigned* sched;        // work-stealing tile scheduler: this launch's tile counter (starts at 0), or null = static stride
    unsigned* sched_clear;  // a counter slot of a FUTURE launch, zeroed by this one
};

struct WorkItem {
    int m_blk, n_blk, kb0, kb1;
};

// num_m_blocks counts tiles of G x 128 rows; the returned m_blk is THIS CTA's 128-row block.
template <int G>
__device__ __forceinline__ WorkItem decode_work(const GemmKernelParams& p, int w, int rank) {
    // split index is the slowest so that concurrently running CTAs share the same K-slice (L2 reuse in wgrad);
    // within a split, n is fastest so CTAs running together share the A row panel.
    int tiles = p.num_m_blocks * p.num_n_blocks;
    int split = w / tiles;
    int tile = w - split * tiles;
    WorkItem it;
    const int mt = tile / p.num_n_blocks;
    it.m_blk = mt * G + rank;
    it.n_blk = tile - mt * p.num_n_blocks;
    it.kb0 = split * p.kb_per_split;
    it.kb1 = min(it.kb0 + p.kb_per_split, p.num_k_blocks);
    return it;
}

// EPI_T >= 0: the epilogue is fixed at compile time (the hot combinations); -1: taken from p.epi at run time.
// BNT = tile width in columns: 256, or 192 (pairs only, N % 192 == 0) for the N = 768 GEMMs at few token rows: 50 x 3 tiles
// of 256 columns are 2.03 waves on 74 CTA pairs (3 waves paid), 50 x 4 tiles of 192 columns are 2.7 waves of 3/4 the work.
// Each CTA of a pair then stages 96 columns of B (K-major: a 96-row box; MN-major: two 64-column boxes, the upper half
// of the second one unused) in the same 16 KB slot; the accumulators keep their 256-column stride in TMEM.
template <int A_MN, int B_MN, int G, int EPI_T, int DEEP, int BNT = BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                    const __grid_constant__ CUtensorMap tmAux, const GemmKernelParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned stage buffers
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    using Cfg = GemmCfg<G, DEEP>;
    constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int STAGING = Cfg::STAGING_BYTES, NBUF = STAGING / 2048;
    static_assert(BNT == BN || (BNT == 192 && G == 2 && (EPI_T == VB_EPI_BF16 || EPI_T == VB_EPI_BF16_RESID)),
                  "tile width: 256, or 192 for the CTA-pair kernels of the plain / residual bf16 epilogues");
    constexpr int BHALF = BNT / 2;   // columns per epilogue warp
    constexpr int NCH = BHALF / 32;  // 32-column chunks per epilogue warp and tile
    // bytes one CTA's producer puts into a stage (the MN-major B boxes are 64 columns wide: 96 columns take two)
    constexpr int B_LOAD_BYTES = B_MN == 0 ? (BNT / G) * BK * 2 : ((BNT / G + 63) / 64) * 64 * BK * 2;
    constexpr int STAGE_TX_BYTES = A_STAGE_BYTES + B_LOAD_BYTES;
    uint8_t* staging_base = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging_base + EPI_WARPS * STAGING);
    uint64_t* full_bar = bars;                       // [STAGES]  TMA -> MMA (G = 2: the leader's, fed by both CTAs)
    uint64_t* empty_bar = bars + MAX_STAGES;         // [STAGES]  MMA -> TMA (G = 2: multicast commit to both CTAs)
    uint64_t* tfull_bar = bars + 2 * MAX_STAGES;     // [2]       MMA -> epilogue (G = 2: multicast)
    uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;  // [2]     epilogue -> MMA (G = 2: both CTAs' warps arrive on the leader's)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
    uint64_t* aux_bar = bars + 2 * MAX_STAGES + 6;   // [EPI_WARPS][4]  TMA (aux chunk) -> epilogue warp, AUX_TMA only
    uint64_t* sched_full = bars + 2 * MAX_STAGES + 6 + EPI_WARPS * 4;  // [SCHED_SLOTS] fetcher -> every role of both CTAs
    uint64_t* sched_empty = sched_full + SCHED_SLOTS;                  // [SCHED_SLOTS] (leader's) every role -> fetcher
    volatile int* sched_ids = reinterpret_cast<volatile int*>(sched_empty + SCHED_SLOTS);  // [SCHED_SLOTS] tile index or -1
    // The residual / saved-GELU' operand of the epilogue comes in through TMA, straight into the staging buffer the output
    // chunk leaves from: two chunks ahead, no per-lane strided global loads, the result overwrites it in place.
    constexpr bool AUX_TMA = G == 2 && DEEP == 1 && (EPI_T == VB_EPI_BF16_RESID || EPI_T == VB_EPI_BF16_MULAUX || EPI_T == VB_EPI_BF16_ROWDOT);
    // rank of this CTA in its pair (0 = leader: issues the MMAs); work is distributed over pairs
    const int rank = G == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    const int unit = static_cast<int>(blockIdx.x) / G, num_units = static_cast<int>(gridDim.x) / G;

    // warp index through a shuffle: provably warp-uniform, so per-warp addresses / coordinates live in uniform registers
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmC2);
        if (AUX_TMA) tma_prefetch_desc(&tmAux);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], EPI_WARPS * G);
        }
        if (AUX_TMA)
            for (int a = 0; a < EPI_WARPS * 4; ++a) mbar_init(&aux_bar[a], 1);
        for (int a = 0; a < SCHED_SLOTS; ++a) {
            mbar_init(&sched_full[a], 1);
            // consumers of a tile index: TMA warp + 8 epilogue warps of each CTA, + the leader's MMA warp
            mbar_init(&sched_empty[a], G * (1 + EPI_WARPS) + 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (G == 1) {
            tmem_alloc(tmem_ptr_smem, TMEM_COLS);
            tmem_relinquish();
        } else {
            tmem_alloc_pair(tmem_ptr_smem, TMEM_COLS);
            tmem_relinquish_pair();
        }
    }
    tc_fence_before();
    if (G == 1)
        __syncthreads();
    else
        cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive / complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // The launch gives every thread 168 registers; the TMA / MMA warpgroup needs few, the two epilogue warpgroups hold two
    // 32-register accumulator chunks plus the epilogue operands: 128 x 56 + 256 x 224 = 64 512 registers.

    const int total_work = p.num_m_blocks * p.num_n_blocks * p.split_k;

    // Tile order. Static: pair u takes tiles u, u + num_units, ... Work-stealing (p.sched != null): warp 3 of the leader CTA
    // draws tile indices from a global counter and hands each one to every role of both CTAs through a 4-slot ring, so a
    // CTA pair that becomes resident late (SMs held by another kernel, e.g. an overlapped all-reduce) finds the remaining
    // tiles already taken instead of running its whole static share after everybody else has finished.
    const bool dyn = p.sched != nullptr;
    struct TileIter {
        int w, slot;
        uint32_t ph;
    };
    auto next_tile = [&](TileIter& ti) -> int {
        if (!dyn) {
            const int r = ti.w < total_work ? ti.w : -1;
            ti.w += num_units;
            return r;
        }
        mbar_wait(&sched_full[ti.slot], ti.ph, 6);
        const int tile = sched_ids[ti.slot];
        __syncwarp();
        if (elect_one()) {
            if (G == 1)
                mbar_arrive(&sched_empty[ti.slot]);
            else
                mbar_arrive_leader(&sched_empty[ti.slot]);
        }
        if (++ti.slot == SCHED_SLOTS) ti.slot = 0, ti.ph ^= 1;
        return tile;
    };
    // The producer and MMA warps run their loops with ALL 32 lanes in uniform control flow and elect one lane only
    // around the asynchronous instructions: addresses / descriptors stay in uniform registers. (With the whole loop
    // under `if (lane == 0)` the compiler wrapped every UTCHMMA / UTMALDG in a divergence "waterfall" loop of ~20
    // instructions, which made the single issuing thread the bottleneck.)
    if (warp < EPI_FIRST_WARP) {
    setmaxnreg_dec<56>();
    if (warp == 0) {
        // =========================== TMA producer ===========================
        int stage = 0;
        uint32_t phase = 0;
        auto load = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
            if (G == 1)
                tma_load_2d(dst, map, bar, c0, c1);
            else
                tma_load_2d_pair(dst, map, bar, c0, c1);
        };
        TileIter ti = {unit, 0, 0};
        for (int w = next_tile(ti); w >= 0; w = next_tile(ti)) {
            const WorkItem it = decode_work<G>(p, w, rank);
            const int n0 = it.n_blk * BNT + rank * (BNT / G);  // this CTA's share of the B tile
            for (int kb = it.kb0; kb < it.kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_STAGE_BYTES;
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], STAGE_TX_BYTES * G);
                    if (A_MN == 0) {
                        load(sa, &tmA, &full_bar[stage], kb * BK, it.m_blk * BM);
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            load(sa + i * (64 * BK * 2), &tmA, &full_bar[stage], it.m_blk * BM + i * 64, kb * BK);
                    }
                    if (B_MN == 0) {
                        load(sb, &tmB, &full_bar[stage], kb * BK, n0);
                    } else {
#pragma unroll
                        for (int i = 0; i < (BNT / G + 63) / 64; ++i)
                            load(sb + i * (64 * BK * 2), &tmB, &full_bar[stage], n0 + i * 64, kb * BK);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // =========================== MMA issuer (leader CTA only when G = 2) ===========================
        constexpr uint32_t idesc = make_idesc_bf16(BM * G, BNT, A_MN, B_MN);
        // descriptor = constant high word (SBO 1024 B, version 1, SWIZZLE_128B) + low word (address, LBO)
        constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t lbo_a = A_MN == 0 ? 16u : (64u * BK * 2u), lbo_b = B_MN == 0 ? 16u : (64u * BK * 2u);
        constexpr uint32_t kstep_a = (A_MN == 0 ? 32u : 2048u) >> 4, kstep_b = (B_MN == 0 ? 32u : 2048u) >> 4;
        const uint32_t smem_lo = smem_u32(smem) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        TileIter ti = {unit, 0, 0};
        for (int w = next_tile(ti); w >= 0; w = next_tile(ti)) {
            const WorkItem it = decode_work<G>(p, w, rank);
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = it.kb0; kb < it.kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase, 3);
                tc_fence_after();
                // K-major : 8-row groups 1024 B apart (SBO), step 16 elements = 32 B inside the swizzle span
                // MN-major: 64-element MN chunks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO), step 16 k-rows = 2048 B
                const uint32_t a_lo = (smem_lo + stage * (STAGE_BYTES >> 4)) | ((lbo_a >> 4) << 16);
                const uint32_t b_lo = (smem_lo + stage * (STAGE_BYTES >> 4) + (A_STAGE_BYTES >> 4)) | ((lbo_b >> 4) << 16);
                const uint32_t first = kb > it.kb0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        if (G == 1)
                            umma_bf16_ss(tmem_d, make_desc(a_lo + k * kstep_a, desc_hi), make_desc(b_lo + k * kstep_b, desc_hi),
                                         idesc, k > 0 ? 1u : first);
                        else
                            umma_bf16_ss_pair(tmem_d, make_desc(a_lo + k * kstep_a, desc_hi),
                                              make_desc(b_lo + k * kstep_b, desc_hi), idesc, k > 0 ? 1u : first);
                    }
                    // smem slot is free once these MMAs have read it (G = 2: in both CTAs)
                    if (G == 1)
                        umma_commit(&empty_bar[stage]);
                    else
                        umma_commit_pair(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) {  // accumulator complete
                if (G == 1)
                    umma_commit(&tfull_bar[acc]);
                else
                    umma_commit_pair(&tfull_bar[acc]);
            }
            __syncwarp();
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (warp == 3 && rank == 0 && dyn) {
        // =========================== tile fetcher (work-stealing scheduler) ===========================
        if (lane == 0) {
            if (blockIdx.x == 0) *p.sched_clear = 0u;  // the counter of a launch far in the future
            int slot = 0;
            uint32_t ph = 0;
            while (true) {
                mbar_wait(&sched_empty[slot], ph ^ 1, 7);
                const unsigned t = atomicAdd(p.sched, 1u);
                const int tile = t < static_cast<unsigned>(total_work) ? static_cast<int>(t) : -1;
                const uint32_t id_addr = smem_u32(const_cast<int*>(&sched_ids[slot])), full_addr = smem_u32(&sched_full[slot]);
                if (G == 1) {
                    sched_ids[slot] = tile;
                    mbar_arrive(&sched_full[slot]);  // (release at CTA scope orders the store before it)
                } else {
                    for (int r = 0; r < G; ++r) {  // asynchronous store + transaction count: waiters need no cluster-scope acquire
                        const uint32_t bar_r = mapa_u32(full_addr, r);
                        mbar_arrive_expect_tx_cluster(bar_r, 4);
                        st_async_u32(mapa_u32(id_addr, r), static_cast<uint32_t>(tile), bar_r);
                    }
                }
                if (tile < 0) break;
                if (++slot == SCHED_SLOTS) slot = 0, ph ^= 1;
            }
        }
        __syncwarp();
    }
    } else {
        // =========================== epilogue ===========================
        setmaxnreg_inc<224>();
        // Per tile each warp drains a 32-row x 128-column block of the accumulator in four 32-column chunks (BNT = 192: 96
        // columns, three chunks). The TMEM load
        // of chunk c + 1 is in flight while chunk c is processed (two register buffers), and the accumulator is handed back
        // to the MMA warp as soon as the last load has landed, before that chunk's math and stores.
        const int ew = warp - EPI_FIRST_WARP;
        const int q = warp & 3;   // TMEM lane quarter this warp may access
        const int hf = ew >> 2;   // which half of the tile's columns
        uint8_t* stg = staging_base + ew * STAGING;
        const int epi = EPI_T >= 0 ? EPI_T : p.epi;
        const bool has_aux = (epi == VB_EPI_BF16_RESID || epi == VB_EPI_BF16_DGELU || epi == VB_EPI_BF16_MULAUX || epi == VB_EPI_BF16_ROWDOT);
        const bool two = (epi == VB_EPI_BF16_GELU || epi == VB_EPI_BF16_GELU_GRAD) && p.has_out2;
        const bool add_bias = p.bias != nullptr && epi != VB_EPI_F32_ADD && epi != VB_EPI_SUMSQ && epi != VB_EPI_BF16_DGELU &&
                              epi != VB_EPI_BF16_MULAUX;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t ring = 0;  // staging buffers used so far (2 KB units for bf16 chunks, 4 KB for two-output / fp32 chunks)
        // AUX_TMA: chunk number t (counted over this warp's tiles, 4 per tile) lives in staging buffer t % 4; its aux
        // operand is requested while chunk t - 2 is processed, after the store that last read the buffer (chunk t - 4).
        uint64_t* my_aux_bar = aux_bar + ew * 4;
        auto request_aux = [&](const WorkItem& wi, int c, uint32_t t) {
            if (elect_one()) {
                tma_store_wait_read<1>();
                uint64_t* bar = my_aux_bar + (t & 3);
                mbar_arrive_expect_tx(bar, 2048);
                tma_load_2d(stg + (t & 3) * 2048, &tmAux, bar, wi.n_blk * BNT + hf * BHALF + c * 32, wi.m_blk * BM + q * 32);
            }
            __syncwarp();
        };
        uint32_t tchunk = 0;
        TileIter ti = {unit, 0, 0};
        int w = next_tile(ti);
        int w_next = w >= 0 ? next_tile(ti) : -1;  // the epilogue looks one tile ahead (aux prefetch)
        if (AUX_TMA && w >= 0) {
            const WorkItem first = decode_work<G>(p, w, rank);
            request_aux(first, 0, 0);
            request_aux(first, 1, 1);
        }
        for (; w >= 0; w = w_next, w_next = (w >= 0 ? next_tile(ti) : -1)) {
            const WorkItem it = decode_work<G>(p, w, rank);
            const bool have_next = w_next >= 0;
            const WorkItem nxt_it = (AUX_TMA && have_next) ? decode_work<G>(p, w_next, rank) : it;
            const int row0 = it.m_blk * BM + q * 32;
            const int row = row0 + lane;
            const int colbase = it.n_blk * BNT + hf * BHALF;
            const bool row_ok = row < p.M;

            uint4 aux_cur[4], aux_nxt[4];
            auto load_aux = [&](uint4(&dst)[4], int col0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(0, 0, 0, 0);
                if (has_aux && row_ok) {
                    const bf16* src = p.aux + (long long)row * p.ld_aux + col0;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col0 + j * 8 < p.N) dst[j] = __ldg(reinterpret_cast<const uint4*>(src) + j);
                }
            };
            if (!AUX_TMA) load_aux(aux_cur, colbase);  // issued before the accumulator wait: latency hides behind the MMA
            if (!AUX_TMA && has_aux && have_next) {
                // The aux rows of this CTA's NEXT tile go to L2 now: a whole main loop ahead of their use, so the
                // per-chunk loads above hit L2 (~300 cycles) instead of HBM (~1500), which had made the short-K
                // residual / GELU' epilogues latency-bound (proj forward: 136 us against 93 us without epilogue).
                const WorkItem nx = decode_work<G>(p, w_next, rank);
                const int nrow = nx.m_blk * BM + q * 32 + lane, ncol = nx.n_blk * BNT + hf * BHALF;
                if (nrow < p.M && ncol < p.N) {
                    const bf16* pa = p.aux + (long long)nrow * p.ld_aux + ncol;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
                    if (ncol + 64 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 64));
                }
            }

            mbar_wait(&tfull_bar[acc], acc_phase, 4);
            tc_fence_after();

            float sumsq_local = 0.f;
            float rowdot = 0.f;  // ROWDOT: this row's partial sum over the current 64-column group (two chunks)
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + hf * BHALF;

            // one 32-column chunk, accumulator values in v (as loaded from TMEM)
            auto process = [&](uint32_t(&v)[32], const int c) {
                const int col0 = colbase + c * 32;
                uint8_t* abuf = stg + (tchunk & 3) * 2048;  // AUX_TMA: aux chunk in, output chunk out
                if (AUX_TMA) {
                    // the aux operand of the chunk after next; then wait for this chunk's (every chunk is requested and
                    // awaited, also one that lies outside the matrix: the TMA unit zero-fills it)
                    if (c + 2 < NCH)
                        request_aux(it, c + 2, tchunk + 2);
                    else if (have_next)
                        request_aux(nxt_it, c + 2 - NCH, tchunk + 2);
                    mbar_wait(my_aux_bar + (tchunk & 3), (tchunk >> 2) & 1, 5);
                }
                ++tchunk;
                if (!(col0 < p.N && row0 < p.M)) return;  // (G = 2: the odd CTA's rows of the last tile may all lie beyond M)
                float f[32];
                if (add_bias) {
                    if (col0 + 32 <= p.N) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
                            const uint64_t s0 = add_f32x2(pack_f32x2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                                                          pack_f32x2(b4.x, b4.y));
                            const uint64_t s1 = add_f32x2(pack_f32x2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])),
                                                          pack_f32x2(b4.z, b4.w));
                            unpack_f32x2(s0, f[4 * j], f[4 * j + 1]);
                            unpack_f32x2(s1, f[4 * j + 2], f[4 * j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + (col0 + j < p.N ? __ldg(p.bias + col0 + j) : 0.f);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                }
                if (epi == VB_EPI_SUMSQ) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N) sumsq_local += f[j] * f[j];
                } else if (epi == VB_EPI_F32 || epi == VB_EPI_F32_ADD) {
                    // 32 rows x 128 B, 16-byte chunk index XOR (row & 7) == TMA SWIZZLE_128B
                    uint8_t* b0 = stg + (ring % (NBUF / 2)) * 4096;
                    ++ring;
                    if (elect_one()) tma_store_wait_read<NBUF / 2 - 1>();
                    __syncwarp();
                    const uint32_t rowaddr = smem_u32(b0) + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t a = rowaddr + ((j ^ (lane & 7)) << 4);
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(f[4 * j]),
                                     "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3])
                                     : "memory");
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (elect_one()) {
                        if (epi == VB_EPI_F32)
                            tma_store_2d(&tmC, b0, col0, row0);
                        else
                            tma_reduce_add_2d(&tmC, b0, col0, row0);
                        tma_store_commit();
                    }
                } else {
                    // bf16 outputs: 32 rows x 64 B, chunk index XOR ((row >> 1) & 3) == TMA SWIZZLE_64B
                    uint32_t o[16], o2[16];
                    uint32_t axs[16];
                    if (AUX_TMA) {
                        // this lane's row of the aux chunk: 64 B, 16-byte pieces XOR-swizzled like the output (SWIZZLE_64B)
                        const uint32_t ra = smem_u32(abuf) + lane * 64;
                        const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(axs[4 * j]), "=r"(axs[4 * j + 1]), "=r"(axs[4 * j + 2]), "=r"(axs[4 * j + 3])
                                         : "r"(ra + ((j ^ sw) << 4))
                                         : "memory");
                    }
                    if (epi == VB_EPI_BF16_RESID) {
                        const uint32_t* ax = AUX_TMA ? axs : reinterpret_cast<const uint32_t*>(aux_cur);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 r = unpack_bf16x2(ax[j]);
                            o[j] = pack_bf16x2(f[2 * j] + r.x, f[2 * j + 1] + r.y);
                        }
                    } else if (epi == VB_EPI_BF16_DGELU) {
                        const uint32_t* ax = reinterpret_cast<const uint32_t*>(aux_cur);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 z = unpack_bf16x2(ax[j]);
                            o[j] = pack_bf16x2(f[2 * j] * dgelu_erf(z.x), f[2 * j + 1] * dgelu_erf(z.y));
                        }
                    } else if (epi == VB_EPI_BF16_MULAUX) {
                        const uint32_t* ax = AUX_TMA ? axs : reinterpret_cast<const uint32_t*>(aux_cur);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 g = unpack_bf16x2(ax[j]);
                            o[j] = pack_bf16x2(f[2 * j] * g.x, f[2 * j + 1] * g.y);
                        }
                    } else if (epi == VB_EPI_BF16_ROWDOT) {
                        const uint32_t* ax = AUX_TMA ? axs : reinterpret_cast<const uint32_t*>(aux_cur);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            o[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                            const float2 v = unpack_bf16x2(o[j]), w = unpack_bf16x2(ax[j]);  // the values as stored
                            rowdot = fmaf(v.x, w.x, fmaf(v.y, w.y, rowdot));
                        }
                        if (c & 1) {  // second half of a 64-column group: one plain store per (row, group)
                            if (row_ok) {
                                const int sample = row / p.rows_per_sample, q = row - sample * p.rows_per_sample;
                                p.sumsq[((long long)sample * p.n_groups + (col0 >> 6)) * p.rows_per_sample + q] = rowdot;
                            }
                            rowdot = 0.f;
                        }
                    } else if (epi == VB_EPI_BF16_GELU_GRAD) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) gelu_and_grad_erf_x2(f[2 * j], f[2 * j + 1], o[j], o2[j]);
                    } else if (epi == VB_EPI_BF16_GELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            // GELU is applied to the bf16-rounded pre-activation so that forward and the
                            // saved z used by backward agree exactly
                            const uint32_t zz = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                            const float2 z = unpack_bf16x2(zz);
                            o2[j] = zz;
                            o[j] = pack_bf16x2(gelu_erf(z.x), gelu_erf(z.y));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                    }
                    if (p.out_colsum != nullptr) {
                        // Column sums of this warp's 32 x 32 output block (as rounded to bf16, i.e. exactly what a
                        // separate pass over the stored tensor would add up): butterfly transpose-reduce over the 32
                        // lanes (= rows), 31 shuffles, lane c ends with column c; one reduction per column. Replaces a
                        // full HBM pass over the output (dz of fc1: 620 MB per layer) by ~125 instructions per chunk.
                        float cs[32];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 t = unpack_bf16x2(o[j]);
                            cs[2 * j] = row_ok ? t.x : 0.f;
                            cs[2 * j + 1] = row_ok ? t.y : 0.f;
                        }
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) {
                            const bool hi = (lane & off) != 0;
#pragma unroll
                            for (int i = 0; i < off; ++i) {
                                const float send = hi ? cs[i] : cs[i + off];
                                const float keep = hi ? cs[i + off] : cs[i];
                                cs[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                        if (col0 + lane < p.N) atomicAdd(p.out_colsum + col0 + lane, cs[0]);
                    }
                    uint8_t* b0;
                    if (AUX_TMA) {
                        b0 = abuf;  // free: the store that last read it was awaited before the aux load was issued
                        ++ring;
                    } else if (two) {
                        b0 = stg + (ring % (NBUF / 2)) * 4096;
                        if (elect_one()) tma_store_wait_read<NBUF / 2 - 1>();
                    } else {
                        b0 = stg + (ring % NBUF) * 2048;
                        if (elect_one()) tma_store_wait_read<NBUF - 1>();
                    }
                    if (!AUX_TMA) ++ring;
                    __syncwarp();
                    const uint32_t rowaddr = smem_u32(b0) + lane * 64;
                    const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t a = rowaddr + ((j ^ sw) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(o[4 * j]),
                                     "r"(o[4 * j + 1]), "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                                     : "memory");
                        if (two)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(o2[4 * j]),
                                         "r"(o2[4 * j + 1]), "r"(o2[4 * j + 2]), "r"(o2[4 * j + 3])
                                         : "memory");
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (elect_one()) {
                        tma_store_2d(&tmC, b0, col0, row0);
                        if (two) tma_store_2d(&tmC2, b0 + 2048, col0, row0);
                        tma_store_commit();
                    }
                }
            };

            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(tacc, va);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (!AUX_TMA && c < NCH - 1) load_aux(aux_nxt, colbase + c * 32 + 32);
                if ((c & 1) == 0) {
                    tmem_ld_wait_x32(va);
                    if (c < NCH - 1) tmem_ld_32x32b_x32(tacc + (c + 1) * 32, vb);
                } else {
                    tmem_ld_wait_x32(vb);
                    if (c < NCH - 1) tmem_ld_32x32b_x32(tacc + (c + 1) * 32, va);
                }
                if (c == NCH - 1) {
                    // every tcgen05.ld of this accumulator has completed: hand it back to the MMA warp now
                    tc_fence_before();
                    __syncwarp();
                    if (elect_one()) {
                        if (G == 1)
                            mbar_arrive(&tempty_bar[acc]);
                        else
                            mbar_arrive_leader(&tempty_bar[acc]);
                    }
                }
                if ((c & 1) == 0)
                    process(va, c);
                else
                    process(vb, c);
#pragma unroll
                for (int j = 0; j < 4; ++j) aux_cur[j] = aux_nxt[j];
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;

            if (epi == VB_EPI_SUMSQ && colbase < p.N) {
                const int group = colbase / p.cols_per_group;
                const int sample = row_ok ? row / p.rows_per_sample : -1;
                const int s0 = __shfl_sync(0xffffffffu, sample, 0);
                const bool uniform = __all_sync(0xffffffffu, sample == s0);
                if (uniform) {
                    const float t = warp_sum(sumsq_local);
                    if (lane == 0 && s0 >= 0) atomicAdd(p.sumsq + (long long)s0 * p.n_groups + group, t);
                } else if (row_ok) {
                    atomicAdd(p.sumsq + (long long)sample * p.n_groups + group, sumsq_local);
                }
            }
        }
        if (elect_one()) tma_store_wait_all<0>();
    }

    tc_fence_before();
    if (G == 1)
        __syncthreads();
    else
        cluster_sync_all();  // neither CTA may leave (or free TMEM) while the pair's MMAs / remote arrives are in flight
    if (warp == 2) {
        tc_fence_after();
        if (G == 1)
            tmem_dealloc(tmem_base, TMEM_COLS);
        else
            tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

template <int A_MN, int B_MN, int G, int EPI_T, int DEEP, int BNT = BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                       const CUtensorMap& tmAux, const GemmKernelParams& p, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    auto kern = gemm_tcgen05_kernel<A_MN, B_MN, G, EPI_T, DEEP, BNT>;
    constexpr int GEMM_SMEM_BYTES = GemmCfg<G, DEEP>::SMEM_BYTES;
    if (dev < 64 && !attr_set[dev]) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
        attr_set[dev] = true;
    }
    const int total_work = p.num_m_blocks * p.num_n_blocks * p.split_k;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = GEMM_SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = G > 1 ? 1 : 0;
    // CTAs (G = 1) or CTA pairs (G = 2) that can be resident at once. The tile loop strides statically over the grid, so a
    // pair that only becomes resident after another one has EXITED would serialise its whole share: never launch more pairs
    // than the device can co-schedule (asked once per kernel and device; 74 on a full B200).
    static int max_units[64] = {0};
    int units = num_sms() / G;
    if (G > 1 && dev < 64) {
        if (max_units[dev] == 0) {
            int n = 0;
            cfg.gridDim = dim3(units * G);
            if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
                (void)cudaGetLastError();
                n = units;
            }
            max_units[dev] = n;
        }
        if (max_units[dev] < units) units = max_units[dev];
    }
    const int grid = (total_work < units ? total_work : units) * G;
    cfg.gridDim = dim3(grid);
    VB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmC2, tmAux, p));
    VB_CHECK_LAUNCH();
    return VB_OK;
}

// Work-stealing tile scheduler: a ring of per-launch tile counters per device. Launch number s uses slot s % R (zero: cleared
// by launch s - R/2, or by the initial memset) and clears slot (s + R/2) % R for a later launch, so there is no per-launch
// memset and no host-side bookkeeping of counter values; correct as long as fewer than R/2 GEMM launches of one device are
// in flight at once. Off by default (VB_GEMM_DYNAMIC=1 or vb_set_gemm_scheduler(1) turns it on).
constexpr int SCHED_RING = 1024;
static unsigned* g_sched_ring[64] = {nullptr};
static unsigned long long g_sched_seq[64] = {0};
static int g_dynamic = -1;
static int dynamic_enabled() {
    if (g_dynamic < 0) {
        const char* e = getenv("VB_GEMM_DYNAMIC");
        g_dynamic = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return g_dynamic;
}
static int sched_slots(unsigned** cur, unsigned** clear, cudaStream_t stream) {
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    *cur = *clear = nullptr;
    if (!dynamic_enabled() || dev >= 64) return VB_OK;
    // A captured launch would bake its counter slot into the graph: every replay would find the slot where the previous
    // replay left it (tiles skipped, wrong results). Launches under stream capture keep the static tile order.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    VB_CHECK_CUDA(cudaStreamIsCapturing(stream, &cap));
    if (cap != cudaStreamCaptureStatusNone) return VB_OK;
    if (g_sched_ring[dev] == nullptr) {
        VB_CHECK_CUDA(cudaMalloc(&g_sched_ring[dev], SCHED_RING * sizeof(unsigned)));
        VB_CHECK_CUDA(cudaMemset(g_sched_ring[dev], 0, SCHED_RING * sizeof(unsigned)));
        VB_CHECK_CUDA(cudaDeviceSynchronize());
    }
    const unsigned long long s = g_sched_seq[dev]++;
    *cur = g_sched_ring[dev] + (s % SCHED_RING);
    *clear = g_sched_ring[dev] + ((s + SCHED_RING / 2) % SCHED_RING);
    return VB_OK;
}

// 0 = single CTAs (128 x 256); 1 = CTA pairs (256 x 256 tiles, cta_group::2) with the ring / staging split chosen per
// epilogue; 2 / 3 = pairs with the 6-stage / 5-stage split forced everywhere (measurement only).
// Default from VB_GEMM_CTA_PAIR (1 if unset).
static int g_cta_pair = -1;
static int cta_pair_enabled() {
    if (g_cta_pair < 0) {
        const char* e = getenv("VB_GEMM_CTA_PAIR");
        g_cta_pair = (e != nullptr && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 1;
    }
    return g_cta_pair;
}

// Tile width of the CTA-pair kernels: 0 = chosen per launch (192 where it saves a wave, see vb_gemm_bf16), 192 / 256 = forced
// wherever a 192-column variant exists (tests, A/B measurements). Default from VB_GEMM_TILE_N (0 if unset).
static int g_tile_n = -1;
static int tile_n_mode() {
    if (g_tile_n < 0) {
        const char* e = getenv("VB_GEMM_TILE_N");
        const int v = e ? atoi(e) : 0;
        g_tile_n = (v == 192 || v == 256) ? v : 0;
    }
    return g_tile_n;
}

}  // namespace vb

extern "C" void vb_set_gemm_cta_pair(int mode) { vb::g_cta_pair = (mode >= 0 && mode <= 3) ? mode : 1; }
extern "C" int vb_get_gemm_cta_pair(void) { return vb::cta_pair_enabled(); }
extern "C" void vb_set_gemm_tile_n(int tile_n) { vb::g_tile_n = (tile_n == 192 || tile_n == 256) ? tile_n : 0; }
extern "C" int vb_get_gemm_tile_n(void) { return vb::tile_n_mode(); }
extern "C" void vb_set_gemm_scheduler(int dynamic) { vb::g_dynamic = dynamic ? 1 : 0; }
extern "C" int vb_get_gemm_scheduler(void) { return vb::dynamic_enabled(); }

// How a GEMM is mapped onto the persistent grid: tile mapping, tile width, split-K. Pure host arithmetic on the shapes (no
// pointer of `a` is dereferenced, no device is needed: without one the grid is sized for 148 SMs), shared by vb_gemm_bf16 and
// vb_gemm_plan so that the choices can be inspected and tested without a GPU.
static void make_plan(const vb_gemm_args* a, vb_gemm_plan_t* plan) {
    using namespace vb;
    const int epi = a->epilogue;
    // a CTA pair per 256 x 256 tile whenever there is more than one 128-row block to pair up
    const int G = (cta_pair_enabled() != 0 && a->m > BM) ? 2 : 1;
    const int m_tiles = (a->m + BM * G - 1) / (BM * G), k_blocks = (a->k + BK - 1) / BK;
    const int units = num_sms() / G;
    // Tile width. The N = 768 GEMMs of a ViT-B block (proj / fc2 forward, the fc1 / qkv dgrads) have 3 tiles of 256 columns
    // per row block: at 64 images per GPU that is 150 tiles = 2.03 waves on 74 pairs, a third of the launch spent on the last
    // 3 tiles. 192-column tiles (4 per row block) cost ~0.78 of a 256-column tile each; they are taken when the waves they
    // need are cheaper by more than 3 % (batch 64: 3 x 0.78 against 3; batch 512: 22 x 0.78 against 16, so 256 stays).
    int bnt = BN;
    const bool has192 = G == 2 && a->a_layout == 0 && a->n % 192 == 0 &&
                        ((a->b_layout == 0 && epi == VB_EPI_BF16_RESID) || (a->b_layout == 1 && epi == VB_EPI_BF16 && a->out_colsum == nullptr));
    if (has192) {
        const int waves256 = (m_tiles * ((a->n + BN - 1) / BN) + units - 1) / units, waves192 = (m_tiles * (a->n / 192) + units - 1) / units;
        if (tile_n_mode() == 192 || (tile_n_mode() == 0 && 0.78 * waves192 < 0.97 * waves256)) bnt = 192;
    }
    const int n_tiles = (a->n + bnt - 1) / bnt;
    int split_k = a->split_k;
    if (split_k <= 0) {
        // auto (VB_EPI_F32_ADD only): the smallest split whose work items fill >= 90 % of the CTAs / pairs of its last
        // wave (fewer splits = fewer fp32 reduce-adds), else the best-filling one; at least 8 k-blocks per split
        split_k = 1;
        if (epi == VB_EPI_F32_ADD) {
            const int tiles = m_tiles * n_tiles;
            const int max_split = k_blocks / 8 < 1 ? 1 : (k_blocks / 8 > 64 ? 64 : k_blocks / 8);
            double best = 0.0;
            for (int s = 1; s <= max_split; ++s) {
                const int work = tiles * s, waves = (work + units - 1) / units;
                const double fill = static_cast<double>(work) / (static_cast<double>(waves) * units);
                if (fill > best + 1e-9) {
                    best = fill;
                    split_k = s;
                }
                if (fill >= 0.9) break;
            }
        }
    }
    const int sk = split_k > k_blocks ? k_blocks : split_k;
    const int kb_per_split = (k_blocks + sk - 1) / sk;
    plan->cta_pair = G == 2 ? 1 : 0;
    plan->tile_m = BM * G;
    plan->tile_n = bnt;
    plan->m_tiles = m_tiles;
    plan->n_tiles = n_tiles;
    plan->split_k = split_k;
    plan->k_blocks_per_split = kb_per_split;
    plan->units = units;
    const int work = m_tiles * n_tiles * ((k_blocks + kb_per_split - 1) / kb_per_split);  // every split is non-empty
    plan->waves = (work + units - 1) / units;
}

extern "C" int vb_gemm_plan(const vb_gemm_args* a, vb_gemm_plan_t* plan) {
    using namespace vb;
    VB_CHECK_ARG(a != nullptr && plan != nullptr, "vb_gemm_plan: null argument");
    VB_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0, "vb_gemm_plan: bad shape m=%d n=%d k=%d", a->m, a->n, a->k);
    VB_CHECK_ARG(a->epilogue >= VB_EPI_BF16 && a->epilogue <= VB_EPI_BF16_ROWDOT, "vb_gemm_plan: bad epilogue %d", a->epilogue);
    make_plan(a, plan);
    return VB_OK;
}

extern "C" int vb_gemm_bf16(const vb_gemm_args* a, vb_stream_t stream_) {
    using namespace vb;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VB_CHECK_ARG(a != nullptr, "vb_gemm_bf16: null args");
    VB_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0, "vb_gemm_bf16: bad shape m=%d n=%d k=%d", a->m, a->n, a->k);
    VB_CHECK_ARG(a->a && a->b, "vb_gemm_bf16: null operand");
    VB_CHECK_ARG(a->a_layout == 0 || a->a_layout == 1, "vb_gemm_bf16: bad a_layout %d", a->a_layout);
    VB_CHECK_ARG(a->b_layout == 0 || a->b_layout == 1, "vb_gemm_bf16: bad b_layout %d", a->b_layout);
    VB_CHECK_ARG(a->epilogue >= VB_EPI_BF16 && a->epilogue <= VB_EPI_BF16_ROWDOT, "vb_gemm_bf16: bad epilogue %d",
                 a->epilogue);
    VB_CHECK_ARG(a->n % 8 == 0, "vb_gemm_bf16: n=%d must be a multiple of 8", a->n);
    const int epi = a->epilogue;
    vb_gemm_plan_t plan;
    make_plan(a, &plan);
    const int G = plan.cta_pair ? 2 : 1, bnt = plan.tile_n, m_tiles = plan.m_tiles, n_tiles = plan.n_tiles;
    const int k_blocks = (a->k + BK - 1) / BK, split_k = plan.split_k;
    VB_CHECK_ARG(split_k == 1 || epi == VB_EPI_F32_ADD, "vb_gemm_bf16: split_k > 1 needs VB_EPI_F32_ADD");
    if (epi == VB_EPI_BF16_ROWDOT)
        VB_CHECK_ARG(a->sumsq && a->rows_per_sample > 0 && a->cols_per_group == 64 && a->n % 64 == 0 && a->n_groups == a->n / 64,
                     "vb_gemm_bf16: ROWDOT needs sumsq, rows_per_sample, cols_per_group == 64, n %% 64 == 0, n_groups == n / 64");
    if (epi == VB_EPI_BF16_RESID || epi == VB_EPI_BF16_DGELU || epi == VB_EPI_BF16_MULAUX || epi == VB_EPI_BF16_ROWDOT)
        VB_CHECK_ARG(a->aux != nullptr && a->ld_aux % 8 == 0 && (reinterpret_cast<uintptr_t>(a->aux) & 15) == 0,
                     "vb_gemm_bf16: aux must be non-null, 16B aligned, ld multiple of 8");
    if (epi == VB_EPI_SUMSQ)
        VB_CHECK_ARG(a->sumsq && a->rows_per_sample > 0 && a->cols_per_group > 0 && a->cols_per_group % 128 == 0 &&
                         a->n_groups > 0,
                     "vb_gemm_bf16: SUMSQ needs sumsq, rows_per_sample, cols_per_group %% 128 == 0, n_groups");
    else
        VB_CHECK_ARG(a->out != nullptr, "vb_gemm_bf16: null out");
    if (epi == VB_EPI_BF16_GELU_GRAD) VB_CHECK_ARG(a->out2 != nullptr, "vb_gemm_bf16: GELU_GRAD epilogue needs out2");
    if (a->bias) VB_CHECK_ARG((reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "vb_gemm_bf16: bias must be 16B aligned");
    if (a->out_colsum)
        VB_CHECK_ARG(epi == VB_EPI_BF16 || epi == VB_EPI_BF16_RESID || epi == VB_EPI_BF16_DGELU || epi == VB_EPI_BF16_MULAUX ||
                         epi == VB_EPI_BF16_ROWDOT,
                     "vb_gemm_bf16: out_colsum needs a single-output bf16 epilogue");

    CUtensorMap tmA, tmB, tmC, tmC2;
    int rc;
    if (a->a_layout == 0)
        rc = make_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->a, a->k, a->m, a->lda * 2, BK, BM,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    else
        rc = make_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->a, a->m, a->k, a->lda * 2, 64, BK,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (a->b_layout == 0)
        rc = make_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->b, a->k, a->n, a->ldb * 2, BK, bnt / G,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    else
        rc = make_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->b, a->n, a->k, a->ldb * 2, 64, BK,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (epi == VB_EPI_SUMSQ) {
        tmC = tmA;  // unused
        tmC2 = tmA;
    } else if (epi == VB_EPI_F32 || epi == VB_EPI_F32_ADD) {
        rc = make_tensor_map_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->out, a->n, a->m, a->ld_out * 4, 32, 32,
                                CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        tmC2 = tmC;
    } else {
        rc = make_tensor_map_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->out, a->n, a->m, a->ld_out * 2, 32, 32,
                                CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        if ((epi == VB_EPI_BF16_GELU || epi == VB_EPI_BF16_GELU_GRAD) && a->out2 != nullptr) {
            rc = make_tensor_map_2d(&tmC2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->out2, a->n, a->m, a->ld_out2 * 2, 32,
                                    32, CU_TENSOR_MAP_SWIZZLE_64B);
            if (rc) return rc;
        } else {
            tmC2 = tmC;
        }
    }

    CUtensorMap tmAux = tmC;  // residual / saved-GELU' operand, read through TMA by the pair kernels of those epilogues
    if (epi == VB_EPI_BF16_RESID || epi == VB_EPI_BF16_MULAUX || epi == VB_EPI_BF16_ROWDOT) {
        rc = make_tensor_map_2d(&tmAux, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a->aux, a->n, a->m, a->ld_aux * 2, 32, 32,
                                CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }

    GemmKernelParams p;
    p.M = a->m;
    p.N = a->n;
    p.K = a->k;
    p.num_m_blocks = m_tiles;
    p.num_n_blocks = n_tiles;
    p.num_k_blocks = k_blocks;
    int sk = split_k > p.num_k_blocks ? p.num_k_blocks : split_k;
    p.kb_per_split = (p.num_k_blocks + sk - 1) / sk;
    p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;  // every split is non-empty
    p.epi = epi;
    p.bias = a->bias;
    p.aux = static_cast<const bf16*>(a->aux);
    p.ld_aux = a->ld_aux;
    p.sumsq = a->sumsq;
    p.rows_per_sample = a->rows_per_sample;
    p.cols_per_group = a->cols_per_group;
    p.n_groups = a->n_groups;
    p.has_out2 = a->out2 != nullptr;
    p.out_colsum = a->out_colsum;
    rc = sched_slots(&p.sched, &p.sched_clear, stream);
    if (rc) return rc;

    // the hot combinations of the training step get an epilogue fixed at compile time, everything else the generic kernel
#define VB_LAUNCH(AL, BL, GG, EP, DP) return launch_gemm<AL, BL, GG, EP, DP>(tmA, tmB, tmC, tmC2, tmAux, p, stream)
#define VB_LAUNCH_PAIR(AL, BL, EP, DP)                     \
    do {                                                   \
        if (deep_mode == 0) VB_LAUNCH(AL, BL, 2, EP, 0);   \
        if (deep_mode == 1) VB_LAUNCH(AL, BL, 2, EP, 1);   \
        VB_LAUNCH(AL, BL, 2, EP, DP);                      \
    } while (0)
    const int al = a->a_layout, bl = a->b_layout;
    const int deep_mode = cta_pair_enabled() == 2 ? 0 : (cta_pair_enabled() == 3 ? 1 : -1);
    if (bnt == 192) {
        if (bl == 0) {
            if (a->k >= 2048) return launch_gemm<0, 0, 2, VB_EPI_BF16_RESID, 0, 192>(tmA, tmB, tmC, tmC2, tmAux, p, stream);
            return launch_gemm<0, 0, 2, VB_EPI_BF16_RESID, 1, 192>(tmA, tmB, tmC, tmC2, tmAux, p, stream);
        }
        return launch_gemm<0, 1, 2, VB_EPI_BF16, 0, 192>(tmA, tmB, tmC, tmC2, tmAux, p, stream);
    }
    if (G == 2) {
        if (al == 0 && bl == 0) {
            if (epi == VB_EPI_BF16) VB_LAUNCH_PAIR(0, 0, VB_EPI_BF16, 0);
            if (epi == VB_EPI_BF16_RESID) {
                // short K (proj): the epilogue paces the tile, take the TMA-fed residual (5-stage ring); long K (fc2): the
                // main loop does, keep the 6-stage ring
                if (a->k >= 2048) VB_LAUNCH_PAIR(0, 0, VB_EPI_BF16_RESID, 0);
                VB_LAUNCH_PAIR(0, 0, VB_EPI_BF16_RESID, 1);
            }
            if (epi == VB_EPI_BF16_GELU_GRAD) VB_LAUNCH_PAIR(0, 0, VB_EPI_BF16_GELU_GRAD, 1);
            VB_LAUNCH(0, 0, 2, -1, 0);
        }
        if (al == 0 && bl == 1) {
            if (epi == VB_EPI_BF16) VB_LAUNCH_PAIR(0, 1, VB_EPI_BF16, 0);
            if (epi == VB_EPI_BF16_MULAUX) VB_LAUNCH_PAIR(0, 1, VB_EPI_BF16_MULAUX, 1);
            if (epi == VB_EPI_BF16_ROWDOT) VB_LAUNCH_PAIR(0, 1, VB_EPI_BF16_ROWDOT, 1);
            VB_LAUNCH(0, 1, 2, -1, 0);
        }
        if (al == 1 && bl == 1) {
            if (epi == VB_EPI_F32_ADD) VB_LAUNCH_PAIR(1, 1, VB_EPI_F32_ADD, 0);
            VB_LAUNCH(1, 1, 2, -1, 0);
        }
        VB_LAUNCH(1, 0, 2, -1, 0);
    }
    if (al == 0 && bl == 0) VB_LAUNCH(0, 0, 1, -1, 0);
    if (al == 0 && bl == 1) VB_LAUNCH(0, 1, 1, -1, 0);
    if (al == 1 && bl == 1) VB_LAUNCH(1, 1, 1, -1, 0);
    VB_LAUNCH(1, 0, 1, -1, 0);
#undef VB_LAUNCH_PAIR
#undef VB_LAUNCH
}
