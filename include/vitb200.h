/*
 * vitb200.h — C ABI of libvitb200.so, the sm_100a (B200) kernel library under the ViT hot path of
 * ambroiseodt/vit-plasticity (ViT Block forward/backward + per-component plasticity estimator).
 *
 * The reference has no FFI of its own: every arithmetic op on the path is a PyTorch ATen call made from
 * `src/vitef/models/transformer/architecture.py` / `transformer/utils.py`. Each entry point below names the
 * reference call site(s) it replaces. The reference-side binding (a ctypes stub inside a
 * torch.autograd.Function) is shown in INTEGRATION.md and implemented in vit_plasticity_b200/_lib.py.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - "bf16" buffers are raw 16-bit bfloat16; row-major; leading dimensions (ld*) are in ELEMENTS;
 *   - every call is asynchronous on `stream` (a cudaStream_t), allocates nothing that outlives the call,
 *     and returns 0 on success / non-zero on error (message: vb_last_error(), thread-local);
 *   - nothing here falls back to a CPU or library (cuBLAS/cuDNN) path.
 */
#ifndef VITB200_H
#define VITB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vb_stream_t; /* cudaStream_t */

#define VB_OK 0
#define VB_ERR_INVALID 1
#define VB_ERR_CUDA 2
#define VB_ERR_UNSUPPORTED 3

int vb_version(void);
const char* vb_last_error(void);
/* number of kernel launches issued by this library in this process (bench.py's `gpu_launches`) */
int64_t vb_launch_count(void);
void vb_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05/TMEM fed by TMA:  C[M,N] = epilogue( sum_k A[m,k] * B[n,k] ), bf16 in, fp32 accumulate.
 * Replaces nn.Linear / nn.Conv2d(k=s=P) forward (architecture.py:205,236,295,297; transformer/utils.py:91)
 * and their autograd dgrad / wgrad (`loss.backward()`, apps/vit/train.py:270).
 *
 *   a_layout / b_layout: 0 = "K-major": A stored [M,K] (B stored [N,K]), contraction index contiguous;
 *                        1 = "MN-major": A stored [K,M] (B stored [K,N]), contraction index strided.
 *   forward  y = x W^T        : A = x  [M,K]  layout 0,  B = W  [N,K]        layout 0
 *   dgrad    dx = dy W        : A = dy [M,N'] layout 0,  B = W  [N',K'] seen as [K=N', N=K'] layout 1
 *   wgrad    dW = dy^T x      : A = dy [tokens,N'] seen as [K=tokens, M=N'] layout 1,
 *                               B = x  [tokens,K'] seen as [K=tokens, N=K'] layout 1
 * ------------------------------------------------------------------------------------------------ */
enum vb_epilogue {
    VB_EPI_BF16 = 0,       /* out = acc (+ bias)                                   out: bf16 [M,N]          */
    VB_EPI_BF16_RESID = 1, /* out = acc (+ bias) + aux                             aux: bf16 [M,N]          */
    VB_EPI_BF16_GELU = 2,  /* out2 = z = acc (+ bias); out = gelu_erf(z)           out, out2: bf16 [M,N]    */
    VB_EPI_BF16_DGELU = 3, /* out = acc * gelu_erf'(aux)                           aux = z: bf16 [M,N]      */
    VB_EPI_F32 = 4,        /* out = acc (+ bias)                                   out: f32 [M,N]           */
    VB_EPI_F32_ADD = 5,    /* out += acc   (TMA reduce-add; split_k >= 1)          out: f32 [M,N]           */
    VB_EPI_SUMSQ = 6,      /* sumsq[row / rows_per_sample, col / cols_per_group] += acc^2 ; nothing stored  */
    VB_EPI_BF16_GELU_GRAD = 7, /* z = acc (+ bias); out = gelu_erf(z); out2 = gelu_erf'(z)  (training forward of fc1:
                                  the derivative is saved instead of z, so the backward epilogue is one multiply) */
    VB_EPI_BF16_MULAUX = 8, /* out = acc * aux                                      aux: bf16 [M,N] (= gelu'(z))     */
    VB_EPI_BF16_ROWDOT = 9  /* out = acc; sumsq[(row / rows_per_sample) * n_groups + col / 64][row % rows_per_sample] =
                               sum over the 64-column group of bf16(out) * aux  (cols_per_group must be 64). With
                               out = dO (proj dgrad) and aux = O this is the attention backward's delta[b, h, q], taken
                               from the epilogue registers instead of a pass over dO and O. Plain stores: every (row, group)
                               is written exactly once.                                                      */
};

typedef struct vb_gemm_args {
    const void* a;      /* bf16 */
    const void* b;      /* bf16 */
    int64_t lda;        /* row stride of the stored A matrix, elements */
    int64_t ldb;
    int32_t a_layout;   /* 0 K-major, 1 MN-major */
    int32_t b_layout;
    int32_t m, n, k;
    int32_t epilogue;   /* enum vb_epilogue */
    const float* bias;  /* f32 [N] or NULL */
    const void* aux;    /* bf16 [M,N] (residual or pre-activation) or NULL */
    int64_t ld_aux;
    void* out;          /* bf16 or f32 [M,N] (unused for SUMSQ) */
    int64_t ld_out;
    void* out2;         /* bf16 [M,N], GELU (may be NULL: z not stored) and GELU_GRAD only */
    int64_t ld_out2;
    float* sumsq;       /* f32 [n_samples, n_groups], SUMSQ (accumulated into; caller zeroes); ROWDOT: f32 [n_samples, n_groups,
                           rows_per_sample], overwritten */
    int32_t rows_per_sample;
    int32_t cols_per_group; /* multiple of 128 */
    int32_t n_groups;
    int32_t split_k;    /* >= 1 (> 1 only with VB_EPI_F32_ADD); <= 0: chosen by the library to fill the SMs */
    float* out_colsum;  /* f32 [N] or NULL (single-output bf16 epilogues): out_colsum[n] += sum_m out[m,n], reduced from the
                           epilogue registers. With out = dz (fc2 dgrad x gelu') this is fc1's bias gradient, without the
                           extra pass over dz. Caller zeroes (or accumulates). */
} vb_gemm_args;

int vb_gemm_bf16(const vb_gemm_args* args, vb_stream_t stream);
/* How vb_gemm_bf16 would map `args` onto its persistent grid (host arithmetic on shapes / layouts / epilogue only: no pointer
 * of `args` is read, nothing is launched, no device is needed — without one the grid is sized for 148 SMs). */
typedef struct vb_gemm_plan_t {
    int32_t cta_pair;           /* 1: a CTA pair per tile (tcgen05 cta_group::2), 0: one CTA per tile */
    int32_t tile_m, tile_n;     /* rows x columns of one output tile: 256 or 128 x 256 or 192 */
    int32_t m_tiles, n_tiles;   /* tiles along M and N */
    int32_t split_k;            /* K splits (1 unless VB_EPI_F32_ADD) */
    int32_t k_blocks_per_split; /* 64-deep k-blocks per split */
    int32_t units;              /* CTAs or CTA pairs the grid is sized for */
    int32_t waves;              /* ceil(m_tiles * n_tiles * splits / units): what the launch pays for */
} vb_gemm_plan_t;
int vb_gemm_plan(const vb_gemm_args* args, vb_gemm_plan_t* plan);
/* Tile mapping of vb_gemm_bf16: 1 (default; env VB_GEMM_CTA_PAIR=0..3 overrides) = a CTA pair per 256 x 256 tile with
 * tcgen05.mma.cta_group::2 (each SM stages half of the B tile), 0 = one CTA per 128 x 256 tile; 2 / 3 = pairs with the
 * 6-stage / 5-stage shared-memory split forced for every epilogue. Results are bit-identical per output element for
 * split_k == 1 (same k order); the setting exists for measurement and bisection. */
void vb_set_gemm_cta_pair(int mode);
int vb_get_gemm_cta_pair(void);
/* Tile order of vb_gemm_bf16: 0 (default; env VB_GEMM_DYNAMIC=1 overrides) = every CTA pair strides statically over the
 * tiles, 1 = work stealing: tile indices are drawn from a per-launch device counter, so CTAs that become resident late
 * (SMs held by a concurrent kernel such as an overlapped all-reduce) do not delay the launch. Same results either way.
 * Launches issued under CUDA stream capture always use the static order (a replayed graph would reuse the counter). */
void vb_set_gemm_scheduler(int dynamic);
int vb_get_gemm_scheduler(void);
/* Tile width of the CTA-pair kernels: 0 (default; env VB_GEMM_TILE_N=192|256 overrides) = chosen per launch: 192-column
 * tiles where they save a wave of the persistent grid (the N = 768 GEMMs of a ViT-B block at <= 128 images per GPU: proj /
 * fc2 forward with the residual epilogue, the fc1 / qkv dgrads), else 256; 192 / 256 = forced wherever the 192-column
 * variant exists. Bit-identical results per output element (same k order). */
void vb_set_gemm_tile_n(int tile_n);
int vb_get_gemm_tile_n(void);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (nn.LayerNorm, transformer/utils.py:293; used at architecture.py:347,349 and utils.py:396)
 * one warp per row, 128-bit accesses, fp32 statistics. x, y: bf16 [rows, cols]; gamma/beta f32 [cols].
 * ------------------------------------------------------------------------------------------------ */
int vb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                     int32_t rows, int32_t cols, float eps, vb_stream_t stream);
/* dx = (dres ? dres : 0) + LN'(dy); dgamma/dbeta accumulated (+=) into f32 [cols] when non-NULL.
 * `dres_colsum` (f32 [cols] or NULL; needs dres): += column sums of dres, taken while dres passes through the kernel. In a
 * pre-norm block dres of the second norm's backward is the block's incoming gradient and dres of the first norm's backward
 * is the gradient of the attention branch's output, so these are the bias gradients of the two Linear layers that write
 * the residual stream (architecture.py:236,297) without a pass of their own.
 * vb_layernorm_bwd_workspace_bytes is kept for callers of the first version of this interface and returns 0. */
int64_t vb_layernorm_bwd_workspace_bytes(int32_t cols);
int vb_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                     const void* dres, void* dx, float* dgamma, float* dbeta, float* dres_colsum, int32_t rows,
                     int32_t cols, vb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention core (architecture.py:212-233): softmax(q k^T / sqrt(d)) v, non-causal, no dropout.
 * qkv: bf16 [batch*seq, 3*E] as produced by the fused qkv Linear (column = which*E + head*d + j);
 * out: bf16 [batch*seq, E] (heads merged); lse: f32 [batch, heads, seq] (log-sum-exp of scaled scores).
 * head_dim must be 64.
 * ------------------------------------------------------------------------------------------------ */
int vb_attention_fwd(const void* qkv, void* out, float* lse, int32_t batch, int32_t seq, int32_t heads,
                     int32_t head_dim, vb_stream_t stream);
/* dqkv: bf16 [batch*seq, 3*E]; dout: bf16 [batch*seq, E]; workspace: caller buffer of
 * vb_attention_bwd_workspace_bytes(batch, seq, heads) bytes (holds delta = rowsum(dout * out), f32 [batch, heads, seq]). */
int64_t vb_attention_bwd_workspace_bytes(int32_t batch, int32_t seq, int32_t heads);
int vb_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, void* workspace,
                     int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream);
/* Same, and dbias[3*E] += column sums of dqkv (the qkv Linear's bias gradient; caller zeroes or accumulates), reduced
 * from the accumulators while they are drained instead of by a second pass over dqkv. */
int vb_attention_bwd_bias(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dbias,
                          void* workspace, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream);
/* Same with delta ALREADY in `workspace` (f32 [batch, heads, seq], e.g. from the VB_EPI_BF16_ROWDOT epilogue of the GEMM
 * that produced dout): `out` is not read, no delta pass is launched. seq <= 208 only; dbias may be NULL. */
int vb_attention_bwd_with_delta(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias,
                                const void* workspace, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                                vb_stream_t stream);
/* Same, but only the QUERY third of the qkv bias gradient is reduced here: dbias_q[0:E] += column sums of dQ (dbias_q is the
 * start of the [3*E] bias gradient; its other two thirds are not touched). In exact arithmetic the key third is zero (a
 * per-query shift of the scores does not change the softmax: sum over keys of dS = 0; the reference's fp32 value is
 * round-off) and the value third equals the column sums of dout (softmax rows sum to one), which the GEMM that produced
 * dout delivers from its epilogue (vb_gemm_args.out_colsum). The per-key-tile drain of dV / dK is then stores only. */
int vb_attention_bwd_with_delta_qbias(const void* qkv, const void* dout, const float* lse, void* dqkv, float* dbias_q,
                                      const void* workspace, int32_t batch, int32_t seq, int32_t heads, int32_t head_dim,
                                      vb_stream_t stream);
/* Paired attention for the plasticity estimator: runs the core on qkv_a and qkv_b (same shapes) and writes
 * delta = attn(qkv_a) - attn(qkv_b), subtracted in fp32 before the bf16 down-cast. */
int vb_attention_pair_delta(const void* qkv_a, const void* qkv_b, int64_t ld_qkv, void* delta, int64_t ld_delta,
                            int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream);
/* All layers of the estimator in one launch (architecture.py:877-881 feeds the same embedding to every block):
 * qkv_a / qkv_b hold every layer's projection side by side, [batch*seq, ld_qkv] with feature = layer*3E + which*E +
 * head*d + j; delta: bf16 [layers, batch*seq, E]. seq <= 208. */
int vb_attention_pair_delta_layers(const void* qkv_a, const void* qkv_b, int64_t ld_qkv, void* delta, int32_t layers,
                                   int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream);

/* Perturbation form of the paired attention (apps/vit/analysis.py:216-233 evaluated on a pair (x, x + dx);
 * apps/plots/loss_landscape.py:180-191 style magnitude sweeps): qkv_a = W t_a + b are the projections of the base tokens,
 * dqkv = W (t_b - t_a) those of the token DIFFERENCE (no bias), same layout as above. Writes
 * delta = attn(t_b) - attn(t_a), bf16 [layers, batch*seq, E], with the score difference, the probability difference and
 * the output difference each formed from the small operands directly (never as a difference of two rounded
 * forward passes), so the relative accuracy does not depend on |t_b - t_a| / |t_a|. seq <= 208, head_dim 64. */
int vb_attention_perturb_delta_layers(const void* qkv_a, const void* dqkv, int64_t ld_qkv, void* delta, int32_t layers,
                                      int32_t batch, int32_t seq, int32_t heads, int32_t head_dim, vb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Element-wise / data-movement helpers on the path
 * ------------------------------------------------------------------------------------------------ */
/* f32 -> bf16 over n elements (weights shadow copy; activations at the module boundary) */
int vb_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vb_stream_t stream);
int vb_cast_bf16_to_f32(const void* src, float* dst, int64_t n, vb_stream_t stream);
/* im2col for non-overlapping P x P patches (transformer/utils.py:91,114): img f32 [N,C,H,W] ->
 * patches bf16 [N*(H/P)*(W/P), C*P*P], column = c*P*P + py*P + px. If img2 != NULL, writes img - img2
 * (subtracted in fp32). */
int vb_im2col_patches(const float* img, const float* img2, void* patches, int32_t n, int32_t c, int32_t h,
                      int32_t w, int32_t p, vb_stream_t stream);
/* Device-side input pipeline: the reference's PIL / torchvision preprocessing of uint8 images (data/images/utils.py:337-366:
 * Resize -> CenterCrop -> ToTensor -> Normalize for eval / analysis / probing; RandomResizedCrop -> RandomHorizontalFlip ->
 * ToTensor -> Normalize for training), bit for bit: Pillow's two-pass 8-bit bilinear ImagingResample with the 22-bit fixed-point
 * tap tables the host precomputes, then the 3 x 256 table lut[c][v] = fp32((fp32(v) / 255 - mean[c]) / std[c]).
 *   src        uint8 [n, src_h, src_w, 3] (HWC)
 *   params     int32 [n, 8] = {top, left, crop_h, crop_w, flip, row_table, col_table, 0} per image, or NULL (whole image, table 0)
 *   tab_bounds int32 [n_tables, out, 2] = (first source index, tap count);  tab_coef int32 [n_tables, out, ksize]
 *   max_src_rows_per_strip: the largest number of source rows any strip of output rows needs (strip = patch rows when
 *              out_patches is given, else 16), sizes the shared-memory intermediate
 *   out_f32    f32 [n, 3, out, out] or NULL;  out_patches bf16 [n * (out/patch)^2, 3 * patch * patch] or NULL (same layout
 *              as vb_im2col_patches) */
int vb_preprocess_u8(const uint8_t* src, int32_t n, int32_t src_h, int32_t src_w, const int32_t* params,
                     const int32_t* tab_bounds, const int32_t* tab_coef, int32_t n_tables, int32_t ksize,
                     int32_t max_src_rows_per_strip, const float* lut, int32_t out, float* out_f32, void* out_patches,
                     int32_t patch, vb_stream_t stream);
/* tokens[b, 0, :] = cls + pos[0]; tokens[b, 1+i, :] = patch_out[b*np + i, :] + pos[1+i]
 * (architecture.py:666-675). patch_out bf16 [batch*np, E] -> tokens bf16 [batch*(np+1), E];
 * tokens_f32 (optional, may be NULL) receives the same values before the bf16 rounding. */
int vb_assemble_tokens(const void* patch_out, const float* patch_out_f32, const float* cls, const float* pos,
                       void* tokens, float* tokens_f32, int32_t batch, int32_t np, int32_t e, vb_stream_t stream);
/* backward of the above: dpatch_out (bf16 [batch*np, E]) = dtokens rows 1..np; dcls += sum_b dtokens[b,0];
 * dpos += sum_b dtokens[b]. dcls/dpos may be NULL (frozen embedding). */
int vb_assemble_tokens_bwd(const void* dtokens, void* dpatch_out, float* dcls, float* dpos, int32_t batch,
                           int32_t np, int32_t e, vb_stream_t stream);
/* out[c] += sum_r x[r, c]  (bias gradients). x bf16 [rows, cols] with row stride ldx. */
int vb_colsum_bf16(const void* x, int64_t ldx, float* out, int32_t rows, int32_t cols, vb_stream_t stream);
/* y = a + b (bf16), n elements */
int vb_add_bf16(const void* a, const void* b, void* y, int64_t n, vb_stream_t stream);

/* Probe pooling (apps/vit/linear_probing.py:92-103,109-112): pooled[n,:] = x[n,0,:] (cls_pooling) or the mean over the
 * seq tokens, optionally L2-normalised per row. x: bf16 [n, seq, dim]; pooled: f32 [n, dim]. Keeps the 8 taps per block
 * of get_probes on the device: (n, dim) rows go to the host instead of (n, seq, dim) tensors. */
int vb_pool_tokens(const void* x, float* pooled, int32_t n, int32_t seq, int32_t dim, int32_t cls_pooling, int32_t normalize,
                   vb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Plasticity reductions (apps/vit/analysis.py:68 `distance`; apps/plots/analysis.py:97 ratio)
 * ------------------------------------------------------------------------------------------------ */
/* out[s] += sum over the sample's rows/cols of (a - b)^2, a/b f32 [n_samples*rows_per_sample, cols];
 * b may be NULL (then a is already a difference). */
int vb_rowsumsq_diff_f32(const float* a, const float* b, float* out, int32_t n_samples, int32_t rows_per_sample,
                         int32_t cols, vb_stream_t stream);
/* LayerNorm plasticity for all norms at once: u[s, d] = sum_l (zhat_a[s,l,d] - zhat_b[s,l,d])^2 where zhat is
 * the normalised (no affine) row; the per-norm squared distance is then sum_d gamma[d]^2 u[s,d]
 * (LN_i(a) - LN_i(b) = gamma_i * (zhat_a - zhat_b); beta cancels). a, b: f32 [n_samples*rows, cols]. */
int vb_layernorm_pair_sqdiff(const float* a, const float* b, float* u, int32_t n_samples, int32_t rows_per_sample,
                             int32_t cols, float eps, vb_stream_t stream);

/* Perturbation form of vb_layernorm_pair_sqdiff: the second input is b = a + scale * d (d: f32 token difference);
 * zhat_b - zhat_a is evaluated without cancellation (see layernorm.cu), d = 0 gives exactly 0. */
int vb_layernorm_delta_sqdiff(const float* a, const float* d, float scale, float* u, int32_t n_samples,
                              int32_t rows_per_sample, int32_t cols, float eps, vb_stream_t stream);
/* dst = bf16(scale * src), n bf16 elements (the sweep re-uses the unit-noise projections for every magnitude). */
int vb_scale_bf16(const void* src, void* dst, int64_t n, float scale, vb_stream_t stream);
/* Perturbation directions of the sweep, generated on the device: out f32 [n_images, elems_per_image] standard normals
 * from Philox4x32-10 keyed by `seed` with the image index (first_image + row) in the counter, so an image's noise does not
 * depend on batching or sharding. Replaces the host-side torch.randn of the reference-style sweep
 * (apps/plots/loss_landscape.py:172-191). Integer stream and transform: oracle/philox_oracle.py. */
int vb_philox_normal_f32(float* out, int64_t n_images, int64_t elems_per_image, uint64_t seed, uint64_t first_image,
                         vb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused optimizer step on a flat gradient arena (apps/vit/train.py:277-283: clip_grad_norm_ + optimizer.step with
 * src/vitef/optim.py:76-82 torch.optim.SGD(momentum)). The trainable parameters keep their own fp32 allocations (the
 * reference's state_dict); their gradients and momentum buffers live back to back in two flat f32 arenas.
 * ------------------------------------------------------------------------------------------------ */
/* partials[b] = sum of x[i]^2 over block b's grid-stride share, b < n_partials (x 16-byte aligned). No atomics: the
 * total is bit-reproducible, so data-parallel replicas derive identical clip coefficients. */
int vb_sumsq_partials_f32(const float* x, int64_t n, float* partials, int32_t n_partials, vb_stream_t stream);
/* chunk_table: DEVICE array of n_chunks records {float* param; int64 arena_offset; int32 count; int32 pad; bf16* shadow}
 * (32 bytes), chunk starts 16-byte aligned. With norm = sqrt(sum of the partials) and coef = min(1, max_norm / (norm + 1e-6)):
 *   g = coef * grad + weight_decay * p;  v = first_step ? g : momentum * v + g (skipped if momentum == 0);  p -= lr * v.
 * shadow (optional per chunk): the same elements of the parameter's bf16 copy (the GEMM operand) are rewritten from the
 * updated values in the same pass. *grad_norm_out (optional) = norm, the value train.py logs as grad_norm.
 * hyper_dev (optional, DEVICE f32[4] = {lr, momentum, weight_decay, max_norm}) overrides the by-value hyper-parameters: a
 * step captured in a CUDA graph then follows an LR schedule without re-capture. */
int vb_sgd_momentum_clip_step(const void* chunk_table, int32_t n_chunks, const float* grad_arena, float* momentum_arena,
                              const float* sumsq_partials, int32_t n_partials, float* grad_norm_out, float max_norm, float lr,
                              float momentum, float weight_decay, int32_t first_step, const float* hyper_dev,
                              vb_stream_t stream);

/* AdamW (torch.optim.AdamW, amsgrad off: src/vitef/optim.py:83-88) over the same arenas, clip coefficient as above:
 *   g = coef * grad;  p *= 1 - lr * weight_decay;  m += (1 - beta1)(g - m);  v = beta2 v + (1 - beta2) g^2;
 *   p -= (lr / bias_correction1) * m / (sqrt(v) / sqrt(bias_correction2) + eps),  bias_correction_i = 1 - beta_i^step.
 * hyper_dev (optional, DEVICE f32[5] = {lr, weight_decay, max_norm, lr / bias_correction1, 1 / sqrt(bias_correction2)})
 * overrides the by-value ones (CUDA-graph replay across steps). */
int vb_adamw_clip_step(const void* chunk_table, int32_t n_chunks, const float* grad_arena, float* exp_avg_arena,
                       float* exp_avg_sq_arena, const float* sumsq_partials, int32_t n_partials, float* grad_norm_out,
                       float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                       float bias_correction1, float bias_correction2, const float* hyper_dev, vb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VITB200_H */
