/*
 * ccdm_b200.h -- C ABI of libccdm_b200.so, the sm_100a implementation of the CCDM
 * reverse-process hot path.
 *
 * The reference (LarsDoorenbos/ccdm-stochastic-segmentation) is pure
 * Python/PyTorch and has no FFI of its own; its boundary is the Python object
 * protocol `DenoisingModel.forward` (ddpm/models/diffusion_denoising.py:144).
 * Each entry point below names the reference lines it replaces; INTEGRATION.md
 * shows the ctypes binding (ccdm_b200/_lib.py) that sits between them.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer except `ops`, `err` strings
 *    and handles is a DEVICE pointer owned by the caller.
 *  - no allocation of caller-visible memory, no host synchronisation, every
 *    launch goes to the `stream` argument (a cudaStream_t passed as void*).
 *  - return 0 on success, negative on error; ccdm_last_error() describes it.
 *  - fp32 activations (`CCDM_DT_F32`, the FFMA kernels) are NHWC ("pixel-major,
 *    channel-minor"); bf16 activations (`CCDM_DT_BF16`, the tensor-core kernels)
 *    are PLANE-MAJOR [B][C/8][H][W][8]: one 16-byte row of a tcgen05 "K-major,
 *    no swizzle" core matrix per (8-channel plane, pixel), so a TMA box of a
 *    tile lands in shared memory already in the MMA operand layout and a warp
 *    of epilogue threads writes 512 contiguous bytes per plane.  fp32 logits
 *    [B,H,W,K] are NHWC in both modes.
 *  - label maps are uint8 [B,H,W]; GroupNorm statistics travel
 *    beside every activation as per-(sample,channel) double2 {sum, sum of
 *    squares} written by the producing kernel.
 *  - one process per GPU; a plan is not re-entrant.
 */
#ifndef CCDM_B200_H
#define CCDM_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCDM_ABI_VERSION 4

/* storage types of activations / packed conv weights */
#define CCDM_DT_F32 0
#define CCDM_DT_BF16 1
/* "fp16x2": every value v is stored as TWO fp16 numbers hi = fp16(16 v), lo = fp16(16 v - hi) (22 significand bits,
 * error <= max(2^-23 |v|, 2^-29)), plane-major like bf16 with the hi and lo planes of an 8-channel group adjacent:
 * [B][C/8][2][H][W][8].  The tensor-core kernels multiply such operands with three fp16 MMAs (hi*hi + hi*lo + lo*hi,
 * fp32 accumulation in TMEM): fp32-grade products at tensor-core speed -- the "exact" tensor-core mode. */
#define CCDM_DT_F16X2 2
#define CCDM_F16X2_SCALE_LOG2 4 /* stored = 2^4 * value */

/* last-step / draw modes of the categorical head (diffusion_denoising.py:206-212) */
#define CCDM_DRAW_SAMPLE 0     /* t > 1: x_{t-1} ~ Cat(p)  == argmax p/E            */
#define CCDM_DRAW_MAJORITY 1   /* t == 1, step_T_sample None|"majority": argmax p   */
#define CCDM_DRAW_CONFIDENCE 2 /* t == 1, "confidence": normalised probabilities    */
#define CCDM_DRAW_X0 3         /* no posterior: emit softmax(logits) (forward_step) */
#define CCDM_DRAW_POSTERIOR 4  /* standalone API only: emit the raw (unclamped) theta_post_prob (:128) */

/* noise sources of the draw */
#define CCDM_NOISE_TENSOR 0 /* explicit E ~ Exp(1), fp32 [B*H*W, K] (torch generator parity) */
#define CCDM_NOISE_PHILOX 1 /* in-kernel Philox4x32-10 keyed by (seed, global sample, draw, pixel) */

/* One row per chain step, resident on the device; kernels index it with the
 * device-side step counter so one captured CUDA graph replays for every t. */
typedef struct ccdm_step_entry {
    float t;            /* timestep fed to the UNet (diffusion_denoising.py:191-194) */
    float alpha_t;      /* alphas[t-1], 0 at t==1   (:110-113)                        */
    float cumalpha_tm1; /* cumalphas[t-2], 1 at t==1 (:111-113)                        */
    int32_t mode;       /* CCDM_DRAW_*                                                */
    uint32_t draw;      /* Philox draw index (chain step ordinal)                     */
    int32_t emb_row;    /* row of the timestep-embedding table this step uses         */
    int32_t pad0, pad1;
} ccdm_step_entry;

/* ---- program ops: one fused kernel launch each ---------------------------- */
#define CCDM_OP_INPUT_CONV 1 /* unet.py:760 + input_blocks[0] (:516-518)                 */
#define CCDM_OP_CONV 2       /* GN+SiLU+conv3x3/1x1 (+emb +bias +residual/1x1 skip)     */
#define CCDM_OP_ATTENTION 3  /* QKVAttentionLegacy (:343-360)                           */
#define CCDM_OP_HEAD 4       /* softmax + theta_post_prob + clamp + draw                */
#define CCDM_OP_ENCODE_INPUT 5 /* bf16 mode: one-hot(labels) ++ image (unet.py:760) as a plane-major bf16 tensor of
                                  Cout = ceil16(K + C_img) channels (zero padded), the input of input_blocks[0] run as
                                  an ordinary tensor-core conv */

typedef struct ccdm_op {
    int32_t kind; /* CCDM_OP_* */
    int32_t dtype; /* CCDM_DT_* of activations in/out */
    int32_t B, Hin, Win, Hout, Wout;
    int32_t C0, C1;     /* channels of the two concatenated sources (C1 may be 0) */
    int32_t Cout;
    int32_t ksize;      /* 1 or 3 */
    int32_t stride;     /* 1 or 2 (Downsample, unet.py:136-139) */
    int32_t upsample;   /* 1: nearest x2 in front of the conv (Upsample, :106-116) */
    int32_t gn;         /* 1: GroupNorm(32 groups, eps 1e-5) on the input (nn.py:93-100) */
    int32_t silu;       /* 1: SiLU after the norm */
    int32_t S0, S1;     /* channels of the 1x1-skip sources (ResBlock.skip_connection, :221-228) */
    int32_t heads, head_dim; /* attention */
    int32_t K;          /* classes (input conv, head) */
    int32_t C_img;      /* image channels (input conv) */
    int32_t emb_off;    /* column of this block in the embedding table, -1: none */
    int32_t emb_cols;   /* row length of the embedding table */
    int32_t emb_bstride;/* rows between consecutive samples (0: same t for the batch) */
    int32_t noise_mode; /* CCDM_NOISE_* (head) */
    int32_t sample0;    /* global index of local sample 0 (Philox counter; sharding) */
    int32_t out_dtype;  /* CCDM_DT_* of `out` (logits stay fp32 in bf16 mode) */
    int32_t src_kind;   /* 0: src0/src1 NHWC activations; 1: one-hot(labels_in) ++ image (unet.py:760) */
    int32_t exact;      /* 1: fp32 FFMA kernels (parity mode); 0: tensor-core kernels where available; HEAD: 0 lets sampling
                         * steps use approximate exp2/log2/reciprocal (same Philox bits, no IEEE divisions) */
    int32_t acc_shift;  /* fp16x2 convs: packed weights are scaled by 2^(acc_shift - CCDM_F16X2_SCALE_LOG2); the epilogue
                         * multiplies the accumulator by 2^-acc_shift (powers of two: exact) */
    int32_t img_rep;    /* input conv / encode_input: consecutive samples that share ONE conditioning image (`image` then has
                         * B / img_rep entries and sample b reads entry b / img_rep); 0 or 1: one image per sample.  Replaces
                         * the evaluators' image.repeat_interleave(N) (evaluate_lidc_uncertainty.py:96) */
    /* Layout of stat0 / stat1.  st_slots[i] == 0: double2 [B, C] {sum, sum of squares}, folded by the producer.
     * st_slots[i] > 0 ("deferred fold", tensor-core producers): fp32 per-CTA partial rows [B][st_slots][st_rows][2] exactly
     * as the producer's epilogue wrote them (ccdm_conv_stat_layout); the consumer folds rows
     * [0, CTAs that touched sample b) in order, in double -- same sums, no ticket / fence / atomic in the producer. */
    int32_t st_slots[2], st_ips[2], st_items[2], st_grid[2], st_rows[2];
    int32_t tile_batch; /* tensor-core convs: batch size the TILE SELECTION assumes (0: this op's own B).  The tile height is chosen
                         * from the number of work items, i.e. from the batch; a fixed value makes the tiling -- and with it the
                         * fp32 per-item partial sums of the GroupNorm statistics -- independent of how a batch is split over
                         * calls or GPUs (bit-identical results in the fp16x2 mode), at the price of fewer, larger items when the
                         * real batch is small */
    int32_t gn_cpg;     /* GroupNorm: channels per group of the input (0: (C0 + C1) / 32, nn.py:93-100).  Set when the op sees only PART
                         * of a normalised concatenation: the per-step half of input_blocks[10] after the constant DINO channels were
                         * folded into a per-chain map (unet.py:770-788) */
    int32_t gn_off;     /* ... and the number of channels of that concatenation in front of this op's channel 0: channel c belongs to
                         * group (c + gn_off) / gn_cpg; groups cut by the op's channel range are only used with zero weights */
    uint64_t seed;      /* Philox key */
    /* device pointers (0 = absent) */
    uint64_t src0, src1;       /* inputs NHWC                                             */
    uint64_t stat0, stat1;     /* double2 [B,C] stats of src0/src1 (gn=1)                 */
    uint64_t gamma, beta;      /* fp32 [C0+C1]                                            */
    uint64_t weight;           /* packed [tap][CinP][CoutP], CinP=ceil8(Cin), CoutP=ceil32(Cout), zero padded */
    uint64_t bias;             /* fp32 [CoutP] (conv bias, + skip bias folded in)         */
    uint64_t emb;              /* fp32 table [rows, emb_cols]                             */
    uint64_t skip0, skip1;     /* raw NHWC inputs of the fused 1x1 skip conv              */
    uint64_t skip_w;           /* packed [S0+S1][CoutP]                                   */
    uint64_t res;              /* identity residual NHWC [B,Hout,Wout,Cout]               */
    uint64_t out;              /* NHWC output                                             */
    uint64_t ostat;            /* double2 [B,Cout] stats of `out` (0: not needed, or deferred fold: only `part` is written) */
    uint64_t part;             /* fp32 per-CTA / per-tile partial stats (scratch, or the tensor's statistics in deferred mode) */
    uint64_t ticket;           /* uint32 [B] arrival counters (self-resetting)            */
    uint64_t labels_in;        /* uint8 [B,H,W]   (input conv, head)                      */
    uint64_t labels_out;       /* uint8 [B,H,W]   (head)                                  */
    uint64_t image;            /* fp32 NCHW [B,C_img,H,W] (input conv)                    */
    uint64_t noise;            /* fp32 [B*H*W,K]  (head, CCDM_NOISE_TENSOR)               */
    uint64_t probs_out;        /* fp32 [B,H,W,K]  (head: confidence / x0 output, or 0)    */
    uint64_t noise_out;        /* fp32 [B*H*W,K]  (head: export of the E used, or 0)      */
    uint64_t steps;            /* const ccdm_step_entry* table                            */
    uint64_t step_ptr;         /* int32* device step counter                              */
} ccdm_op;

/* ---- library ---------------------------------------------------------------- */
int ccdm_abi_version(void);
const char *ccdm_last_error(void);
/* struct sizes, so a foreign-language binding can verify its mirror of the structs */
size_t ccdm_sizeof_op(void);
size_t ccdm_sizeof_step_entry(void);
/* floats of `part` scratch a conv op with statistics needs */
size_t ccdm_conv_part_floats(int B, int Hout, int Wout, int Cout);
/* the same, for the kernel `op` will actually be dispatched to */
size_t ccdm_op_part_floats(const ccdm_op *op);
/* 1 if `op` (exact == 0, dtype bf16 or fp16x2) runs on the tcgen05 kernel (conv_tma.cu); its `weight` / `skip_w`
 * must then be packed [cc][Cin/8][tap][NT][8] bf16 (cc = chunk of NT = ccdm_conv_tc_nt(Cout, taps) output
 * channels of ceil16(Cout)) instead of fp32 [tap][CinP][CoutP]: every (chunk, 8-channel plane, tap)
 * is NT rows of 16 bytes, the UMMA K-major no-swizzle canonical form, and a K chunk of planes is one
 * contiguous block (one cp.async.bulk).  fp16x2: [cc][Cin/8][tap][2][NT][8] fp16 -- the NT hi rows of a
 * (plane, tap), then its NT lo rows -- of the weights scaled by 2^(acc_shift - 4). */
int ccdm_conv_uses_tc(const ccdm_op *op);
/* same (kept for tools; there is one tensor-core conv kernel) */
int ccdm_conv_uses_tma(const ccdm_op *op);
int ccdm_conv_tc_nt(int Cout, int taps, int x3); /* taps = 1 (1x1), 9 (3x3) or 16 (sub-pixel taps of an upsampling conv); x3 = 1 for fp16x2 */
/* Tile / pipeline configuration the tcgen05 kernel would use for `op` (introspection for DESIGN.md, the
 * bench and tests): out16 = {PL, R, Wt, MB, WN, NT, n_cc, NS, resident, acc2, tmem_cols, tiles, n_items,
 * grid, smem bytes, K chunks per item}.  Returns -1 if `op` does not run on that kernel. */
int ccdm_conv_tc_config(const ccdm_op *op, int32_t *out16);
/* Deferred-fold layout of the statistics a tensor-core conv `op` writes to `part`: out5 = {slots, items per sample,
 * items, grid, row length}; the consumer's st_* fields.  Returns -1 if `op` does not run on a tensor-core kernel. */
int ccdm_conv_stat_layout(const ccdm_op *op, int32_t *out5);
/* 0 if the current device is compute capability 10.x, negative otherwise. */
int ccdm_check_device(void);

/* ---- single kernels (unit-parity surface) ---------------------------------- */

/* Launch one op.  Replaces, depending on op->kind: unet.py:760+516-518 (input
 * conv), :242-262 / :106-116 / :144-146 / :305-311 (fused conv variants),
 * :343-360 (attention), :701-707 + diffusion_denoising.py:197-212 (head). */
int ccdm_launch_op(const ccdm_op *op, void *stream);

/* Timestep-embedding table.  Replaces nn.py:103-121 + unet.py:506-510,758 and
 * every ResBlock's emb_layers (:205-211,251): out[r, :] = W_all * silu(time_embed(
 * sinusoid(t[r]))) + b_all, with W_all [cols, 4*mc] the row-concatenation of all
 * emb_layers.1.weight.  t: fp32 [rows]. */
int ccdm_time_table(const float *t, int rows, int model_channels, const float *te0_w, const float *te0_b,
                    const float *te2_w, const float *te2_b, const float *w_all, const float *b_all, int cols,
                    float *out, void *stream);

/* One-hot (or any score map) -> labels: argmax over K of a strided fp32
 * [B,K,H,W] tensor (strides in elements).  Replaces the `.argmax(dim=1)` the
 * callers apply to x_T (the chain keeps x_t as uint8 labels). */
int ccdm_onehot_to_labels(const float *x, int64_t sb, int64_t sk, int64_t sh, int64_t sw, int B, int K, int H,
                          int W, uint8_t *labels, void *stream);

/* labels -> one-hot int64 [B,H,W,K] (max_prob_sample, one_hot_categorical.py:46-50). */
int ccdm_labels_to_onehot_i64(const uint8_t *labels, size_t n_pix, int K, int64_t *out, void *stream);

/* NCHW fp32 -> NHWC (fp32|bf16|fp16x2) + per-(sample,channel) stats; used once per chain
 * for the feature condition (unet.py:784-786).  rep > 1: `src` holds B / rep entries and sample b converts entry
 * b / rep (N samples per image without a repeat_interleave copy). */
int ccdm_nchw_to_nhwc_stats(const float *src, int B, int C, int H, int W, int dtype, void *dst, double *stat, int rep,
                            void *stream);

/* Vote over the N samples of each image (evaluate_lidc_uncertainty.py:103-125 `prediction.reshape(B, N, ...)`,
 * eval_cdm.py:176-193 `predict_multiple`'s running mean): labels uint8 [B_img * N, n_pix] (image-major, sample-minor) ->
 * freq fp32 [B_img, K, n_pix] = fraction of the N samples that chose class k (the mean of the one-hot maps; NCHW like the
 * reference's result) and, if `majority` is not NULL, uint8 [B_img, n_pix] = argmax_k freq (first maximum). */
int ccdm_vote(const uint8_t *labels, int B_img, int N, size_t n_pix, int K, float *freq, uint8_t *majority, void *stream);

/* Standalone posterior + draw on probabilities theta fp32 [n_pix_total, K]
 * (diffusion_denoising.py:99-128,204-212; one_hot_categorical.py:30-54).
 * labels_in/labels_out uint8; noise fp32 [n, K] or NULL with philox; probs_out
 * (normalised clamped posterior) and noise_out optional. */
int ccdm_posterior_draw(const float *theta, const uint8_t *labels_in, size_t n_pix_per_sample, int B, int K,
                        float alpha_t, float cumalpha_tm1, int mode, int noise_mode, const float *noise,
                        uint64_t seed, uint32_t draw, uint32_t sample0, uint8_t *labels_out, float *probs_out,
                        float *noise_out, void *stream);

/* Raw Philox words, layout [n_samples, n_pix, K] (tests the counter layout). */
int ccdm_philox_bits(uint64_t seed, uint32_t draw, uint32_t sample0, uint32_t n_samples, uint32_t n_pix, int K,
                     uint32_t *bits, void *stream);

/* x_T: uniform labels from Philox (draw index `draw`), the device-side
 * equivalent of OneHotCategoricalBCHW(logits=zeros).sample(). */
int ccdm_uniform_labels(uint64_t seed, uint32_t draw, uint32_t sample0, uint32_t n_samples, uint32_t n_pix,
                        int K, uint8_t *labels, void *stream);

/* Pairwise segmentation distance between two sets of label maps of the same images -- the kernel under the LIDC metrics
 * (SURVEY.md 8f-4): ddpm/utils.py:129-142 `iou` + `batched_distance` (also evaluate_lidc_uncertainty.py:27-41), which
 * `calc_batched_generalised_energy_distance` (:145-158) and `batched_hungarian_matching` (:161-174) are built on.
 * x uint8 [B, N, n_pix], y uint8 [B, M, n_pix] -> dist double [B, N, M]:
 *   dist[b,n,m] = 1 - mean over classes c = 1..K-1 (background excluded) of |x==c & y==c| / |x==c | y==c|, 0/0 := 1.
 * The reference builds [B,N,M,n_pix,K] boolean broadcasts on the host; here one CTA per (b, n, m) counts in integers and
 * does the final divisions in double, so the result equals numpy's to the last bits of the class mean. */
int ccdm_pairwise_distance(const uint8_t *x, const uint8_t *y, int B, int N, int M, size_t n_pix, int K, double *dist, void *stream);

/* ---- DINO ViT condition encoder ------------------------------- */
/* The kernels under ccdm_b200.models.condition_encoder.DinoViT: ddpm/models/condition_encoder.py:26-46 (DinoViT.forward),
 * ddpm/models/dino.py:211-229,279-309 (ViTExtractor._extract_features / extract_descriptors, 'key' facet of one block) and the
 * torch.hub ViT they run (facebookresearch/dino vision_transformer.py).  Tokens are fp16x2 plane-major tensors
 * [B][C/8][2][T][8] (CCDM_DT_F16X2; T = 1 + patches, token 0 = cls); the attention of a block is CCDM_OP_ATTENTION on the
 * [B][3C/8][2][T][8] qkv tensor (Hin = 1, Win = T, head_dim 64) with channels ordered h*3d + {q,k,v}*d + i. */

/* Output-channel tile of ccdm_vit_linear: w_packed is [Cout/NT][Cin/8][2][NT][8] fp16 = per (tile, 8-channel group) the NT hi
 * rows, then the NT lo rows of 2^(acc_shift - 4) * W (nn.Linear weight [Cout, Cin]). */
int ccdm_vit_linear_nt(void);
/* out = x W^T + bias, optionally GELU (erf form) and / or + residual: Attention.qkv / .proj, Mlp.fc1 / .fc2 of a ViT block
 * (vision_transformer.py Block.forward), and the 'key' rows of blocks[layer].attn.qkv (dino.py:172-176).  tcgen05: per K step
 * A_hi x [W_hi; W_lo] and A_lo x W_hi, fp32 accumulation in TMEM.  Cin % 32 == 0, Cout % NT == 0; out must not alias x / residual. */
int ccdm_vit_linear(const void *x, const void *w_packed, const float *bias, const void *residual, int B, int T, int Cin, int Cout,
                    int gelu, int acc_shift, void *out, void *stream);
/* nn.LayerNorm(C, eps) over the channels of every token (Block.norm1 / norm2). */
int ccdm_vit_layernorm(const void *x, const float *gamma, const float *beta, int B, int T, int C, float eps, void *out, void *stream);
/* PatchEmbed (Conv2d(3, D, patch, stride)) + cls token + position embedding (VisionTransformer.prepare_tokens): image fp32 NCHW
 * [B,3,H,W]; w_t = proj.weight transposed to [3*patch*patch][D]; pos [1 + hp*wp][D] already resized (ccdm_vit_pos_embed);
 * tokens [B][D/8][2][1 + hp*wp][8], hp = 1 + (H - patch) / stride. */
int ccdm_vit_patch_embed(const float *image, const float *w_t, const float *bias, const float *cls, const float *pos, int B, int H,
                         int W, int patch, int stride, int D, void *tokens, void *stream);
/* interpolate_pos_encoding (dino.py:86-117): bicubic resize (align_corners False, A = -0.75) of the [n_side, n_side] grid of
 * pos [1 + n_side^2][D] to [hp, wp] with the caller's scale factors ((hp + 0.1) / n_side, (wp + 0.1) / n_side); row 0 copied. */
int ccdm_vit_pos_embed(const float *pos, int n_side, int D, int hp, int wp, double scale_h, double scale_w, float *out, void *stream);
/* extract_descriptors (dino.py:291-301): key [B][C/8][2][T][8] (channel h*d + i) -> out fp32 NCHW [B, C, Ho, Wo], channel
 * i*heads + h, cls dropped, bilinear (align_corners False) from the [hp, wp] patch grid (identity when Ho, Wo == hp, wp). */
int ccdm_vit_descriptor(const void *key, int B, int T, int heads, int head_dim, int hp, int wp, int Ho, int Wo, float *out,
                        void *stream);

/* ---- programs: the whole reverse step as one launch sequence / CUDA graph --- */
typedef struct ccdm_plan ccdm_plan;

/* Copies `ops` (host array).  The op sequence is one reverse step:
 * UNet forward (unet.py:744-808) + posterior + draw. */
ccdm_plan *ccdm_plan_create(const ccdm_op *ops, int n_ops);
void ccdm_plan_destroy(ccdm_plan *plan);
int ccdm_plan_num_launches(const ccdm_plan *plan);
/* Rewrites per-run fields (head noise mode / seed / sample0 / export pointers);
 * invalidates a captured graph if they changed. */
int ccdm_plan_set_noise(ccdm_plan *plan, int noise_mode, uint64_t seed, int32_t sample0, const float *noise,
                        float *noise_out);
/* One reverse step: replays the captured graph (captured on first use when
 * use_graph != 0), then advances the device step counter. */
int ccdm_plan_step(ccdm_plan *plan, int use_graph, void *stream);
/* n_steps reverse steps back to back (the T-loop of forward_denoising, :189-212). */
int ccdm_plan_run(ccdm_plan *plan, int n_steps, int use_graph, void *stream);

/* Measurement aid (bench.py): device time of every op of the plan, each replayed `iters` times as a
 * one-node CUDA graph between two events; ms_per_op has ccdm_plan_num_launches() entries (host memory). */
int ccdm_plan_profile(ccdm_plan *plan, int iters, float *ms_per_op, void *stream);

/* Debug aid: %globaltimer stamps (ns) CTA 0 of the most recent conv_tma launch took at its milestones
 * {start, setup done, GN affine ready, first TMA landed, first stage transformed, first item's MMAs committed,
 * first item's epilogue done, statistics flushed, end}; returns the number of slots written. */
int ccdm_debug_conv_trace(unsigned long long *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* CCDM_B200_H */
