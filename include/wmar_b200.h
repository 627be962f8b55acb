/*
 * wmar_b200.h -- C-ABI of the B200-native watermarked autoregressive image-generation hot path.
 *
 * Drop-in boundary for the hot path of facebookresearch/wmar (reference file:line cited per entry point).
 * Plain pointers and sizes only; no torch / C++ types.  All `d_` pointers are DEVICE pointers on the current CUDA
 * device, all `h_` pointers are HOST pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default
 * stream).  Every function returns 0 on success or a negative wmar_status; the message of the last failure on the
 * calling thread is available from wmar_last_error().  Nothing here throws, exits, or synchronises the device unless
 * stated.  One handle per (process, device); calls on one handle must be serialised by the caller
 * (the reference is single-threaded per GPU as well).
 *
 * Ownership: the caller owns every buffer it passes in.  Handles own only their scratch (KV cache, activations,
 * split-K workspaces) and BORROW weight pointers: weights live in the caller's tensors (the reference patches
 * nn.Module weights after construction -- generate.py:327-332 -- so the host shim re-packs and calls *_create again
 * after any patch).
 */
#ifndef WMAR_B200_H
#define WMAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum wmar_status {
    WMAR_OK = 0,
    WMAR_ERR_INVALID = -1, /* bad argument (shape / enum / NULL)            */
    WMAR_ERR_CUDA = -2,    /* CUDA runtime error, see wmar_last_error()      */
    WMAR_ERR_RANGE = -3,   /* context sum / token id outside the table       */
    WMAR_ERR_NOMEM = -4,
    WMAR_ERR_SHORT = -5    /* detect(): len(codes) <= context size (reference raises ValueError,
                              gentime_watermark.py:287-291) */
};

/* wmar/watermarking/gentime_watermark.py:94-106 */
enum wmar_seed_strategy { WMAR_SEED_FIXED = 0, WMAR_SEED_LINEAR = 1, WMAR_SEED_SPATIAL = 2 };
enum wmar_split_strategy { WMAR_SPLIT_RANDOM = 0, WMAR_SPLIT_RANDOM_STRATIFIED = 1 };

#define WMAR_SALT_KEY 15485863ull /* gentime_watermark.py:121 */

int wmar_version(void);
const char *wmar_last_error(void);
/* number of kernels launched by this library since load (all handles); bench.py reports it as gpu_launches */
uint64_t wmar_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Greenlist engine.  Replaces GentimeWatermark._split_with_seed + _get_greenlist_ids_for_context
 * (gentime_watermark.py:161-226): seed = (salt * sum(ctx)) mod (2^64-1) -> MT19937 (torch CPU generator) ->
 * randperm(alive), randperm(dead) -> first int(n_alive*gamma) alive + int(V*gamma)-that dead ids are green.
 * Output: a bitmask TABLE [n_rows][words], words = ceil(V/32), row s = greenlist for context SUM s (only the sum
 * enters the seed); FIXED seeding uses n_rows = 1 (seed 0).  Bit (id & 31) of word (id >> 5) is set iff id is green.
 * alive/dead are int64 id lists exactly as armm_wrapper.py:42-55 builds them (alive in file order, dead sorted).
 * ---------------------------------------------------------------------------------------------------------- */
int wmar_greenlist_build_host(int64_t vocab_size, double gamma, int split_strategy, int seed_strategy,
                              uint64_t salt_key, const int64_t *h_alive, int64_t n_alive, const int64_t *h_dead,
                              int64_t n_dead, int64_t n_rows, uint32_t *h_table_out, int n_threads);
/* same table built by a CUDA kernel (one CTA per row) straight into HBM */
int wmar_greenlist_build_device(int64_t vocab_size, double gamma, int split_strategy, int seed_strategy,
                                uint64_t salt_key, const int64_t *d_alive, int64_t n_alive, const int64_t *d_dead,
                                int64_t n_dead, int64_t n_rows, uint32_t *d_table_out, void *stream);

/* Watermark description shared by the operators below. */
typedef struct wmar_wm_params {
    const uint32_t *d_table; /* [n_rows][ceil(V/32)] from wmar_greenlist_build_*; NULL = no watermark */
    int64_t n_rows;
    int64_t vocab_size;
    int seed_strategy;   /* wmar_seed_strategy */
    int context_size;    /* h */
    int spatial_dim;     /* 16 (Taming/RAR) or 32 (Chameleon), gentime_watermark.py:153 */
    float delta;
    double gamma;
} wmar_wm_params;

/*
 * The logit-processor operator: GentimeWatermark._process_logits (gentime_watermark.py:229-271), the callback invoked
 * at mingpt.py:350, rar.py:451, chameleon.py:320.  In place: logits[b, green(ctx_b)] += delta.  Rows whose history is
 * shorter than the context are left untouched (the reference swallows the ValueError, :268-270).
 * d_past_ids int64 [B][t] with row stride past_stride (elements); d_logits fp32 [B][V] contiguous.
 */
int wmar_wm_process_logits(const wmar_wm_params *wm, const int64_t *d_past_ids, int64_t B, int64_t t,
                           int64_t past_stride, float *d_logits, void *stream);

/* Sampling parameters of one step (mingpt.py:326-368 / rar.py:446-454). */
typedef struct wmar_sample_params {
    float temperature;
    int top_k;        /* <= 0: no top-k  (HF TopKLogitsWarper, keeps ties)                         */
    double top_p;     /* <= 0 or >= 1: no top-p (HF TopPLogitsWarper, ascending sort, keep last); double so
                         that (float)(1 - top_p) equals torch's scalar cast                        */
    int greedy;       /* 1: first arg-max of softmax (sample_logits=False, mingpt.py:360-361)      */
    uint64_t seed;    /* Philox key used when d_noise == NULL and !greedy                          */
    /* rng_mode 1 (d_noise == NULL, !greedy): draw q exactly as torch.multinomial would on this device -- the Exp(1)
     * tensor of `empty_like(probs).exponential_(1)` on torch's CUDA generator (ATen distribution_nullary_kernel:
     * Philox4x32-10 keyed by the generator seed, subsequence = thread index of a 256-thread grid of
     * min(SMs * threads_per_SM / 256, ceil(numel / 1024)) blocks, 4 values per Philox call, offset advancing by
     * 4 * ceil(numel / (4 * threads)) per call).  seed = the generator's seed, torch_offset = its Philox offset before
     * the first step, torch_threads = 256 * grid, torch_numel = rows * row length of the probability tensor
     * (mingpt.py:363, rar.py:454, token_selector.py:26-47).  The caller advances the generator afterwards. */
    int rng_mode;
    int torch_threads;
    uint64_t torch_offset;
    int64_t torch_numel;
    int64_t torch_rowlen;   /* row length of the probability tensor (the full vocabulary)              */
} wmar_sample_params;

/*
 * Fused operator for one step on given logits: +delta on green -> /T -> top-k -> top-p -> softmax -> multinomial.
 * d_noise fp32 [B][V] = q ~ Exp(1) as drawn by torch.multinomial (argmax(p / q)); NULL = in-kernel Philox.
 * d_logits is not modified.  d_out_ids int64 [B].
 */
int wmar_wm_sample(const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_past_ids, int64_t B,
                   int64_t t, int64_t past_stride, const float *d_logits, const float *d_noise, int64_t *d_out_ids,
                   void *stream);

/* Test hook of rng_mode 1: d_out fp32 [rows][rowlen] = the Exp(1) tensor torch's CUDA generator in state (seed, offset)
 * draws for `empty(rows, rowlen).exponential_(1)`; threads = 256 * min(SMs * max_threads_per_SM / 256, ceil(numel / 1024)). */
int wmar_debug_torch_exponential(uint64_t seed, uint64_t offset, int64_t rows, int64_t rowlen, int threads, float *d_out,
                                 void *stream);

/* Reads and clears the device-side error flag set by the operators above (synchronises `stream`):
 * WMAR_OK, or WMAR_ERR_RANGE when a context sum fell outside the table / the top-p candidate set overflowed. */
int wmar_check_device_flag(void *stream);

/*
 * Detector: GentimeWatermark.detect + _score_ngrams_in_passage (gentime_watermark.py:285-344).  One CTA per passage.
 * d_codes int64 [B][L].  Outputs (device, any may be NULL): n_green int32[B], n_scored int32[B] (unique n-grams),
 * z f64[B] = (n_green - gamma T)/sqrt(T gamma (1-gamma)), p f64[B] = I_gamma(n_green, T - n_green + 1) = P[Bin(T,gamma)
 * >= n_green] (scipy.special.betainc at :338), mask int8 [B][mask_stride] (-1 unscored/repeat, 0 red, 1 green; entry
 * layout as the reference's list: h leading -1 then one entry per n-gram), mask_len int32[B].
 * Returns WMAR_ERR_SHORT if L - h < 1.
 */
int wmar_detect(const wmar_wm_params *wm, const int64_t *d_codes, int64_t B, int64_t L, int32_t *d_n_green,
                int32_t *d_n_scored, double *d_z, double *d_p, int8_t *d_mask, int64_t mask_stride,
                int32_t *d_mask_len, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Taming minGPT decode engine.  Replaces sample_with_past (mingpt.py:326-368) + GPT.forward_with_past (:183-214) +
 * Block / CausalSelfAttention (:42-122) with the watermark operator and the sampler fused behind the lm_head, the
 * whole `steps`-token loop enqueued on `stream` with no host round trip per token.
 * Weights are fp32, borrowed, passed as a table of device pointers in this order:
 *   [0] tok_emb [V][d]   [1] pos_emb [block_size][d]
 *   per layer l (12 entries, base 2 + 12 l): ln1.weight, ln1.bias, Wqkv [3d][d] (rows: query|key|value), bqkv [3d],
 *     Wproj [d][d], bproj, ln2.weight, ln2.bias, W1 [4d][d], b1, W2 [d][4d], b2
 *   then ln_f.weight, ln_f.bias, head.weight [V][d]
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct wmar_gpt_config {
    int vocab_size, block_size, n_layer, n_head, n_embd;
    int max_batch; /* rows per sampling call, <= 16 in this version */
} wmar_gpt_config;

typedef struct wmar_gpt wmar_gpt;

int wmar_gpt_create(const wmar_gpt_config *cfg, const void *const *d_weights, int n_weights, wmar_gpt **out);
void wmar_gpt_destroy(wmar_gpt *g);
/*
 * d_cond int64 [B] class ids (the conditioning token, taming_wrapper.py:62); d_noise fp32 [steps][B][V] or NULL;
 * d_out_codes int64 [B][steps].  d_out_logits (optional, may be NULL) fp32 [steps][B][V] receives the raw lm_head
 * logits of every step (before the watermark), for numerics tests.
 */
int wmar_gpt_sample(wmar_gpt *g, const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_cond,
                    int64_t B, int64_t steps, const float *d_noise, int64_t *d_out_codes, float *d_out_logits,
                    void *stream);
/* algorithmic bytes moved by one sampling call of B rows x steps (weights once per step + KV), for the roofline */
double wmar_gpt_algorithmic_bytes(const wmar_gpt *g, int64_t B, int64_t steps);
/* kernels launched per decode step by this engine */
int wmar_gpt_launches_per_step(const wmar_gpt *g);
/*
 * Static work plan of the persistent decode-step kernel (WMAR_STEP=pstep: one launch per token step whose CTAs each
 * stream a fixed share of every weight matrix, csrc/pstep.cuh).  Host logic only, no CUDA call: fills h_progs_out with
 * G records of wmar_pstep_prog_bytes() bytes (layout: csrc/pstep_plan.h PsProg) and h_info_out[8] with {Kp, KC, NBn,
 * stages per packed layer, stages of the packed head, max load, min load, 0}; wmar_pstep_stage_src tells which weights
 * ring stage `s` of CTA `cta` holds (out4 = {phase or -1 for fc2, n16 tile, k chunk, fc2 n-block}).  Replaces the
 * per-layer launch chain of mingpt.py:183-214; exported so that the plan's invariants are unit-tested without a GPU.
 */
int wmar_pstep_prog_bytes(void);
int wmar_pstep_plan_debug(int G, int d, int H, int V, void *h_progs_out, long long *h_info_out);
int wmar_pstep_stage_src(int G, int d, int H, int V, int cta, int s, int head, int *out4);

/* ------------------------------------------------------------------------------------------------------------
 * RAR decode engine.  Replaces RAR.generate (deps/rar/modeling/rar.py:408-459) + forward_fn (:319-405) + Block /
 * Attention / FinalLayer (:56-183) as called by RarARMMWrapper.sample (wmar/models/rar_wrapper.py:89-107): classifier-
 * free guidance u + (c - u) * guidance_scale over rows (cond | none-cond), watermark on the guided logits, /T, softmax,
 * multinomial; the whole loop enqueued on `stream`.  Weights fp32, borrowed, in this order:
 *   [0] cls_token [d]  [1] embeddings.weight [codebook+1+n_classes+1][d]  [2] pos_embed [>= seq+2][d]
 *   [3] target_aware_pos_embed [>= seq+2][d]  [4] timesteps_embeddings [>= seq+1][d]
 *   per block l (18 entries, base 5 + 18 l): norm1.weight, norm1.bias, attn.qkv.weight [3d][d], attn.qkv.bias,
 *     attn.q_norm.weight, attn.q_norm.bias, attn.k_norm.weight, attn.k_norm.bias, attn.proj.weight, attn.proj.bias,
 *     norm2.weight, norm2.bias, mlp.fc1.weight [mlp][d], mlp.fc1.bias, mlp.fc2.weight [d][mlp], mlp.fc2.bias,
 *     adaLN_modulation.1.weight [6d][d], adaLN_modulation.1.bias
 *   then adaln_before_head.adaLN_modulation.1.weight [2d][d], .bias, lm_head.weight [codebook][d], lm_head.bias
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct wmar_rar_config {
    int codebook_size, n_classes, image_seq_len, n_layer, n_head, hidden, mlp;
    int max_batch; /* images per sampling call, <= 8 (2B guided rows <= 16) */
} wmar_rar_config;

typedef struct wmar_rar wmar_rar;

int wmar_rar_create(const wmar_rar_config *cfg, const void *const *d_weights, int n_weights, wmar_rar **out);
void wmar_rar_destroy(wmar_rar *g);
/* d_cond int64 [B] ImageNet class ids; d_noise fp32 [steps][B][V] or NULL; d_out_ids int64 [B][steps];
 * d_out_logits (optional) fp32 [steps][B][V] = the guided logits of every step before the watermark. */
int wmar_rar_sample(wmar_rar *g, const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_cond,
                    int64_t B, int64_t steps, float guidance_scale, const float *d_noise, int64_t *d_out_ids,
                    float *d_out_logits, void *stream);
double wmar_rar_algorithmic_bytes(const wmar_rar *g, int64_t B, int64_t steps);

/* A single skinny GEMM (the dominant kernel), exposed for unit tests and the roofline microbenchmark:
 * y[16][N] = x[16][K] . W[N][K]^T + bias, fp32 in/out, 3xTF32 tensor-core products with fp32 accumulation. */
int wmar_skinny_gemm(const float *d_x, const float *d_w, const float *d_bias, float *d_y, int64_t N, int64_t K,
                     int split_k, void *stream);

/* The same GEMM with bf16 weights (Chameleon / Anole): y = bf16(bf16(x) . W^T), fp32 accumulation, y holds bf16 values. */
int wmar_skinny_gemm_bf16(const float *d_x, const void *d_w_bf16, float *d_y, int64_t N, int64_t K, int split_k,
                          void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * Chameleon / Anole-7B image-token decode engine.
 * Replaces, for image generation, ImageDecoder.__init__/__next__ (deps/chameleon/inference/chameleon.py:299-389), the
 * ChameleonGenerator step (generation.py:68-103), ChameleonModelAdapter (model_adapter.py:51-118, the per-row key
 * ranges of BlockDiagonalCausalWithOffsetPaddedKeysMask) and Transformer.forward_with_attn_bias
 * (transformer.py:97-337), with the logits processors of chameleon.py:312-327 (InBatchInstructCFGLogitsProcessor,
 * the watermark callback, AllowOnlyTokensLogitsProcessor, temperature, top-p) and the
 * ReplicatedInputTokenSelector(Multinomial | Argmax, n=3) of token_selector.py:26-47 fused behind the output head.
 * The model is bf16 (matrices as bf16, norm parameters as fp32 holding the bf16 values), weights borrowed, in this order:
 *   [0] tok_embeddings.weight bf16 [V][d]
 *   per layer l (10 entries, base 1 + 10 l): attention_norm.weight f32 [d], attention.wqkv.weight bf16 [(H+2Hkv)128][d],
 *     attention.q_normalization.weight / .bias f32 [128], attention.k_normalization.weight / .bias f32 [128],
 *     attention.wo.weight bf16 [d][H 128], ffn_norm.weight f32 [d], feed_forward.w13.weight bf16 [2F][d] with its rows
 *     INTERLEAVED per 64-row tile (rows 64T..64T+31 = w1 rows 32T..32T+31, rows 64T+32..64T+63 = w3 rows 32T..32T+31, so
 *     that the GEMM epilogue can form silu(x1) * x3), feed_forward.w2.weight bf16 [d][F]
 *   then norm.weight f32 [d], output.weight bf16 [V][d]
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct wmar_cham_config {
    int vocab_size, dim, n_layer, n_head, n_kv_head;
    int ffn_hidden;       /* F: the FeedForward hidden size after the multiple_of rounding (11008 for 7B)          */
    int max_seq;          /* cache positions per row: longest prompt + generated tokens                            */
    int max_batch;        /* images per call, <= 8 (n_groups * B guided rows <= 16)                                */
    int image_token_lo, image_token_hi; /* allowed ids [lo, hi): vocab.image_tokens (4..8195 for Chameleon)        */
    float norm_eps, rope_theta;
    int qk_norm;          /* ModelArgs.qk_normalization                                                            */
} wmar_cham_config;

typedef struct wmar_cham wmar_cham;

int wmar_cham_create(const wmar_cham_config *cfg, const void *const *d_weights, int n_weights, wmar_cham **out);
void wmar_cham_destroy(wmar_cham *g);
/* d_prompts int64 [n_groups*B][max_prompt] (rows: B full-conditioned, B image-conditioned, B unconditioned prompts, each
 * ending in <boi>, left-aligned), d_prompt_len int32 [n_groups*B], p_max = the longest prompt.  n_groups = 3, or 2 when
 * the image-conditioned rows equal the unconditioned ones (text-only prompts: the rows are then [full | unconditioned]
 * and the guidance reads the unconditioned logits for both; bit-identical to the 3-group result, 8 images per call).  d_noise fp32 [steps][B][V] (the
 * Exp(1) draws of torch.multinomial over the full vocabulary) or NULL; d_out_ids int64 [B][steps];
 * d_out_logits (optional) fp32 [steps][B][hi-lo] = the mixed logits of the image-token window before the watermark. */
int wmar_cham_sample(wmar_cham *g, const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_prompts,
                     const int32_t *d_prompt_len, int64_t max_prompt, int64_t p_max, int64_t B, int n_groups,
                     float guidance_text, float guidance_image, int64_t steps, const float *d_noise, int64_t *d_out_ids,
                     float *d_out_logits, void *stream);
double wmar_cham_algorithmic_bytes(const wmar_cham *g, int64_t B, int n_groups, int64_t p_max, int64_t steps);
int wmar_cham_launches_per_pass(const wmar_cham *g);
/* The logits-processor + token-selector operator of ImageDecoder on GIVEN logits (chameleon.py:312-346, generation.py:
 * 86-97): d_logits3 fp32 [3B][V] (full | image-conditioned | unconditioned rows) -> mixed = u + s_img (i-u) + s_txt (f-i)
 * -> watermark -> allow only ids [lo, hi) -> /T -> top-p -> softmax -> multinomial (d_noise fp32 [B][V] Exp(1), or
 * NULL = Philox) or argmax.  d_mixed fp32 [B][hi-lo] receives the mixed logits of the window; d_out_ids int64 [B]. */
int wmar_cham_select(const wmar_wm_params *wm, const wmar_sample_params *sp, const float *d_logits3, int64_t B, int64_t V,
                     int64_t image_token_lo, int64_t image_token_hi, float guidance_text, float guidance_image,
                     const int64_t *d_past_ids, int64_t t, int64_t past_stride, const float *d_noise, int64_t *d_out_ids,
                     float *d_mixed, void *stream);

/* ------------------------------------------------------------------------------------------------------------
 * VQGAN tokenizer (Taming / Chameleon family and MaskGIT / RAR family).
 * Replaces codes_to_images / images_to_codes (taming_wrapper.py:79-92, rar_wrapper.py:109-128) and below them
 * Decoder/Encoder.forward (taming model.py:343-538; maskgit_vqgan.py:160-245), the quant convs (vqgan.py:64-73)
 * and the codebook arg-min / gather (quantize.py:272-331; maskgit_vqgan.py:283-321).
 * Weights: fp32, borrowed, as a table of device pointers in the order produced by the host shim
 * (wmar_b200/models/vqgan_pack.py documents it; conv weights are pre-transposed to [Cout][ky][kx][Cin]).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct wmar_vqgan_config {
    int family;            /* 0 = Taming/Chameleon (attention, conv down, biases), 1 = MaskGIT (RAR)      */
    int ch;                /* base channels (128)                                                          */
    int n_levels;          /* len(ch_mult)                                                                 */
    int ch_mult[8];
    int num_res_blocks;
    int attn_resolution;   /* Taming: resolution at which AttnBlocks are inserted (16); 0 = none           */
    int resolution;        /* image side (256)                                                             */
    int z_channels;        /* 256                                                                          */
    int embed_dim;         /* codebook vector dim (256)                                                    */
    int n_embed;           /* codebook size                                                                */
    int max_batch;
    int precision;         /* 0 = 3xTF32 (fp32-faithful), 1 = 1xTF32 (the reference's cuDNN default),
                              2 = bf16x3 on the tcgen05 3x3 convs (x = x1 + x2 in bf16, three kind::f16 products:
                                  half the tensor time of 3xTF32, ~2^-17 per product), other convs 3xTF32,
                              3 = 2 for the decoder, 0 for the encoder (detection keeps fp32-faithful codes)     */
} wmar_vqgan_config;

typedef struct wmar_vqgan wmar_vqgan;

int wmar_vqgan_create(const wmar_vqgan_config *cfg, const void *const *d_weights, int n_weights, wmar_vqgan **out);
void wmar_vqgan_destroy(wmar_vqgan *v);
/* d_codes int64 [B][s*s] -> d_images fp32 [B][3][S][S] (NCHW, [-1,1], clamped) */
int wmar_vqgan_decode(wmar_vqgan *v, const int64_t *d_codes, int64_t B, float *d_images, void *stream);
/* d_images fp32 [B][3][S][S] in [-1,1] -> d_codes int64 [B][s*s] */
int wmar_vqgan_encode(wmar_vqgan *v, const float *d_images, int64_t B, int64_t *d_codes, void *stream);
double wmar_vqgan_flops(const wmar_vqgan *v, int decode);

/* ------------------------------------------------------------------------------------------------------------
 * Evaluation augmentations: the transforms applied between codes_to_images and images_to_codes in
 * fill_batch_log (generate.py:142-163).  Images fp32 NCHW [B][3][H][W] in [0,1], out of place.  Replaces
 * wmar/augmentations/valuemetric.py:76-137 and geometric.py:26-117 (torchvision.transforms.functional below them).
 *   BRIGHTNESS       params {factor}                                  clamp(factor * x, 0, 1)
 *   GAUSSIAN_NOISE   params {std}, d_aux = N(0,1) noise [B][3][H][W]   clamp(x + noise * std, 0, 1)
 *   GAUSSIAN_BLUR    params {k}, d_aux = k x k normalised weights       reflect pad, depthwise correlation, clamp
 *   HFLIP
 *   AFFINE_NEAREST   params = inverse affine [2][3] (torchvision F.rotate / F.affine convention), output OH x OW
 *   CROP_RESIZE      params {h2, w2}: bilinear resize of x[:, :, :h2, :w2] back to H x W
 *   CROP_PAD         params {h2, w2}: x[:, :, :h2, :w2] zero padded back to H x W
 * ---------------------------------------------------------------------------------------------------------- */
enum wmar_aug_op {
    WMAR_AUG_BRIGHTNESS = 0, WMAR_AUG_GAUSSIAN_NOISE = 1, WMAR_AUG_GAUSSIAN_BLUR = 2, WMAR_AUG_HFLIP = 3,
    WMAR_AUG_AFFINE_NEAREST = 4, WMAR_AUG_CROP_RESIZE = 5, WMAR_AUG_CROP_PAD = 6
};
int wmar_augment(int op, const float *d_in, float *d_out, int64_t B, int64_t H, int64_t W, int64_t OH, int64_t OW,
                 const float *params, int n_params, const float *d_aux, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WMAR_B200_H */
