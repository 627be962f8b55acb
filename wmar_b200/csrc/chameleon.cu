// Chameleon / Anole-7B image-token decode engine: ImageDecoder + ChameleonGenerator + ChameleonModelAdapter +
// Transformer.forward_with_attn_bias (deps/chameleon/inference/chameleon.py:299-389, generation.py:68-103,
// model_adapter.py:51-118, transformer.py:97-337) as a replayed CUDA graph of hand-written kernels, no host work and no
// Python per token.
//
// Rows: R = 3B (full-conditioned rows 0..B-1, image-conditioned B..2B-1, unconditioned 2B..3B-1, chameleon.py:351-372),
// R <= 16 = the M of the skinny GEMM.  Every row has its own prompt; the reference right-aligns them and gives each row
// its own key range (BlockDiagonalCausalWithOffsetPaddedKeysMask: positions start at 0 per row).  Here the rows are
// right-aligned IN TIME: row r starts at pass Pmax - P_r, so that all rows consume their last prompt token (<boi>) in
// pass Pmax - 1, and pass i feeds row r its token p = i - (Pmax - P_r) at rotary position p.  With a KV cache a causal
// prefill is the same as feeding the prompt one position at a time.
// Per pass and layer (the model is bf16: every op rounds its output to bf16, activations live in fp32 buffers):
//     qkv  = bf16( RMSNorm(x) Wqkv^T )                                   transformer.py:111,238   (GEMM, RMS prologue)
//     q,k  = LayerNorm_hd(q), LayerNorm_hd(k)  (qk_normalization)        :116-123
//     q,k  = RoPE(q,k; interleaved pairs, theta, position p) ; append k,v to the bf16 cache     :130-138 (xformers)
//     y    = softmax(q K^T / sqrt(hd)) V  over keys 0..p of the row      :149-156
//     x    = bf16( x + bf16(y Wo^T) )                                    :158,238-244
//     h    = bf16( bf16(silu(x1)) * x3 ),  [x1|x3] = bf16( RMSNorm(x) W13^T )  (fused into the w13 epilogue)   :215-218
//     x    = bf16( x + bf16( h W2^T ) )                                    :219,245
//   logits = bf16( RMSNorm(x) Wout^T ).float()                           :314-319
// Sampling (passes >= Pmax - 1), chameleon.py:312-327 + generation.py:86-97:
//     mixed = u + s_img (i - u) + s_txt (f - i)  ->  +delta on green  ->  -inf outside the image tokens  ->  / T
//     -> top-p -> softmax -> multinomial (or argmax) on the B primary rows; the token is fed to all three row groups.
#include <vector>

#include "gemm_bf16.cuh"
#include "sample.cuh"

using namespace wmar;

namespace wmar {
int make_sample_args(const wmar_wm_params *wm, const wmar_sample_params *sp, int V, SampleArgs *out);
int *device_err_flag();
}  // namespace wmar

namespace {

struct ChamCall {
    SampleArgs sa;
    const int64_t *prompts;     // [3B][max_prompt]
    const int32_t *prompt_len;  // [3B]
    const float *noise;         // [steps][B][V] or null
    int64_t *out_ids;           // [B][steps]
    float *out_logits;          // [steps][B][W] mixed logits of the image-token window (before the watermark) or null
    float s_txt, s_img;
    int B, steps, max_prompt, p_max;
    int n_groups;   // 3: full | image-conditioned | unconditioned rows; 2: the image-conditioned rows ARE the unconditioned
                    // ones (text-only prompts: both reduce to <s> <boi>, chameleon.py:351-372), computed once
};

struct ChamLayer {
    const float *attn_norm;
    const __nv_bfloat16 *wqkv;
    const float *qn_g, *qn_b, *kn_g, *kn_b;
    const __nv_bfloat16 *wo;
    const float *ffn_norm;
    const __nv_bfloat16 *w13, *w2;
};

constexpr int CH_ATT_THREADS = 256;
constexpr int CH_HD = 128;

}  // namespace

struct wmar_cham {
    wmar_cham_config cfg;
    int n_sms, d, hd, F, W;   // W = number of allowed (image) tokens
    const __nv_bfloat16 *tok_emb, *wout;
    const float *norm_w;
    std::vector<ChamLayer> layers;
    float *x, *qkv, *y, *h13, *logits, *guided, *ws;
    size_t ws_bytes;
    __nv_bfloat16 *kcache, *vcache;
    float2 *stats;
    unsigned *counters;
    int64_t *seq;     // [B][max_seq]: tokens of the primary rows (prompt, then the generated ids) = the watermark's past_ids
    int *rowpos;      // [16] position of each row in this pass (-1 = not started yet)
    int *pass;        // device pass counter
    ChamCall *d_call, *h_call;
    cudaEvent_t call_done;
    bool call_pending;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    size_t graph_smem;
    int graph_B;
    int s_qkv, s_wo, s_w13, s_w2, s_out;
    int launches_per_pass;
};

namespace {

__global__ void cham_init_kernel(const ChamCall *cp, int64_t *seq, int max_seq, int *pass) {
    // seq[b] = prompt of the primary row b
    const int b = blockIdx.x;
    if (b < cp->B) {
        const int P = cp->prompt_len[b];
        for (int i = threadIdx.x; i < P; i += blockDim.x) seq[(size_t)b * max_seq + i] = cp->prompts[(size_t)b * cp->max_prompt + i];
    }
    if (b == 0 && threadIdx.x == 0) *pass = 0;
}
__global__ void cham_advance_kernel(int *pass) { *pass += 1; }

// x[r] = tok_embeddings[token of row r in this pass] (bf16 -> fp32), (mean, M2) partials per 64-column tile, rowpos[r]
__global__ void __launch_bounds__(256) cham_embed_kernel(const ChamCall *cp, const int *pass, const int64_t *seq, int max_seq,
                                                         const __nv_bfloat16 *__restrict__ emb, int d, int V,
                                                         float *__restrict__ x, float2 *__restrict__ stats, int *rowpos) {
    const int r = blockIdx.x, i = *pass, B = cp->B;
    int p = -1;
    long long tok = 0;
    if (r < cp->n_groups * B) {
        const int P = cp->prompt_len[r];
        p = i - (cp->p_max - P);
        if (p >= 0) {
            if (p < P) tok = cp->prompts[(size_t)r * cp->max_prompt + p];
            else {
                const int b = r % B;                      // the generated ids are shared by the three row groups
                const int Pb = cp->prompt_len[b];
                tok = seq[(size_t)b * max_seq + Pb + (p - P)];
            }
            if (tok < 0 || tok >= V) tok = 0;
        }
    }
    if (threadIdx.x == 0) rowpos[r] = p;
    const bool valid = p >= 0;
    const int lane16 = threadIdx.x & 15;
    for (int c0 = (threadIdx.x >> 4) * 64; c0 < d; c0 += (blockDim.x >> 4) * 64) {
        const int c = c0 + lane16 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const uint2 e = *reinterpret_cast<const uint2 *>(emb + (size_t)tok * d + c);
            const __nv_bfloat162 e0 = *reinterpret_cast<const __nv_bfloat162 *>(&e.x), e1 = *reinterpret_cast<const __nv_bfloat162 *>(&e.y);
            v = make_float4(__low2float(e0), __high2float(e0), __low2float(e1), __high2float(e1));
        }
        *reinterpret_cast<float4 *>(x + (size_t)r * d + c) = v;
        float s = v.x + v.y + v.z + v.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 64.0f);
        float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane16 == 0) stats[(c0 / 64) * 16 + r] = make_float2(mean, q);
    }
}

// One CTA per (q head, row).  qk LayerNorm, RoPE at the row's position, cache append (by the first q head of each kv
// group), attention over keys 0..p.  head_dim 128: sixteen lanes cover a key (one 16-byte load = 8 dims per lane), K and V loops software-pipelined.
__global__ void __launch_bounds__(CH_ATT_THREADS, 4) cham_attn_kernel(const float *__restrict__ qkv, int ld_qkv, int H, int Hkv, int T,
                                                                   int qk_norm, float theta,
                                                                   const float *__restrict__ qn_g, const float *__restrict__ qn_b,
                                                                   const float *__restrict__ kn_g, const float *__restrict__ kn_b,
                                                                   __nv_bfloat16 *__restrict__ kcache, __nv_bfloat16 *__restrict__ vcache,
                                                                   int layer, const int *rowpos, float *__restrict__ y, int ld_y) {
    constexpr int HD = CH_HD;
    __shared__ __align__(16) float sq[HD], sk[HD], sv[HD];
    __shared__ float sc[4096];
    __shared__ __align__(16) float part[8][HD];
    __shared__ float red[8];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int h = blockIdx.x, r = blockIdx.y, p = rowpos[r];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *yo = y + (size_t)r * ld_y + h * HD;
    if (p < 0 || p >= T) {   // row not started yet: defined (zero) output, nothing cached
        if (tid < HD) yo[tid] = 0.f;
        return;
    }
    const int grp = H / Hkv, hk = h / grp;
    const float *q = qkv + (size_t)r * ld_qkv + h * HD;
    const float *kn = qkv + (size_t)r * ld_qkv + (H + hk) * HD;
    const float *vn = qkv + (size_t)r * ld_qkv + (H + Hkv + hk) * HD;
    const size_t base = (((size_t)layer * 16 + r) * Hkv + hk) * (size_t)T * HD;
    __nv_bfloat16 *K = kcache + base, *V = vcache + base;
    if (warp < 2) {
        // warp 0: q, warp 1: the new k.  nn.LayerNorm(head_dim) on a bf16 tensor (fp32 math, bf16 result), then RoPE
        // on adjacent pairs (x0 + i x1) * exp(i p theta^(-2j/hd)) in fp32, rounded to bf16 (xformers rope_padded)
        const float *src = warp == 0 ? q : kn;
        const float *gg = warp == 0 ? qn_g : kn_g, *bb = warp == 0 ? qn_b : kn_b;
        float4 v4 = *reinterpret_cast<const float4 *>(src + 4 * lane);
        float vals[4] = {v4.x, v4.y, v4.z, v4.w};
        if (qk_norm) {
            const float mean = warp_sum(vals[0] + vals[1] + vals[2] + vals[3]) * (1.0f / HD);
            float v2 = 0.f;
#pragma unroll
            for (int e = 0; e < 4; e++) { const float dd = vals[e] - mean; v2 += dd * dd; }
            v2 = warp_sum(v2);
            const float rstd = 1.0f / sqrtf(v2 * (1.0f / HD) + 1e-5f);
#pragma unroll
            for (int e = 0; e < 4; e++) vals[e] = bf16r((vals[e] - mean) * rstd * gg[4 * lane + e] + bb[4 * lane + e]);
        }
        float out[4];
#pragma unroll
        for (int pr = 0; pr < 2; pr++) {
            const int jpair = 2 * lane + pr;                       // pair index 0..63
            const float freq = powf(theta, -2.0f * (float)jpair / (float)HD);
            float sn, cs;
            sincosf((float)p * freq, &sn, &cs);
            const float x0 = vals[2 * pr], x1 = vals[2 * pr + 1];
            out[2 * pr] = bf16r(x0 * cs - x1 * sn);
            out[2 * pr + 1] = bf16r(x0 * sn + x1 * cs);
        }
        float *dst = warp == 0 ? sq : sk;
        *reinterpret_cast<float4 *>(dst + 4 * lane) = make_float4(out[0], out[1], out[2], out[3]);
        if (warp == 1 && h % grp == 0) {
            __nv_bfloat162 a = __floats2bfloat162_rn(out[0], out[1]), b = __floats2bfloat162_rn(out[2], out[3]);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
            *reinterpret_cast<uint2 *>(K + (size_t)p * HD + 4 * lane) = pk;
        }
    } else if (warp == 2) {
        const float4 v4 = *reinterpret_cast<const float4 *>(vn + 4 * lane);
        *reinterpret_cast<float4 *>(sv + 4 * lane) = v4;          // already bf16 values (GEMM epilogue)
        if (h % grp == 0) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v4.x, v4.y), b = __floats2bfloat162_rn(v4.z, v4.w);
            uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
            *reinterpret_cast<uint2 *>(V + (size_t)p * HD + 4 * lane) = pk;
        }
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)HD);
    const int nk = p + 1;
    // SIXTEEN lanes per key: a lane owns 8 consecutive dims (one 16-byte load of a bf16 cache row), a warp instruction
    // covers TWO keys (half = lane >> 4), a round of BATCH instructions 16 keys per warp = 128 keys per CTA; the dot
    // product needs 4 shuffle steps.  Twice the bytes in flight per load instruction of the 32-lane layout.
    const int half = lane >> 4, l16 = lane & 15;
    float q8[8];
    {
        const float4 qa = *reinterpret_cast<const float4 *>(sq + 8 * l16), qb = *reinterpret_cast<const float4 *>(sq + 8 * l16 + 4);
        q8[0] = qa.x; q8[1] = qa.y; q8[2] = qa.z; q8[3] = qa.w; q8[4] = qb.x; q8[5] = qb.y; q8[6] = qb.z; q8[7] = qb.w;
    }
    constexpr int BATCH = 8, ROUND = 16 * BATCH;      // keys per CTA per round: 8 warps x 2 halves x BATCH
    auto cvt8 = [&](const uint4 u, float (&f)[8]) {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; e++) { f[2 * e] = __uint_as_float(w[e] << 16); f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
    };
    auto own8 = [&](const float *src, float (&f)[8]) {
        const float4 a4 = *reinterpret_cast<const float4 *>(src + 8 * l16), b4 = *reinterpret_cast<const float4 *>(src + 8 * l16 + 4);
        f[0] = a4.x; f[1] = a4.y; f[2] = a4.z; f[3] = a4.w; f[4] = b4.x; f[5] = b4.y; f[6] = b4.z; f[7] = b4.w;
    };
    const int jw = 2 * warp + half;                    // this half-warp's first key; its keys: jw + 16 u + ROUND * round
    // software pipeline: the rows of the NEXT round are requested before this round's dot products and shuffle reductions
    uint4 kraw[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        const int j = jw + 16 * u;
        kraw[u] = j < p ? *reinterpret_cast<const uint4 *>(K + (size_t)j * HD + 8 * l16) : make_uint4(0u, 0u, 0u, 0u);
    }
    for (int j0 = 0; j0 < nk; j0 += ROUND) {          // block-uniform trip count: the shuffles below use the full mask
        float sd[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + jw + 16 * u;
            float k8[8];
            if (j == p) own8(sk, k8);                  // this pass's own key
            else cvt8(kraw[u], k8);
            const int jn = j + ROUND;                  // slot u is free again: request its row of the next round
            kraw[u] = jn < p ? *reinterpret_cast<const uint4 *>(K + (size_t)jn * HD + 8 * l16) : make_uint4(0u, 0u, 0u, 0u);
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 8; e++) d = fmaf(q8[e], k8[e], d);
            sd[u] = d;
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < BATCH; u++) sd[u] += __shfl_xor_sync(0xffffffffu, sd[u], o);
        if (l16 == 0) {
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                if (j0 + jw + 16 * u < nk) sc[j0 + jw + 16 * u] = sd[u] * scale;
        }
    }
    // the first round of V rows is requested now: it lands while the softmax statistics are reduced
    uint4 vraw[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        const int j = jw + 16 * u;
        vraw[u] = j < p ? *reinterpret_cast<const uint4 *>(V + (size_t)j * HD + 8 * l16) : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < nk; j += CH_ATT_THREADS) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += CH_ATT_THREADS) {
        const float e = expf(sc[j] - m);
        sc[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) sum += red[w];
    const float inv = 1.0f / sum;
    float acc8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < nk; j0 += ROUND) {
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + jw + 16 * u;
            float v8[8];
            if (j == p) own8(sv, v8);
            else cvt8(vraw[u], v8);
            const int jn = j + ROUND;
            vraw[u] = jn < p ? *reinterpret_cast<const uint4 *>(V + (size_t)jn * HD + 8 * l16) : make_uint4(0u, 0u, 0u, 0u);
            const float pr = j < nk ? sc[j] * inv : 0.f;
#pragma unroll
            for (int e = 0; e < 8; e++) acc8[e] = fmaf(pr, v8[e], acc8[e]);
        }
    }
    // the two halves of a warp hold different keys of the same dims
#pragma unroll
    for (int e = 0; e < 8; e++) acc8[e] += __shfl_xor_sync(0xffffffffu, acc8[e], 16);
    if (half == 0) {
        *reinterpret_cast<float4 *>(&part[warp][8 * l16]) = make_float4(acc8[0], acc8[1], acc8[2], acc8[3]);
        *reinterpret_cast<float4 *>(&part[warp][8 * l16 + 4]) = make_float4(acc8[4], acc8[5], acc8[6], acc8[7]);
    }
    __syncthreads();
    if (tid < HD) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) o += part[w][tid];
        yo[tid] = bf16r(o);
    }
}

// mixed = u + s_img (i - u) + s_txt (f - i) over the allowed (image token) window [lo, lo + W) (logits_processor.py:312-335,
// 135-151: everything outside the window becomes -inf after the watermark, so only the window is ever sampled from)
__global__ void __launch_bounds__(256) cham_guide_kernel(const ChamCall *cp, const float *__restrict__ logits, int V, int lo, int W,
                                                         float *__restrict__ guided, const int *pass) {
    const int b = blockIdx.x, i = *pass;
    const int s = i - (cp->p_max - 1);
    if (s < 0 || s >= cp->steps) return;
    const int B = cp->B;
    const float *lf = logits + (size_t)b * V + lo, *li = logits + (size_t)(B + b) * V + lo,
                *lu = logits + (size_t)((cp->n_groups - 1) * B + b) * V + lo;   // 2 groups: li == lu
    const float s_txt = cp->s_txt, s_img = cp->s_img;
    for (int v = threadIdx.x; v < W; v += blockDim.x) {
        const float u = lu[v], im = li[v], f = lf[v];
        const float mixed = (u + s_img * (im - u)) + s_txt * (f - im);
        guided[(size_t)b * W + v] = mixed;
        if (cp->out_logits != nullptr) cp->out_logits[((size_t)s * B + b) * W + v] = mixed;
    }
}

__global__ void __launch_bounds__(SAMPLE_THREADS, 1) cham_sample_kernel(const ChamCall *cp, const float *__restrict__ guided, int W,
                                                                         int V, int lo, int64_t *seq, int max_seq, const int *pass,
                                                                         int *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.x, i = *pass;
    const int s = i - (cp->p_max - 1);
    if (s < 0 || s >= cp->steps) return;
    const SampleArgs a = cp->sa;
    const int B = cp->B;
    const float *noise = cp->noise ? cp->noise + ((size_t)s * B + b) * V + lo : nullptr;
    const int Pb = cp->prompt_len[b];
    // past_ids of the reference = the whole input row (prompt + generated ids), generation.py:88
    int id = sample_row(a, guided + (size_t)b * W, seq + (size_t)b * max_seq, (long long)(Pb + s), noise,
                        ((unsigned long long)s << 32) | (unsigned)b, err, smem_raw);
    if (threadIdx.x == 0) {
        seq[(size_t)b * max_seq + Pb + s] = id;
        cp->out_ids[(size_t)b * cp->steps + s] = id;
    }
}

// stand-alone form of the two kernels above for given logits (wmar_cham_select)
__global__ void __launch_bounds__(256) cham_mix_kernel(const float *__restrict__ logits, int B, int V, int lo, int W, float s_txt,
                                                       float s_img, float *__restrict__ mixed) {
    const int b = blockIdx.x;
    const float *lf = logits + (size_t)b * V + lo, *li = logits + (size_t)(B + b) * V + lo, *lu = logits + (size_t)(2 * B + b) * V + lo;
    for (int v = threadIdx.x; v < W; v += blockDim.x) {
        const float u = lu[v], im = li[v], f = lf[v];
        mixed[(size_t)b * W + v] = (u + s_img * (im - u)) + s_txt * (f - im);
    }
}
__global__ void __launch_bounds__(SAMPLE_THREADS, 1) cham_select_kernel(SampleArgs a, const float *__restrict__ mixed, int V, int lo,
                                                                         const int64_t *__restrict__ past, long long t,
                                                                         long long past_stride, const float *__restrict__ noise,
                                                                         int64_t *__restrict__ out_ids, int *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.x;
    int id = sample_row(a, mixed + (size_t)b * a.V, past ? past + (long long)b * past_stride : nullptr, t,
                        noise ? noise + (size_t)b * V + lo : nullptr, (unsigned long long)b, err, smem_raw);
    if (threadIdx.x == 0) out_ids[b] = id;
}

void cham_free_graph(wmar_cham *g) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    g->exec = nullptr;
    g->graph = nullptr;
}

int cham_enqueue_pass(wmar_cham *g, int B, int n_groups, size_t sample_smem, cudaStream_t s) {
    const wmar_cham_config &c = g->cfg;
    const int d = g->d, H = c.n_head, Hkv = c.n_kv_head, V = c.vocab_size, F = g->F;
    const int qkv_n = (H + 2 * Hkv) * g->hd;
    const int stat_tiles = d / 64;
    int rc, launches = 0;
    cham_embed_kernel<<<16, 256, 0, s>>>(g->d_call, g->pass, g->seq, c.max_seq, g->tok_emb, d, V, g->x, g->stats, g->rowpos);
    WMAR_LAUNCH_CHECK();
    launches++;
    // flag-carrying split-K hand-off (gemm.cuh): epoch = the pass counter, salt = launch index within the pass
    static const bool ll_on = []() { const char *e = getenv("WMAR_LL"); return !(e && e[0] == '0'); }();
    WMAR_REQUIRE(4 * c.n_layer + 1 < 1024, "too many GEMM launches per pass for the hand-off flag");
    unsigned salt = 0;
    auto ll = [&](Bf16GemmArgs &q) { if (ll_on) { q.ll_epoch = g->pass; q.ll_salt = ++salt; } };
    for (int l = 0; l < c.n_layer; l++) {
        const ChamLayer &L = g->layers[l];
        Bf16GemmArgs a{};
        a.ws = g->ws; a.counters = g->counters; a.eps = c.norm_eps;
        a.X = g->x; a.ldx = d; a.W = L.wqkv; a.Y = g->qkv; a.ldy = qkv_n; a.N = qkv_n; a.K = d; a.splits = g->s_qkv;
        a.rms_w = L.attn_norm; a.stats_in = g->stats; a.n_stat_tiles = stat_tiles;
        ll(a);
        if ((rc = launch_skinny_gemm_bf16(BPRO_RMS, BEPI_STORE, a, s))) return rc;
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)H, (unsigned)(n_groups * B), 1);
            cfg.blockDim = dim3(CH_ATT_THREADS, 1, 1);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, cham_attn_kernel, (const float *)g->qkv, qkv_n, H, Hkv, c.max_seq, c.qk_norm,
                                               c.rope_theta, L.qn_g, L.qn_b, L.kn_g, L.kn_b, g->kcache, g->vcache, l,
                                               (const int *)g->rowpos, g->y, H * g->hd));
            g_launches.fetch_add(1);
        }
        Bf16GemmArgs o{};
        o.ws = g->ws; o.counters = g->counters;
        o.X = g->y; o.ldx = H * g->hd; o.W = L.wo; o.Y = g->x; o.ldy = d; o.N = d; o.K = H * g->hd; o.splits = g->s_wo;
        o.resid = g->x; o.ld_resid = d; o.stats_out = g->stats;
        ll(o);
        if ((rc = launch_skinny_gemm_bf16(BPRO_NONE, BEPI_RESID, o, s))) return rc;
        Bf16GemmArgs f{};
        f.ws = g->ws; f.counters = g->counters; f.eps = c.norm_eps;
        // w13 rows are interleaved by the host packer (32 x1 rows, then their 32 x3 rows, per 64-row tile): the epilogue
        // writes h = silu(x1) * x3 directly
        f.X = g->x; f.ldx = d; f.W = L.w13; f.Y = g->h13; f.ldy = F; f.N = 2 * F; f.K = d; f.splits = g->s_w13;
        f.rms_w = L.ffn_norm; f.stats_in = g->stats; f.n_stat_tiles = stat_tiles;
        ll(f);
        if ((rc = launch_skinny_gemm_bf16(BPRO_RMS, BEPI_SWIGLU, f, s))) return rc;
        Bf16GemmArgs w{};
        w.ws = g->ws; w.counters = g->counters;
        w.X = g->h13; w.ldx = F; w.W = L.w2; w.Y = g->x; w.ldy = d; w.N = d; w.K = F; w.splits = g->s_w2;
        w.resid = g->x; w.ld_resid = d; w.stats_out = g->stats;
        ll(w);
        if ((rc = launch_skinny_gemm_bf16(BPRO_NONE, BEPI_RESID, w, s))) return rc;
        launches += 5;
    }
    Bf16GemmArgs hd{};
    hd.ws = g->ws; hd.counters = g->counters; hd.eps = c.norm_eps;
    hd.X = g->x; hd.ldx = d; hd.W = g->wout; hd.Y = g->logits; hd.ldy = V; hd.N = V; hd.K = d; hd.splits = g->s_out;
    hd.rms_w = g->norm_w; hd.stats_in = g->stats; hd.n_stat_tiles = stat_tiles;
    ll(hd);
        if ((rc = launch_skinny_gemm_bf16(BPRO_RMS, BEPI_STORE, hd, s))) return rc;
    int *err = device_err_flag();
    WMAR_REQUIRE(err != nullptr, "cannot allocate the device error flag");
    cham_guide_kernel<<<B, 256, 0, s>>>(g->d_call, g->logits, V, c.image_token_lo, g->W, g->guided, g->pass);
    WMAR_LAUNCH_CHECK();
    cham_sample_kernel<<<B, SAMPLE_THREADS, sample_smem, s>>>(g->d_call, g->guided, g->W, V, c.image_token_lo, g->seq, c.max_seq,
                                                              g->pass, err);
    WMAR_LAUNCH_CHECK();
    cham_advance_kernel<<<1, 1, 0, s>>>(g->pass);
    WMAR_LAUNCH_CHECK();
    launches += 4;
    g->launches_per_pass = launches;
    return WMAR_OK;
}

}  // namespace

extern "C" {

int wmar_cham_create(const wmar_cham_config *cfg, const void *const *d_weights, int n_weights, wmar_cham **out) {
    WMAR_REQUIRE(cfg != nullptr && d_weights != nullptr && out != nullptr, "NULL argument");
    WMAR_REQUIRE(cfg->n_head > 0 && cfg->n_kv_head > 0 && cfg->n_head % cfg->n_kv_head == 0, "n_head must be a multiple of n_kv_head");
    WMAR_REQUIRE(cfg->dim % cfg->n_head == 0 && cfg->dim / cfg->n_head == CH_HD, "this engine supports head_dim 128 (Chameleon)");
    WMAR_REQUIRE(cfg->dim % 64 == 0 && cfg->vocab_size % 64 == 0 && cfg->ffn_hidden % 64 == 0, "dim, vocab and ffn_hidden must be multiples of 64");
    WMAR_REQUIRE(cfg->ffn_hidden % 32 == 0, "ffn_hidden must be a multiple of 32 (w13 row interleave)");
    WMAR_REQUIRE(cfg->max_seq >= 2 && cfg->max_seq <= 4096, "max_seq must be in [2,4096]");
    WMAR_REQUIRE(cfg->max_batch >= 1 && 2 * cfg->max_batch <= 16, "max_batch must be in [1,8] (2B or 3B guided rows <= 16)");
    WMAR_REQUIRE(cfg->image_token_lo >= 0 && cfg->image_token_hi > cfg->image_token_lo && cfg->image_token_hi <= cfg->vocab_size,
                 "bad image token range");
    WMAR_REQUIRE(n_weights == 1 + 10 * cfg->n_layer + 2, "weight table has the wrong number of entries");
    for (int i = 0; i < n_weights; i++) WMAR_REQUIRE(d_weights[i] != nullptr, "NULL weight pointer");
    wmar_cham *g = new (std::nothrow) wmar_cham();
    if (!g) return set_error(WMAR_ERR_NOMEM, "out of host memory%s%s");
    g->cfg = *cfg;
    g->d = cfg->dim; g->hd = CH_HD; g->F = cfg->ffn_hidden; g->W = cfg->image_token_hi - cfg->image_token_lo;
    int dev = 0;
    WMAR_CUDA_CHECK(cudaGetDevice(&dev));
    WMAR_CUDA_CHECK(cudaDeviceGetAttribute(&g->n_sms, cudaDevAttrMultiProcessorCount, dev));
    auto WB = [&](int i) { return reinterpret_cast<const __nv_bfloat16 *>(d_weights[i]); };
    auto WF = [&](int i) { return reinterpret_cast<const float *>(d_weights[i]); };
    g->tok_emb = WB(0);
    g->layers.resize(cfg->n_layer);
    for (int l = 0; l < cfg->n_layer; l++) {
        const int b = 1 + 10 * l;
        g->layers[l] = ChamLayer{WF(b), WB(b + 1), WF(b + 2), WF(b + 3), WF(b + 4), WF(b + 5), WB(b + 6), WF(b + 7), WB(b + 8), WB(b + 9)};
    }
    const int tb = 1 + 10 * cfg->n_layer;
    g->norm_w = WF(tb); g->wout = WB(tb + 1);
    const int d = g->d, H = cfg->n_head, Hkv = cfg->n_kv_head, V = cfg->vocab_size, F = g->F;
    const int qkv_n = (H + 2 * Hkv) * CH_HD;
    g->s_qkv = pick_splits_bf16(qkv_n, d, g->n_sms);
    g->s_wo = pick_splits_bf16(d, H * CH_HD, g->n_sms);
    g->s_w13 = pick_splits_bf16(2 * F, d, g->n_sms);
    g->s_w2 = pick_splits_bf16(d, F, g->n_sms);
    g->s_out = pick_splits_bf16(V, d, g->n_sms);
    WMAR_REQUIRE(d % (g->s_qkv * BG_KI) == 0 && (H * CH_HD) % (g->s_wo * BG_KI) == 0 && F % (g->s_w2 * BG_KI) == 0,
                 "dim / ffn_hidden must be multiples of 32");
    size_t ws_floats = 1;
    int max_tiles = 1;
    auto upd = [&](int N, int S) {
        size_t n = 2 * (size_t)(N / GEMM_NT) * S * GEMM_M * GEMM_NT;   // x2: {value, flag} words of the split-K hand-off
        if (n > ws_floats) ws_floats = n;
        if (N / GEMM_NT > max_tiles) max_tiles = N / GEMM_NT;
    };
    upd(qkv_n, g->s_qkv); upd(d, g->s_wo); upd(2 * F, g->s_w13); upd(d, g->s_w2); upd(V, g->s_out);
    const size_t kv_elems = (size_t)cfg->n_layer * 16 * Hkv * cfg->max_seq * CH_HD;
    WMAR_CUDA_CHECK(cudaMalloc(&g->x, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->qkv, sizeof(float) * 16 * qkv_n));
    WMAR_CUDA_CHECK(cudaMalloc(&g->y, sizeof(float) * 16 * H * CH_HD));
    WMAR_CUDA_CHECK(cudaMalloc(&g->h13, sizeof(float) * 16 * 2 * F));
    WMAR_CUDA_CHECK(cudaMalloc(&g->logits, sizeof(float) * 16 * (size_t)V));
    WMAR_CUDA_CHECK(cudaMalloc(&g->guided, sizeof(float) * (size_t)cfg->max_batch * g->W));
    WMAR_CUDA_CHECK(cudaMalloc(&g->kcache, sizeof(__nv_bfloat16) * kv_elems));
    WMAR_CUDA_CHECK(cudaMalloc(&g->vcache, sizeof(__nv_bfloat16) * kv_elems));
    WMAR_CUDA_CHECK(cudaMalloc(&g->ws, sizeof(float) * ws_floats));
    g->ws_bytes = sizeof(float) * ws_floats;
    WMAR_CUDA_CHECK(cudaMalloc(&g->stats, sizeof(float2) * (d / 64) * 16));
    WMAR_CUDA_CHECK(cudaMalloc(&g->counters, sizeof(unsigned) * max_tiles));
    WMAR_CUDA_CHECK(cudaMalloc(&g->seq, sizeof(int64_t) * (size_t)cfg->max_batch * cfg->max_seq));
    WMAR_CUDA_CHECK(cudaMalloc(&g->rowpos, sizeof(int) * 16));
    WMAR_CUDA_CHECK(cudaMalloc(&g->pass, sizeof(int)));
    WMAR_CUDA_CHECK(cudaMalloc(&g->d_call, sizeof(ChamCall)));
    WMAR_CUDA_CHECK(cudaMallocHost(&g->h_call, sizeof(ChamCall)));
    WMAR_CUDA_CHECK(cudaEventCreateWithFlags(&g->call_done, cudaEventDisableTiming));
    g->call_pending = false;
    WMAR_REQUIRE(device_err_flag() != nullptr, "cannot allocate the device error flag");
    WMAR_CUDA_CHECK(cudaMemset(g->counters, 0, sizeof(unsigned) * max_tiles));
    WMAR_CUDA_CHECK(cudaMemset(g->x, 0, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->qkv, 0, sizeof(float) * 16 * qkv_n));
    WMAR_CUDA_CHECK(cudaMemset(g->y, 0, sizeof(float) * 16 * H * CH_HD));
    WMAR_CUDA_CHECK(cudaMemset(g->h13, 0, sizeof(float) * 16 * 2 * F));
    WMAR_CUDA_CHECK(cudaMemset(g->logits, 0, sizeof(float) * 16 * (size_t)V));
    WMAR_CUDA_CHECK(cudaMemset(g->seq, 0, sizeof(int64_t) * (size_t)cfg->max_batch * cfg->max_seq));
    WMAR_CUDA_CHECK(cudaMemset(g->stats, 0, sizeof(float2) * (d / 64) * 16));
    WMAR_CUDA_CHECK(cudaMemset(g->rowpos, 0xff, sizeof(int) * 16));
    g->graph = nullptr; g->exec = nullptr; g->graph_smem = 0; g->graph_B = 0;
    g->launches_per_pass = 5 * cfg->n_layer + 5;
    // the zero-fills above went to the legacy default stream: finish them before the handle can be used from a
    // non-blocking stream (a second engine lane otherwise saw them land in the middle of its first generation)
    WMAR_CUDA_CHECK(cudaDeviceSynchronize());
    *out = g;
    return WMAR_OK;
}

void wmar_cham_destroy(wmar_cham *g) {
    if (!g) return;
    cudaDeviceSynchronize();
    cham_free_graph(g);
    cudaFree(g->x); cudaFree(g->qkv); cudaFree(g->y); cudaFree(g->h13); cudaFree(g->logits); cudaFree(g->guided);
    cudaFree(g->kcache); cudaFree(g->vcache); cudaFree(g->ws); cudaFree(g->stats); cudaFree(g->counters);
    cudaFree(g->seq); cudaFree(g->rowpos); cudaFree(g->pass); cudaFree(g->d_call); cudaFreeHost(g->h_call);
    cudaEventDestroy(g->call_done);
    delete g;
}

int wmar_cham_sample(wmar_cham *g, const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_prompts,
                     const int32_t *d_prompt_len, int64_t max_prompt, int64_t p_max, int64_t B, int n_groups, float guidance_text,
                     float guidance_image, int64_t steps, const float *d_noise, int64_t *d_out_ids, float *d_out_logits,
                     void *stream) {
    WMAR_REQUIRE(g != nullptr && sp != nullptr && d_prompts != nullptr && d_prompt_len != nullptr && d_out_ids != nullptr, "NULL argument");
    WMAR_REQUIRE(n_groups == 2 || n_groups == 3, "n_groups must be 2 (text-only prompts) or 3");
    WMAR_REQUIRE(B >= 1 && B <= g->cfg.max_batch && n_groups * B <= 16, "batch exceeds max_batch or 16 guided rows");
    WMAR_REQUIRE(p_max >= 1 && p_max <= max_prompt, "p_max (longest prompt) must be in [1, max_prompt]");
    WMAR_REQUIRE(steps >= 1 && p_max + steps <= g->cfg.max_seq, "prompt + steps exceeds max_seq");
    cudaStream_t s = as_stream(stream);
    const int V = g->cfg.vocab_size, W = g->W;
    wmar_wm_params wm_local{};
    wm_local.vocab_size = W;
    bool has_wm = wm != nullptr && wm->d_table != nullptr;
    if (has_wm) {
        WMAR_REQUIRE(wm->vocab_size == V, "watermark vocab_size != model vocab");
        wm_local = *wm;
        wm_local.vocab_size = W;   // the sampler works on the image-token window; the table keeps its full-vocab rows
    }
    SampleArgs sa;
    int rc = make_sample_args(&wm_local, sp, W, &sa);
    if (rc) return rc;
    sa.id_base = g->cfg.image_token_lo;
    sa.table_V = V;
    const size_t smem = sample_smem_bytes(W, sa.cand_cap);
    if (g->call_pending) WMAR_CUDA_CHECK(cudaEventSynchronize(g->call_done));
    g->h_call->sa = sa;
    g->h_call->prompts = d_prompts;
    g->h_call->prompt_len = d_prompt_len;
    g->h_call->noise = d_noise;
    g->h_call->out_ids = d_out_ids;
    g->h_call->out_logits = d_out_logits;
    g->h_call->s_txt = guidance_text;
    g->h_call->s_img = guidance_image;
    g->h_call->B = (int)B;
    g->h_call->steps = (int)steps;
    g->h_call->max_prompt = (int)max_prompt;
    g->h_call->p_max = (int)p_max;
    g->h_call->n_groups = n_groups;
    WMAR_CUDA_CHECK(cudaMemcpyAsync(g->d_call, g->h_call, sizeof(ChamCall), cudaMemcpyHostToDevice, s));
    WMAR_CUDA_CHECK(cudaEventRecord(g->call_done, s));
    g->call_pending = true;

    if (g->exec == nullptr || g->graph_smem != smem || g->graph_B != (int)B * 4 + n_groups) {
        cham_free_graph(g);
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(cham_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaStream_t cs;
        WMAR_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        WMAR_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = cham_enqueue_pass(g, (int)B, n_groups, smem, cs);
        cudaError_t e = cudaStreamEndCapture(cs, &g->graph);
        cudaStreamDestroy(cs);
        if (rc) { if (g->graph) cudaGraphDestroy(g->graph); g->graph = nullptr; return rc; }
        if (e != cudaSuccess) return set_error(WMAR_ERR_CUDA, "cudaStreamEndCapture: %s%s", cudaGetErrorString(e));
        WMAR_CUDA_CHECK(cudaGraphInstantiate(&g->exec, g->graph, 0));
        g->graph_smem = smem;
        g->graph_B = (int)B * 4 + n_groups;
    }
    // the pass counter restarts at 0: stale {value, flag} words of the previous generation must not match
    WMAR_CUDA_CHECK(cudaMemsetAsync(g->ws, 0, g->ws_bytes, s));
    cham_init_kernel<<<(unsigned)B, 64, 0, s>>>(g->d_call, g->seq, g->cfg.max_seq, g->pass);
    WMAR_LAUNCH_CHECK();
    const int64_t passes = p_max + steps - 1;
    for (int64_t i = 0; i < passes; i++) {
        WMAR_CUDA_CHECK(cudaGraphLaunch(g->exec, s));
        g_launches.fetch_add((uint64_t)g->launches_per_pass);
    }
    return WMAR_OK;
}

double wmar_cham_algorithmic_bytes(const wmar_cham *g, int64_t B, int n_groups, int64_t p_max, int64_t steps) {
    if (!g) return 0.0;
    const double d = g->d, V = g->cfg.vocab_size, L = g->cfg.n_layer, F = g->F, H = g->cfg.n_head, Hkv = g->cfg.n_kv_head;
    // bf16 parameters streamed once per pass: per layer wqkv + wo + w13 + w2, plus the output head
    const double P = L * ((H + 2 * Hkv) * 128.0 * d + H * 128.0 * d + 3.0 * F * d) + V * d;
    double kv = 0.0;  // bf16 K and V rows read per pass and row (upper bound: every row at the longest prompt)
    const double passes = (double)(p_max + steps - 1);
    for (int64_t i = 0; i < (int64_t)passes; i++) kv += 2.0 * L * Hkv * 128.0 * (double)(i + 1);
    return 2.0 * (P * passes + kv * (double)n_groups * (double)B);
}

int wmar_cham_launches_per_pass(const wmar_cham *g) { return g ? g->launches_per_pass : 0; }

int wmar_cham_select(const wmar_wm_params *wm, const wmar_sample_params *sp, const float *d_logits3, int64_t B, int64_t V,
                     int64_t image_token_lo, int64_t image_token_hi, float guidance_text, float guidance_image,
                     const int64_t *d_past_ids, int64_t t, int64_t past_stride, const float *d_noise, int64_t *d_out_ids,
                     float *d_mixed, void *stream) {
    WMAR_REQUIRE(sp != nullptr && d_logits3 != nullptr && d_out_ids != nullptr && d_mixed != nullptr && B > 0, "bad arguments");
    WMAR_REQUIRE(image_token_lo >= 0 && image_token_hi > image_token_lo && image_token_hi <= V, "bad image token range");
    const int W = (int)(image_token_hi - image_token_lo);
    wmar_wm_params wm_local{};
    wm_local.vocab_size = W;
    if (wm != nullptr && wm->d_table != nullptr) {
        WMAR_REQUIRE(wm->vocab_size == V, "watermark vocab_size != logits width");
        wm_local = *wm;
        wm_local.vocab_size = W;
    }
    SampleArgs sa;
    int rc = make_sample_args(&wm_local, sp, W, &sa);
    if (rc) return rc;
    sa.id_base = (int)image_token_lo;
    sa.table_V = (int)V;
    int *err = device_err_flag();
    WMAR_REQUIRE(err != nullptr, "cannot allocate the device error flag");
    cudaStream_t s = as_stream(stream);
    cham_mix_kernel<<<(unsigned)B, 256, 0, s>>>(d_logits3, (int)B, (int)V, (int)image_token_lo, W, guidance_text, guidance_image, d_mixed);
    WMAR_LAUNCH_CHECK();
    const size_t smem = sample_smem_bytes(W, sa.cand_cap);
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(cham_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cham_select_kernel<<<(unsigned)B, SAMPLE_THREADS, smem, s>>>(sa, d_mixed, (int)V, (int)image_token_lo, d_past_ids, t, past_stride,
                                                                d_noise, d_out_ids, err);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

}  // extern "C"
