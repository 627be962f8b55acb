// RAR decode engine: RAR.generate (deps/rar/modeling/rar.py:408-459) with classifier-free guidance, as a replayed CUDA
// graph of hand-written kernels -- no host work per token.
//
// Rows: R = 2B (cond rows 0..B-1, none-cond rows B..2B-1, rar.py:436-440), R <= 16 = the M of the skinny GEMM.
// The reference's step 0 runs the two prefix positions (cls token, condition) through a 2x2 causal mask and every later
// step only the newest position against the KV cache (rar.py:379-385).  With a cache a causal forward is the same as
// feeding the positions one at a time, so the engine runs 257 single-position passes i = 0..256:
//     x_i  = tok_i + pos_embed[i] + (i >= 1 ? target_aware_pos_embed[i+1] : 0)                   rar.py:342-371
//            tok_0 = cls_token, tok_1 = emb[cond], tok_{2+j} = emb[id_j]
//     c_i  = emb[cond] + timesteps_embeddings[i]                                                  rar.py:377
//     per block: mod = Linear(SiLU(c_i)) -> (sh1, sc1, g1, sh2, sc2, g2)                          rar.py:180
//                qkv = (LN(x)(1+sc1)+sh1) Wqkv^T + b ; q,k = LN_hd(q), LN_hd(k) ; append k,v     rar.py:90-118,181
//                x  += g1 * (softmax(q K^T / sqrt(hd)) V Wproj^T + b)
//                x  += g2 * (GELU((LN(x)(1+sc2)+sh2) W1^T + b1) W2^T + b2)                        rar.py:182
//     head:  (scale, shift) = Linear(SiLU(c_i)) ; logits = (LN_noaffine(x)(1+scale)+shift) Wlm^T + b  rar.py:131-134
//     pass i >= 1 samples image token i-1:  u + (c - u) * cfg -> +delta on green(ids) -> /T -> softmax -> multinomial
// KV cache: fp32 [layer][row16][head][seq+2][hd], one contiguous stream per (layer,row,head).
//
// adaLN hoisting: c_i depends only on (class, pass index), never on the tokens, so all modulation vectors of a
// generation are computed BEFORE the token loop by one M-large GEMM per layer over the (B + 1) distinct condition rows
// x (steps + 1) passes (rar.py:180,131-134,377), and a pass only gathers its 16 x 6d slice per layer (491 KB at XL)
// instead of streaming the 6 d^2 adaLN weights of every layer (1.27 of 3.80 GB per pass at XL).  WMAR_RAR_HOIST=0 keeps
// the per-pass GEMMs on their forked graph branch (round 1).
#include <algorithm>
#include <vector>

#include "gemm.cuh"
#include "sample.cuh"
#include "conv_igemm.cuh"   // the 128 x 64 tile 3xTF32 GEMM, used for the hoisted adaLN tables

using namespace wmar;

namespace wmar {
int make_sample_args(const wmar_wm_params *wm, const wmar_sample_params *sp, int V, SampleArgs *out);
int *device_err_flag();
}  // namespace wmar

namespace {

struct RarCall {
    SampleArgs sa;
    const int64_t *cond;   // [B] class ids
    const float *noise;    // [steps][B][V] or null
    int64_t *out_ids;      // [B][steps]
    float *out_logits;     // [steps][B][V] guided logits (before the watermark) or null
    float cfg_scale;
    int B, steps;
};

struct RarLayer {
    const float *n1_g, *n1_b, *wqkv, *bqkv, *qn_g, *qn_b, *kn_g, *kn_b, *wproj, *bproj, *n2_g, *n2_b, *w1, *b1, *w2, *b2,
        *wada, *bada;
};

constexpr int RAR_ATT_THREADS = 256;

}  // namespace

struct wmar_rar {
    wmar_rar_config cfg;
    int n_sms, d, hd, T;  // T = image_seq_len + 2 cache slots
    const float *cls_token, *emb, *pos_embed, *ta_pos_embed, *ts_embed, *whada, *bhada, *wlm, *blm;
    std::vector<RarLayer> layers;
    float *x, *csilu, *mod, *qkv, *y, *hbuf, *hmod, *logits, *guided, *kcache, *vcache, *ws;
    float2 *stats;
    unsigned *counters;
    int64_t *ids;   // [8][seq]: generated image tokens per cond row (the watermark's past_ids, rar.py:451)
    int *pos;       // device pass counter i
    RarCall *d_call, *h_call;
    cudaEvent_t call_done;
    bool call_pending;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    size_t graph_smem;
    int graph_B;
    int s_ada, s_qkv, s_proj, s_fc1, s_fc2, s_hada, s_lm;
    // the adaLN modulation GEMMs depend only on the conditioning, not on x: they run on a forked branch of the step
    // graph, off the critical path, with their own split-K workspace; mod holds one [16][6d] block per layer
    float *ws_side;
    size_t ws_bytes, ws_side_bytes;
    unsigned *counters_side;
    int launches_per_pass;
    // hoisted adaLN tables: ada_in [Mpad][d] = SiLU(c) of row m = pass * (B + 1) + r' (r' < B: cond row, r' == B: the
    // none-cond row); modtab [L][Mpad][6d], hmodtab [Mpad][2d]
    bool hoist;
    int m_pad;
    float *ada_in, *modtab, *hmodtab;
};

namespace {

__global__ void rar_init_kernel(int *pos) { *pos = 0; }
__global__ void rar_advance_kernel(int *pos) { *pos += 1; }

// Rows of the hoisted adaLN GEMM: m = pass * (B + 1) + r', SiLU(emb[cond row] + timesteps_embeddings[pass]) (rar.py:377,180);
// rows beyond (steps + 1) * (B + 1) are zero padding up to a multiple of the GEMM's 128-row tile.
__global__ void __launch_bounds__(256) rar_ada_input_kernel(const RarCall *cp, int codebook, int n_classes,
                                                            const float *__restrict__ emb, const float *__restrict__ ts_embed,
                                                            int d, float *__restrict__ ada_in) {
    const int m = blockIdx.x, B = cp->B, R1 = B + 1;
    const int pass = m / R1, r = m - pass * R1;
    const bool valid = pass <= cp->steps;
    long long cond_row = codebook + 1 + n_classes;
    if (valid && r < B) {
        long long cls = cp->cond[r];
        if (cls < 0 || cls >= n_classes) cls = 0;
        cond_row = codebook + 1 + cls;
    }
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const float4 e = *reinterpret_cast<const float4 *>(emb + (size_t)cond_row * d + c);
            const float4 ts = *reinterpret_cast<const float4 *>(ts_embed + (size_t)pass * d + c);
            float cc[4] = {e.x + ts.x, e.y + ts.y, e.z + ts.z, e.w + ts.w};
#pragma unroll
            for (int k = 0; k < 4; k++) cc[k] = cc[k] / (1.0f + expf(-cc[k]));  // SiLU, same expression as rar_embed_kernel
            o = make_float4(cc[0], cc[1], cc[2], cc[3]);
        }
        *reinterpret_cast<float4 *>(ada_in + (size_t)m * d + c) = o;
    }
}

// This pass's modulation vectors: mod[l][r][6d] <- modtab[l][pass * (B + 1) + r'][6d] for the 2B rows (r' = r for the cond
// rows, B for every none-cond row), and hmod[r][2d] likewise.  grid = (L + 1, 16).
__global__ void __launch_bounds__(256) rar_mod_gather_kernel(const RarCall *cp, const int *pos, int seq, int m_pad, int d,
                                                             int n_layer, const float *__restrict__ modtab,
                                                             const float *__restrict__ hmodtab, float *__restrict__ mod,
                                                             float *__restrict__ hmod) {
    const int l = blockIdx.x, r = blockIdx.y, i = *pos, B = cp->B;
    if (i > seq || r >= 2 * B) return;
    const size_t m = (size_t)i * (B + 1) + (r < B ? r : B);
    const int n = (l < n_layer ? 6 : 2) * d;
    const float *src = l < n_layer ? modtab + ((size_t)l * m_pad + m) * n : hmodtab + m * n;
    float *dst = l < n_layer ? mod + ((size_t)l * 16 + r) * n : hmod + (size_t)r * n;
    for (int c = threadIdx.x * 4; c < n; c += blockDim.x * 4)
        *reinterpret_cast<float4 *>(dst + c) = *reinterpret_cast<const float4 *>(src + c);
}

// x and SiLU(c) of pass i = *pos for the 16 rows (rows >= 2B zero), plus LN (mean, M2) partials of x per 64-col tile.
__global__ void __launch_bounds__(256) rar_embed_kernel(const RarCall *cp, const int *pos, const int64_t *ids, int seq,
                                                        int codebook, int n_classes, const float *__restrict__ cls_token,
                                                        const float *__restrict__ emb, const float *__restrict__ pos_embed,
                                                        const float *__restrict__ ta_pos, const float *__restrict__ ts_embed,
                                                        int d, float *__restrict__ x, float *__restrict__ csilu,
                                                        float2 *__restrict__ stats) {
    const int r = blockIdx.x, i = *pos, B = cp->B;
    if (i > seq) return;
    const bool valid = r < 2 * B;
    const int b = r < B ? r : r - B;
    long long cond_row = 0, tok_row = -1;
    if (valid) {
        long long cls = cp->cond[b];
        if (cls < 0 || cls >= n_classes) cls = 0;
        cond_row = r < B ? (codebook + 1 + cls) : (codebook + 1 + n_classes);  // rar.py:303-313
        if (i == 1) tok_row = cond_row;
        else if (i >= 2) {
            tok_row = ids[(size_t)b * seq + (i - 2)];
            if (tok_row < 0 || tok_row >= codebook) tok_row = 0;
        }
    }
    const int lane16 = threadIdx.x & 15;
    for (int c0 = (threadIdx.x >> 4) * 64; c0 < d; c0 += (blockDim.x >> 4) * 64) {
        const int c = c0 + lane16 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), cs = v;
        if (valid) {
            float4 t = tok_row < 0 ? *reinterpret_cast<const float4 *>(cls_token + c)
                                   : *reinterpret_cast<const float4 *>(emb + (size_t)tok_row * d + c);
            float4 p = *reinterpret_cast<const float4 *>(pos_embed + (size_t)i * d + c);
            v = make_float4(t.x + p.x, t.y + p.y, t.z + p.z, t.w + p.w);
            if (i >= 1) {
                float4 a = *reinterpret_cast<const float4 *>(ta_pos + (size_t)(i + 1) * d + c);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
            float4 e = *reinterpret_cast<const float4 *>(emb + (size_t)cond_row * d + c);
            float4 ts = *reinterpret_cast<const float4 *>(ts_embed + (size_t)i * d + c);
            float cc[4] = {e.x + ts.x, e.y + ts.y, e.z + ts.z, e.w + ts.w};
#pragma unroll
            for (int k = 0; k < 4; k++) cc[k] = cc[k] / (1.0f + expf(-cc[k]));  // SiLU
            cs = make_float4(cc[0], cc[1], cc[2], cc[3]);
        }
        *reinterpret_cast<float4 *>(x + (size_t)r * d + c) = v;
        *reinterpret_cast<float4 *>(csilu + (size_t)r * d + c) = cs;
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

// One CTA per (head, row): LayerNorm(q), LayerNorm(k_new) over head_dim, append k,v at slot i, attend over 0..i.
// qkv / the caches are NOT __restrict__: loads through read-only (__restrict__ const) pointers may be hoisted above
// griddepcontrol.wait, i.e. read q,k,v before the producing GEMM has written them (seen in gpt.cu).
__global__ void __launch_bounds__(RAR_ATT_THREADS) rar_attn_kernel(const float *qkv, int d, int H, int hd, int T,
                                                                   const float *__restrict__ qn_g, const float *__restrict__ qn_b,
                                                                   const float *__restrict__ kn_g, const float *__restrict__ kn_b,
                                                                   float *kcache, float *vcache,
                                                                   int layer, const int *pos, float *__restrict__ y) {
    __shared__ float sq[128];
    __shared__ float sc[1280];
    __shared__ float part[8][128];
    __shared__ float red[8];
    // programmatic dependent launch: the proj GEMM may start (and request its first weights) while this kernel runs
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int h = blockIdx.x, r = blockIdx.y, i = *pos;   // the pass counter was advanced at the end of the previous pass
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (((size_t)layer * 16 + r) * H + h) * (size_t)T * hd;
    float *K = kcache + base, *V = vcache + base;
    // K and V rows of EARLIER passes are final: ask L2 for them before waiting for this pass's q,k,v, so that the loads
    // after the wait are L2 hits (the HBM round trip overlaps the tail of the qkv GEMM, as in attn_decode_kernel)
    if (i > 0 && i < T) {
        const size_t bytes = (size_t)i * hd * sizeof(float);
        const char *kp = reinterpret_cast<const char *>(K), *vp = reinterpret_cast<const char *>(V);
        for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)RAR_ATT_THREADS * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + off));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // qkv comes from the previous kernel
    if (i >= T) return;
    const float *q = qkv + (size_t)r * 3 * d + h * hd;
    const float *kn = q + d, *vn = q + 2 * d;
    if (warp < 2) {  // warp 0: q, warp 1: new k  (nn.LayerNorm(hd, eps 1e-6), rar.py:82-83,103)
        const float *src = warp == 0 ? q : kn;
        const float *g = warp == 0 ? qn_g : kn_g, *bb = warp == 0 ? qn_b : kn_b;
        float s = 0.f;
        for (int c = lane; c < hd; c += 32) s += src[c];
        s = warp_sum(s);
        const float mean = s / (float)hd;
        float v2 = 0.f;
        for (int c = lane; c < hd; c += 32) { float dd = src[c] - mean; v2 += dd * dd; }
        v2 = warp_sum(v2);
        const float rstd = 1.0f / sqrtf(v2 / (float)hd + 1e-6f);
        for (int c = lane; c < hd; c += 32) {
            float o = (src[c] - mean) * rstd * g[c] + bb[c];
            if (warp == 0) sq[c] = o;
            else K[(size_t)i * hd + c] = o;
        }
    } else if (warp == 2) {
        for (int c = lane; c < hd; c += 32) V[(size_t)i * hd + c] = vn[c];
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)hd);
    const int nk = i + 1;
    const int hd4 = hd >> 2;
    // memory-latency bound: 8 keys per warp in flight at a time
    constexpr int BATCH = 8;
    const int lc = lane < hd4 ? lane : 0;
    const float4 q4 = *reinterpret_cast<const float4 *>(sq + 4 * lc);
    for (int j0 = warp; j0 < nk; j0 += 8 * BATCH) {
        float4 k4[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + 8 * u;
            k4[u] = (j < nk && lane < hd4) ? *reinterpret_cast<const float4 *>(K + (size_t)j * hd + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sd[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) sd[u] = q4.x * k4[u].x + q4.y * k4[u].y + q4.z * k4[u].z + q4.w * k4[u].w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < BATCH; u++) sd[u] += __shfl_xor_sync(0xffffffffu, sd[u], o);
        if (lane == 0) {
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                if (j0 + 8 * u < nk) sc[j0 + 8 * u] = sd[u] * scale;
        }
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < nk; j += RAR_ATT_THREADS) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += RAR_ATT_THREADS) {
        float e = expf(sc[j] - m);
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
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j0 = warp; j0 < nk; j0 += 8 * BATCH) {
        float4 v4[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + 8 * u;
            v4[u] = (j < nk && lane < hd4) ? *reinterpret_cast<const float4 *>(V + (size_t)j * hd + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + 8 * u;
            const float p = j < nk ? sc[j] * inv : 0.f;
            acc.x += p * v4[u].x; acc.y += p * v4[u].y; acc.z += p * v4[u].z; acc.w += p * v4[u].w;
        }
    }
    if (lane < hd4) *reinterpret_cast<float4 *>(&part[warp][4 * lane]) = acc;
    __syncthreads();
    if (tid < hd) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) o += part[w][tid];
        y[(size_t)r * d + h * hd + tid] = o;
    }
}

// Second form of the attention kernel (head_dim % 16 == 0: 32, 64, 80, 128): FOUR lanes per key, each lane owning F4
// consecutive float4 of the key row, eight keys per warp instruction and KB keys per lane group in flight, so that all
// (up to 256) cached keys of a pass are requested in ONE round trip -- and requested BEFORE griddepcontrol.wait: K rows of
// earlier passes are final, they land in registers while the qkv GEMM is still running.  The kernel above needs a whole
// warp (20 of 32 lanes busy at head_dim 80) and eight round trips per 512 keys.
template <int F4>
__global__ void __launch_bounds__(RAR_ATT_THREADS, 2) rar_attn4_kernel(const float *qkv, int d, int H, int T,
                                                                    const float *__restrict__ qn_g, const float *__restrict__ qn_b,
                                                                    const float *__restrict__ kn_g, const float *__restrict__ kn_b,
                                                                    float *kcache, float *vcache,
                                                                    int layer, const int *pos, float *__restrict__ y) {
    constexpr int HD = 16 * F4, KB = F4 <= 4 ? 4 : 2, GROUPS = RAR_ATT_THREADS / 4;   // 64 lane groups x KB keys per sweep
    __shared__ __align__(16) float sq[128];
    __shared__ float sc[1280];
    __shared__ __align__(16) float part[GROUPS / 2][HD];
    __shared__ float red[8];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int h = blockIdx.x, r = blockIdx.y, i = *pos;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid >> 2, sub = tid & 3;
    const size_t base = (((size_t)layer * 16 + r) * H + h) * (size_t)T * HD;
    float *K = kcache + base, *V = vcache + base;
    const int nprev = i < T ? i : 0;                   // cached keys 0 .. i-1 (this pass's own key comes after the wait)
    float4 k4[KB][F4];
#pragma unroll
    for (int u = 0; u < KB; u++) {
        const int j = grp + u * GROUPS;
#pragma unroll
        for (int f = 0; f < F4; f++)
            k4[u][f] = j < nprev ? *reinterpret_cast<const float4 *>(K + (size_t)j * HD + (sub * F4 + f) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (nprev > 0) {   // V (and the K rows past the first sweep): L2 prefetch, read after the softmax
        const size_t bytes = (size_t)nprev * HD * sizeof(float);
        const char *vp = reinterpret_cast<const char *>(V), *kp = reinterpret_cast<const char *>(K);
        for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)RAR_ATT_THREADS * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + off));
            if (off >= (size_t)GROUPS * KB * HD * sizeof(float)) asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + off));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // qkv comes from the previous kernel
    if (i >= T) return;
    const float *q = qkv + (size_t)r * 3 * d + h * HD;
    const float *kn = q + d, *vn = q + 2 * d;
    if (warp < 2) {  // warp 0: q, warp 1: new k  (nn.LayerNorm(hd, eps 1e-6), rar.py:82-83,103)
        const float *src = warp == 0 ? q : kn;
        const float *g = warp == 0 ? qn_g : kn_g, *bb = warp == 0 ? qn_b : kn_b;
        float s = 0.f;
        for (int c = lane; c < HD; c += 32) s += __ldcg(src + c);
        s = warp_sum(s);
        const float mean = s / (float)HD;
        float v2 = 0.f;
        for (int c = lane; c < HD; c += 32) { float dd = __ldcg(src + c) - mean; v2 += dd * dd; }
        v2 = warp_sum(v2);
        const float rstd = 1.0f / sqrtf(v2 / (float)HD + 1e-6f);
        for (int c = lane; c < HD; c += 32) {
            float o = (__ldcg(src + c) - mean) * rstd * g[c] + bb[c];
            if (warp == 0) sq[c] = o;
            else K[(size_t)i * HD + c] = o;
        }
    } else if (warp == 2) {
        for (int c = lane; c < HD; c += 32) V[(size_t)i * HD + c] = __ldcg(vn + c);
    }
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)HD);
    const int nk = i + 1;
    float4 q4[F4];
#pragma unroll
    for (int f = 0; f < F4; f++) q4[f] = *reinterpret_cast<const float4 *>(sq + (sub * F4 + f) * 4);
    for (int jb = 0; jb < nk; jb += GROUPS * KB) {      // block-uniform trip count
#pragma unroll
        for (int u = 0; u < KB; u++) {
            const int j = jb + grp + u * GROUPS;
            float sd = 0.f;
#pragma unroll
            for (int f = 0; f < F4; f++) {
                float4 kk;
                if (jb == 0 && j < nprev) kk = k4[u][f];
                else kk = j < nk ? *reinterpret_cast<const float4 *>(K + (size_t)j * HD + (sub * F4 + f) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                sd += q4[f].x * kk.x + q4[f].y * kk.y + q4[f].z * kk.z + q4[f].w * kk.w;
            }
            sd += __shfl_xor_sync(0xffffffffu, sd, 1);
            sd += __shfl_xor_sync(0xffffffffu, sd, 2);
            if (sub == 0 && j < nk) sc[j] = sd * scale;
        }
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < nk; j += RAR_ATT_THREADS) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; w++) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += RAR_ATT_THREADS) {
        float e = expf(sc[j] - m);
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
    float4 acc[F4];
#pragma unroll
    for (int f = 0; f < F4; f++) acc[f] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jb = 0; jb < nk; jb += GROUPS * KB) {
        float4 v4[KB][F4];
#pragma unroll
        for (int u = 0; u < KB; u++) {
            const int j = jb + grp + u * GROUPS;
#pragma unroll
            for (int f = 0; f < F4; f++)
                v4[u][f] = j < nk ? *reinterpret_cast<const float4 *>(V + (size_t)j * HD + (sub * F4 + f) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < KB; u++) {
            const int j = jb + grp + u * GROUPS;
            const float p = j < nk ? sc[j] * inv : 0.f;
#pragma unroll
            for (int f = 0; f < F4; f++) {
                acc[f].x += p * v4[u][f].x; acc[f].y += p * v4[u][f].y; acc[f].z += p * v4[u][f].z; acc[f].w += p * v4[u][f].w;
            }
        }
    }
    // 64 lane groups hold partial outputs over disjoint key sets: pairs of groups first (shuffle across lanes 4 apart),
    // then 32 partial rows through shared memory, summed in a fixed order
#pragma unroll
    for (int f = 0; f < F4; f++) {
        acc[f].x += __shfl_xor_sync(0xffffffffu, acc[f].x, 4); acc[f].y += __shfl_xor_sync(0xffffffffu, acc[f].y, 4);
        acc[f].z += __shfl_xor_sync(0xffffffffu, acc[f].z, 4); acc[f].w += __shfl_xor_sync(0xffffffffu, acc[f].w, 4);
    }
    if ((lane & 4) == 0) {
#pragma unroll
        for (int f = 0; f < F4; f++) *reinterpret_cast<float4 *>(&part[grp >> 1][(sub * F4 + f) * 4]) = acc[f];
    }
    __syncthreads();
    if (tid < HD) {
        float o = 0.f;
#pragma unroll 8
        for (int w = 0; w < GROUPS / 2; w++) o += part[w][tid];
        y[(size_t)r * d + h * HD + tid] = o;
    }
}

// guided = u + (c - u) * cfg (rar.py:441) for cond row b; a separate launch so that the sampler below may read the row
// through the read-only path.
__global__ void __launch_bounds__(256) rar_guide_kernel(const RarCall *cp, const float *__restrict__ logits,
                                                        float *__restrict__ guided, int V, const int *pos) {
    const int b = blockIdx.x, i = *pos;
    if (i < 1 || i > cp->steps) return;
    const int s = i - 1, B = cp->B;
    const float *lc = logits + (size_t)b * V, *lu = logits + (size_t)(B + b) * V;
    const float cfg = cp->cfg_scale;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
        const float u = lu[v];
        const float gl = u + (lc[v] - u) * cfg;
        guided[(size_t)b * V + v] = gl;
        if (cp->out_logits != nullptr) cp->out_logits[((size_t)s * B + b) * V + v] = gl;
    }
}

// watermark + sampler on the guided logits of cond row b; pass i samples image token i-1 (rar.py:446-454).
__global__ void __launch_bounds__(SAMPLE_THREADS, 1) rar_sample_kernel(const RarCall *cp, const float *__restrict__ guided,
                                                                        int64_t *ids, int seq, const int *pos, int *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.x, i = *pos;
    if (i < 1 || i > cp->steps) return;
    const int s = i - 1;
    const SampleArgs a = cp->sa;
    const int B = cp->B;
    const float *noise = cp->noise ? cp->noise + ((size_t)s * B + b) * a.V : nullptr;
    // past_ids of the reference = the image tokens generated so far (rar.py:451), length s
    int id = sample_row(a, guided + (size_t)b * a.V, ids + (size_t)b * seq, (long long)s, noise,
                        ((unsigned long long)s << 32) | (unsigned)b, err, smem_raw);
    if (threadIdx.x == 0) {
        ids[(size_t)b * seq + s] = id;
        cp->out_ids[(size_t)b * cp->steps + s] = id;
    }
}

void rar_free_graph(wmar_rar *g) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    g->exec = nullptr;
    g->graph = nullptr;
}

// events of the fork / join edges captured into the step graph (capture-time only objects, shared by all handles)
cudaEvent_t g_fork_event() {
    static cudaEvent_t e = nullptr;
    if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    return e;
}
cudaEvent_t g_layer_event(int l) {
    static std::vector<cudaEvent_t> ev;
    while ((int)ev.size() <= l) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ev.push_back(e);
    }
    return ev[l];
}

// `side`: second capturing stream for the forked adaLN branch (null: everything in order on `s`)
int rar_enqueue_pass(wmar_rar *g, int B, size_t sample_smem, cudaStream_t s, cudaStream_t side) {
    const wmar_rar_config &c = g->cfg;
    const int d = g->d, H = c.n_head, V = c.codebook_size, mlp = c.mlp;
    const int stat_tiles = d / 64;
    int rc, launches = 0;
    rar_embed_kernel<<<16, 256, 0, s>>>(g->d_call, g->pos, g->ids, c.image_seq_len, c.codebook_size, c.n_classes,
                                        g->cls_token, g->emb, g->pos_embed, g->ta_pos_embed, g->ts_embed, d, g->x, g->csilu,
                                        g->stats);
    WMAR_LAUNCH_CHECK();
    launches++;
    // ---- all adaLN modulations of this pass: gathered from the hoisted tables, or (WMAR_RAR_HOIST=0) computed by the
    // per-pass GEMMs on a forked branch (they read SiLU(c) only) ----
    const bool capturing = side != nullptr && !g->hoist;
    cudaStream_t ms = capturing ? side : s;
    const size_t mod_ld = (size_t)16 * 6 * d;
    if (capturing) {
        WMAR_CUDA_CHECK(cudaEventRecord(g_fork_event(), s));
        WMAR_CUDA_CHECK(cudaStreamWaitEvent(side, g_fork_event(), 0));
    }
    // flag-carrying split-K hand-off (gemm.cuh): epoch = the pass counter, salt = launch index within the pass
    static const bool ll_on = []() { const char *e = getenv("WMAR_LL"); return !(e && e[0] == '0'); }();
    WMAR_REQUIRE(5 * c.n_layer + 2 < 1024, "too many GEMM launches per pass for the hand-off flag");
    unsigned salt = 0;
    auto ll = [&](GemmArgs &q) { if (ll_on) { q.ll_epoch = g->pos; q.ll_salt = ++salt; } };
    if (g->hoist) {
        rar_mod_gather_kernel<<<dim3((unsigned)c.n_layer + 1, 16), 256, 0, s>>>(g->d_call, g->pos, c.image_seq_len, g->m_pad, d,
                                                                             c.n_layer, g->modtab, g->hmodtab, g->mod, g->hmod);
        WMAR_LAUNCH_CHECK();
        launches++;
    } else {
        for (int l = 0; l < c.n_layer; l++) {
            const RarLayer &L = g->layers[l];
            GemmArgs m{};
            m.ws = g->ws_side; m.counters = g->counters_side;
            m.X = g->csilu; m.ldx = d; m.W = L.wada; m.bias = L.bada; m.Y = g->mod + l * mod_ld; m.ldy = 6 * d; m.N = 6 * d; m.K = d;
            m.splits = g->s_ada;
            ll(m);
            if ((rc = launch_skinny_gemm(PRO_NONE, EPI_STORE, m, ms))) return rc;
            if (capturing) WMAR_CUDA_CHECK(cudaEventRecord(g_layer_event(l), side));
            launches++;
        }
        GemmArgs hm{};
        hm.ws = g->ws_side; hm.counters = g->counters_side;
        hm.X = g->csilu; hm.ldx = d; hm.W = g->whada; hm.bias = g->bhada; hm.Y = g->hmod; hm.ldy = 2 * d; hm.N = 2 * d; hm.K = d;
        hm.splits = g->s_hada;
        ll(hm);
        if ((rc = launch_skinny_gemm(PRO_NONE, EPI_STORE, hm, ms))) return rc;
        if (capturing) WMAR_CUDA_CHECK(cudaEventRecord(g_layer_event(c.n_layer), side));
        launches++;
    }
    for (int l = 0; l < c.n_layer; l++) {
        const RarLayer &L = g->layers[l];
        float *mod = g->mod + l * mod_ld;
        if (capturing) WMAR_CUDA_CHECK(cudaStreamWaitEvent(s, g_layer_event(l), 0));
        GemmArgs a{};
        a.ws = g->ws; a.counters = g->counters; a.eps = 1e-6f;
        a.X = g->x; a.ldx = d; a.W = L.wqkv; a.bias = L.bqkv; a.Y = g->qkv; a.ldy = 3 * d; a.N = 3 * d; a.K = d;
        a.splits = g->s_qkv; a.ln_g = L.n1_g; a.ln_b = L.n1_b; a.stats_in = g->stats; a.n_stat_tiles = stat_tiles;
        a.mod_shift = mod; a.mod_scale = mod + d; a.ld_mod = 6 * d;
        ll(a);
        if ((rc = launch_skinny_gemm(PRO_ADALN, EPI_STORE, a, s))) return rc;
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)H, (unsigned)(2 * B), 1);
            cfg.blockDim = dim3(RAR_ATT_THREADS, 1, 1);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            static const bool attn4 = []() { const char *e = getenv("WMAR_RAR_ATTN4"); return !(e && e[0] == '0'); }();
#define WMAR_RAR_ATTN4(F4)                                                                                                      \
    WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rar_attn4_kernel<F4>, (const float *)g->qkv, d, H, g->T, (const float *)L.qn_g,    \
                                       (const float *)L.qn_b, (const float *)L.kn_g, (const float *)L.kn_b, g->kcache,         \
                                       g->vcache, l, (const int *)g->pos, g->y))
            if (attn4 && g->hd == 32) WMAR_RAR_ATTN4(2);
            else if (attn4 && g->hd == 64) WMAR_RAR_ATTN4(4);
            else if (attn4 && g->hd == 80) WMAR_RAR_ATTN4(5);
            else if (attn4 && g->hd == 128) WMAR_RAR_ATTN4(8);
            else
                WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rar_attn_kernel, (const float *)g->qkv, d, H, g->hd, g->T,
                                                   (const float *)L.qn_g, (const float *)L.qn_b, (const float *)L.kn_g, (const float *)L.kn_b,
                                                   g->kcache, g->vcache, l, (const int *)g->pos, g->y));
#undef WMAR_RAR_ATTN4
        }
        GemmArgs p{};
        p.ws = g->ws; p.counters = g->counters;
        p.X = g->y; p.ldx = d; p.W = L.wproj; p.bias = L.bproj; p.Y = g->x; p.ldy = d; p.N = d; p.K = d;
        p.splits = g->s_proj; p.resid = g->x; p.ld_resid = d; p.gate = mod + 2 * d; p.ld_gate = 6 * d;
        p.stats_out = g->stats;
        ll(p);
        if ((rc = launch_skinny_gemm(PRO_NONE, EPI_GATE_RESID, p, s))) return rc;
        GemmArgs f{};
        f.ws = g->ws; f.counters = g->counters; f.eps = 1e-6f;
        f.X = g->x; f.ldx = d; f.W = L.w1; f.bias = L.b1; f.Y = g->hbuf; f.ldy = mlp; f.N = mlp; f.K = d;
        f.splits = g->s_fc1; f.ln_g = L.n2_g; f.ln_b = L.n2_b; f.stats_in = g->stats; f.n_stat_tiles = stat_tiles;
        f.mod_shift = mod + 3 * d; f.mod_scale = mod + 4 * d; f.ld_mod = 6 * d;
        ll(f);
        if ((rc = launch_skinny_gemm(PRO_ADALN, EPI_GELU, f, s))) return rc;
        GemmArgs o{};
        o.ws = g->ws; o.counters = g->counters;
        o.X = g->hbuf; o.ldx = mlp; o.W = L.w2; o.bias = L.b2; o.Y = g->x; o.ldy = d; o.N = d; o.K = mlp;
        o.splits = g->s_fc2; o.resid = g->x; o.ld_resid = d; o.gate = mod + 5 * d; o.ld_gate = 6 * d;
        o.stats_out = g->stats;
        ll(o);
        if ((rc = launch_skinny_gemm(PRO_NONE, EPI_GATE_RESID, o, s))) return rc;
        launches += 5;
    }
    if (capturing) WMAR_CUDA_CHECK(cudaStreamWaitEvent(s, g_layer_event(c.n_layer), 0));   // joins the forked branch
    GemmArgs lm{};
    lm.ws = g->ws; lm.counters = g->counters; lm.eps = 1e-6f;
    lm.X = g->x; lm.ldx = d; lm.W = g->wlm; lm.bias = g->blm; lm.Y = g->logits; lm.ldy = V; lm.N = V; lm.K = d;
    lm.splits = g->s_lm; lm.ln_g = nullptr; lm.ln_b = nullptr; lm.stats_in = g->stats; lm.n_stat_tiles = stat_tiles;
    lm.mod_scale = g->hmod; lm.mod_shift = g->hmod + d; lm.ld_mod = 2 * d;  // scale FIRST (rar.py:131)
    ll(lm);
        if ((rc = launch_skinny_gemm(PRO_ADALN, EPI_STORE, lm, s))) return rc;
    int *err = device_err_flag();
    WMAR_REQUIRE(err != nullptr, "cannot allocate the device error flag");
    rar_guide_kernel<<<B, 256, 0, s>>>(g->d_call, g->logits, g->guided, V, g->pos);
    WMAR_LAUNCH_CHECK();
    rar_sample_kernel<<<B, SAMPLE_THREADS, sample_smem, s>>>(g->d_call, g->guided, g->ids, c.image_seq_len, g->pos, err);
    WMAR_LAUNCH_CHECK();
    rar_advance_kernel<<<1, 1, 0, s>>>(g->pos);
    WMAR_LAUNCH_CHECK();
    launches += 4;
    g->launches_per_pass = launches;
    return WMAR_OK;
}

}  // namespace

extern "C" {

int wmar_rar_create(const wmar_rar_config *cfg, const void *const *d_weights, int n_weights, wmar_rar **out) {
    WMAR_REQUIRE(cfg != nullptr && d_weights != nullptr && out != nullptr, "NULL argument");
    WMAR_REQUIRE(cfg->hidden % 64 == 0 && cfg->hidden % cfg->n_head == 0, "hidden must be a multiple of 64 and of n_head");
    WMAR_REQUIRE(cfg->mlp % 64 == 0 && cfg->codebook_size % 64 == 0, "mlp % 64 == 0 and codebook % 64 == 0 required");
    WMAR_REQUIRE((2 * cfg->hidden) % CV_BN == 0 && cfg->hidden % CV_BK == 0, "hidden must tile the adaLN table GEMM");
    const int hd = cfg->hidden / cfg->n_head;
    WMAR_REQUIRE(hd % 4 == 0 && hd <= 128, "head_dim must be a multiple of 4 and <= 128");
    WMAR_REQUIRE(cfg->image_seq_len >= 1 && cfg->image_seq_len + 2 <= 1280, "image_seq_len out of range");
    WMAR_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch <= 8, "max_batch must be in [1,8] (2B rows <= 16)");
    WMAR_REQUIRE(n_weights == 5 + 18 * cfg->n_layer + 4, "weight table has the wrong number of entries");
    for (int i = 0; i < n_weights; i++) WMAR_REQUIRE(d_weights[i] != nullptr, "NULL weight pointer");
    wmar_rar *g = new (std::nothrow) wmar_rar();
    if (!g) return set_error(WMAR_ERR_NOMEM, "out of host memory%s%s");
    g->cfg = *cfg;
    g->d = cfg->hidden; g->hd = hd; g->T = cfg->image_seq_len + 2;
    int dev = 0;
    WMAR_CUDA_CHECK(cudaGetDevice(&dev));
    WMAR_CUDA_CHECK(cudaDeviceGetAttribute(&g->n_sms, cudaDevAttrMultiProcessorCount, dev));
    auto W = [&](int i) { return reinterpret_cast<const float *>(d_weights[i]); };
    g->cls_token = W(0); g->emb = W(1); g->pos_embed = W(2); g->ta_pos_embed = W(3); g->ts_embed = W(4);
    g->layers.resize(cfg->n_layer);
    for (int l = 0; l < cfg->n_layer; l++) {
        int b = 5 + 18 * l;
        g->layers[l] = RarLayer{W(b), W(b + 1), W(b + 2), W(b + 3), W(b + 4), W(b + 5), W(b + 6), W(b + 7), W(b + 8),
                                W(b + 9), W(b + 10), W(b + 11), W(b + 12), W(b + 13), W(b + 14), W(b + 15), W(b + 16), W(b + 17)};
    }
    int b = 5 + 18 * cfg->n_layer;
    g->whada = W(b); g->bhada = W(b + 1); g->wlm = W(b + 2); g->blm = W(b + 3);
    const int d = g->d, V = cfg->codebook_size, mlp = cfg->mlp;
    g->s_ada = pick_splits(6 * d, d, g->n_sms);
    g->s_qkv = pick_splits(3 * d, d, g->n_sms);
    g->s_proj = pick_splits(d, d, g->n_sms);
    g->s_fc1 = pick_splits(mlp, d, g->n_sms);
    g->s_fc2 = pick_splits(d, mlp, g->n_sms);
    g->s_hada = pick_splits(2 * d, d, g->n_sms);
    g->s_lm = pick_splits(V, d, g->n_sms);
    size_t ws_floats = 1;
    int max_tiles = 1;
    auto upd = [&](int N, int K, int S) {
        size_t n = gemm_ws_floats(N, K, S, g->n_sms);
        if (n > ws_floats) ws_floats = n;
        if (N / GEMM_NT > max_tiles) max_tiles = N / GEMM_NT;
    };
    upd(6 * d, d, g->s_ada); upd(3 * d, d, g->s_qkv); upd(d, d, g->s_proj); upd(mlp, d, g->s_fc1); upd(d, mlp, g->s_fc2);
    upd(2 * d, d, g->s_hada); upd(V, d, g->s_lm);
    const size_t kv_elems = (size_t)cfg->n_layer * 16 * d * g->T;
    WMAR_CUDA_CHECK(cudaMalloc(&g->x, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->csilu, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->mod, sizeof(float) * 16 * 6 * d * (size_t)cfg->n_layer));
    {
        const size_t side_ws = std::max(gemm_ws_floats(6 * d, d, g->s_ada, g->n_sms), gemm_ws_floats(2 * d, d, g->s_hada, g->n_sms));
        WMAR_CUDA_CHECK(cudaMalloc(&g->ws_side, sizeof(float) * (side_ws ? side_ws : 1)));
        g->ws_side_bytes = sizeof(float) * (side_ws ? side_ws : 1);
        WMAR_CUDA_CHECK(cudaMalloc(&g->counters_side, sizeof(unsigned) * (6 * d / GEMM_NT)));
        WMAR_CUDA_CHECK(cudaMemset(g->counters_side, 0, sizeof(unsigned) * (6 * d / GEMM_NT)));
    }
    WMAR_CUDA_CHECK(cudaMalloc(&g->qkv, sizeof(float) * 16 * 3 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->y, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->hbuf, sizeof(float) * 16 * mlp));
    WMAR_CUDA_CHECK(cudaMalloc(&g->hmod, sizeof(float) * 16 * 2 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->logits, sizeof(float) * 16 * V));
    WMAR_CUDA_CHECK(cudaMalloc(&g->guided, sizeof(float) * 8 * V));
    WMAR_CUDA_CHECK(cudaMalloc(&g->kcache, sizeof(float) * kv_elems));
    WMAR_CUDA_CHECK(cudaMalloc(&g->vcache, sizeof(float) * kv_elems));
    WMAR_CUDA_CHECK(cudaMalloc(&g->ws, sizeof(float) * ws_floats));
    g->ws_bytes = sizeof(float) * ws_floats;
    WMAR_CUDA_CHECK(cudaMalloc(&g->stats, sizeof(float2) * (d / 64) * 16));
    WMAR_CUDA_CHECK(cudaMalloc(&g->counters, sizeof(unsigned) * max_tiles));
    WMAR_CUDA_CHECK(cudaMalloc(&g->ids, sizeof(int64_t) * 8 * cfg->image_seq_len));
    WMAR_CUDA_CHECK(cudaMalloc(&g->pos, sizeof(int)));
    WMAR_CUDA_CHECK(cudaMalloc(&g->d_call, sizeof(RarCall)));
    WMAR_CUDA_CHECK(cudaMallocHost(&g->h_call, sizeof(RarCall)));
    WMAR_CUDA_CHECK(cudaEventCreateWithFlags(&g->call_done, cudaEventDisableTiming));
    g->call_pending = false;
    WMAR_REQUIRE(device_err_flag() != nullptr, "cannot allocate the device error flag");
    WMAR_CUDA_CHECK(cudaMemset(g->counters, 0, sizeof(unsigned) * max_tiles));
    WMAR_CUDA_CHECK(cudaMemset(g->x, 0, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->csilu, 0, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->mod, 0, sizeof(float) * 16 * 6 * d * (size_t)cfg->n_layer));
    WMAR_CUDA_CHECK(cudaMemset(g->qkv, 0, sizeof(float) * 16 * 3 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->y, 0, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->hbuf, 0, sizeof(float) * 16 * mlp));
    WMAR_CUDA_CHECK(cudaMemset(g->hmod, 0, sizeof(float) * 16 * 2 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->ids, 0, sizeof(int64_t) * 8 * cfg->image_seq_len));
    WMAR_CUDA_CHECK(cudaMemset(g->stats, 0, sizeof(float2) * (d / 64) * 16));
    g->graph = nullptr; g->exec = nullptr; g->graph_smem = 0; g->graph_B = 0;
    g->launches_per_pass = 6 * cfg->n_layer + 6;
    {
        const char *e = getenv("WMAR_RAR_HOIST");
        g->hoist = !(e && e[0] == '0');
        g->ada_in = g->modtab = g->hmodtab = nullptr;
        g->m_pad = ((cfg->image_seq_len + 1) * (cfg->max_batch + 1) + CV_BM - 1) / CV_BM * CV_BM;
        if (g->hoist) {
            WMAR_CUDA_CHECK(cudaMalloc(&g->ada_in, sizeof(float) * (size_t)g->m_pad * d));
            WMAR_CUDA_CHECK(cudaMalloc(&g->modtab, sizeof(float) * (size_t)cfg->n_layer * g->m_pad * 6 * d));
            WMAR_CUDA_CHECK(cudaMalloc(&g->hmodtab, sizeof(float) * (size_t)g->m_pad * 2 * d));
            const size_t smem = sizeof(float) * 2 * (CV_BM + CV_BN) * CV_LD;
            WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
    }
    // the zero-fills above went to the legacy default stream: finish them before the handle can be used from a
    // non-blocking stream (a second engine lane otherwise saw them land in the middle of its first generation)
    WMAR_CUDA_CHECK(cudaDeviceSynchronize());
    *out = g;
    return WMAR_OK;
}

void wmar_rar_destroy(wmar_rar *g) {
    if (!g) return;
    cudaDeviceSynchronize();
    rar_free_graph(g);
    cudaFree(g->ws_side); cudaFree(g->counters_side);
    cudaFree(g->ada_in); cudaFree(g->modtab); cudaFree(g->hmodtab);
    cudaFree(g->x); cudaFree(g->csilu); cudaFree(g->mod); cudaFree(g->qkv); cudaFree(g->y); cudaFree(g->hbuf);
    cudaFree(g->hmod); cudaFree(g->logits); cudaFree(g->guided); cudaFree(g->kcache); cudaFree(g->vcache); cudaFree(g->ws);
    cudaFree(g->stats); cudaFree(g->counters); cudaFree(g->ids); cudaFree(g->pos); cudaFree(g->d_call);
    cudaFreeHost(g->h_call);
    cudaEventDestroy(g->call_done);
    delete g;
}

int wmar_rar_sample(wmar_rar *g, const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_cond, int64_t B,
                    int64_t steps, float guidance_scale, const float *d_noise, int64_t *d_out_ids, float *d_out_logits,
                    void *stream) {
    WMAR_REQUIRE(g != nullptr && sp != nullptr && d_cond != nullptr && d_out_ids != nullptr, "NULL argument");
    WMAR_REQUIRE(B >= 1 && B <= g->cfg.max_batch, "batch exceeds max_batch");
    WMAR_REQUIRE(steps >= 1 && steps <= g->cfg.image_seq_len, "steps must be in [1, image_seq_len]");
    cudaStream_t s = as_stream(stream);
    wmar_wm_params wm_local{};
    wm_local.vocab_size = g->cfg.codebook_size;
    if (wm != nullptr && wm->d_table != nullptr) wm_local = *wm;
    SampleArgs sa;
    int rc = make_sample_args(&wm_local, sp, g->cfg.codebook_size, &sa);
    if (rc) return rc;
    const size_t smem = sample_smem_bytes(g->cfg.codebook_size, sa.cand_cap);
    if (g->call_pending) WMAR_CUDA_CHECK(cudaEventSynchronize(g->call_done));
    g->h_call->sa = sa;
    g->h_call->cond = d_cond;
    g->h_call->noise = d_noise;
    g->h_call->out_ids = d_out_ids;
    g->h_call->out_logits = d_out_logits;
    g->h_call->cfg_scale = guidance_scale;
    g->h_call->B = (int)B;
    g->h_call->steps = (int)steps;
    WMAR_CUDA_CHECK(cudaMemcpyAsync(g->d_call, g->h_call, sizeof(RarCall), cudaMemcpyHostToDevice, s));
    WMAR_CUDA_CHECK(cudaEventRecord(g->call_done, s));
    g->call_pending = true;
    if (g->exec == nullptr || g->graph_smem != smem || g->graph_B != (int)B) {
        rar_free_graph(g);
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(rar_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaStream_t cs;
        WMAR_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        WMAR_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        cudaStream_t side;
        WMAR_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        rc = rar_enqueue_pass(g, (int)B, smem, cs, side);
        cudaError_t e = cudaStreamEndCapture(cs, &g->graph);
        cudaStreamDestroy(cs);
        cudaStreamDestroy(side);
        if (rc) { if (g->graph) cudaGraphDestroy(g->graph); g->graph = nullptr; return rc; }
        if (e != cudaSuccess) return set_error(WMAR_ERR_CUDA, "cudaStreamEndCapture: %s%s", cudaGetErrorString(e));
        WMAR_CUDA_CHECK(cudaGraphInstantiate(&g->exec, g->graph, 0));
        g->graph_smem = smem;
        g->graph_B = (int)B;
    }
    // the pass counter restarts at 0: stale {value, flag} words of the previous generation must not match
    WMAR_CUDA_CHECK(cudaMemsetAsync(g->ws, 0, g->ws_bytes, s));
    WMAR_CUDA_CHECK(cudaMemsetAsync(g->ws_side, 0, g->ws_side_bytes, s));
    rar_init_kernel<<<1, 1, 0, s>>>(g->pos);
    WMAR_LAUNCH_CHECK();
    if (g->hoist) {
        // every modulation vector of this generation, before the token loop: one M-large 3xTF32 GEMM per layer over the
        // (steps + 1) x (B + 1) distinct (pass, condition) rows
        const int d = g->d;
        const int M = ((int)steps + 1) * ((int)B + 1), Mp = (M + CV_BM - 1) / CV_BM * CV_BM;
        WMAR_REQUIRE(Mp <= g->m_pad, "adaLN table too small");
        rar_ada_input_kernel<<<Mp, 256, 0, s>>>(g->d_call, g->cfg.codebook_size, g->cfg.n_classes, g->emb, g->ts_embed, d, g->ada_in);
        WMAR_LAUNCH_CHECK();
        const size_t smem = sizeof(float) * 2 * (CV_BM + CV_BN) * CV_LD;
        for (int l = 0; l <= g->cfg.n_layer; l++) {
            const bool head = l == g->cfg.n_layer;
            ConvArgs a{};
            a.in = g->ada_in; a.w = head ? g->whada : g->layers[l].wada; a.bias = head ? g->bhada : g->layers[l].bada;
            a.resid = nullptr;
            a.out = head ? g->hmodtab : g->modtab + (size_t)l * g->m_pad * 6 * d;
            a.B = 1; a.Hs = 1; a.Ws = Mp; a.Cin = d; a.Ho = 1; a.Wo = Mp; a.Cout = (head ? 2 : 6) * d; a.Cout_pad = a.Cout;
            a.ks = 1; a.stride = 1; a.pad = 0; a.up = 0; a.out_scale = 1.f;
            conv_igemm_kernel<1><<<dim3((unsigned)(Mp / CV_BM), (unsigned)(a.Cout / CV_BN)), CV_THREADS, smem, s>>>(a);
            WMAR_LAUNCH_CHECK();
        }
    }
    for (int64_t i = 0; i <= steps; i++) {  // pass 0 = cls token (fills the cache only), pass i >= 1 samples token i-1
        WMAR_CUDA_CHECK(cudaGraphLaunch(g->exec, s));
        g_launches.fetch_add((uint64_t)g->launches_per_pass);
    }
    return WMAR_OK;
}

double wmar_rar_algorithmic_bytes(const wmar_rar *g, int64_t B, int64_t steps) {
    if (!g) return 0.0;
    const double d = g->d, V = g->cfg.codebook_size, L = g->cfg.n_layer, mlp = g->cfg.mlp;
    // dense parameters streamed once per pass: per block 4 d^2 + 2 d mlp (+ 6 d^2 adaLN when not hoisted)
    // + biases / norms, head V d (+ 2 d^2 adaLN when not hoisted)
    const double ada = g->hoist ? 0.0 : 1.0;
    const double P = L * ((4.0 + 6.0 * ada) * d * d + 2.0 * d * mlp + 9.0 * d + mlp + 6.0 * d * ada + 4.0 * d + 4.0 * (d / g->cfg.n_head)) +
                     ada * (2.0 * d * d + 2.0 * d) + V * d + V;
    double kv = 0.0;  // per row: i+1 keys read at pass i, one appended
    for (int64_t i = 0; i <= steps; i++) kv += 2.0 * L * d * (double)(i + 1) + 2.0 * L * d;
    double hoisted = 0.0;
    if (g->hoist) {
        // once per generation: the adaLN weights read once, the tables written once; per pass: 2B rows x (6 d L + 2 d) read
        const double rows = (double)(steps + 1) * (double)(B + 1);
        hoisted = (6.0 * d * d + 6.0 * d) * L + 2.0 * d * d + 2.0 * d + rows * (6.0 * d * L + 2.0 * d) +
                  (double)(steps + 1) * 2.0 * (double)B * (6.0 * d * L + 2.0 * d);
    }
    return 4.0 * (P * (double)(steps + 1) + kv * 2.0 * (double)B + hoisted);
}

}  // extern "C"
