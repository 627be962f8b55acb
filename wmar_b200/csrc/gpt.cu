// Taming minGPT decode engine: the whole sample_with_past loop (mingpt.py:326-368) as a replayed CUDA graph of
// hand-written kernels, no host work per token.
//
// Per step (one graph replay):   for each of L layers
//     qkv   = LN1(x) Wqkv^T + b              skinny GEMM, LayerNorm fused in the prologue      (mingpt.py:69-78,117)
//     y     = softmax(q K^T / sqrt(hd)) V     decode attention, appends k,v to the cache in place (:80-90)
//     x    += y Wproj^T + b                   skinny GEMM, residual epilogue + LN statistics     (:93-94,119)
//     m     = GELU(LN2(x) W1^T + b1)          skinny GEMM, LN prologue, erf-GELU epilogue        (:105-108)
//     x    += m W2^T + b2                     skinny GEMM, residual epilogue + LN statistics     (:109,120)
//   logits = LNf(x) Whead^T                   skinny GEMM                                        (:206-207)
//   id     = sample(logits)                   watermark bias + /T + top-k + top-p + multinomial  (:349-363)
//   x      = tok_emb[id] + pos_emb[t+1]       embedding for the next step                        (:186-200)
// Alternative path (WMAR_STEP=fused, gpt_fused.cuh): per layer  attention block kernel (tcgen05 + TMA + 8-CTA clusters)
// -> residual reduce -> MLP block kernel -> residual reduce, chained with programmatic dependent launch.  Same token
// ids; measured 4 % slower than the per-GEMM path on B200 in round 1 (profiles/r01_fused_summary.md), so not the default.
// KV cache layout in HBM: K,V fp32 [layer][row(16)][head][block_size][head_dim] -- one contiguous stream per
// (layer,row,head), so the attention kernel reads 2*(t+1)*hd*4 bytes per CTA with fully coalesced 256 B rows.
#include <algorithm>
#include <vector>

#include "gemm.cuh"
#include "gpt_fused.cuh"
#include "pstep_host.h"
#include "sample.cuh"

using namespace wmar;

namespace wmar {
int make_sample_args(const wmar_wm_params *wm, const wmar_sample_params *sp, int V, SampleArgs *out);
int *device_err_flag();
}  // namespace wmar

namespace {

// Per-call parameters read by the kernels of the (pre-captured) step graph.
struct CallParams {
    SampleArgs sa;
    const int64_t *cond;
    const float *noise;   // [steps][B][V] or null
    int64_t *out_codes;   // [B][steps]
    float *out_logits;    // [steps][B][V] or null
    int B, steps;
};

struct Layer {
    const float *ln1_g, *ln1_b, *wqkv, *bqkv, *wproj, *bproj, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
};

}  // namespace

struct wmar_gpt {
    wmar_gpt_config cfg;
    int n_sms;
    const float *tok_emb, *pos_emb, *lnf_g, *lnf_b, *head;
    std::vector<Layer> layers;
    // device scratch
    float *x, *qkv, *y, *hbuf, *logits, *kcache, *vcache, *ws;
    size_t ws_bytes;
    float2 *stats;
    unsigned *counters;
    int64_t *seq;      // [16][block_size + 1]: conditioning token then the generated ids
    int *step;         // device step counter
    CallParams *d_call;
    CallParams *h_call;  // pinned staging
    cudaEvent_t call_done;
    bool call_pending;
    // cached step graph
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    size_t graph_smem;
    int graph_B;
    int splits_qkv, splits_proj, splits_fc1, splits_fc2, splits_head;
    int launches_per_step;
    // fused block kernels (gpt_fused.cuh): two kernels + two reductions per layer instead of five GEMMs + attention
    bool fused;
    unsigned long long *d_trace;   // probe only (WMAR_STEP_TRACE=<step>): [3][48] stamps: attention + MLP block of layer n_layer/2, attention block of the next layer
    int trace_step;
    float *ws2;        // second partial buffer (MLP block)
    float *hpart;      // K-split partials of the fc1 tiles
    unsigned *hflag;   // [n_layer][4d/128] arrival counters
    // persistent step kernel (pstep.cuh): opt-in (WMAR_STEP=pstep, measured slower than the per-GEMM graph), one launch per token step
    PstepState *pstep;
    int *ticket;       // arrival counter of gpt_tail_kernel (self-resetting)
};

namespace {

constexpr int ATT_THREADS = 256;

// seq[b][0] = cond[b], step = 0
__global__ void init_call_kernel(const CallParams *cp, int64_t *seq, int seq_ld, int *step) {
    const int b = threadIdx.x;
    if (b < cp->B) seq[(size_t)b * seq_ld] = cp->cond[b];
    if (b == 0) *step = 0;
}

// x[b] = tok_emb[seq[b][t]] + pos_emb[t]  (t = *step), plus (mean, M2) partials per 64-column tile; rows >= B are zero
__global__ void __launch_bounds__(256) embed_kernel(const CallParams *cp, const int64_t *seq, int seq_ld,
                                                    const int *step, const float *__restrict__ tok_emb,
                                                    const float *__restrict__ pos_emb, int d, int block_size, int V,
                                                    float *__restrict__ x, float2 *__restrict__ stats) {
    const int b = blockIdx.x, t = *step;
    if (t >= block_size) return;  // after the last token there is no next position
    const bool valid = b < cp->B;
    long long id = valid ? seq[(size_t)b * seq_ld + t] : 0;
    if (id < 0 || id >= V) id = 0;
    const int lane16 = threadIdx.x & 15;
    for (int c0 = (threadIdx.x >> 4) * 64; c0 < d; c0 += (blockDim.x >> 4) * 64) {
        // 16 lanes own one 64-column tile
        const int c = c0 + lane16 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            float4 e = *reinterpret_cast<const float4 *>(tok_emb + (size_t)id * d + c);
            float4 p = *reinterpret_cast<const float4 *>(pos_emb + (size_t)t * d + c);
            v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
        }
        *reinterpret_cast<float4 *>(x + (size_t)b * d + c) = v;
        float s = v.x + v.y + v.z + v.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 64.0f);
        float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane16 == 0) stats[(c0 / 64) * 16 + b] = make_float2(mean, q);
    }
}

// One CTA per (head, row).  Appends this step's k,v to the cache and attends over keys 0..t.
template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const float *qkv, int d, int H, int T,
                                                                  float *kcache, float *vcache,
                                                                  int layer, const int *step, float *__restrict__ y) {
    // NOTE: qkv / the caches are NOT __restrict__: the compiler may hoist loads through read-only (__restrict__ const)
    // pointers above griddepcontrol.wait, i.e. read q,k,v before the producing GEMM has written them.
    static_assert(HD == 64, "Taming head_dim");
    __shared__ float sc[1024];                      // scores / probabilities (T <= 1024)
    __shared__ __align__(16) float part[ATT_THREADS / 16][HD];
    __shared__ float red[ATT_THREADS / 32];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int h = blockIdx.x, b = blockIdx.y, t = *step;   // the step counter was advanced before the previous kernel
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = tid & 15, grp = tid >> 4;       // 16 lanes per key, ATT_THREADS/16 keys in flight
    constexpr int GROUPS = ATT_THREADS / 16;
    constexpr int BATCH = 8;                        // keys per group in flight at a time (the phase is latency bound)
    const float *q = qkv + (size_t)b * 3 * d + h * HD;
    const float *kn = q + d, *vn = q + 2 * d;
    const size_t base = (((size_t)layer * 16 + b) * H + h) * (size_t)T * HD;
    float *K = kcache + base, *Vc = vcache + base;
    const int nk = t + 1;
    // K and V rows of EARLIER steps are final: the first 128 keys of both are fetched before waiting for this step's
    // q,k,v (two of the four memory round trips of the kernel overlap the tail of the qkv GEMM)
    float4 k4a[BATCH], v4a[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; u++) {
        const int j = grp + u * GROUPS;
        k4a[u] = j < t ? *reinterpret_cast<const float4 *>(K + (size_t)j * HD + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        v4a[u] = j < t ? *reinterpret_cast<const float4 *>(Vc + (size_t)j * HD + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // keys past the register batch: ask L2 for their lines now, so that the in-loop loads after the wait are L2 hits
    if (t > GROUPS * BATCH) {
        const size_t bytes = (size_t)(t - GROUPS * BATCH) * HD * sizeof(float);
        const char *kp = reinterpret_cast<const char *>(K + (size_t)GROUPS * BATCH * HD);
        const char *vp = reinterpret_cast<const char *>(Vc + (size_t)GROUPS * BATCH * HD);
        for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)ATT_THREADS * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + off));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // qkv comes from the previous kernel
    const float4 q4 = __ldcg(reinterpret_cast<const float4 *>(q + 4 * sub));
    const float4 kn4 = __ldcg(reinterpret_cast<const float4 *>(kn + 4 * sub)), vn4 = __ldcg(reinterpret_cast<const float4 *>(vn + 4 * sub));
    if (tid < 16) *reinterpret_cast<float4 *>(K + (size_t)t * HD + 4 * tid) = kn4;             // append for later steps
    else if (tid < 32) *reinterpret_cast<float4 *>(Vc + (size_t)t * HD + 4 * (tid - 16)) = vn4;
    const float scale = 1.0f / sqrtf((float)HD);
    for (int jb = 0; jb < nk; jb += GROUPS * BATCH) {   // warp-uniform trip count: the shuffles below use the full mask
        const int j0 = jb + grp;
        float4 k4[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + u * GROUPS;
            if (j == t) k4[u] = kn4;                                    // this step's own key (not read back from the cache)
            else if (jb == 0) k4[u] = k4a[u];
            else k4[u] = j < t ? *reinterpret_cast<const float4 *>(K + (size_t)j * HD + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sd[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) sd[u] = q4.x * k4[u].x + q4.y * k4[u].y + q4.z * k4[u].z + q4.w * k4[u].w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int u = 0; u < BATCH; u++) sd[u] += __shfl_xor_sync(0xffffffffu, sd[u], o);
        if (sub == 0) {
#pragma unroll
            for (int u = 0; u < BATCH; u++)
                if (j0 + u * GROUPS < nk) sc[j0 + u * GROUPS] = sd[u] * scale;
        }
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < nk; j += ATT_THREADS) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < ATT_THREADS / 32; w++) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += ATT_THREADS) {
        float e = expf(sc[j] - m);
        sc[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_THREADS / 32; w++) sum += red[w];
    const float inv = 1.0f / sum;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jb = 0; jb < nk; jb += GROUPS * BATCH) {
        const int j0 = jb + grp;
        float4 v4[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + u * GROUPS;
            if (j == t) v4[u] = vn4;
            else if (jb == 0) v4[u] = v4a[u];
            else v4[u] = j < t ? *reinterpret_cast<const float4 *>(Vc + (size_t)j * HD + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
            const int j = j0 + u * GROUPS;
            const float p = j < nk ? sc[j] * inv : 0.f;
            acc.x += p * v4[u].x; acc.y += p * v4[u].y; acc.z += p * v4[u].z; acc.w += p * v4[u].w;
        }
    }
    *reinterpret_cast<float4 *>(&part[grp][4 * sub]) = acc;
    __syncthreads();
    if (tid < HD) {
        float o = 0.f;
#pragma unroll
        for (int gI = 0; gI < GROUPS; gI++) o += part[gI][tid];
        y[(size_t)b * d + h * HD + tid] = o;
    }
}

// Second form of the attention kernel: FOUR lanes per key (each lane 4 consecutive float4 of the 64-wide row), 64 lane
// groups x 4 keys = all 256 cached keys of the Taming sequence requested in ONE sweep, before griddepcontrol.wait (K into
// registers, V as an L2 prefetch), and two shuffle steps per score instead of four.  Opt-in (WMAR_ATTN4=1): slower here.
__global__ void __launch_bounds__(ATT_THREADS, 2) attn_decode4_kernel(const float *qkv, int d, int H, int T, float *kcache,
                                                                      float *vcache, int layer, const int *step,
                                                                      float *__restrict__ y) {
    constexpr int HD = 64, F4 = 4, KB = 4, GROUPS = ATT_THREADS / 4;
    __shared__ float sc[1024];
    __shared__ __align__(16) float part[GROUPS / 2][HD];
    __shared__ float red[ATT_THREADS / 32];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int h = blockIdx.x, b = blockIdx.y, t = *step;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = tid >> 2, sub = tid & 3;
    const float *q = qkv + (size_t)b * 3 * d + h * HD;
    const float *kn = q + d, *vn = q + 2 * d;
    const size_t base = (((size_t)layer * 16 + b) * H + h) * (size_t)T * HD;
    float *K = kcache + base, *Vc = vcache + base;
    const int nk = t + 1;
    float4 k4[KB][F4];
#pragma unroll
    for (int u = 0; u < KB; u++) {
        const int j = grp + u * GROUPS;
#pragma unroll
        for (int f = 0; f < F4; f++)
            k4[u][f] = j < t ? *reinterpret_cast<const float4 *>(K + (size_t)j * HD + (sub * F4 + f) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (t > 0) {
        const size_t bytes = (size_t)t * HD * sizeof(float);
        const char *vp = reinterpret_cast<const char *>(Vc), *kp = reinterpret_cast<const char *>(K);
        for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)ATT_THREADS * 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + off));
            if (off >= (size_t)GROUPS * KB * HD * sizeof(float)) asm volatile("prefetch.global.L2 [%0];" ::"l"(kp + off));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // qkv comes from the previous kernel
    float4 q4[F4];
#pragma unroll
    for (int f = 0; f < F4; f++) q4[f] = __ldcg(reinterpret_cast<const float4 *>(q + (sub * F4 + f) * 4));
    if (tid < 16) *reinterpret_cast<float4 *>(K + (size_t)t * HD + 4 * tid) = __ldcg(reinterpret_cast<const float4 *>(kn + 4 * tid));
    else if (tid < 32) *reinterpret_cast<float4 *>(Vc + (size_t)t * HD + 4 * (tid - 16)) = __ldcg(reinterpret_cast<const float4 *>(vn + 4 * (tid - 16)));
    const float scale = 1.0f / sqrtf((float)HD);
    for (int jb = 0; jb < nk; jb += GROUPS * KB) {
#pragma unroll
        for (int u = 0; u < KB; u++) {
            const int j = jb + grp + u * GROUPS;
            float sd = 0.f;
#pragma unroll
            for (int f = 0; f < F4; f++) {
                float4 kk;
                if (j == t) kk = __ldcg(reinterpret_cast<const float4 *>(kn + (sub * F4 + f) * 4));   // this step's own key
                else if (jb == 0) kk = k4[u][f];
                else kk = j < t ? *reinterpret_cast<const float4 *>(K + (size_t)j * HD + (sub * F4 + f) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                sd += q4[f].x * kk.x + q4[f].y * kk.y + q4[f].z * kk.z + q4[f].w * kk.w;
            }
            sd += __shfl_xor_sync(0xffffffffu, sd, 1);
            sd += __shfl_xor_sync(0xffffffffu, sd, 2);
            if (sub == 0 && j < nk) sc[j] = sd * scale;
        }
    }
    __syncthreads();
    float m = -INFINITY;
    for (int j = tid; j < nk; j += ATT_THREADS) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < ATT_THREADS / 32; w++) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = tid; j < nk; j += ATT_THREADS) {
        float e = expf(sc[j] - m);
        sc[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_THREADS / 32; w++) sum += red[w];
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
            for (int f = 0; f < F4; f++) {
                if (j == t) v4[u][f] = __ldcg(reinterpret_cast<const float4 *>(vn + (sub * F4 + f) * 4));
                else v4[u][f] = j < t ? *reinterpret_cast<const float4 *>(Vc + (size_t)j * HD + (sub * F4 + f) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
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
        y[(size_t)b * d + h * HD + tid] = o;
    }
}

// lm_head epilogue: watermark + sampler for row b; writes the id into seq / out_codes and (optionally) the raw logits
__global__ void __launch_bounds__(SAMPLE_THREADS, 1) gpt_sample_kernel(const CallParams *cp, const float *__restrict__ logits,
                                                                        int64_t *seq, int seq_ld, const int *step, int *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.x, t = *step;
    const SampleArgs a = cp->sa;
    const float *row = logits + (size_t)b * a.V;
    if (cp->out_logits != nullptr) {
        float *dst = cp->out_logits + ((size_t)t * cp->B + b) * a.V;
        for (int v = threadIdx.x; v < a.V; v += SAMPLE_THREADS) dst[v] = row[v];
    }
    const float *noise = cp->noise ? cp->noise + ((size_t)t * cp->B + b) * a.V : nullptr;
    // past_ids of the reference = [cond, ids...] (mingpt.py:328,350), length t + 1
    int id = sample_row(a, row, seq + (size_t)b * seq_ld, (long long)t + 1, noise,
                        ((unsigned long long)t << 32) | (unsigned)b, err, smem_raw);
    if (threadIdx.x == 0) {
        seq[(size_t)b * seq_ld + t + 1] = id;
        cp->out_codes[(size_t)b * cp->steps + t] = id;
    }
}

__global__ void advance_step_kernel(int *step) { *step += 1; }

// The whole tail of a token step in ONE launch (default; WMAR_TAIL=0 keeps sample / advance / embed as three kernels):
// CTA b samples row b (watermark + warpers + multinomial, as gpt_sample_kernel), then embeds the sampled token for the
// NEXT position (x[b] = tok_emb[id] + pos_emb[t + 1] and the LayerNorm partials, as embed_kernel) -- a row only needs
// its own id -- and the last CTA to finish advances the step counter (ticket).  grid = 16: rows >= B are zeroed.
__global__ void __launch_bounds__(SAMPLE_THREADS, 1) gpt_tail_kernel(const CallParams *cp, const float *__restrict__ logits,
                                                                      int64_t *seq, int seq_ld, int *step, int *ticket, int *err,
                                                                      const float *__restrict__ tok_emb,
                                                                      const float *__restrict__ pos_emb, int d, int block_size,
                                                                      float *__restrict__ x, float2 *__restrict__ stats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ int s_id;
    const int b = blockIdx.x, t = *step;
    const SampleArgs a = cp->sa;
    const bool valid = b < cp->B;
    if (valid) {
        const float *row = logits + (size_t)b * a.V;
        if (cp->out_logits != nullptr) {
            float *dst = cp->out_logits + ((size_t)t * cp->B + b) * a.V;
            for (int v = threadIdx.x; v < a.V; v += SAMPLE_THREADS) dst[v] = row[v];
        }
        const float *noise = cp->noise ? cp->noise + ((size_t)t * cp->B + b) * a.V : nullptr;
        const int id = sample_row(a, row, seq + (size_t)b * seq_ld, (long long)t + 1, noise,
                                  ((unsigned long long)t << 32) | (unsigned)b, err, smem_raw);
        if (threadIdx.x == 0) {
            seq[(size_t)b * seq_ld + t + 1] = id;
            cp->out_codes[(size_t)b * cp->steps + t] = id;
            s_id = id;
        }
    }
    __syncthreads();
    const int tn = t + 1;
    if (tn < block_size) {       // after the last token there is no next position
        long long id = valid ? s_id : 0;
        if (id < 0 || id >= a.V) id = 0;
        const int lane16 = threadIdx.x & 15;
        for (int c0 = (threadIdx.x >> 4) * 64; c0 < d; c0 += (SAMPLE_THREADS >> 4) * 64) {
            const int c = c0 + lane16 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float4 e = *reinterpret_cast<const float4 *>(tok_emb + (size_t)id * d + c);
                const float4 p = *reinterpret_cast<const float4 *>(pos_emb + (size_t)tn * d + c);
                v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
            }
            *reinterpret_cast<float4 *>(x + (size_t)b * d + c) = v;
            float s = v.x + v.y + v.z + v.w;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * (1.0f / 64.0f);
            const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
            float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            if (lane16 == 0) stats[(c0 / 64) * 16 + b] = make_float2(mean, q);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {   // every CTA has read `t`: the step may advance
            *ticket = 0;
            *step = t + 1;
        }
    }
}

int free_graph(wmar_gpt *g) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    g->exec = nullptr;
    g->graph = nullptr;
    return 0;
}

template <int KIND>
int launch_fused_t(const CUtensorMap &m1, const CUtensorMap &m2, const FusedArgs &a, int groups, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(fused_block_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SM_ALLOC));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(KIND == FB_ATT ? FB_ATT_CS : FB_MLP_CS), (unsigned)groups, 1);
    cfg.blockDim = dim3(FB_THREADS, 1, 1);
    cfg.dynamicSmemBytes = FB_SM_ALLOC;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;   // attention block only: the MLP block exchanges through L2
    attr[1].val.clusterDim.x = FB_ATT_CS;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = KIND == FB_ATT ? 2 : 1;
    WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fused_block_kernel<KIND>, m1, m2, a));
    g_launches.fetch_add(1);
    return WMAR_OK;
}

int launch_fused(int kind, const CUtensorMap &m1, const CUtensorMap &m2, const FusedArgs &a, int groups, cudaStream_t s) {
    return kind == FB_ATT ? launch_fused_t<FB_ATT>(m1, m2, a, groups, s) : launch_fused_t<FB_MLP>(m1, m2, a, groups, s);
}

int launch_resid_reduce(const float *ws, int P, const float *bias, float *x, float2 *stats, int d, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(d / 32), 1, 1);
    cfg.blockDim = dim3(RR_THREADS, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, resid_reduce_kernel, ws, P, bias, x, stats, d));
    g_launches.fetch_add(1);
    return WMAR_OK;
}

// The fused block kernels tile this model (and the device can co-schedule their clusters)?
bool fused_eligible(const wmar_gpt_config &c) {
    const int d = c.n_embd, H = c.n_head;
    if (!tc_available() || d % 128 != 0 || d / H != 64 || H % 2 != 0 || c.block_size > FB_ATT_T) return false;
    const int C1 = d / 32, T2 = d / 128;
    if (C1 < FB_ATT_CS || (C1 + FB_ATT_CS - 1) / FB_ATT_CS > FB_ATT_B1_MAX) return false;
    if ((C1 + FB_MLP_CS - 1) / FB_MLP_CS > FB_MLP_B1_MAX || T2 < FB_MLP_CS) return false;
    int dev = 0, max_smem = 0, n_sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    // the CTAs of an fc1 tile wait for each other: the whole MLP grid (one CTA per SM) must be resident at once
    if (cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || FB_MLP_CS * (4 * d / 128) > n_sms) return false;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    if (max_smem < FB_SM_ALLOC) return false;
    if (cudaFuncSetAttribute(fused_block_kernel<FB_ATT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SM_ALLOC) != cudaSuccess ||
        cudaFuncSetAttribute(fused_block_kernel<FB_MLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SM_ALLOC) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(FB_ATT_CS, (unsigned)(H / 2), 1);
    cfg.blockDim = dim3(FB_THREADS, 1, 1);
    cfg.dynamicSmemBytes = FB_SM_ALLOC;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = FB_ATT_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, fused_block_kernel<FB_ATT>, &cfg) != cudaSuccess || n_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    if (getenv("WMAR_STEP_TRACE")) {
        int n_mlp = 0;
        cfg.gridDim = dim3(FB_MLP_CS, (unsigned)(4 * d / 128), 1);
        attr[0].val.clusterDim.x = FB_MLP_CS;
        cudaOccupancyMaxActiveClusters(&n_mlp, fused_block_kernel<FB_MLP>, &cfg);
        fprintf(stderr, "[wmar] max co-resident clusters: attention (x%d) %d of %d, mlp (x%d) %d of %d\n", FB_ATT_CS, n_clusters,
                H / 2, FB_MLP_CS, n_mlp, 4 * d / 128);
    }
    return true;
}

// Enqueue the kernels of one decode step on `s` (used both under stream capture and for direct launches).
int enqueue_step(wmar_gpt *g, int B, size_t sample_smem, cudaStream_t s) {
    const wmar_gpt_config &c = g->cfg;
    const int d = c.n_embd, H = c.n_head, V = c.vocab_size;
    const int stat_tiles = d / 64;
    int rc;
    int launches = 0;
    if (g->pstep) {
        int *perr = device_err_flag();
        WMAR_REQUIRE(perr != nullptr, "cannot allocate the device error flag");
        if ((rc = pstep_enqueue(g->pstep, B, perr, s))) return rc;
        gpt_sample_kernel<<<B, SAMPLE_THREADS, sample_smem, s>>>(g->d_call, g->logits, g->seq, c.block_size + 1, g->step, perr);
        WMAR_LAUNCH_CHECK();
        advance_step_kernel<<<1, 1, 0, s>>>(g->step);
        WMAR_LAUNCH_CHECK();
        g->launches_per_step = 3;
        return WMAR_OK;
    }
    if (g->fused) {
        const int P_att = (H / 2) * 4, P_mlp = 4 * d / 128;
        for (int l = 0; l < c.n_layer; l++) {
            const Layer &L = g->layers[l];
            CUtensorMap mqkv, mproj, mw1, mw2;
            if ((rc = tc_weight_map(L.wqkv, 3 * d, d, &mqkv))) return rc;
            if ((rc = tc_weight_map(L.wproj, d, d, &mproj))) return rc;
            if ((rc = tc_weight_map(L.w1, 4 * d, d, &mw1))) return rc;
            if ((rc = tc_weight_map(L.w2, d, 4 * d, &mw2))) return rc;
            FusedArgs a{};
            a.x = g->x; a.d = d; a.stats_in = g->stats; a.n_stat_tiles = l == 0 ? d / 64 : d / 32; a.eps = 1e-5f;
            a.ln_g = L.ln1_g; a.ln_b = L.ln1_b; a.bias1 = L.bqkv; a.ws = g->ws; a.C1 = d / 32; a.T2 = d / 128;
            a.kcache = g->kcache; a.vcache = g->vcache; a.step = g->step; a.H = H; a.T = c.block_size; a.layer = l;
            { const char *e = getenv("WMAR_FB_DBG"); a.dbg = e ? atoi(e) : 0; }
            if (g->d_trace && l == c.n_layer / 2) { a.trace = g->d_trace; a.trace_step = g->trace_step; }
            if (g->d_trace && l == c.n_layer / 2 + 1) { a.trace = g->d_trace + 96; a.trace_step = g->trace_step; }
            if ((rc = launch_fused(FB_ATT, mqkv, mproj, a, H / 2, s))) return rc;
            if ((rc = launch_resid_reduce(g->ws, P_att, L.bproj, g->x, g->stats, d, s))) return rc;
            FusedArgs m{};
            m.x = g->x; m.d = d; m.stats_in = g->stats; m.n_stat_tiles = d / 32; m.eps = 1e-5f;
            m.ln_g = L.ln2_g; m.ln_b = L.ln2_b; m.bias1 = L.b1; m.ws = g->ws2; m.C1 = d / 32; m.T2 = d / 128;
            m.step = g->step; m.dbg = a.dbg; m.hpart = g->hpart; m.hflag = g->hflag + (size_t)l * (4 * d / 128);
            if (g->d_trace && l == c.n_layer / 2) { m.trace = g->d_trace + 48; m.trace_step = g->trace_step; }
            if ((rc = launch_fused(FB_MLP, mw1, mw2, m, 4 * d / 128, s))) return rc;
            if ((rc = launch_resid_reduce(g->ws2, P_mlp, L.b2, g->x, g->stats, d, s))) return rc;
            launches += 4;
        }
    }
    // flag-carrying split-K hand-off (gemm.cuh): epoch = the token-step counter, salt = launch index within the step;
    // WMAR_LL=0 falls back to the counter hand-off (A/B runs)
    static const bool ll_on = []() { const char *e = getenv("WMAR_LL"); return !(e && e[0] == '0'); }();
    WMAR_REQUIRE(4 * c.n_layer + 1 < 1024, "too many GEMM launches per step for the hand-off flag");
    unsigned salt = 0;
    // (not on the fused path: there the block kernels fill the same workspace with raw fp32 partials every step)
    auto ll = [&](GemmArgs &q) { if (ll_on && !g->fused) { q.ll_epoch = g->step; q.ll_salt = ++salt; } };
    for (int l = 0; l < (g->fused ? 0 : c.n_layer); l++) {
        const Layer &L = g->layers[l];
        GemmArgs a{};
        a.ws = g->ws; a.counters = g->counters; a.eps = 1e-5f;
        // qkv = LN1(x) Wqkv^T + b
        a.X = g->x; a.ldx = d; a.W = L.wqkv; a.bias = L.bqkv; a.Y = g->qkv; a.ldy = 3 * d; a.N = 3 * d; a.K = d;
        a.splits = g->splits_qkv; a.ln_g = L.ln1_g; a.ln_b = L.ln1_b; a.stats_in = g->stats; a.n_stat_tiles = stat_tiles;
        ll(a);
        if ((rc = launch_skinny_gemm(PRO_LN, EPI_STORE, a, s))) return rc;
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)H, (unsigned)B, 1);
            cfg.blockDim = dim3(ATT_THREADS, 1, 1);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            // measured on Taming C2: 2896 us / token with the 4-lane kernel vs 2742 with the 16-lane one (which also holds
            // the first 128 V rows in registers before the wait) -> opt-in only; the RAR engine's twin is its default
            static const bool attn4 = []() { const char *e = getenv("WMAR_ATTN4"); return e && e[0] == '1'; }();
            if (attn4)
                WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_decode4_kernel, (const float *)g->qkv, d, H, c.block_size, g->kcache,
                                                   g->vcache, l, (const int *)g->step, g->y));
            else
                WMAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, attn_decode_kernel<64>, (const float *)g->qkv, d, H, c.block_size, g->kcache,
                                                   g->vcache, l, (const int *)g->step, g->y));
            g_launches.fetch_add(1);
        }
        // x += y Wproj^T + b
        GemmArgs p{};
        p.ws = g->ws; p.counters = g->counters;
        p.X = g->y; p.ldx = d; p.W = L.wproj; p.bias = L.bproj; p.Y = g->x; p.ldy = d; p.N = d; p.K = d;
        p.splits = g->splits_proj; p.resid = g->x; p.ld_resid = d; p.stats_out = g->stats;
        ll(p);
        if ((rc = launch_skinny_gemm(PRO_NONE, EPI_RESID, p, s))) return rc;
        // m = GELU(LN2(x) W1^T + b1)
        GemmArgs f{};
        f.ws = g->ws; f.counters = g->counters; f.eps = 1e-5f;
        f.X = g->x; f.ldx = d; f.W = L.w1; f.bias = L.b1; f.Y = g->hbuf; f.ldy = 4 * d; f.N = 4 * d; f.K = d;
        f.splits = g->splits_fc1; f.ln_g = L.ln2_g; f.ln_b = L.ln2_b; f.stats_in = g->stats; f.n_stat_tiles = stat_tiles;
        ll(f);
        if ((rc = launch_skinny_gemm(PRO_LN, EPI_GELU, f, s))) return rc;
        // x += m W2^T + b2
        GemmArgs o{};
        o.ws = g->ws; o.counters = g->counters;
        o.X = g->hbuf; o.ldx = 4 * d; o.W = L.w2; o.bias = L.b2; o.Y = g->x; o.ldy = d; o.N = d; o.K = 4 * d;
        o.splits = g->splits_fc2; o.resid = g->x; o.ld_resid = d; o.stats_out = g->stats;
        ll(o);
        if ((rc = launch_skinny_gemm(PRO_NONE, EPI_RESID, o, s))) return rc;
        launches += 5;
    }
    GemmArgs hd{};
    hd.ws = g->ws; hd.counters = g->counters; hd.eps = 1e-5f;
    hd.X = g->x; hd.ldx = d; hd.W = g->head; hd.bias = nullptr; hd.Y = g->logits; hd.ldy = V; hd.N = V; hd.K = d;
    hd.splits = g->splits_head; hd.ln_g = g->lnf_g; hd.ln_b = g->lnf_b; hd.stats_in = g->stats;
    hd.n_stat_tiles = g->fused ? d / 32 : stat_tiles;
    ll(hd);
    if ((rc = launch_skinny_gemm(PRO_LN, EPI_STORE, hd, s))) return rc;
    launches += 1;
    int *err = device_err_flag();
    WMAR_REQUIRE(err != nullptr, "cannot allocate the device error flag");
    static const bool one_tail = []() { const char *e = getenv("WMAR_TAIL"); return !(e && e[0] == '0'); }();
    if (one_tail) {
        gpt_tail_kernel<<<16, SAMPLE_THREADS, sample_smem, s>>>(g->d_call, g->logits, g->seq, c.block_size + 1, g->step, g->ticket, err,
                                                               g->tok_emb, g->pos_emb, d, c.block_size, g->x, g->stats);
        WMAR_LAUNCH_CHECK();
        launches += 1;
    } else {
        gpt_sample_kernel<<<B, SAMPLE_THREADS, sample_smem, s>>>(g->d_call, g->logits, g->seq, c.block_size + 1, g->step, err);
        WMAR_LAUNCH_CHECK();
        advance_step_kernel<<<1, 1, 0, s>>>(g->step);
        WMAR_LAUNCH_CHECK();
        embed_kernel<<<16, 256, 0, s>>>(g->d_call, g->seq, c.block_size + 1, g->step, g->tok_emb, g->pos_emb, d, c.block_size,
                                        V, g->x, g->stats);
        WMAR_LAUNCH_CHECK();
        launches += 3;
    }
    g->launches_per_step = launches;
    return WMAR_OK;
}

}  // namespace

extern "C" {

int wmar_gpt_create(const wmar_gpt_config *cfg, const void *const *d_weights, int n_weights, wmar_gpt **out) {
    WMAR_REQUIRE(cfg != nullptr && d_weights != nullptr && out != nullptr, "NULL argument");
    WMAR_REQUIRE(cfg->n_embd % 64 == 0 && cfg->n_embd % cfg->n_head == 0, "n_embd must be a multiple of 64 and of n_head");
    WMAR_REQUIRE(cfg->n_embd / cfg->n_head == 64, "this engine supports head_dim 64 (Taming)");
    WMAR_REQUIRE(cfg->n_embd % 128 == 0 && cfg->vocab_size % 64 == 0, "n_embd % 128 == 0 and vocab % 64 == 0 required");
    WMAR_REQUIRE(cfg->block_size >= 1 && cfg->block_size <= 1024, "block_size must be in [1,1024]");
    WMAR_REQUIRE(cfg->max_batch >= 1 && cfg->max_batch <= 16, "max_batch must be in [1,16]");
    WMAR_REQUIRE(n_weights == 2 + 12 * cfg->n_layer + 3, "weight table has the wrong number of entries");
    for (int i = 0; i < n_weights; i++) WMAR_REQUIRE(d_weights[i] != nullptr, "NULL weight pointer");
    wmar_gpt *g = new (std::nothrow) wmar_gpt();
    if (!g) return set_error(WMAR_ERR_NOMEM, "out of host memory%s%s");
    g->cfg = *cfg;
    int dev = 0;
    WMAR_CUDA_CHECK(cudaGetDevice(&dev));
    WMAR_CUDA_CHECK(cudaDeviceGetAttribute(&g->n_sms, cudaDevAttrMultiProcessorCount, dev));
    auto W = [&](int i) { return reinterpret_cast<const float *>(d_weights[i]); };
    g->tok_emb = W(0);
    g->pos_emb = W(1);
    g->layers.resize(cfg->n_layer);
    for (int l = 0; l < cfg->n_layer; l++) {
        int b = 2 + 12 * l;
        g->layers[l] = Layer{W(b), W(b + 1), W(b + 2), W(b + 3), W(b + 4), W(b + 5), W(b + 6), W(b + 7), W(b + 8), W(b + 9), W(b + 10), W(b + 11)};
    }
    int b = 2 + 12 * cfg->n_layer;
    g->lnf_g = W(b); g->lnf_b = W(b + 1); g->head = W(b + 2);
    const int d = cfg->n_embd, V = cfg->vocab_size;
    g->splits_qkv = pick_splits(3 * d, d, g->n_sms);
    g->splits_proj = pick_splits(d, d, g->n_sms);
    g->splits_fc1 = pick_splits(4 * d, d, g->n_sms);
    g->splits_fc2 = pick_splits(d, 4 * d, g->n_sms);
    g->splits_head = pick_splits(V, d, g->n_sms);
    {   // probe only: WMAR_SPLITS="qkv,proj,fc1,fc2,head" overrides (0 keeps the pick)
        const char *e = getenv("WMAR_SPLITS");
        int v5[5] = {0, 0, 0, 0, 0};
        if (e && sscanf(e, "%d,%d,%d,%d,%d", &v5[0], &v5[1], &v5[2], &v5[3], &v5[4]) >= 1) {
            if (v5[0] > 0) g->splits_qkv = v5[0];
            if (v5[1] > 0) g->splits_proj = v5[1];
            if (v5[2] > 0) g->splits_fc1 = v5[2];
            if (v5[3] > 0) g->splits_fc2 = v5[3];
            if (v5[4] > 0) g->splits_head = v5[4];
        }
    }
    size_t ws_floats = 0;
    auto upd = [&](int N, int K, int S) { size_t n = gemm_ws_floats(N, K, S, g->n_sms); if (n > ws_floats) ws_floats = n; };
    upd(3 * d, d, g->splits_qkv); upd(d, d, g->splits_proj); upd(4 * d, d, g->splits_fc1); upd(d, 4 * d, g->splits_fc2);
    upd(V, d, g->splits_head);
    const size_t kv_elems = (size_t)cfg->n_layer * 16 * d * cfg->block_size;
    int max_tiles = (V > 4 * d ? V : 4 * d) / GEMM_NT;
    WMAR_CUDA_CHECK(cudaMalloc(&g->x, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->qkv, sizeof(float) * 16 * 3 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->y, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->hbuf, sizeof(float) * 16 * 4 * d));
    WMAR_CUDA_CHECK(cudaMalloc(&g->logits, sizeof(float) * 16 * V));
    WMAR_CUDA_CHECK(cudaMalloc(&g->kcache, sizeof(float) * kv_elems));
    WMAR_CUDA_CHECK(cudaMalloc(&g->vcache, sizeof(float) * kv_elems));
    WMAR_CUDA_CHECK(cudaMalloc(&g->ws, sizeof(float) * (ws_floats ? ws_floats : 1)));
    g->ws_bytes = sizeof(float) * (ws_floats ? ws_floats : 1);
    WMAR_CUDA_CHECK(cudaMalloc(&g->stats, sizeof(float2) * (d / 32) * 16));
    WMAR_CUDA_CHECK(cudaMalloc(&g->counters, sizeof(unsigned) * max_tiles));
    WMAR_CUDA_CHECK(cudaMalloc(&g->seq, sizeof(int64_t) * 16 * (cfg->block_size + 1)));
    WMAR_CUDA_CHECK(cudaMalloc(&g->step, sizeof(int)));
    WMAR_CUDA_CHECK(cudaMalloc(&g->ticket, sizeof(int)));
    WMAR_CUDA_CHECK(cudaMemset(g->ticket, 0, sizeof(int)));
    WMAR_CUDA_CHECK(cudaMalloc(&g->d_call, sizeof(CallParams)));
    WMAR_CUDA_CHECK(cudaMallocHost(&g->h_call, sizeof(CallParams)));
    WMAR_CUDA_CHECK(cudaEventCreateWithFlags(&g->call_done, cudaEventDisableTiming));
    g->call_pending = false;
    WMAR_REQUIRE(device_err_flag() != nullptr, "cannot allocate the device error flag");  // not allowed under capture
    WMAR_CUDA_CHECK(cudaMemset(g->counters, 0, sizeof(unsigned) * max_tiles));
    WMAR_CUDA_CHECK(cudaMemset(g->x, 0, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->qkv, 0, sizeof(float) * 16 * 3 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->y, 0, sizeof(float) * 16 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->hbuf, 0, sizeof(float) * 16 * 4 * d));
    WMAR_CUDA_CHECK(cudaMemset(g->seq, 0, sizeof(int64_t) * 16 * (cfg->block_size + 1)));
    WMAR_CUDA_CHECK(cudaMemset(g->stats, 0, sizeof(float2) * (d / 32) * 16));
    g->graph = nullptr; g->exec = nullptr; g->graph_smem = 0; g->graph_B = 0;
    g->launches_per_step = 5 * cfg->n_layer + 4;
    // fused block kernels whenever the model tiles (WMAR_STEP=graph forces the per-GEMM path)
    g->fused = false; g->ws2 = nullptr; g->hpart = nullptr; g->hflag = nullptr; g->d_trace = nullptr; g->trace_step = -1;
    g->pstep = nullptr;
    {
        const char *e = getenv("WMAR_STEP_TRACE");
        if (e) {
            g->trace_step = atoi(e);
            WMAR_CUDA_CHECK(cudaMalloc(&g->d_trace, sizeof(unsigned long long) * 144));
            WMAR_CUDA_CHECK(cudaMemset(g->d_trace, 0, sizeof(unsigned long long) * 144));
            const unsigned long long big = ~0ull;   // the "min" slots
            for (int k = 0; k < 3; k++) {
                WMAR_CUDA_CHECK(cudaMemcpy(g->d_trace + 48 * k + 40, &big, 8, cudaMemcpyHostToDevice));
                WMAR_CUDA_CHECK(cudaMemcpy(g->d_trace + 48 * k + 42, &big, 8, cudaMemcpyHostToDevice));
            }
        }
    }
    {
        // WMAR_STEP=pstep -> the persistent step kernel (pstep.cuh); default / WMAR_STEP=graph -> per-GEMM graph,
        // WMAR_STEP=fused -> block kernels.  The default is whichever path measures fastest.
        const char *e = getenv("WMAR_STEP");
        const bool want_pstep = e && e[0] == 'p';
        if (want_pstep && pstep_eligible(*cfg, g->n_sms)) {
            std::vector<PstepWeights::Layer> pl((size_t)cfg->n_layer);
            for (int l = 0; l < cfg->n_layer; l++) {
                const Layer &Ls = g->layers[l];
                pl[l] = PstepWeights::Layer{Ls.ln1_g, Ls.ln1_b, Ls.wqkv, Ls.bqkv, Ls.wproj, Ls.bproj, Ls.ln2_g, Ls.ln2_b, Ls.w1, Ls.b1, Ls.w2, Ls.b2};
            }
            PstepWeights pw{g->tok_emb, g->pos_emb, g->lnf_g, g->lnf_b, g->head, pl.data()};
            int prc = pstep_create(*cfg, g->n_sms, pw, g->kcache, g->vcache, g->logits, g->step, g->seq, cfg->block_size + 1, &g->pstep);
            if (prc != WMAR_OK) { wmar_gpt_destroy(g); return prc; }
            g->launches_per_step = 3;
        }
    }
    {
        const char *e = getenv("WMAR_STEP");
        const bool want_fused = e && e[0] == 'f';   // WMAR_STEP=fused: tcgen05 / cluster block kernels (round 1 experiment)
        if (want_fused && fused_eligible(*cfg)) {
            const size_t P = (size_t)std::max((cfg->n_head / 2) * 4, 4 * d / 128);
            if (ws_floats < P * 16 * d) {
                cudaFree(g->ws);
                WMAR_CUDA_CHECK(cudaMalloc(&g->ws, sizeof(float) * P * 16 * d));
                g->ws_bytes = sizeof(float) * P * 16 * d;
            }
            WMAR_CUDA_CHECK(cudaMalloc(&g->ws2, sizeof(float) * P * 16 * d));
            WMAR_CUDA_CHECK(cudaMalloc(&g->hpart, sizeof(float) * (size_t)(4 * d / 128) * FB_MLP_CS * 16 * 128));
            WMAR_CUDA_CHECK(cudaMalloc(&g->hflag, sizeof(unsigned) * (size_t)cfg->n_layer * (4 * d / 128)));
            g->fused = true;
            g->launches_per_step = 4 * cfg->n_layer + 4;
        }
    }
    // the zero-fills above went to the legacy default stream: finish them before the handle can be used from a
    // non-blocking stream (a second engine lane otherwise saw them land in the middle of its first generation)
    WMAR_CUDA_CHECK(cudaDeviceSynchronize());
    *out = g;
    return WMAR_OK;
}

void wmar_gpt_destroy(wmar_gpt *g) {
    if (!g) return;
    cudaDeviceSynchronize();
    free_graph(g);
    cudaFree(g->x); cudaFree(g->qkv); cudaFree(g->y); cudaFree(g->hbuf); cudaFree(g->logits);
    cudaFree(g->kcache); cudaFree(g->vcache); cudaFree(g->ws); cudaFree(g->stats); cudaFree(g->counters);
    cudaFree(g->seq); cudaFree(g->step); cudaFree(g->ticket); cudaFree(g->d_call); cudaFreeHost(g->h_call);
    cudaFree(g->ws2); cudaFree(g->d_trace); cudaFree(g->hpart); cudaFree(g->hflag);
    pstep_destroy(g->pstep);
    cudaEventDestroy(g->call_done);
    delete g;
}

int wmar_gpt_sample(wmar_gpt *g, const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_cond,
                    int64_t B, int64_t steps, const float *d_noise, int64_t *d_out_codes, float *d_out_logits,
                    void *stream) {
    WMAR_REQUIRE(g != nullptr && sp != nullptr && d_cond != nullptr && d_out_codes != nullptr, "NULL argument");
    WMAR_REQUIRE(B >= 1 && B <= g->cfg.max_batch, "batch exceeds max_batch");
    WMAR_REQUIRE(steps >= 1 && steps <= g->cfg.block_size, "steps must be in [1, block_size]");
    cudaStream_t s = as_stream(stream);
    wmar_wm_params wm_local{};
    wm_local.vocab_size = g->cfg.vocab_size;
    if (wm != nullptr && wm->d_table != nullptr) wm_local = *wm;
    SampleArgs sa;
    int rc = make_sample_args(&wm_local, sp, g->cfg.vocab_size, &sa);
    if (rc) return rc;
    const size_t smem = sample_smem_bytes(g->cfg.vocab_size, sa.cand_cap);
    // the pinned staging struct may still be in flight from the previous call
    if (g->call_pending) WMAR_CUDA_CHECK(cudaEventSynchronize(g->call_done));
    g->h_call->sa = sa;
    g->h_call->cond = d_cond;
    g->h_call->noise = d_noise;
    g->h_call->out_codes = d_out_codes;
    g->h_call->out_logits = d_out_logits;
    g->h_call->B = (int)B;
    g->h_call->steps = (int)steps;
    WMAR_CUDA_CHECK(cudaMemcpyAsync(g->d_call, g->h_call, sizeof(CallParams), cudaMemcpyHostToDevice, s));
    WMAR_CUDA_CHECK(cudaEventRecord(g->call_done, s));
    g->call_pending = true;

    if (g->exec == nullptr || g->graph_smem != smem || g->graph_B != (int)B) {
        free_graph(g);
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(gpt_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(gpt_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaStream_t cs;
        WMAR_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        WMAR_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        rc = enqueue_step(g, (int)B, smem, cs);
        cudaError_t e = cudaStreamEndCapture(cs, &g->graph);
        cudaStreamDestroy(cs);
        if (rc) { if (g->graph) cudaGraphDestroy(g->graph); g->graph = nullptr; return rc; }
        if (e != cudaSuccess) return set_error(WMAR_ERR_CUDA, "cudaStreamEndCapture: %s%s", cudaGetErrorString(e));
        WMAR_CUDA_CHECK(cudaGraphInstantiate(&g->exec, g->graph, 0));
        g->graph_smem = smem;
        g->graph_B = (int)B;
    }
    // the step counter restarts at 0: stale {value, flag} words of the previous generation must not match
    if (g->pstep) { if ((rc = pstep_reset(g->pstep, s))) return rc; }
    else if (!g->fused) WMAR_CUDA_CHECK(cudaMemsetAsync(g->ws, 0, g->ws_bytes, s));
    if (g->fused)
        WMAR_CUDA_CHECK(cudaMemsetAsync(g->hflag, 0, sizeof(unsigned) * (size_t)g->cfg.n_layer * (4 * g->cfg.n_embd / 128), s));
    init_call_kernel<<<1, 32, 0, s>>>(g->d_call, g->seq, g->cfg.block_size + 1, g->step);
    WMAR_LAUNCH_CHECK();
    if (!g->pstep) {   // the persistent kernel embeds the token itself
        embed_kernel<<<16, 256, 0, s>>>(g->d_call, g->seq, g->cfg.block_size + 1, g->step, g->tok_emb, g->pos_emb,
                                        g->cfg.n_embd, g->cfg.block_size, g->cfg.vocab_size, g->x, g->stats);
        WMAR_LAUNCH_CHECK();
    }
    for (int64_t t = 0; t < steps; t++) {
        WMAR_CUDA_CHECK(cudaGraphLaunch(g->exec, s));
        g_launches.fetch_add((uint64_t)g->launches_per_step);
    }
    return WMAR_OK;
}

double wmar_gpt_algorithmic_bytes(const wmar_gpt *g, int64_t B, int64_t steps) {
    if (!g) return 0.0;
    const double d = g->cfg.n_embd, V = g->cfg.vocab_size, L = g->cfg.n_layer;
    // dense parameters streamed once per step: per layer 12 d^2 + 13 d, head V d, ln_f 2 d
    const double P = L * (12.0 * d * d + 13.0 * d) + V * d + 2.0 * d;
    double kv = 0.0;  // K and V rows read per step: t+1 keys at step t, plus the appended row written
    for (int64_t t = 0; t < steps; t++) kv += 2.0 * L * d * (double)(t + 1) + 2.0 * L * d;
    return 4.0 * (P * (double)steps + kv * (double)B);
}

int wmar_gpt_launches_per_step(const wmar_gpt *g) { return g ? g->launches_per_step : 0; }

/* probe only (WMAR_PSTEP_TRACE): globaltimer stamps of the last token step, [G][512]; returns G or a negative status */
int wmar_gpt_debug_pstep_trace(const wmar_gpt *g, unsigned long long *out, int64_t cap) {
    if (!g || !g->pstep || !out) return -1;
    return pstep_trace(g->pstep, out, (size_t)cap);
}

/* probe only: copies the [3][48] globaltimer stamps of the traced step to the host */
int wmar_gpt_debug_trace(const wmar_gpt *g, unsigned long long *out) {
    WMAR_REQUIRE(g != nullptr && g->d_trace != nullptr && out != nullptr, "tracing is off (WMAR_STEP_TRACE)");
    WMAR_CUDA_CHECK(cudaDeviceSynchronize());
    WMAR_CUDA_CHECK(cudaMemcpy(out, g->d_trace, sizeof(unsigned long long) * 144, cudaMemcpyDeviceToHost));
    return WMAR_OK;
}

}  // extern "C"
