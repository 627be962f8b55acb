// Fused sampling operator for one row of logits, one CTA per row:
//   +delta on green(ctx)  ->  /T  ->  top-k (keep ties)  ->  top-p (ascending sort, softmax, cumsum, keep last)
//   ->  softmax  ->  argmax(p / q)  (== torch.multinomial(p, 1) with q ~ Exp(1))  or first arg-max when greedy.
// Reference order: mingpt.py:349-363 (Taming), rar.py:441-454 (RAR); warper bodies: transformers logits_process.py
// TopKLogitsWarper / TopPLogitsWarper.  See oracle/sampling.py for the CPU restatement this is tested against.
#pragma once
#include "common.cuh"

namespace wmar {

constexpr int SAMPLE_THREADS = 1024;
constexpr int SAMPLE_WARPS = SAMPLE_THREADS / 32;

struct SampleArgs {
    // watermark
    const uint32_t *table;  // may be null
    long long n_rows;
    int V;
    int seed_strategy, h, spatial_dim;
    float delta;
    // sampler
    float temperature;
    int top_k;
    float top_p_threshold;  // (float)(1 - top_p); < 0 disables top-p
    int greedy;
    unsigned long long seed;
    int cand_cap;  // capacity of the candidate arrays (power of two), 0 when top-p is off
    // id window (Chameleon: only the image tokens [id_base, id_base + V) are allowed, logits_processor.py:135-151):
    // logits_row / noise_row point at the window, the greenlist table still covers the full vocabulary table_V
    int id_base;   // 0 = no window
    int table_V;   // 0 = V
    // rng_mode 1: replicate torch's CUDA generator (see wmar_sample_params): q of element (row b, column c) of the
    // [rows][rowlen] tensor drawn at call `s` = transform(Philox(seed, subsequence = li % T, counter = offset / 4 +
    // s * iters + (li / T) / 4)[(li / T) % 4]) with li = b * rowlen + c, T = torch_threads, iters = ceil(numel / 4T)
    int rng_mode;
    unsigned torch_threads, torch_iters;
    unsigned long long torch_offset;
    long long torch_rowlen;
};

__host__ __device__ inline size_t sample_smem_bytes(int V, int cand_cap) {
    return sizeof(float) * (size_t)V + sizeof(uint32_t) * SAMPLE_WARPS * 256 + (sizeof(float) + sizeof(int)) * (size_t)cand_cap + 256;
}

__device__ __forceinline__ uint32_t order_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// Philox4x32-10 (Salmon et al.), used only when no pre-drawn noise is supplied.
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
}
__device__ __forceinline__ float philox_exp1(unsigned long long seed, unsigned long long stream, uint32_t idx) {
    uint32_t c0 = idx, c1 = (uint32_t)stream, c2 = (uint32_t)(stream >> 32), c3 = 0x9E3779B9u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    float u = ((float)(c0 >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    return -logf(u);
}

// q ~ Exp(1) exactly as `empty_like(probs).exponential_(1)` yields it on torch's CUDA generator (ATen
// DistributionTemplates.h distribution_nullary_kernel + transformation::exponential, curand Philox4_32_10 /
// curand_uniform4): bit-identical to the tensor torch.multinomial divides by.
__device__ __forceinline__ float torch_cuda_exp1(const SampleArgs &a, unsigned call, unsigned row, unsigned col) {
    const unsigned long long li = (unsigned long long)row * (unsigned long long)a.torch_rowlen + col;
    const unsigned T = a.torch_threads;
    const unsigned idx = (unsigned)(li % T), q = (unsigned)(li / T);
    const unsigned long long ctr = a.torch_offset / 4ull + (unsigned long long)call * a.torch_iters + (q >> 2);
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = idx, c3 = 0u;
    uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    const unsigned ii = q & 3u;
    const uint32_t x = ii == 0 ? c0 : ii == 1 ? c1 : ii == 2 ? c2 : c3;
    const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);            // _curand_uniform: (0, 1]
    // at::log<float> on the device is the fast approximation __logf (ATen/NumericUtils.h:150-160), not logf
    const float lg = u >= 1.0f - 1.1920928955078125e-07f / 2.0f ? -1.1920928955078125e-07f / 2.0f : __logf(u);
    return -1.0f / 1.0f * lg;
}

// Block-wide arg-max with first-index tie break.  red_* : smem scratch of SAMPLE_WARPS entries each.
__device__ __forceinline__ int block_argmax(float v, int idx, float *red_v, int *red_i) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { red_v[warp] = v; red_i[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        v = lane < SAMPLE_WARPS ? red_v[lane] : -INFINITY;
        idx = lane < SAMPLE_WARPS ? red_i[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (lane == 0) red_i[0] = idx;
    }
    __syncthreads();
    int r = red_i[0];
    __syncthreads();
    return r;
}

__device__ __forceinline__ float block_max(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < SAMPLE_WARPS ? red[lane] : -INFINITY;
        v = warp_max(v);
        if (lane == 0) red[0] = v;
    }
    __syncthreads();
    float r = red[0];
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < SAMPLE_WARPS ? red[lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) red[0] = v;
    }
    __syncthreads();
    float r = red[0];
    __syncthreads();
    return r;
}

// logits_row: V floats (global); past: this row's history (length t); noise_row: V floats or null.
// Returns the sampled id (valid in every thread).  err: global int flag (bit 0 = ctx out of table, bit 1 = candidate
// overflow in top-p).
__device__ inline int sample_row(const SampleArgs &a, const float *__restrict__ logits_row, const int64_t *past, long long t,
                          const float *__restrict__ noise_row, unsigned long long noise_stream, int *err,
                          uint8_t *smem_raw) {
    const int V = a.V;
    float *vals = reinterpret_cast<float *>(smem_raw);
    uint32_t *whist = reinterpret_cast<uint32_t *>(vals + V);            // [SAMPLE_WARPS][256]
    float *cval = reinterpret_cast<float *>(whist + SAMPLE_WARPS * 256);  // [cand_cap]
    int *cidx = reinterpret_cast<int *>(cval + a.cand_cap);               // [cand_cap]
    float *red_v = reinterpret_cast<float *>(cidx + a.cand_cap);          // [32]
    int *red_i = reinterpret_cast<int *>(red_v + 32);                     // [32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- 1. watermark bias + temperature
    const uint32_t *row = nullptr;
    if (a.table != nullptr) {
        long long s = context_sum(past, t, a.seed_strategy, a.h, a.spatial_dim);
        if (s >= 0) {
            if (s < a.n_rows) row = a.table + s * (long long)(((a.table_V > 0 ? a.table_V : V) + 31) / 32);
            else if (tid == 0) atomicOr(err, 1);
        }
    }
    for (int v = tid; v < V; v += SAMPLE_THREADS) {
        float l = logits_row[v];
        const int gv = v + a.id_base;
        if (row != nullptr && ((row[gv >> 5] >> (gv & 31)) & 1u)) l += a.delta;
        vals[v] = l / a.temperature;
    }
    __syncthreads();

    // ---- 2. top-k: radix-select the k-th largest key, then drop everything strictly below it
    if (a.top_k > 0 && a.top_k < V) {
        uint32_t prefix = 0, mask = 0;
        int remaining = a.top_k;
        for (int pass = 0; pass < 4; pass++) {
            const int shift = 24 - 8 * pass;
            for (int i = tid; i < SAMPLE_WARPS * 256; i += SAMPLE_THREADS) whist[i] = 0;
            __syncthreads();
            for (int v = tid; v < V; v += SAMPLE_THREADS) {
                uint32_t k = order_key(vals[v]);
                if ((k & mask) == prefix) atomicAdd(&whist[warp * 256 + ((k >> shift) & 255u)], 1u);
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t s = 0;
                for (int w = 0; w < SAMPLE_WARPS; w++) s += whist[w * 256 + tid];
                whist[tid] = s;  // bin totals in warp-0's slice (each thread only touched its own column)
            }
            __syncthreads();
            if (tid == 0) {
                int cum = 0, bin = 255;
                for (; bin > 0; bin--) {
                    int c = (int)whist[bin];
                    if (cum + c >= remaining) break;
                    cum += c;
                }
                red_i[0] = bin;
                red_i[1] = remaining - cum;
            }
            __syncthreads();
            prefix |= ((uint32_t)red_i[0]) << shift;
            mask |= 255u << shift;
            remaining = red_i[1];
            __syncthreads();
        }
        const float kth = key_to_float(prefix);
        for (int v = tid; v < V; v += SAMPLE_THREADS)
            if (vals[v] < kth) vals[v] = -INFINITY;
        __syncthreads();
    }

    // ---- 3. top-p over the finite candidates
    if (a.top_p_threshold >= 0.f && a.cand_cap > 0) {
        int *counter = red_i + 2;
        if (tid == 0) *counter = 0;
        __syncthreads();
        for (int v = tid; v < V; v += SAMPLE_THREADS) {
            float l = vals[v];
            if (l > -INFINITY) {
                int p = atomicAdd(counter, 1);
                if (p < a.cand_cap) { cval[p] = l; cidx[p] = v; }
            }
        }
        __syncthreads();
        const int n = *counter;
        if (n > a.cand_cap) {
            if (tid == 0) atomicOr(err, 2);
        } else if (n > 1) {
            int np2 = 1;
            while (np2 < n) np2 <<= 1;
            for (int i = n + tid; i < np2; i += SAMPLE_THREADS) { cval[i] = INFINITY; cidx[i] = 0x7fffffff; }
            __syncthreads();
            // bitonic sort ascending by (value, index)
            for (int k = 2; k <= np2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < np2; i += SAMPLE_THREADS) {
                        int ixj = i ^ j;
                        if (ixj > i) {
                            float va = cval[i], vb = cval[ixj];
                            int ia = cidx[i], ib = cidx[ixj];
                            bool a_gt_b = (va > vb) || (va == vb && ia > ib);
                            bool up = ((i & k) == 0);
                            if (a_gt_b == up) { cval[i] = vb; cval[ixj] = va; cidx[i] = ib; cidx[ixj] = ia; }
                        }
                    }
                    __syncthreads();
                }
            }
            // softmax of the sorted logits (fp32), cumulative sum in fp64 (torch CPU cumsum accumulates float in
            // double), threshold compare in fp32, the largest is always kept (min_tokens_to_keep = 1)
            const float mx = cval[n - 1];
            float part = 0.f;
            for (int i = tid; i < n; i += SAMPLE_THREADS) part += expf(cval[i] - mx);
            const float sum = block_sum(part, red_v);
            if (tid == 0) {
                double cum = 0.0;
                for (int i = 0; i < n - 1; i++) {
                    float p = expf(cval[i] - mx) / sum;
                    cum += (double)p;
                    if ((float)cum <= a.top_p_threshold) vals[cidx[i]] = -INFINITY;
                    else break;  // cumulative sums are non-decreasing: nothing further is removed
                }
            }
            __syncthreads();
        }
    }

    // ---- 4. softmax + multinomial (argmax p/q) or greedy
    float m = -INFINITY;
    for (int v = tid; v < V; v += SAMPLE_THREADS) m = fmaxf(m, vals[v]);
    m = block_max(m, red_v);
    float part = 0.f;
    for (int v = tid; v < V; v += SAMPLE_THREADS) part += expf(vals[v] - m);
    const float sum = block_sum(part, red_v);
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int v = tid; v < V; v += SAMPLE_THREADS) {
        float p = expf(vals[v] - m) / sum;
        float sc;
        if (a.greedy) sc = p;
        else {
            float q;
            if (noise_row != nullptr) q = noise_row[v];
            else if (a.rng_mode == 1) q = torch_cuda_exp1(a, (unsigned)(noise_stream >> 32), (unsigned)noise_stream, (unsigned)(v + a.id_base));
            else q = philox_exp1(a.seed, noise_stream, (uint32_t)v);
            sc = p / q;
        }
        if (sc > best) { best = sc; best_i = v; }  // ascending v within a thread keeps the first index on ties
    }
    return block_argmax(best, best_i, red_v, red_i) + a.id_base;
}

}  // namespace wmar
