// C-ABI entry points for the watermark logit processor and the fused sampling operator.
#include <mutex>

#include "sample.cuh"

using namespace wmar;

namespace {

// GentimeWatermark._process_logits (gentime_watermark.py:229-271): one CTA per row, in place.
__global__ void __launch_bounds__(256) process_logits_kernel(const uint32_t *__restrict__ table, long long n_rows, int V,
                                                             int seed_strategy, int h, int spatial_dim, float delta,
                                                             const int64_t *__restrict__ past, long long t,
                                                             long long past_stride, float *__restrict__ logits, int *err) {
    const int b = blockIdx.x;
    long long s = context_sum(past + (long long)b * past_stride, t, seed_strategy, h, spatial_dim);
    if (s < 0) return;  // the reference skips rows whose history is too short (:268-270)
    if (s >= n_rows) {
        if (threadIdx.x == 0) atomicOr(err, 1);
        return;
    }
    const uint32_t *row = table + s * (long long)((V + 31) / 32);
    float *l = logits + (long long)b * V;
    for (int w = threadIdx.x; w < (V + 31) / 32; w += blockDim.x) {
        uint32_t bits = row[w];
        while (bits) {
            int k = __ffs(bits) - 1;
            bits &= bits - 1;
            int v = w * 32 + k;
            if (v < V) l[v] += delta;
        }
    }
}

__global__ void torch_exp_fill_kernel(SampleArgs a, long long numel, float *__restrict__ out) {
    const long long li = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (li < numel) out[li] = torch_cuda_exp1(a, 0u, (unsigned)(li / a.torch_rowlen), (unsigned)(li % a.torch_rowlen));
}

__global__ void __launch_bounds__(SAMPLE_THREADS, 1) wm_sample_kernel(SampleArgs a, const float *__restrict__ logits,
                                                                       const int64_t *__restrict__ past, long long t,
                                                                       long long past_stride,
                                                                       const float *__restrict__ noise,
                                                                       int64_t *__restrict__ out_ids, int *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.x;
    int id = sample_row(a, logits + (long long)b * a.V, past ? past + (long long)b * past_stride : nullptr, t,
                        noise ? noise + (long long)b * a.V : nullptr, (unsigned long long)b, err, smem_raw);
    if (threadIdx.x == 0) out_ids[b] = id;
}

constexpr int MAX_DEVICES = 64;
int *g_err_flags[MAX_DEVICES] = {};  // one device int per CUDA device, lazily allocated on the device that is current
std::mutex g_err_mutex;

int *cur_err_flag() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return nullptr;
    return g_err_flags[dev];
}

int ensure_err_flag() {
    int dev = 0;
    WMAR_CUDA_CHECK(cudaGetDevice(&dev));
    WMAR_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "device index out of range");
    std::lock_guard<std::mutex> lock(g_err_mutex);
    if (g_err_flags[dev] == nullptr) {
        int *p = nullptr;
        WMAR_CUDA_CHECK(cudaMalloc(&p, sizeof(int)));
        WMAR_CUDA_CHECK(cudaMemset(p, 0, sizeof(int)));
        g_err_flags[dev] = p;
    }
    return WMAR_OK;
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

namespace wmar {

int make_sample_args(const wmar_wm_params *wm, const wmar_sample_params *sp, int V, SampleArgs *out) {
    SampleArgs a{};
    a.V = V;
    if (wm != nullptr && wm->d_table != nullptr) {
        WMAR_REQUIRE(wm->vocab_size == V, "watermark vocab_size != logits width");
        WMAR_REQUIRE(wm->seed_strategy >= 0 && wm->seed_strategy <= 2, "bad seed strategy");
        WMAR_REQUIRE(wm->seed_strategy != WMAR_SEED_SPATIAL || wm->context_size == 1 || wm->context_size == 3,
                     "Spatial seeding only implemented for context size in [1,3]");
        a.table = wm->d_table;
        a.n_rows = wm->n_rows;
        a.seed_strategy = wm->seed_strategy;
        a.h = wm->context_size;
        a.spatial_dim = wm->spatial_dim > 0 ? wm->spatial_dim : 16;
        a.delta = wm->delta;
    }
    WMAR_REQUIRE(sp->temperature > 0.f, "temperature must be > 0");
    a.temperature = sp->temperature;
    a.top_k = sp->top_k;
    a.greedy = sp->greedy;
    a.seed = sp->seed;
    a.rng_mode = sp->rng_mode;
    if (sp->rng_mode == 1) {
        WMAR_REQUIRE(sp->torch_threads > 0 && sp->torch_threads % 256 == 0 && sp->torch_numel > 0 && sp->torch_rowlen > 0 &&
                     sp->torch_offset % 4 == 0, "bad torch generator replication parameters");
        a.torch_threads = (unsigned)sp->torch_threads;
        a.torch_iters = (unsigned)((sp->torch_numel - 1) / ((int64_t)sp->torch_threads * 4) + 1);
        a.torch_offset = sp->torch_offset;
        a.torch_rowlen = sp->torch_rowlen;
    }
    const bool use_top_p = sp->top_p > 0.0 && sp->top_p < 1.0;
    a.top_p_threshold = use_top_p ? (float)(1.0 - sp->top_p) : -1.f;
    if (use_top_p) {
        int cap = (sp->top_k > 0 && sp->top_k < V) ? next_pow2(sp->top_k * 2 > 2048 ? sp->top_k * 2 : 2048) : next_pow2(V);
        if (cap > next_pow2(V)) cap = next_pow2(V);
        a.cand_cap = cap;
    }
    WMAR_REQUIRE(sample_smem_bytes(V, a.cand_cap) <= 227 * 1024, "vocab too large for the shared-memory sampler");
    *out = a;
    return WMAR_OK;
}

}  // namespace wmar

extern "C" {

int wmar_wm_process_logits(const wmar_wm_params *wm, const int64_t *d_past_ids, int64_t B, int64_t t,
                           int64_t past_stride, float *d_logits, void *stream) {
    WMAR_REQUIRE(wm != nullptr && wm->d_table != nullptr, "no greenlist table");
    WMAR_REQUIRE(d_logits != nullptr && B > 0 && t >= 0, "bad logits / batch");
    WMAR_REQUIRE(d_past_ids != nullptr || t == 0, "past_ids is NULL");
    WMAR_REQUIRE(wm->seed_strategy >= 0 && wm->seed_strategy <= 2, "Invalid seed strategy");
    WMAR_REQUIRE(wm->seed_strategy != WMAR_SEED_SPATIAL || wm->context_size == 1 || wm->context_size == 3,
                 "Spatial seeding only implemented for context size in [1,3]");
    int rc = ensure_err_flag();
    if (rc) return rc;
    process_logits_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(
        wm->d_table, wm->n_rows, (int)wm->vocab_size, wm->seed_strategy, wm->context_size,
        wm->spatial_dim > 0 ? wm->spatial_dim : 16, wm->delta, d_past_ids, t, past_stride, d_logits, cur_err_flag());
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

int wmar_wm_sample(const wmar_wm_params *wm, const wmar_sample_params *sp, const int64_t *d_past_ids, int64_t B,
                   int64_t t, int64_t past_stride, const float *d_logits, const float *d_noise, int64_t *d_out_ids,
                   void *stream) {
    WMAR_REQUIRE(sp != nullptr && d_logits != nullptr && d_out_ids != nullptr && B > 0, "bad arguments");
    int64_t V = (wm != nullptr && wm->d_table != nullptr) ? wm->vocab_size : 0;
    WMAR_REQUIRE(V > 0 || wm != nullptr, "vocab size unknown: pass wm with vocab_size set (d_table may be NULL)");
    if (V == 0) V = wm->vocab_size;
    SampleArgs a;
    int rc = make_sample_args(wm, sp, (int)V, &a);
    if (rc) return rc;
    rc = ensure_err_flag();
    if (rc) return rc;
    size_t smem = sample_smem_bytes((int)V, a.cand_cap);
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(wm_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wm_sample_kernel<<<(unsigned)B, SAMPLE_THREADS, smem, as_stream(stream)>>>(a, d_logits, d_past_ids, t, past_stride,
                                                                              d_noise, d_out_ids, cur_err_flag());
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

/* test hook: out[numel] = the Exp(1) tensor torch's CUDA generator (seed, offset) would draw for a [rows][rowlen] tensor
 * on a device with `threads` = 256 * grid generator threads (rng_mode 1 of the sampler, element by element) */
int wmar_debug_torch_exponential(uint64_t seed, uint64_t offset, int64_t rows, int64_t rowlen, int threads, float *d_out,
                                 void *stream) {
    WMAR_REQUIRE(d_out != nullptr && rows > 0 && rowlen > 0 && threads > 0 && threads % 256 == 0, "bad arguments");
    SampleArgs a{};
    a.seed = seed; a.rng_mode = 1; a.torch_threads = (unsigned)threads; a.torch_offset = offset; a.torch_rowlen = rowlen;
    a.torch_iters = (unsigned)((rows * rowlen - 1) / ((int64_t)threads * 4) + 1);
    torch_exp_fill_kernel<<<(unsigned)((rows * rowlen + 255) / 256), 256, 0, as_stream(stream)>>>(a, rows * rowlen, d_out);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

/* Reads and clears the error flag of the current device (synchronises the stream): 0 ok, bit 0 = context sum outside
 * the greenlist table, bit 1 = more top-p candidates than the shared-memory sorter holds, bit 2 = a bounded wait of the
 * persistent step kernel expired. */
int wmar_check_device_flag(void *stream) {
    int *g_err_flag = cur_err_flag();
    if (g_err_flag == nullptr) return 0;
    int v = 0;
    WMAR_CUDA_CHECK(cudaMemcpyAsync(&v, g_err_flag, sizeof(int), cudaMemcpyDeviceToHost, as_stream(stream)));
    WMAR_CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    if (v != 0) {
        WMAR_CUDA_CHECK(cudaMemsetAsync(g_err_flag, 0, sizeof(int), as_stream(stream)));
        if (v & 4) {
            char codes[64];
            snprintf(codes, sizeof(codes), "wait codes 0x%x", (unsigned)v >> 8);
            return set_error(WMAR_ERR_CUDA, "persistent step kernel: a wait timed out (%s)%s", codes);
        }
        return set_error(WMAR_ERR_RANGE, "%s%s", (v & 1) ? "context sum outside the greenlist table; " : "",
                         (v & 2) ? "top-p candidate overflow" : "");
    }
    return WMAR_OK;
}

}  // extern "C"

namespace wmar {
int *device_err_flag() {
    if (ensure_err_flag() != WMAR_OK) return nullptr;
    return cur_err_flag();
}
}  // namespace wmar
