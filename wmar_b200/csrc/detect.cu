// Detector: GentimeWatermark.detect / _score_ngrams_in_passage (wmar/watermarking/gentime_watermark.py:285-344),
// ngrams (:33-44) and spatial_ngrams (:47-88).  One CTA per passage: the (h+1)-grams are de-duplicated by tuple
// comparison in shared memory, each unique n-gram looks up one bit of the greenlist table row of its context sum,
// and the CTA reduces (n_green, T), then evaluates z and the binomial tail p = I_gamma(n_green, T - n_green + 1) in fp64.
#include "common.cuh"

using namespace wmar;

namespace wmar {
int *device_err_flag();
}

namespace {

constexpr int DT = 256;
constexpr int MAX_N = 8;  // max n-gram length (h + 1)

struct DetectArgs {
    const uint32_t *table;
    long long n_rows;
    int V, seed_strategy, h;
    double gamma;
    int L, sq, n_ngrams;
};

// position in the passage of element k of n-gram a
__device__ __forceinline__ int ngram_pos(const DetectArgs &a, int g, int k) {
    if (a.seed_strategy != WMAR_SEED_SPATIAL) return g + k;
    const int sq = a.sq;
    if (a.h == 1) {
        // order of spatial_ngrams(n=2): row-major over (i,j) skipping (0,0); (i,0) pairs with the token above,
        // (i,j>0) with the token to the left
        int cell = g + 1;  // (0,0) yields nothing
        int i = cell / sq, j = cell % sq;
        if (j == 0) return k == 0 ? (i - 1) * sq : i * sq;
        return k == 0 ? i * sq + j - 1 : i * sq + j;
    }
    // h == 3: 2x2 blocks, row-major over (i,j) in [0,sq-1)^2: TL, TR, BL, BR
    int i = g / (sq - 1), j = g % (sq - 1);
    return (i + (k >> 1)) * sq + j + (k & 1);
}

__global__ void __launch_bounds__(DT) detect_kernel(DetectArgs a, const int64_t *__restrict__ codes,
                                                    int32_t *__restrict__ out_green, int32_t *__restrict__ out_scored,
                                                    double *__restrict__ out_z, double *__restrict__ out_p,
                                                    int8_t *__restrict__ out_mask, long long mask_stride,
                                                    int32_t *__restrict__ out_mask_len, int *err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    long long *tok = reinterpret_cast<long long *>(smem_raw);  // [L]
    __shared__ int s_green[DT / 32], s_scored[DT / 32];
    __shared__ double s_red[DT / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.h + 1;
    for (int i = tid; i < a.L; i += DT) tok[i] = codes[(long long)b * a.L + i];
    __syncthreads();

    int my_green = 0, my_scored = 0;
    const long long words = (a.V + 31) / 32;
    for (int g = tid; g < a.n_ngrams; g += DT) {
        long long mine[MAX_N];
#pragma unroll
        for (int k = 0; k < MAX_N; k++)
            if (k < n) mine[k] = tok[ngram_pos(a, g, k)];
        bool dup = false;
        for (int o = 0; o < g && !dup; o++) {
            bool same = true;
#pragma unroll
            for (int k = 0; k < MAX_N; k++)
                if (k < n && same) same = (tok[ngram_pos(a, o, k)] == mine[k]);
            dup = same;
        }
        int8_t m = -1;
        if (!dup) {
            long long s = 0;
            if (a.seed_strategy != WMAR_SEED_FIXED)
                for (int k = 0; k < a.h; k++) s += mine[k];
            int gbit = 0;
            if (s < 0 || s >= a.n_rows) {
                atomicOr(err, 1);
            } else {
                long long tgt = mine[a.h];
                if (tgt >= 0 && tgt < a.V) gbit = (int)((a.table[s * words + (tgt >> 5)] >> (tgt & 31)) & 1u);
            }
            my_scored++;
            my_green += gbit;
            m = (int8_t)gbit;
        }
        if (out_mask) out_mask[(long long)b * mask_stride + a.h + g] = m;
    }
    if (out_mask)
        for (int k = tid; k < a.h; k += DT) out_mask[(long long)b * mask_stride + k] = -1;
    // reduce counts
    for (int o = 16; o > 0; o >>= 1) {
        my_green += __shfl_xor_sync(0xffffffffu, my_green, o);
        my_scored += __shfl_xor_sync(0xffffffffu, my_scored, o);
    }
    if (lane == 0) { s_green[warp] = my_green; s_scored[warp] = my_scored; }
    __syncthreads();
    int n_green = 0, T = 0;
    for (int w = 0; w < DT / 32; w++) { n_green += s_green[w]; T += s_scored[w]; }

    // p = P[Bin(T, gamma) >= n_green] = sum_{k=n_green}^{T} C(T,k) gamma^k (1-gamma)^(T-k)   (== betainc, :338)
    double p;
    if (n_green <= 0) p = 1.0;
    else if (a.gamma <= 0.0) p = 0.0;
    else if (a.gamma >= 1.0) p = 1.0;
    else {
        const double lg = log(a.gamma), l1g = log1p(-a.gamma), lT = lgamma((double)T + 1.0);
        double part = 0.0;
        for (int k = n_green + tid; k <= T; k += DT)
            part += exp(lT - lgamma((double)k + 1.0) - lgamma((double)(T - k) + 1.0) + k * lg + (T - k) * l1g);
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) s_red[warp] = part;
        __syncthreads();
        p = 0.0;
        for (int w = 0; w < DT / 32; w++) p += s_red[w];
        if (p > 1.0) p = 1.0;
    }
    if (tid == 0) {
        if (out_green) out_green[b] = n_green;
        if (out_scored) out_scored[b] = T;
        if (out_z) out_z[b] = ((double)n_green - a.gamma * T) / sqrt((double)T * a.gamma * (1.0 - a.gamma));
        if (out_p) out_p[b] = p;
        if (out_mask_len) out_mask_len[b] = a.h + a.n_ngrams;
    }
}

}  // namespace

extern "C" int wmar_detect(const wmar_wm_params *wm, const int64_t *d_codes, int64_t B, int64_t L, int32_t *d_n_green,
                           int32_t *d_n_scored, double *d_z, double *d_p, int8_t *d_mask, int64_t mask_stride,
                           int32_t *d_mask_len, void *stream) {
    WMAR_REQUIRE(wm != nullptr && wm->d_table != nullptr, "no greenlist table");
    WMAR_REQUIRE(d_codes != nullptr && B > 0 && L > 0 && L <= 16384, "bad codes");
    WMAR_REQUIRE(wm->context_size >= 0 && wm->context_size + 1 <= MAX_N, "context size too large");
    WMAR_REQUIRE(wm->seed_strategy >= 0 && wm->seed_strategy <= 2, "Invalid seed strategy");
    if (L - wm->context_size < 1)
        return set_error(WMAR_ERR_SHORT, "Must have at least 1 token to score after the first min_context_len=%s tokens%s",
                         "h");
    DetectArgs a{};
    a.table = wm->d_table;
    a.n_rows = wm->n_rows;
    a.V = (int)wm->vocab_size;
    a.seed_strategy = wm->seed_strategy;
    a.h = wm->context_size;
    a.gamma = wm->gamma;
    a.L = (int)L;
    if (wm->seed_strategy == WMAR_SEED_SPATIAL) {
        int sq = 0;
        while ((int64_t)(sq + 1) * (sq + 1) <= L) sq++;
        WMAR_REQUIRE((int64_t)sq * sq == L, "Sequence must be a square");
        WMAR_REQUIRE(a.h == 1 || a.h == 3, "spatial n-grams only support n=4 (2x2 blocks) or n=2 (1x2 blocks)");
        a.sq = sq;
        a.n_ngrams = a.h == 1 ? sq * sq - 1 : (sq - 1) * (sq - 1);
    } else {
        a.n_ngrams = (int)L - a.h;
    }
    WMAR_REQUIRE(d_mask == nullptr || mask_stride >= a.h + a.n_ngrams, "mask_stride too small");
    int *err = device_err_flag();
    WMAR_REQUIRE(err != nullptr, "cannot allocate the device error flag");
    size_t smem = sizeof(long long) * (size_t)L;
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(detect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    detect_kernel<<<(unsigned)B, DT, smem, as_stream(stream)>>>(a, d_codes, d_n_green, d_n_scored, d_z, d_p, d_mask,
                                                               mask_stride, d_mask_len, err);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}
