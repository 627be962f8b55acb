// Shared helpers for the wmar_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/wmar_b200.h"

namespace wmar {

extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launches;

inline int set_error(int code, const char *fmt, const char *a = "", const char *b = "") {
    snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
    return code;
}

#define WMAR_CUDA_CHECK(expr)                                                                       \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) return wmar::set_error(WMAR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

#define WMAR_REQUIRE(cond, msg)                                                   \
    do {                                                                          \
        if (!(cond)) return wmar::set_error(WMAR_ERR_INVALID, "%s (%s)", msg, #cond); \
    } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define WMAR_LAUNCH_CHECK()                    \
    do {                                       \
        wmar::g_launches.fetch_add(1);         \
        WMAR_CUDA_CHECK(cudaGetLastError());   \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// (salt * ctx_sum) mod (2^64 - 1) with Python big-int semantics (gentime_watermark.py:225)
__host__ __device__ __forceinline__ uint64_t context_seed(uint64_t salt, uint64_t s) {
#ifdef __CUDA_ARCH__
    uint64_t lo = salt * s, hi = __umul64hi(salt, s);
#else
    unsigned __int128 p = (unsigned __int128)salt * (unsigned __int128)s;
    uint64_t lo = (uint64_t)p, hi = (uint64_t)(p >> 64);
#endif
    // x = hi*2^64 + lo == hi + lo (mod 2^64-1)
    uint64_t r = lo + hi;
    if (r < lo) r += 1;                       // carry: 2^64 == 1 (mod 2^64-1)
    if (r == 0xffffffffffffffffull) r = 0;
    return r;
}

// Context selection of _process_logits (gentime_watermark.py:233-263).  past = history of ONE row (length t).
// Returns the context sum (>= 0), or -1 if the reference skips the row.
__device__ __forceinline__ long long context_sum(const int64_t *past, long long t, int seed_strategy, int h,
                                                 int spatial_dim) {
    if (seed_strategy == WMAR_SEED_FIXED) return 0;
    if (seed_strategy == WMAR_SEED_LINEAR) {
        if (t < h) return -1;
        long long s = 0;
        for (int i = 0; i < h; i++) s += past[t - h + i];
        return s;
    }
    // SPATIAL
    if (h == 3) {
        if (t < spatial_dim + 1) return -1;
        return past[t - spatial_dim - 1] + past[t - spatial_dim] + past[t - 1];
    }
    if (t < h) return -1;
    if (t % spatial_dim == 0) return (t >= spatial_dim) ? past[t - spatial_dim] : 0;
    return past[t - 1];
}

}  // namespace wmar
