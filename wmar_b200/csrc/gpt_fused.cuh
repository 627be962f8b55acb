// Fused transformer-block kernels for the Taming minGPT decode step (mingpt.py:98-122 with the KV cache of :42-95).
//
// A token step is a chain of 5 dependent skinny GEMMs + attention per layer; as separate kernels each one pays
// launch, ramp-up and split-K tail latency (~9 us fixed, measured) for ~3.5 us of HBM time.  Here a layer is TWO
// kernels whose inner dependencies never leave the chip:
//
//   fused_block_kernel<FB_ATT>   cluster of 8 CTAs per head PAIR (grid 8 x H/2)
//       stage 1  [q;k;v](pair) = LN1(x) W^T      3 UMMA tiles of 128 rows, K split 8 ways over the cluster
//       exchange partial tiles stay in shared memory; every CTA pulls + sums (fixed order) the q,k,v of ITS four
//                (head,row) units from the 8 peers through distributed shared memory, appends k,v to the cache
//       attend   softmax(q K^T / sqrt(64)) V for its four units; the K and V rows of earlier steps travel through the
//                SAME shared-memory ring as the weights (1-D bulk copies queued by the producer right behind the
//                stage-1 tiles), so they are on chip before q exists
//       exchange y of the pair is pulled from the peers' shared memory
//       stage 2  partial proj:  y_pair[16 x 32-k slice] . Wproj[:, slice]^T  for 6 of the 12 output tiles
//   fused_block_kernel<FB_MLP>   cluster of 3 CTAs per 128 fc1 columns (grid 3 x 4d/128)
//       stage 1  h_tile = LN2(x) W1[tile]^T       one UMMA tile, K split 3 ways; partials summed through DSMEM,
//                + bias, erf-GELU  ->  B operand of stage 2 (never leaves shared memory)
//       stage 2  partial fc2:  h_tile . W2[:, tile]^T   for 4 of the 12 output tiles per CTA
//   resid_reduce_kernel          x += bias + sum_p partial_p  (fixed order, deterministic) + LayerNorm (mean, M2) partials
//
// Inside a CTA the machinery is that of gemm_tc.cuh: a TMA producer thread streams 128 x 32 fp32 weight tiles (16 KB,
// SWIZZLE_128B, L2 evict-first) of BOTH stages back-to-back into an 8-deep ring -- it depends on nothing, so under
// programmatic dependent launch the weights are in flight before the previous kernel has finished; two groups of
// converter warps split every tile into hi = rna_tf32(w) / lo = w - hi in TENSOR MEMORY; one thread issues
// tcgen05.mma.kind::tf32 (A = weights from TMEM, B = [x_hi ; x_lo] from shared memory, 3xTF32 = fp32-faithful);
// accumulators are double-buffered in TMEM and drained by four epilogue warps with tcgen05.ld.
#pragma once
#include "gemm_tc.cuh"

namespace wmar {

enum { FB_ATT = 0, FB_MLP = 1 };

constexpr int FB_NS = 8;            // shared-memory stages of raw W tiles (128 KB in flight per SM)
constexpr int FB_NTA = 4;           // TMEM stages of split W
constexpr int FB_THREADS = 704;     // warps: 0 TMA, 1 MMA, 2-9 converters (2 groups), 10-13 epilogue, 14-21 workers
constexpr int FB_W_CONV = 2, FB_W_EPI = 10, FB_W_WORK = 14;
constexpr int FB_WORKERS = 256;
constexpr int FB_TMEM_A0 = 256;     // TMEM columns [0,256): 2 accumulator buffers x 4 sets x 32; [256,512): 4 W stages
constexpr int FB_ATT_CS = 8;        // cluster sizes
constexpr int FB_MLP_CS = 3;
constexpr int FB_ATT_B1_MAX = 8;    // stage-1 activation chunks (32 k each) one CTA may own
constexpr int FB_MLP_B1_MAX = 16;
constexpr int FB_ATT_U = 4;         // (head,row) units per CTA: 2 heads x 16 rows / 8 CTAs
constexpr int FB_ATT_T = 256;       // max positions

// shared-memory map (bytes, after 1024 B alignment)
constexpr int FB_SM_A = 0;
constexpr int FB_SM_B1 = FB_SM_A + FB_NS * TC_A_BYTES;                      // 131072
// attention kernel
constexpr int FB_SM_ATT_B2 = FB_SM_B1 + FB_ATT_B1_MAX * TC_B_BYTES;          // 1 chunk
constexpr int FB_SM_ATT_PART = FB_SM_ATT_B2 + TC_B_BYTES;                    // float[3][16][128] partial q/k/v tiles
constexpr int FB_SM_ATT_QKV = FB_SM_ATT_PART + 3 * 16 * 128 * 4;             // float[4][192] reduced q,k,v of my units
constexpr int FB_SM_ATT_Y = FB_SM_ATT_QKV + FB_ATT_U * 192 * 4;              // float[4][64]
constexpr int FB_SM_ATT_SC = FB_SM_ATT_Y + FB_ATT_U * 64 * 4;                // float[4][256]
constexpr int FB_SM_ATT_PV = FB_SM_ATT_SC + FB_ATT_U * FB_ATT_T * 4;         // float[4][16][64]
constexpr int FB_SM_ATT_INV = FB_SM_ATT_PV + FB_ATT_U * 16 * 64 * 4;         // float[4]
constexpr int FB_SM_ATT_END = FB_SM_ATT_INV + 16;
// MLP kernel
constexpr int FB_SM_MLP_B2 = FB_SM_B1 + FB_MLP_B1_MAX * TC_B_BYTES;          // 4 chunks
constexpr int FB_SM_MLP_PART = FB_SM_MLP_B2 + 4 * TC_B_BYTES;                // float[16][128]
constexpr int FB_SM_MLP_END = FB_SM_MLP_PART + 16 * 128 * 4;
// common tail
constexpr int FB_SM_TAIL = FB_SM_MLP_END > FB_SM_ATT_END ? FB_SM_MLP_END : FB_SM_ATT_END;
constexpr int FB_SM_LNP = FB_SM_TAIL;                                        // float[2][16 * 32] LN gamma, beta of my k range
constexpr int FB_SM_ROWSTATS = FB_SM_LNP + 2 * FB_MLP_B1_MAX * TC_KC * 4;    // float2[16]
constexpr int FB_SM_BAR = FB_SM_ROWSTATS + 128;
constexpr int FB_N_BARS = 3 * FB_NS + 2 * FB_NTA + 2 + 4 + 2;
constexpr int FB_SM_MISC = FB_SM_BAR + 8 * FB_N_BARS;
constexpr int FB_SM_BYTES = FB_SM_MISC + 16;
constexpr int FB_SM_ALLOC = FB_SM_BYTES + 1024;
static_assert((2 * FB_ATT_U) % FB_NS == 0, "the K/V chunk count must be a multiple of the ring depth");
static_assert(FB_SM_ALLOC <= 232448, "fused block kernel exceeds 227 KB of shared memory");

struct FusedArgs {
    const float *x;            // [16][d] residual stream (input of the LayerNorm)
    int d;
    const float2 *stats_in;    // [n_stat_tiles][16] (mean, M2) partials of x
    int n_stat_tiles;
    float eps;
    const float *ln_g, *ln_b;
    const float *bias1;        // ATT: bqkv [3d]; MLP: b1 [4d]
    float *ws;                 // stage-2 partials [P][16][d]
    int C1;                    // k chunks of stage 1 (d / 32)
    int T2;                    // output tiles of stage 2 (d / 128)
    // attention only
    float *kcache, *vcache;
    const int *step;
    int H, T, layer;
    unsigned long long *trace;   // probe only: [32] globaltimer stamps of CTA (0,0) at step trace_step (null in production)
    int trace_step;
    // MLP only: the three CTAs of a tile exchange their K-split partials through L2 (no cluster: only 45 of the 48
    // 3-CTA clusters fit on a B200 at once, and a second wave doubles the kernel)
    float *hpart;                // [tiles][3][16][128]
    unsigned *hflag;             // [tiles] monotonic arrival counters of this layer (zeroed per sample call)
    int dbg;                     // probe only (results wrong): bit0 skip MMA issue, bit1 skip conversion, bit3 hi MMAs only, bit4 N=16
};

__device__ __forceinline__ unsigned long long fb_gtime() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
    return v;
}

struct FbSched { int c1_begin, n1, g1, t2_begin, g2, n2; };

template <int KIND>
__device__ __forceinline__ FbSched fb_sched(const FusedArgs &a, int j) {
    FbSched s;
    if (KIND == FB_ATT) {
        s.c1_begin = a.C1 * j / FB_ATT_CS;
        s.n1 = a.C1 * (j + 1) / FB_ATT_CS - s.c1_begin;
        s.g1 = 3;
        const int hf = j >> 2;
        s.t2_begin = a.T2 * hf / 2;
        s.g2 = a.T2 * (hf + 1) / 2 - s.t2_begin;
        s.n2 = 1;
    } else {
        s.c1_begin = a.C1 * j / FB_MLP_CS;
        s.n1 = a.C1 * (j + 1) / FB_MLP_CS - s.c1_begin;
        s.g1 = 1;
        s.t2_begin = a.T2 * j / FB_MLP_CS;
        s.g2 = a.T2 * (j + 1) / FB_MLP_CS - s.t2_begin;
        s.n2 = 4;
    }
    return s;
}

template <int KIND>
__global__ void __launch_bounds__(FB_THREADS, 1)
fused_block_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, const FusedArgs a) {
    using namespace tc05;
    constexpr int CS = KIND == FB_ATT ? FB_ATT_CS : FB_MLP_CS;
    constexpr int SM_B2 = KIND == FB_ATT ? FB_SM_ATT_B2 : FB_SM_MLP_B2;
    constexpr int SM_PART = KIND == FB_ATT ? FB_SM_ATT_PART : FB_SM_MLP_PART;
    extern __shared__ uint8_t fb_smem_raw[];
    const uint32_t smem_base = (smem_u32(fb_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = fb_smem_raw + (smem_base - smem_u32(fb_smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = (int)blockIdx.x;            // K-split rank: == %cluster_ctarank in the attention kernel (cluster dims (8,1,1))
    const int grp_id = blockIdx.y;            // ATT: head pair; MLP: fc1 tile
    const FbSched S = fb_sched<KIND>(a, j);
    const int d = a.d;
    // attention geometry of this CTA: head h of the pair, rows b0..b0+3, t earlier positions in the cache
    int t = 0, h = 0, b0 = 0, nkc = 0;
    size_t cache_base = 0, ustride = 0;
    if (KIND == FB_ATT) {
        t = *a.step;
        h = 2 * grp_id + (j >> 2);
        b0 = 4 * (j & 3);
        nkc = (t + 63) >> 6;                       // 64-key (16 KB) chunks of one unit's K (or V) rows
        ustride = (size_t)a.H * a.T * 64;
        cache_base = (((size_t)a.layer * 16 + b0) * a.H + h) * (size_t)a.T * 64;
    }
    const uint32_t G1 = (uint32_t)(S.g1 * S.n1), NKV = (uint32_t)(2 * FB_ATT_U * nkc);

    const uint32_t bar0 = smem_base + FB_SM_BAR;
    auto a_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (uint32_t)(FB_NS + s); };
    auto ta_full = [&](int u) { return bar0 + 8u * (uint32_t)(2 * FB_NS + u); };
    auto ta_empty = [&](int u) { return bar0 + 8u * (uint32_t)(2 * FB_NS + FB_NTA + u); };
    const uint32_t b1_full = bar0 + 8u * (uint32_t)(2 * FB_NS + 2 * FB_NTA);
    const uint32_t b2_full = b1_full + 8u;
    auto acc_full = [&](int ab) { return b2_full + 8u + 8u * (uint32_t)ab; };
    auto acc_empty = [&](int ab) { return b2_full + 24u + 8u * (uint32_t)ab; };
    const uint32_t xbar1 = b2_full + 40u, xbar2 = xbar1 + 8u;
    // K/V chunks share the ring slots (and a_empty) with the weights but complete on their OWN full barriers, so that
    // each barrier is only ever waited on by one consumer class, phase after phase (no parity aliasing)
    auto kv_full = [&](int s) { return xbar2 + 8u + 8u * (uint32_t)s; };
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + FB_SM_MISC);
    float2 *row_stats = reinterpret_cast<float2 *>(smem + FB_SM_ROWSTATS);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map1);
        tma_prefetch_desc(&map2);
        for (int s = 0; s < FB_NS; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 8); mbar_init(kv_full(s), 1); }
        for (int u = 0; u < FB_NTA; u++) { mbar_init(ta_full(u), 4); mbar_init(ta_empty(u), 1); }
        mbar_init(b1_full, FB_WORKERS / 32);
        mbar_init(b2_full, FB_WORKERS / 32);
        for (int ab = 0; ab < 2; ab++) { mbar_init(acc_full(ab), 1); mbar_init(acc_empty(ab), 4); }
        mbar_init(xbar1, CS);
        mbar_init(xbar2, CS);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(s_tmem));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *s_tmem;
    unsigned long long *trc = (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && *a.step == a.trace_step) ? a.trace : nullptr;
#define FB_TRACE(e) do { if (trc) trc[e] = fb_gtime(); } while (0)
    if (threadIdx.x == 0) FB_TRACE(0);
    // all CTAs: [40] min start, [41] max start, [42] min end, [43] max end
    unsigned long long *trc_all = (a.trace != nullptr && *a.step == a.trace_step) ? a.trace : nullptr;
    if (trc_all && threadIdx.x == 0) {
        const unsigned long long now = fb_gtime();
        atomicMin(trc_all + 40, now);
        atomicMax(trc_all + 41, now);
    }
    // the next kernel of the step may become resident as soon as resources free up and start streaming ITS weights
    if (threadIdx.x == 0) pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer: weights of both stages, waits for nothing but ring slots ==========
        if (lane == 0) {
            uint32_t g = 0;
            FB_TRACE(1);
            for (int gi = 0; gi < S.g1; gi++) {
                const int row = KIND == FB_ATT ? gi * d + grp_id * 128 : grp_id * 128;
                for (int c = 0; c < S.n1; c++, g++) {
                    const uint32_t s = g % FB_NS;
                    mbar_wait(a_empty(s), ((g / FB_NS) & 1) ^ 1);
                    mbar_arrive_expect_tx(a_full(s), TC_A_BYTES);
                    tma_load_2d(smem_base + FB_SM_A + s * TC_A_BYTES, &map1, (S.c1_begin + c) * TC_KC, row, a_full(s),
                                L2_EVICT_FIRST);
                }
            }
            if (KIND == FB_ATT) {
                // K rows of my four units, then their V rows: consumed by the worker warps instead of the converters
                for (int kv = 0; kv < 2; kv++) {
                    const float *cache = (a.dbg & 64) ? a.ws + (size_t)(blockIdx.y * 8 + j) * 4096   // probe: L2-hot source
                                                      : (kv ? a.vcache : a.kcache) + cache_base;
                    for (int u = 0; u < FB_ATT_U; u++) {
                        for (int c = 0; c < nkc; c++, g++) {
                            const uint32_t s = g % FB_NS;
                            const uint32_t bytes = (uint32_t)(t - 64 * c < 64 ? t - 64 * c : 64) * 256u;
                            mbar_wait(a_empty(s), ((g / FB_NS) & 1) ^ 1);
                            mbar_arrive_expect_tx(kv_full(s), bytes);
                            bulk_load(smem_base + FB_SM_A + s * TC_A_BYTES,
                                      cache + ((a.dbg & 64) ? (size_t)0 : (size_t)u * ustride + (size_t)c * 4096), bytes, kv_full(s));
                        }
                    }
                }
            }
            for (int gi = 0; gi < S.g2; gi++) {
                const int row = (S.t2_begin + gi) * 128;
                for (int c = 0; c < S.n2; c++, g++) {
                    const int k = KIND == FB_ATT ? grp_id * 128 + 32 * (j & 3) : grp_id * 128 + 32 * c;
                    const uint32_t s = g % FB_NS;
                    mbar_wait(a_empty(s), ((g / FB_NS) & 1) ^ 1);
                    mbar_arrive_expect_tx(a_full(s), TC_A_BYTES);
                    tma_load_2d(smem_base + FB_SM_A + s * TC_A_BYTES, &map2, k, row, a_full(s), L2_EVICT_FIRST);
                }
            }
            FB_TRACE(2);
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t ID32 = idesc_tf32_m128(32), ID16 = idesc_tf32_m128(16);
            uint32_t g = 0, item = 0;
            const int groups = S.g1 + S.g2;
            for (int gi = 0; gi < groups; gi++, item++) {
                const bool st2 = gi >= S.g1;
                const int nchunks = st2 ? S.n2 : S.n1;
                const uint32_t ab = item & 1;
                mbar_wait(acc_empty(ab), ((item >> 1) & 1) ^ 1);
                if (gi == 0) { mbar_wait(b1_full, 0); FB_TRACE(6); }
                if (gi == S.g1) { mbar_wait(b2_full, 0); FB_TRACE(8); }
                fence_after_sync();
                const uint32_t acc = tmem + ab * 128;
                const uint32_t bbase = smem_base + (st2 ? SM_B2 : FB_SM_B1);
                for (int c = 0; c < nchunks; c++, g++) {
                    const uint32_t u = g % FB_NTA;
                    mbar_wait(ta_full(u), (g / FB_NTA) & 1);
                    fence_after_sync();
                    const uint32_t a_hi = tmem + FB_TMEM_A0 + u * 64, a_lo = a_hi + 32;
                    const uint64_t bd = smem_desc_kmajor_noswz(bbase + c * TC_B_BYTES, 512, 128);
                    if (a.dbg & 8) {          // probe: hi MMAs only (half the instructions)
#pragma unroll
                        for (int ks = 0; ks < 4; ks++)
                            mma_tf32_ts(acc + ks * 32, a_hi + ks * 8, bd + (uint64_t)((ks * 1024) >> 4), ID32, c != 0 ? 1u : 0u);
                    } else if (a.dbg & 16) {  // probe: same instruction count, N = 16 everywhere
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            mma_tf32_ts(acc + ks * 32, a_hi + ks * 8, bd + (uint64_t)((ks * 1024) >> 4), ID16, c != 0 ? 1u : 0u);
                            mma_tf32_ts(acc + ks * 32 + 16, a_lo + ks * 8, bd + (uint64_t)((ks * 1024) >> 4), ID16, 1u);
                        }
                    } else if (!(a.dbg & 1)) {
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t bk = bd + (uint64_t)((ks * 1024) >> 4);
                            mma_tf32_ts(acc + ks * 32, a_hi + ks * 8, bk, ID32, c != 0 ? 1u : 0u);
                            mma_tf32_ts(acc + ks * 32 + 16, a_lo + ks * 8, bk, ID16, 1u);
                        }
                    }
                    mma_commit(ta_empty(u));
                }
                mma_commit(acc_full(ab));
                if (gi == S.g1 - 1) FB_TRACE(7);
            }
            FB_TRACE(9);
        }
        __syncwarp();
    } else if (warp < FB_W_EPI) {
        // ===================== weight converters (thread <-> tile row <-> TMEM lane) =====================
        // two groups of four warps take alternate chunks, so two conversions are always in flight
        const int cg = (warp - FB_W_CONV) >> 2;
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t G = (uint32_t)(S.g1 * S.n1 + S.g2 * S.n2);
        for (uint32_t g = (uint32_t)cg; g < G; g += 2) {
            // NKV is a multiple of the ring depth: weight chunk g always lands in slot g % NS, phase g / NS of a_full
            const uint32_t s = g % FB_NS, u = g % FB_NTA;
            mbar_wait_warp(a_full(s), (g / FB_NS) & 1, lane);
            if (g == 0 && threadIdx.x == FB_W_CONV * 32) FB_TRACE(3);
            const uint8_t *arow = smem + FB_SM_A + s * TC_A_BYTES + r * 128;
            float4 w4[8];
            const bool conv = !(a.dbg & 2);
            if (conv) {
#pragma unroll
                for (int c = 0; c < 8; c++) w4[c] = *reinterpret_cast<const float4 *>(arow + ((c ^ (r & 7)) << 4));
            }
            mbar_wait_warp(ta_empty(u), ((g / FB_NTA) & 1) ^ 1, lane);
            fence_after_sync();
            const uint32_t a_hi = t_lane + FB_TMEM_A0 + u * 64, a_lo = a_hi + 32;
#pragma unroll
            for (int half = 0; half < (conv ? 2 : 0); half++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float wv[4] = {w4[half * 4 + c].x, w4[half * 4 + c].y, w4[half * 4 + c].z, w4[half * 4 + c].w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        // hi = rna_tf32(w) with integer ops; lo = w - hi exactly (the MMA ignores its low 13 bits)
                        const uint32_t h = (__float_as_uint(wv[e]) + 0x1000u) & 0xffffe000u;
                        hi[4 * c + e] = h;
                        lo[4 * c + e] = __float_as_uint(wv[e] - __uint_as_float(h));
                    }
                }
                tmem_st16(a_hi + half * 16, hi);
                tmem_st16(a_lo + half * 16, lo);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cnt(a_empty(s), 2);   // the raw tile is consumed (4 warps x 2 = 8 arrivals; the K/V
                                                             // chunks are released by the 8 worker warps x 1)
            wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(ta_full(u));
            if (threadIdx.x == FB_W_CONV * 32) { if (g == 0) FB_TRACE(4); if (g + 2 >= G) FB_TRACE(5); }
        }
    } else if (warp < FB_W_WORK) {
        // ===================== epilogue (thread <-> output feature of the tile) =====================
        const int ei = threadIdx.x - FB_W_EPI * 32;
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        float *part = reinterpret_cast<float *>(smem + SM_PART);
        const int groups = S.g1 + S.g2;
        const int p_idx = KIND == FB_ATT ? grp_id * 4 + (j & 3) : grp_id;
        for (int gi = 0; gi < groups; gi++) {
            const uint32_t ab = (uint32_t)gi & 1;
            mbar_wait_warp(acc_full(ab), ((uint32_t)gi >> 1) & 1, lane);
            if (ei == 0 && gi == 0) FB_TRACE(17);
            fence_after_sync();
            float v[16], vc[16];
#pragma unroll
            for (int b = 0; b < 16; b++) { v[b] = 0.f; vc[b] = 0.f; }
#pragma unroll
            for (int s4 = 0; s4 < 4; s4++) {
                uint32_t dreg[32];
                tmem_ld32(t_lane + ab * 128 + s4 * 32, dreg);
                wait_ld();
#pragma unroll
                for (int b = 0; b < 16; b++) { v[b] += __uint_as_float(dreg[b]); vc[b] += __uint_as_float(dreg[16 + b]); }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(ab));   // the MMA warp may start the next group in this buffer
#pragma unroll
            for (int b = 0; b < 16; b++) v[b] += vc[b];
            if (gi < S.g1) {
                // stage 1: K-split partial of this tile stays on chip for the cluster exchange
                if (KIND == FB_ATT) {
                    float *dst = part + gi * (16 * 128) + r;
#pragma unroll
                    for (int b = 0; b < 16; b++) dst[b * 128] = v[b];
                    if (gi == S.g1 - 1) {
                        asm volatile("bar.sync 3, 128;" ::: "memory");
                        if (ei < CS) mbar_arrive_remote_release(mapa(xbar1, (uint32_t)ei));
                        if (ei == 0) FB_TRACE(18);
                    }
                } else {
                    float *dst = a.hpart + ((size_t)(grp_id * FB_MLP_CS + j) * 16) * 128 + r;
#pragma unroll
                    for (int b = 0; b < 16; b++) __stcg(dst + b * 128, v[b]);
                    asm volatile("bar.sync 3, 128;" ::: "memory");
                    if (ei == 0) {
                        __threadfence();
                        atomicAdd(a.hflag + grp_id, 1u);
                        FB_TRACE(18);
                    }
                }
            } else {
                // stage 2: partial of the block output, reduced by resid_reduce_kernel
                float *dst = a.ws + ((size_t)p_idx * 16) * d + (size_t)(S.t2_begin + gi - S.g1) * 128 + r;
#pragma unroll
                for (int b = 0; b < 16; b++) __stcg(dst + (size_t)b * d, v[b]);
                if (ei == 0 && gi == groups - 1) FB_TRACE(19);
            }
        }
    } else {
        // ===================== workers: B operands, cluster exchange, attention =====================
        const int wi = threadIdx.x - FB_W_WORK * 32;     // 0..255
        const int wwarp = wi >> 5;
        const int i = wi & 127, xrow = i & 15, xq = i >> 4;
        constexpr int B1MAX = KIND == FB_ATT ? FB_ATT_B1_MAX : FB_MLP_B1_MAX;
        constexpr int NCH = B1MAX / 2;                   // chunks per thread (two groups of 128 threads alternate)
        float *lnp = reinterpret_cast<float *>(smem + FB_SM_LNP);
        {
            // LayerNorm gamma / beta of my k range are weights: staged BEFORE waiting for the activations
            const int nf4 = S.n1 * 8;
            if (i < nf4) {
                const float *src = (wi < 128 ? a.ln_g : a.ln_b) + S.c1_begin * TC_KC + 4 * i;
                *reinterpret_cast<float4 *>(lnp + (wi < 128 ? 0 : FB_MLP_B1_MAX * TC_KC) + 4 * i) = __ldg(reinterpret_cast<const float4 *>(src));
            }
        }
        // stage-1 B operand: LN(x)[16][my k range] -> [x_hi ; x_lo] chunks.  Run TWICE: a dry pass before the dependency
        // wait (its reads of x / stats may be stale, its output is overwritten) pulls the code into the instruction
        // cache and the LN parameters into shared memory while the CTA is idle anyway; the real pass follows the wait.
        auto build_b1 = [&](const bool real) {
            // all loads of the previous kernel's outputs are issued together (one L2 round trip)
            const int rr = wi >> 4, sub = wi & 15;
            const int n_tiles = a.n_stat_tiles;
            float2 sst[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int tl = sub + 16 * k;
                sst[k] = tl < n_tiles ? __ldcg(a.stats_in + tl * 16 + rr) : make_float2(0.f, 0.f);
            }
            float4 xv[NCH];
#pragma unroll
            for (int ci = 0; ci < NCH; ci++) {
                const int ch = (wi >> 7) + 2 * ci;
                xv[ci] = ch < S.n1 ? __ldcg(reinterpret_cast<const float4 *>(a.x + (size_t)xrow * d + (S.c1_begin + ch) * TC_KC + 4 * xq))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // (mean, rstd) per row: 16 consecutive lanes per row
            float n = 0.f, mean = 0.f, m2 = 0.f;
            const float w = (float)(d / n_tiles);
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (sub + 16 * k < n_tiles) chan_combine(n, mean, m2, w, sst[k].x, sst[k].y);
            asm volatile("" ::"f"(mean));
            if (real && wi == 0) FB_TRACE(35);
            asm volatile("" ::"f"(xv[0].x));
            if (real && wi == 0) FB_TRACE(36);
            asm volatile("" ::"f"(xv[NCH - 1].x));
            if (real && wi == 0) FB_TRACE(37);
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float nb = __shfl_xor_sync(0xffffffffu, n, o);
                const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
                const float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
                if ((sub & o) == 0) chan_combine(n, mean, m2, nb, mb, m2b);
                else { float tn = nb, tm = mb, t2 = m2b; chan_combine(tn, tm, t2, n, mean, m2); n = tn; mean = tm; m2 = t2; }
            }
            if (sub == 0) row_stats[rr] = make_float2(mean, 1.0f / sqrtf(m2 / (float)d + a.eps));
            if (real && wi == 0) FB_TRACE(32);
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (real && wi == 0) FB_TRACE(33);
            // ---- stage-1 B operand: LN(x)[16][my k range] -> [x_hi ; x_lo] chunks ----
            const float2 st = row_stats[xrow];
#pragma unroll
            for (int ci = 0; ci < NCH; ci++) {
                const int ch = (wi >> 7) + 2 * ci;
                if (ch < S.n1) {
                    const float4 g4 = *reinterpret_cast<const float4 *>(lnp + ch * TC_KC + 4 * xq);
                    const float4 b4 = *reinterpret_cast<const float4 *>(lnp + FB_MLP_B1_MAX * TC_KC + ch * TC_KC + 4 * xq);
                    const float xs[4] = {(xv[ci].x - st.x) * st.y * g4.x + b4.x, (xv[ci].y - st.x) * st.y * g4.y + b4.y,
                                         (xv[ci].z - st.x) * st.y * g4.z + b4.z, (xv[ci].w - st.x) * st.y * g4.w + b4.w};
                    uint32_t xh[4], xl[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) split_tf32(xs[e], xh[e], xl[e]);
                    uint8_t *bst = smem + FB_SM_B1 + ch * TC_B_BYTES + xq * 512;
                    *reinterpret_cast<uint4 *>(bst + xrow * 16) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
                    *reinterpret_cast<uint4 *>(bst + (16 + xrow) * 16) = make_uint4(xl[0], xl[1], xl[2], xl[3]);
                }
            }
            if (real && wi == 0) FB_TRACE(34);
            fence_proxy_async_smem();
            __syncwarp();
            if (real && lane == 0) mbar_arrive(b1_full);
        };
        if (!(a.dbg & 128)) build_b1(false);
        // everything below reads what the previous kernel produced
        pdl_wait();
        if (wi == 0) FB_TRACE(10);
        build_b1(true);
        if (wi == 0) FB_TRACE(11);
        // ---- the K-split partials of every CTA of the cluster are in shared memory ----
        if (KIND == FB_ATT) {
            mbar_wait_cluster_warp(xbar1, 0, lane);
        } else {
            if (wi == 0) {
                const unsigned target = (unsigned)FB_MLP_CS * (unsigned)(*a.step + 1);
                unsigned v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.hflag + grp_id) : "memory");
                } while (v < target);
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        if (wi == 0) FB_TRACE(12);
        const uint32_t part_addr = smem_base + SM_PART;
        if (KIND == FB_MLP) {
            // h = GELU(b1 + sum_p partial_p): all three CTAs of the tile form the same h, in the same order
            const int n0 = grp_id * 128;
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int idx = wi + it * FB_WORKERS;       // 0..511: (row b, 4 consecutive features)
                const int b = idx & 15, r4 = idx >> 4;
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 pv[FB_MLP_CS];
#pragma unroll
                for (int p = 0; p < FB_MLP_CS; p++)
                    pv[p] = __ldcg(reinterpret_cast<const float4 *>(a.hpart + ((size_t)(grp_id * FB_MLP_CS + p) * 16 + b) * 128 + 4 * r4));
#pragma unroll
                for (int p = 0; p < FB_MLP_CS; p++) { sum.x += pv[p].x; sum.y += pv[p].y; sum.z += pv[p].z; sum.w += pv[p].w; }
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.bias1 + n0 + 4 * r4));
                const float hs[4] = {gelu_erf(sum.x + bb.x), gelu_erf(sum.y + bb.y), gelu_erf(sum.z + bb.z), gelu_erf(sum.w + bb.w)};
                uint32_t xh[4], xl[4];
#pragma unroll
                for (int e = 0; e < 4; e++) split_tf32(hs[e], xh[e], xl[e]);
                uint8_t *bst = smem + SM_B2 + (r4 >> 3) * TC_B_BYTES + (r4 & 7) * 512;
                *reinterpret_cast<uint4 *>(bst + b * 16) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
                *reinterpret_cast<uint4 *>(bst + (16 + b) * 16) = make_uint4(xl[0], xl[1], xl[2], xl[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(b2_full);
            if (wi == 0) FB_TRACE(16);
        } else {
            float *qkv_own = reinterpret_cast<float *>(smem + FB_SM_ATT_QKV);
            float *y_own = reinterpret_cast<float *>(smem + FB_SM_ATT_Y);
            float *sc = reinterpret_cast<float *>(smem + FB_SM_ATT_SC);
            float *pvp = reinterpret_cast<float *>(smem + FB_SM_ATT_PV);
            float *inv_s = reinterpret_cast<float *>(smem + FB_SM_ATT_INV);
            {
                // q,k,v of my units = bias + sum over the 8 K-split partials (rank order)
                const int uu = wi >> 6, dim = wi & 63;
                const int hh = j >> 2;
#pragma unroll
                for (int which = 0; which < 3; which++) {
                    const uint32_t off = (uint32_t)(((which * 16 + b0 + uu) * 128 + hh * 64 + dim) * 4);
                    float pv[CS];
#pragma unroll
                    for (int p = 0; p < CS; p++) pv[p] = ld_dsmem_f1(mapa(part_addr + off, (uint32_t)p));
                    float s = 0.f;
#pragma unroll
                    for (int p = 0; p < CS; p++) s += pv[p];
                    s += __ldg(a.bias1 + (size_t)which * d + h * 64 + dim);
                    qkv_own[uu * 192 + which * 64 + dim] = s;
                    if (which == 1) a.kcache[cache_base + (size_t)uu * ustride + (size_t)t * 64 + dim] = s;
                    if (which == 2) a.vcache[cache_base + (size_t)uu * ustride + (size_t)t * 64 + dim] = s;
                }
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (wi == 0) FB_TRACE(13);
            {
                // ---- attention over the K / V chunks streaming through the ring: 16 lanes per key, 16 key groups;
                // every warp releases a chunk on its own (no CTA barrier per chunk) ----
                const int gk = wi >> 4, sub = wi & 15;
                const float scale = 0.125f;                 // 1 / sqrt(64)
                const int nk = t + 1;
                uint32_t qi = G1;
                for (int u = 0; u < FB_ATT_U; u++) {
                    const float4 qv = *reinterpret_cast<const float4 *>(qkv_own + u * 192 + 4 * sub);
                    for (int c = 0; c < nkc; c++, qi++) {
                        const uint32_t s = qi % FB_NS;
                        mbar_wait_warp(kv_full(s), ((qi - G1) / FB_NS) & 1, lane);
                        if (wi == 0 && qi - G1 < 8) FB_TRACE(24 + (qi - G1));
                        const int valid = t - 64 * c < 64 ? t - 64 * c : 64;
                        const uint8_t *slot = smem + FB_SM_A + s * TC_A_BYTES;
                        float sd[4];
#pragma unroll
                        for (int r4 = 0; r4 < 4; r4++) {
                            const int key = r4 * 16 + gk;
                            const int kk = key < valid ? key : valid - 1;       // clamped: loads stay unconditional
                            const float4 k4 = *reinterpret_cast<const float4 *>(slot + kk * 256 + sub * 16);
                            sd[r4] = qv.x * k4.x + qv.y * k4.y + qv.z * k4.z + qv.w * k4.w;
                        }
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
                            for (int r4 = 0; r4 < 4; r4++) sd[r4] += __shfl_xor_sync(0xffffffffu, sd[r4], o, 16);
                        if (sub == 0) {
#pragma unroll
                            for (int r4 = 0; r4 < 4; r4++)
                                if (r4 * 16 + gk < valid) sc[u * FB_ATT_T + 64 * c + r4 * 16 + gk] = sd[r4] * scale;
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a_empty(s));
                    }
                }
                if (wi == 0) FB_TRACE(21);
                if (wi < 16 * FB_ATT_U) {
                    // this step's own key
                    const int u = wi >> 4;
                    const float4 qv = *reinterpret_cast<const float4 *>(qkv_own + u * 192 + 4 * sub);
                    const float4 k4 = *reinterpret_cast<const float4 *>(qkv_own + u * 192 + 64 + 4 * sub);
                    float sdot = qv.x * k4.x + qv.y * k4.y + qv.z * k4.z + qv.w * k4.w;
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o, 16);
                    if (sub == 0) sc[u * FB_ATT_T + t] = sdot * scale;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (wwarp < FB_ATT_U) {
                    // softmax: warp u owns unit u; each lane first folds its own 8 positions, then one warp reduction
                    float *s1 = sc + wwarp * FB_ATT_T;
                    float e8[FB_ATT_T / 32];
                    float m = -INFINITY;
#pragma unroll
                    for (int w8 = 0; w8 < FB_ATT_T / 32; w8++) {
                        const int jj = 32 * w8 + lane;
                        e8[w8] = jj < nk ? s1[jj] : -INFINITY;
                        m = fmaxf(m, e8[w8]);
                    }
                    m = warp_max(m);
                    float sum = 0.f;
#pragma unroll
                    for (int w8 = 0; w8 < FB_ATT_T / 32; w8++) {
                        const int jj = 32 * w8 + lane;
                        const float e = jj < nk ? expf(e8[w8] - m) : 0.f;
                        if (jj < nk) s1[jj] = e;
                        sum += e;
                    }
                    sum = warp_sum(sum);
                    if (lane == 0) inv_s[wwarp] = 1.0f / sum;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (wi == 0) FB_TRACE(22);
                for (int u = 0; u < FB_ATT_U; u++) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float inv = inv_s[u];
                    for (int c = 0; c < nkc; c++, qi++) {
                        const uint32_t s = qi % FB_NS;
                        mbar_wait_warp(kv_full(s), ((qi - G1) / FB_NS) & 1, lane);
                        const int valid = t - 64 * c < 64 ? t - 64 * c : 64;
                        const uint8_t *slot = smem + FB_SM_A + s * TC_A_BYTES;
                        float pr[4];
                        float4 v4[4];
#pragma unroll
                        for (int r4 = 0; r4 < 4; r4++) {
                            const int key = r4 * 16 + gk;
                            const int kk = key < valid ? key : valid - 1;
                            pr[r4] = key < valid ? sc[u * FB_ATT_T + 64 * c + kk] * inv : 0.f;
                            v4[r4] = *reinterpret_cast<const float4 *>(slot + kk * 256 + sub * 16);
                        }
#pragma unroll
                        for (int r4 = 0; r4 < 4; r4++) {
                            acc.x += pr[r4] * v4[r4].x; acc.y += pr[r4] * v4[r4].y; acc.z += pr[r4] * v4[r4].z; acc.w += pr[r4] * v4[r4].w;
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(a_empty(s));
                    }
                    *reinterpret_cast<float4 *>(pvp + (u * 16 + gk) * 64 + 4 * sub) = acc;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
                {
                    const int u = wi >> 6, dim = wi & 63;
                    float o = 0.f;
#pragma unroll
                    for (int gI = 0; gI < 16; gI++) o += pvp[(u * 16 + gI) * 64 + dim];
                    o += sc[u * FB_ATT_T + t] * inv_s[u] * qkv_own[u * 192 + 128 + dim];
                    y_own[u * 64 + dim] = o;
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            // y of my units is complete (fb_attention ends with a barrier): tell every CTA of the cluster
            if (wi == 0) FB_TRACE(14);
            if (wi < CS) mbar_arrive_remote_release(mapa(xbar2, (uint32_t)wi));
            mbar_wait_cluster_warp(xbar2, 0, lane);
            if (wi == 0) FB_TRACE(15);
            if (wi < 128) {
                // stage-2 B operand: y_pair[16 rows][k slice c of the pair's 128] pulled from the CTAs that own the rows
                const int c = j & 3, hsrc = c >> 1;
                const uint32_t peer = (uint32_t)(4 * hsrc + (xrow >> 2));
                const uint32_t off = (uint32_t)(((xrow & 3) * 64 + 32 * (c & 1) + 4 * xq) * 4);
                const float4 yv = ld_dsmem_f4(mapa(smem_base + FB_SM_ATT_Y + off, peer));
                const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
                uint32_t xh[4], xl[4];
#pragma unroll
                for (int e = 0; e < 4; e++) split_tf32(ys[e], xh[e], xl[e]);
                uint8_t *bst = smem + SM_B2 + xq * 512;
                *reinterpret_cast<uint4 *>(bst + xrow * 16) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
                *reinterpret_cast<uint4 *>(bst + (16 + xrow) * 16) = make_uint4(xl[0], xl[1], xl[2], xl[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(b2_full);
            if (wi == 0) FB_TRACE(16);
        }
    }

    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // peers may still be reading this CTA's shared memory
    if (KIND == FB_ATT) cluster_sync_all();
    if (threadIdx.x == 0) FB_TRACE(20);
    if (trc_all && threadIdx.x == 0) {
        const unsigned long long now = fb_gtime();
        atomicMin(trc_all + 42, now);
        atomicMax(trc_all + 43, now);
    }
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// x[b][n] += bias[n] + sum_p ws[p][b][n]  (fixed order: four contiguous p ranges summed in order, then the four range
// sums in order -> deterministic);  emits (mean, M2) of each row over this CTA's 32 columns for the LayerNorm of the
// next kernel.  grid d / 32, 512 threads: thread <-> (p range, row, 4 consecutive columns); every load of a thread is
// in flight at once (the kernel is pure L2 latency).
constexpr int RR_THREADS = 512;
__global__ void __launch_bounds__(RR_THREADS) resid_reduce_kernel(const float *__restrict__ ws, int P, const float *__restrict__ bias,
                                                                  float *__restrict__ x, float2 *__restrict__ stats_out, int d) {
    using namespace tc05;
    __shared__ float4 red[3][128];
    if (threadIdx.x == 0) pdl_launch_dependents();
    const int pg = threadIdx.x >> 7, i = threadIdx.x & 127;
    const int b = i >> 3, c4 = i & 7;
    const int n = blockIdx.x * 32 + 4 * c4;
    const int p0 = P * pg / 4, p1 = P * (pg + 1) / 4;
    const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + n));
    pdl_wait();
    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pg == 0) r4 = __ldcg(reinterpret_cast<const float4 *>(x + (size_t)b * d + n));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *src = ws + (size_t)b * d + n;
    const size_t pstride = (size_t)16 * d;
    for (int p = p0; p < p1; p += 12) {
        float4 pv[12];
#pragma unroll
        for (int u = 0; u < 12; u++)
            pv[u] = p + u < p1 ? __ldcg(reinterpret_cast<const float4 *>(src + (size_t)(p + u) * pstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 12; u++) { v.x += pv[u].x; v.y += pv[u].y; v.z += pv[u].z; v.w += pv[u].w; }
    }
    if (pg > 0) red[pg - 1][i] = v;
    __syncthreads();
    if (pg > 0) return;
#pragma unroll
    for (int g = 0; g < 3; g++) { const float4 o = red[g][i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    v.x = r4.x + (v.x + bb.x); v.y = r4.y + (v.y + bb.y); v.z = r4.z + (v.z + bb.z); v.w = r4.w + (v.w + bb.w);
    *reinterpret_cast<float4 *>(x + (size_t)b * d + n) = v;
    float s = v.x + v.y + v.z + v.w;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 32.0f);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (c4 == 0) stats_out[blockIdx.x * 16 + b] = make_float2(mean, q);
}

}  // namespace wmar
