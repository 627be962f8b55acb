// tcgen05 skinny GEMM for the decode step:  Y[16][N] = pro(X)[16][K] . W[N][K]^T (+ bias, epilogue), fp32 in/out.
//
// Same contract as skinny_gemm_kernel (gemm.cuh) -- it replaces it for every nn.Linear of the per-token step whose
// N is a multiple of 128 and K a multiple of 32 (mingpt.py:53-60,105-110,142; rar.py:75,79,127-129,173-176,232) --
// but built for Blackwell:
//   * W is the UMMA *A* operand (M = 128 weight rows per CTA), the 16 batch rows are the N side.  A TMA producer
//     thread streams 128 x 32 fp32 tiles of W (16 KB, SWIZZLE_128B tensor map, L2 evict-first) into a 6-deep
//     shared-memory ring, waiting only on ring slots -- never on the activations -- so under programmatic
//     dependent launch the next GEMM's weights are already in flight while the previous kernel finishes.
//   * fp32 fidelity comes from the 3xTF32 split done on chip: four converter warps read each landed tile once
//     (conflict-free through the 128 B swizzle), form hi = rna_tf32(w) and lo = rna_tf32(w - hi) and write both to
//     TENSOR MEMORY with tcgen05.st (thread <-> TMEM lane <-> weight row), so the MMA takes A from TMEM and the
//     shared-memory port only carries each weight byte twice (TMA write + one read).
//   * the activations X (LayerNorm / adaLN applied on the fly from the producer's (mean, M2) partials) are split
//     the same way into a 32-row B operand [x_hi ; x_lo] per 32-k chunk; one elected thread issues, per k8 step,
//     tcgen05.mma.kind::tf32  D[128 x 32] += W_hi . [x_hi | x_lo]^T  and  D[128 x 16] += W_lo . x_hi^T.
//     The accumulator lives in TMEM; the epilogue adds the two column halves (hi.hi + lo.hi) + (hi.lo).
//   * split-K over CTAs with the same deterministic last-arriver reduction, bias / GELU / residual / gate epilogues
//     and (mean, M2) LayerNorm statistics per 64 columns as the mma.sync kernel, so it is a drop-in.
#pragma once
#include "gemm.cuh"
#include "tc05.cuh"

namespace wmar {

constexpr int TC_TILE_N = 128;                       // W rows per CTA = UMMA M
constexpr int TC_KC = 32;                            // k per chunk: 128 B of fp32 = one swizzle row
constexpr int TC_NS = 6;                             // shared-memory stages of raw W tiles
constexpr int TC_NT = 2;                             // TMEM stages of split W (hi|lo) + smem stages of split X
constexpr int TC_THREADS = 192;                      // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 convert/epilogue
constexpr int TC_CONV_THREADS = 128;
constexpr int TC_A_BYTES = TC_TILE_N * TC_KC * 4;    // 16384
constexpr int TC_B_BYTES = 32 * TC_KC * 4;           // 4096: [k/4 (8)][row (32)][4 floats]
constexpr int TC_TMEM_COLS = 256;                    // 4 x 32 accumulator + 2 x 64 operand columns
constexpr int TC_NACC = 4;                           // accumulator sets (one per k8 step of a chunk): the tensor core
                                                     // truncates on accumulate, so chains are kept short and the
                                                     // sets are summed in fp32 (round-to-nearest) by the epilogue
constexpr int TC_ACC_COLS = 32 * TC_NACC;
constexpr int TC_SMEM_A = 0;
constexpr int TC_SMEM_B = TC_SMEM_A + TC_NS * TC_A_BYTES;
constexpr int TC_SMEM_BAR = TC_SMEM_B + TC_NT * TC_B_BYTES;          // a_full[NS] a_empty[NS] t_full[NT] t_empty[NT] acc
constexpr int TC_N_BARS = 2 * TC_NS + 2 * TC_NT + 1;
constexpr int TC_SMEM_MISC = TC_SMEM_BAR + 8 * TC_N_BARS;            // tmem ptr, is_last flag
constexpr int TC_SMEM_ROWSTATS = TC_SMEM_MISC + 16;                  // float2[16]
constexpr int TC_SMEM_STATW = TC_SMEM_ROWSTATS + 16 * 8;             // float2[4 warps][16 rows]
constexpr int TC_SMEM_BYTES = TC_SMEM_STATW + 4 * 16 * 8;
constexpr int TC_SMEM_ALLOC = TC_SMEM_BYTES + 1024;                  // slack for the 1024 B alignment of the ring

struct TcExtra {
    int splits;       // CTAs along K (grid.y)
    int dbg;          // probe only: bit0 skip MMA issue, bit1 skip weight conversion, bit2 skip TMA (results are wrong)
};

// (mean, rstd) per row from the producer's per-64-column (mean, M2) partials; 8 consecutive lanes per row.
__device__ __forceinline__ void tc_combine_row_stats(const float2 *__restrict__ stats_in, int n_tiles, int K, float eps,
                                                     float2 *row_stats, int i) {
    const int r = i >> 3, sub = i & 7;
    float n = 0.f, mean = 0.f, m2 = 0.f;
    const float w = (float)(K / n_tiles);
    for (int tl = sub; tl < n_tiles; tl += 8) {
        float2 s = __ldcg(stats_in + tl * 16 + r);   // not hoistable above griddepcontrol.wait
        chan_combine(n, mean, m2, w, s.x, s.y);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        float nb = __shfl_xor_sync(0xffffffffu, n, o);
        float mb = __shfl_xor_sync(0xffffffffu, mean, o);
        float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
        if ((sub & o) == 0) chan_combine(n, mean, m2, nb, mb, m2b);
        else { float tn = nb, tm = mb, t2 = m2b; chan_combine(tn, tm, t2, n, mean, m2); n = tn; mean = tm; m2 = t2; }
    }
    if (sub == 0) row_stats[r] = make_float2(mean, 1.0f / sqrtf(m2 / (float)K + eps));
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmapW, const GemmArgs a, const TcExtra ex) {
    using namespace tc05;
    extern __shared__ uint8_t tc_smem_raw[];
    const uint32_t smem_base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = tc_smem_raw + (smem_base - smem_u32(tc_smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int C = a.K / TC_KC;
    const int c_begin = (int)(((long long)split * C) / ex.splits);
    const int c_end = (int)(((long long)(split + 1) * C) / ex.splits);
    const int nchunks = c_end - c_begin;

    const uint32_t bar0 = smem_base + TC_SMEM_BAR;
    auto a_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (uint32_t)(TC_NS + s); };
    auto t_full = [&](int u) { return bar0 + 8u * (uint32_t)(2 * TC_NS + u); };
    auto t_empty = [&](int u) { return bar0 + 8u * (uint32_t)(2 * TC_NS + TC_NT + u); };
    const uint32_t acc_full = bar0 + 8u * (uint32_t)(2 * TC_NS + 2 * TC_NT);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + TC_SMEM_MISC);
    int *s_is_last = reinterpret_cast<int *>(smem + TC_SMEM_MISC + 4);
    float2 *row_stats = reinterpret_cast<float2 *>(smem + TC_SMEM_ROWSTATS);
    float2 *stat_w = reinterpret_cast<float2 *>(smem + TC_SMEM_STATW);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmapW);
        for (int s = 0; s < TC_NS; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 4); }
        for (int u = 0; u < TC_NT; u++) { mbar_init(t_full(u), 4); mbar_init(t_empty(u), 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<TC_TMEM_COLS>(smem_u32(s_tmem));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *s_tmem;
    // the dependent kernel (next GEMM of the step) may become resident now and start streaming ITS weights
    if (threadIdx.x == 0) pdl_launch_dependents();

    if (warp == 0) {
        // ===== TMA producer: weights only, never waits for the previous kernel =====
        if (lane == 0) {
            for (int j = 0; j < nchunks; j++) {
                const int s = j % TC_NS;
                mbar_wait(a_empty(s), ((j / TC_NS) & 1) ^ 1);
                if (ex.dbg & 4) { mbar_arrive(a_full(s)); continue; }
                mbar_arrive_expect_tx(a_full(s), TC_A_BYTES);
                tma_load_2d(smem_base + TC_SMEM_A + s * TC_A_BYTES, &tmapW, (c_begin + j) * TC_KC, tile * TC_TILE_N,
                            a_full(s), L2_EVICT_FIRST);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t ID32 = idesc_tf32_m128(32), ID16 = idesc_tf32_m128(16);
            for (int j = 0; j < nchunks; j++) {
                const int u = j % TC_NT;
                mbar_wait(t_full(u), (j / TC_NT) & 1);
                fence_after_sync();
                const uint32_t a_hi = tmem + TC_ACC_COLS + u * 64, a_lo = a_hi + 32;
                const uint64_t bd = smem_desc_kmajor_noswz(smem_base + TC_SMEM_B + u * TC_B_BYTES, 512, 128);
#pragma unroll
                for (int ks = 0; ks < ((ex.dbg & 1) ? 0 : TC_NACC); ks++) {
                    const uint64_t bk = bd + (uint64_t)((ks * 1024) >> 4);
                    // columns [0,16) of set ks: W_hi.x_hi ; columns [16,32): the two small cross terms
                    mma_tf32_ts(tmem + ks * 32, a_hi + ks * 8, bk, ID32, j != 0 ? 1u : 0u);
                    mma_tf32_ts(tmem + ks * 32 + 16, a_lo + ks * 8, bk, ID16, 1u);
                }
                mma_commit(t_empty(u));
            }
            mma_commit(acc_full);
        }
        __syncwarp();
    } else {
        // ===== converter / B builder / epilogue: 128 threads, thread <-> TMEM lane <-> weight row of the tile =====
        const int i = threadIdx.x - 64;                 // 0..127
        const int q = warp & 3;                         // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;                    // tile row == TMEM lane
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const int xrow = i & 15, xq = i >> 4;           // activation element block: row, 16-byte k group (0..7)
        // everything below reads what the previous kernel produced
        pdl_wait();
        if (PRO != PRO_NONE) {
            tc_combine_row_stats(a.stats_in, a.n_stat_tiles, a.K, a.eps, row_stats, i);
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        float2 st = make_float2(0.f, 1.f);
        if (PRO != PRO_NONE) st = row_stats[xrow];
        const float *xp = a.X + (size_t)xrow * a.ldx + (size_t)c_begin * TC_KC + 4 * xq;
        constexpr int PF = 4;                           // activation chunks kept in flight (registers)
        float4 xr[PF];
#pragma unroll
        for (int p = 0; p < PF; p++)
            xr[p] = p < nchunks ? __ldcg(reinterpret_cast<const float4 *>(xp + p * TC_KC)) : make_float4(0.f, 0.f, 0.f, 0.f);

        for (int j0 = 0; j0 < nchunks; j0 += PF) {
#pragma unroll
            for (int p = 0; p < PF; p++) {
                const int j = j0 + p;
                if (j < nchunks) {
                    const int s = j % TC_NS, u = j % TC_NT;
                    // ---- activations of this chunk -> [x_hi ; x_lo] (issued first: independent of the weight tile) ----
                    float xs[4] = {xr[p].x, xr[p].y, xr[p].z, xr[p].w};
                    if (j + PF < nchunks) xr[p] = __ldcg(reinterpret_cast<const float4 *>(xp + (j + PF) * TC_KC));
                    if (PRO != PRO_NONE) {
                        const int k = (c_begin + j) * TC_KC + 4 * xq;
                        float gm[4] = {1.f, 1.f, 1.f, 1.f}, bt[4] = {0.f, 0.f, 0.f, 0.f};
                        if (a.ln_g != nullptr) {
                            float4 g4 = __ldg(reinterpret_cast<const float4 *>(a.ln_g + k));
                            float4 b4 = __ldg(reinterpret_cast<const float4 *>(a.ln_b + k));
                            gm[0] = g4.x; gm[1] = g4.y; gm[2] = g4.z; gm[3] = g4.w;
                            bt[0] = b4.x; bt[1] = b4.y; bt[2] = b4.z; bt[3] = b4.w;
                        }
#pragma unroll
                        for (int e = 0; e < 4; e++) xs[e] = (xs[e] - st.x) * st.y * gm[e] + bt[e];
                        if (PRO == PRO_ADALN) {
                            float4 sc = __ldcg(reinterpret_cast<const float4 *>(a.mod_scale + (size_t)xrow * a.ld_mod + k));
                            float4 sh = __ldcg(reinterpret_cast<const float4 *>(a.mod_shift + (size_t)xrow * a.ld_mod + k));
                            xs[0] = xs[0] * (1.f + sc.x) + sh.x; xs[1] = xs[1] * (1.f + sc.y) + sh.y;
                            xs[2] = xs[2] * (1.f + sc.z) + sh.z; xs[3] = xs[3] * (1.f + sc.w) + sh.w;
                        }
                    }
                    uint32_t xh[4], xl[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) split_tf32(xs[e], xh[e], xl[e]);

                    mbar_wait(t_empty(u), ((j / TC_NT) & 1) ^ 1);   // MMAs that read this TMEM/B stage are done
                    fence_after_sync();
                    uint8_t *bst = smem + TC_SMEM_B + u * TC_B_BYTES + xq * 512;
                    *reinterpret_cast<uint4 *>(bst + xrow * 16) = make_uint4(xh[0], xh[1], xh[2], xh[3]);
                    *reinterpret_cast<uint4 *>(bst + (16 + xrow) * 16) = make_uint4(xl[0], xl[1], xl[2], xl[3]);

                    // ---- weight tile: smem (swizzled rows) -> hi/lo -> tensor memory ----
                    mbar_wait(a_full(s), (j / TC_NS) & 1);
                    const uint8_t *arow = smem + TC_SMEM_A + s * TC_A_BYTES + r * 128;
                    const uint32_t a_hi = t_lane + TC_ACC_COLS + u * 64, a_lo = a_hi + 32;
#pragma unroll
                    for (int half = 0; half < ((ex.dbg & 2) ? 0 : 2); half++) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const int chunk = half * 4 + c;
                            const float4 w4 = *reinterpret_cast<const float4 *>(arow + ((chunk ^ (r & 7)) << 4));
                            split_tf32(w4.x, hi[4 * c + 0], lo[4 * c + 0]);
                            split_tf32(w4.y, hi[4 * c + 1], lo[4 * c + 1]);
                            split_tf32(w4.z, hi[4 * c + 2], lo[4 * c + 2]);
                            split_tf32(w4.w, hi[4 * c + 3], lo[4 * c + 3]);
                        }
                        tmem_st16(a_hi + half * 16, hi);
                        tmem_st16(a_lo + half * 16, lo);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_empty(s));         // the raw tile is consumed: TMA may refill it
                    wait_st();
                    fence_proxy_async_smem();
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(t_full(u));
                }
            }
        }

        // ---- accumulator -> registers ----
        mbar_wait(acc_full, 0);
        fence_after_sync();
        float v[16], vc[16];
#pragma unroll
        for (int b = 0; b < 16; b++) { v[b] = 0.f; vc[b] = 0.f; }
#pragma unroll
        for (int s4 = 0; s4 < TC_NACC; s4++) {
            uint32_t dreg[32];
            tmem_ld32(t_lane + s4 * 32, dreg);
            wait_ld();
#pragma unroll
            for (int b = 0; b < 16; b++) { v[b] += __uint_as_float(dreg[b]); vc[b] += __uint_as_float(dreg[16 + b]); }
        }
#pragma unroll
        for (int b = 0; b < 16; b++) v[b] += vc[b];

        bool run_epilogue = true;
        if (ex.splits > 1) {
            float *wst = a.ws + ((size_t)tile * ex.splits) * (GEMM_M * TC_TILE_N);
            float *mine = wst + (size_t)split * (GEMM_M * TC_TILE_N);
#pragma unroll
            for (int b = 0; b < 16; b++) __stcg(mine + b * TC_TILE_N + r, v[b]);
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (i == 0) {
                unsigned old = atomicAdd(&a.counters[tile], 1u);
                const int last = (old == (unsigned)(ex.splits - 1));
                if (last) a.counters[tile] = 0u;
                *s_is_last = last;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            run_epilogue = (*s_is_last != 0);
            if (run_epilogue) {
                __threadfence();
#pragma unroll
                for (int b = 0; b < 16; b++) v[b] = 0.f;
                for (int sp = 0; sp < ex.splits; sp++) {
                    const float *src = wst + (size_t)sp * (GEMM_M * TC_TILE_N) + r;
#pragma unroll
                    for (int b = 0; b < 16; b++) v[b] += __ldcg(src + b * TC_TILE_N);
                }
            }
        }
        if (run_epilogue) {
            const int n = tile * TC_TILE_N + r;
            const float bias = a.bias != nullptr ? __ldg(a.bias + n) : 0.f;
#pragma unroll
            for (int b = 0; b < 16; b++) {
                float y = v[b] + bias;
                if (EPI == EPI_GELU) y = gelu_erf(y);
                if (EPI == EPI_GATE_RESID) y *= a.gate[(size_t)b * a.ld_gate + n];
                if (EPI == EPI_RESID || EPI == EPI_GATE_RESID) y = a.resid[(size_t)b * a.ld_resid + n] + y;
                a.Y[(size_t)b * a.ldy + n] = y;
                v[b] = y;
            }
            if (a.stats_out != nullptr) {
                // (mean, M2) per row over 64-column tiles: per-warp 32-column partials, then one exact pair merge
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    const float sum = warp_sum(v[b]);
                    const float mean = sum * (1.0f / 32.0f);
                    const float dv = v[b] - mean;
                    const float m2 = warp_sum(dv * dv);
                    if (lane == 0) stat_w[q * 16 + b] = make_float2(mean, m2);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (i < 32) {
                    const int half = i >> 4, b = i & 15;   // 64-column tile `half` of this 128-column tile
                    const float2 p0 = stat_w[(2 * half) * 16 + b], p1 = stat_w[(2 * half + 1) * 16 + b];
                    const float mean = 0.5f * (p0.x + p1.x);
                    const float d = p1.x - p0.x;
                    a.stats_out[(tile * 2 + half) * GEMM_M + b] = make_float2(mean, p0.y + p1.y + d * d * 16.0f);
                }
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 1) tmem_dealloc<TC_TMEM_COLS>(tmem);
}

// host side (gemm_tc.cu)
bool tc_gemm_eligible(const GemmArgs &a);
int tc_pick_splits(int N, int K, int n_sms);
int launch_tc_gemm(int pro, int epi, const GemmArgs &a, cudaStream_t stream);
int tc_weight_map(const float *W, int N, int K, CUtensorMap *out);   // cached TMA descriptor of W[N][K]
bool tc_available();                                                // driver exposes cuTensorMapEncodeTiled
int tc_weight_map_skinny(const float *W, int N, int K, CUtensorMap *out);   // fp32 W[N][K], box 16 x 64, no swizzle (gemm.cuh, TMA ring)
int tc_weight_map_bf16(const void *W, int N, int K, CUtensorMap *out);   // bf16 W[N][K], box 64 x 128 (conv_tc.cuh, bf16x3)
int tc_nhwc_map(const float *base, int B, int H, int W, int C, int bw, int bh, CUtensorMap *out, int estride = 1);   // conv_tc.cuh A operand

}  // namespace wmar
