// Skinny GEMM for the decode step:  Y[16][N] = pro(X)[16][K] . W[N][K]^T (+ bias, epilogue),  fp32 in / fp32 out.
//
// Replaces every nn.Linear of the per-token step (mingpt.py:53-60,105-110,142; rar.py:75,79,127-129,173-176,232).
// The step is HBM-bound on W (each weight is used for only 16 rows), so the kernel is built around streaming W once
// with as many 16-byte loads in flight as possible and doing the math on the tensor pipe:
//   * the 16 batch rows are exactly the M of mma.m16n8k8; W rows map to the MMA's n, and every lane fetches exactly
//     the 16-byte pieces that are ITS B fragments (one piece covers two k8 steps of one n8 tile).  The pieces travel
//     through a per-warp shared-memory ring (3 stages x 4 KB per warp): shared memory is used purely as extra load-queue
//     depth -- up to three iterations of every warp are in flight at once (96 KB per CTA, 192 KB per SM) instead of the
//     two a register double buffer allows, and a lane only ever reads back its own pieces, so there is no CTA barrier and
//     no bank conflict.  The ring is filled by the TMA unit (default since round 2: one cp.async.bulk.tensor.2d box of
//     16 k x 64 rows per warp iteration, completion on a per-(warp, stage) mbarrier) or by per-lane cp.async (round 1,
//     WMAR_GEMM_LOAD=cpasync; measured equal -- the weight stream is not on the critical path),
//   * products are 3xTF32 (hi*hi + hi*lo + lo*hi with fp32 accumulation): fp32-faithful results, which the
//     reference's fp32 (TF32-off) Linear layers require for greedy token parity,
//   * split-K across CTAs (grid.y) fills the 148 SMs even for N = 1536 -- tiles x splits is kept within ONE wave of the
//     2 x 148 resident CTAs (pick_splits, gemm.cu), K is cut into near-equal parts; partial tiles go to an L2-resident workspace
//     as {value, flag} words and the CTA of the last split sums them in split order (deterministic) and runs the
//     epilogue -- no fence, no atomic (GemmArgs::ll_salt; the counter-based last-arriver hand-off is kept as ll_salt = 0),
//   * LayerNorm / adaLN-modulate are applied to X on the fly (prologue) from per-row statistics that the producing
//     GEMM's epilogue emitted as (mean, M2) partials per 64-column tile, combined with Chan's formula.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace wmar {

constexpr int GEMM_THREADS = 256;
constexpr int GEMM_WARPS = GEMM_THREADS / 32;
constexpr int GEMM_NT = 64;      // W rows (output columns) per CTA = 8 n8 tiles
constexpr int GEMM_TILES = GEMM_NT / 8;
constexpr int GEMM_KI = 16;      // k per warp iteration (one LDG.128 per lane per n8 tile)
constexpr int GEMM_M = 16;       // batch rows
constexpr int GEMM_RED_LD = 72;  // padded row of the cross-warp reduction buffer
constexpr int GEMM_STAGES = 3;   // cp.async ring depth per warp
constexpr int GEMM_STAGE_BYTES = GEMM_TILES * 32 * 16;                         // 4 KB: [n8 tile][lane][16 B]
constexpr int GEMM_RING_BYTES = GEMM_WARPS * GEMM_STAGES * GEMM_STAGE_BYTES;   // 96 KB per CTA
constexpr int GEMM_RED_BYTES = GEMM_WARPS * GEMM_M * GEMM_RED_LD * 4;          // aliases the ring after the main loop
constexpr int GEMM_SMEM_BYTES = GEMM_RING_BYTES > GEMM_RED_BYTES ? GEMM_RING_BYTES : GEMM_RED_BYTES;

enum GemmPrologue { PRO_NONE = 0, PRO_LN = 1, PRO_ADALN = 2 };
enum GemmEpilogue { EPI_STORE = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_GATE_RESID = 3 };

struct GemmArgs {
    const float *X; int ldx;
    const float *W;
    const float *bias;
    float *Y; int ldy;
    int N, K, splits;
    // prologue
    const float *ln_g, *ln_b;     // may be null for PRO_ADALN without affine (RAR final layer)
    const float2 *stats_in;       // [n_stat_tiles][16] (mean, M2) of 64-wide column tiles of X
    int n_stat_tiles;
    float eps;
    const float *mod_scale, *mod_shift; int ld_mod;  // adaLN: x*(1+scale[m][k]) + shift[m][k]
    // epilogue
    const float *resid; int ld_resid;
    const float *gate; int ld_gate;
    float2 *stats_out;            // [N/64][16] or null
    float *ws;                    // [N/64][splits][16*64]
    unsigned *counters;           // [N/64], zero-initialised, self-resetting
    // flag-carrying split-K hand-off (ll_salt != 0): every 8-byte word of a partial is {value, flag}, written with
    // one 64-bit store (single-copy atomic), so the partial needs no fence and no counter: the CTA of the LAST split
    // (dispatched after the others) polls the words of the other splits until they carry this launch's flag and sums
    // them in split order.  flag = (*ll_epoch + 1) * 1024 + ll_salt must be unique among all launches that have written
    // into ws since it was last cleared (the engines use their token-step counter and the launch index within the step,
    // and clear ws at the start of every generation).  ll_salt == 0: counter-based hand-off (fence + atomic).
    const int *ll_epoch; unsigned ll_epoch_val; unsigned ll_salt;   // epoch = *ll_epoch, or ll_epoch_val when the pointer is null
    unsigned long long *trace;    // probe only: [16] globaltimer stamps of CTA (0,0) / of the last arriver of tile 0
};
#define GM_TRACE(e) do { if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && (blockIdx.y == 0 || (e) >= 8)) { unsigned long long v_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v_)); a.trace[e] = v_; } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// 16-byte asynchronous copy global -> shared, L2 only (the weights are streamed once)
__device__ __forceinline__ void gm_cp16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void gm_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void gm_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
    hi = tf32_hi(x);
    lo = tf32_hi(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ unsigned long long ll_pack(float v, unsigned flag) {
    return (unsigned long long)__float_as_uint(v) | ((unsigned long long)flag << 32);
}
__device__ __forceinline__ void ll_store2(unsigned long long *p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ll_load2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Chan et al. pairwise combination of (count, mean, M2)
__device__ __forceinline__ void chan_combine(float &n, float &mean, float &m2, float nb, float mb, float m2b) {
    if (nb == 0.f) return;
    float nn = n + nb;
    float d = mb - mean;
    mean = mean + d * (nb / nn);
    m2 = m2 + m2b + d * d * (n * nb / nn);
    n = nn;
}

// Per-row (mean, rstd) of X from the producer's tile partials; result in smem row_stats[16] (x = mean, y = rstd).
// ONE warp per CTA does it: lane = (row, tile parity), every load of the lane in flight at once.  All CTAs of the grid
// read the same few cache lines (n_tiles x 128 B) at the same moment right after the dependency resolves; with every
// warp of every CTA loading them the L2 slices that own those lines serialised ~2000 requests per line (measured:
// 4.4 us from griddepcontrol.wait to the statistics, profiles/r01_final_summary.md) -- one warp per CTA is 8x fewer.
__device__ __forceinline__ void combine_row_stats(const float2 *__restrict__ stats_in, int n_tiles, int K, float eps,
                                                  float2 *row_stats) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x, r = lane & 15, half = lane >> 4;
    constexpr int MAXT = 12;                      // tiles per lane held in registers (n_tiles <= 24 in one pass)
    float n = 0.f, mean = 0.f, m2 = 0.f;
    const float w = (float)(K / n_tiles);
    for (int t0 = 0; t0 < n_tiles; t0 += 2 * MAXT) {
        float2 sv[MAXT];
#pragma unroll
        for (int k = 0; k < MAXT; k++) {
            const int tl = t0 + half + 2 * k;
            sv[k] = tl < n_tiles ? __ldcg(stats_in + tl * 16 + r) : make_float2(0.f, 0.f);   // not hoistable above the wait
        }
        // equal-sized tiles: mean of the tile means, then M2 = sum M2_i + w * sum (mean_i - mean)^2 (no divisions)
        float ms = 0.f, cnt = 0.f;
#pragma unroll
        for (int k = 0; k < MAXT; k++)
            if (t0 + half + 2 * k < n_tiles) { ms += sv[k].x; cnt += 1.f; }
        const float mloc = cnt > 0.f ? ms / cnt : 0.f;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < MAXT; k++)
            if (t0 + half + 2 * k < n_tiles) { const float dd = sv[k].x - mloc; q += sv[k].y + w * dd * dd; }
        chan_combine(n, mean, m2, cnt * w, mloc, q);
    }
    // even tiles (lanes 0-15) and odd tiles (lanes 16-31) of the same row: lower lane is the left operand
    const float nb = __shfl_xor_sync(0xffffffffu, n, 16);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, 16);
    const float m2b = __shfl_xor_sync(0xffffffffu, m2, 16);
    if (half == 0) {
        chan_combine(n, mean, m2, nb, mb, m2b);
        row_stats[r] = make_float2(mean, 1.0f / sqrtf(m2 / (float)K + eps));
    }
}

// TMA = true: the weight ring is filled by the TMA unit instead of per-lane cp.async -- one elected lane per warp issues ONE
// cp.async.bulk.tensor.2d per iteration (box 16 k x 64 rows of W = 4 KB, no swizzle), which lands in the warp's ring stage
// in exactly the [n8 tile][lane][16 B] order the lanes read back (row r of the box = tile r / 8, g = r % 8; its 64 bytes =
// the four t pieces), completion on a per-(warp, stage) mbarrier.  No cross-warp synchronisation either way.
template <int PRO, int EPI, int MODE = 0, bool TMA = false>
__global__ void __launch_bounds__(GEMM_THREADS, 2) skinny_gemm_kernel(GemmArgs a, const __grid_constant__ CUtensorMap wmap) {
    extern __shared__ __align__(128) uint8_t gemm_smem[];
    __shared__ __align__(8) unsigned long long ring_bar[GEMM_WARPS * GEMM_STAGES];
    float *red = reinterpret_cast<float *>(gemm_smem);   // cross-warp reduction buffer: aliases the drained ring
    __shared__ float2 row_stats[GEMM_M];
    __shared__ int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int n0 = tile * GEMM_NT;
    // K is cut into 16-float chunks; chunk c of this split goes to warp c % 8, so the 8 warps of a CTA sweep each
    // weight row in 512-byte contiguous strides (any K that is a multiple of 16 works)
    // split s owns chunks [total * s / splits, total * (s + 1) / splits): equal parts when splits divides the chunk count
    // (every Taming shape), otherwise parts that differ by one chunk -- which lets pick_splits() fit ANY shape into one wave
    const int total_chunks = a.K / GEMM_KI;
    const int c0 = (int)((long long)total_chunks * split / a.splits);
    const int chunks = (int)((long long)total_chunks * (split + 1) / a.splits) - c0;
    const int kw0 = c0 * GEMM_KI + warp * GEMM_KI;
    const int iters = warp < chunks ? (chunks - warp + GEMM_WARPS - 1) / GEMM_WARPS : 0;
    constexpr int KSTEP = GEMM_WARPS * GEMM_KI;

    // issue the weight loads of the first GEMM_STAGES iterations before anything else: they do not depend on the
    // producer kernel.  One commit group per iteration (empty groups keep the count constant).
    const float *wbase = a.W + (size_t)(n0 + g) * a.K + kw0 + 4 * t;
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(gemm_smem) + (uint32_t)(warp * GEMM_STAGES * GEMM_STAGE_BYTES) + (uint32_t)(lane * 16);
    const uint32_t bar0 = tc05::smem_u32(ring_bar) + (uint32_t)(warp * GEMM_STAGES * 8);
    if (TMA) {
        if (lane == 0) {
            if (warp == 0) tc05::tma_prefetch_desc(&wmap);
#pragma unroll
            for (int s0 = 0; s0 < GEMM_STAGES; s0++) tc05::mbar_init(bar0 + 8u * s0, 1);
            tc05::mbar_fence_init();
        }
        __syncwarp();
    }
    auto issue = [&](int it) {
        if (TMA) {
            if (it < iters && lane == 0) {
                const uint32_t st = (uint32_t)(it % GEMM_STAGES);
                tc05::mbar_arrive_expect_tx(bar0 + 8u * st, GEMM_STAGE_BYTES);
                tc05::tma_load_2d(ring - (uint32_t)(lane * 16) + st * GEMM_STAGE_BYTES, &wmap, kw0 + it * KSTEP, n0, bar0 + 8u * st,
                                  tc05::L2_EVICT_FIRST);
            }
            return;
        }
        if (it < iters) {
            const uint32_t dst = ring + (uint32_t)((it % GEMM_STAGES) * GEMM_STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < GEMM_TILES; j++) gm_cp16(dst + j * 512, wbase + (size_t)(8 * j) * a.K + it * KSTEP);
        }
        gm_cp_commit();
    };
#pragma unroll
    for (int s0 = 0; s0 < GEMM_STAGES; s0++) issue(s0);
    // programmatic dependent launch: the next kernel of the step may start (and issue ITS first weight loads) while this
    // one runs; everything below reads what the previous kernel produced
    GM_TRACE(0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    GM_TRACE(1);
    // this launch's hand-off flag (the load is consumed only after the main loop)
    unsigned ll_flag = 0;
    if (a.ll_salt != 0) ll_flag = ((unsigned)(a.ll_epoch != nullptr ? (unsigned)__ldcg(a.ll_epoch) : a.ll_epoch_val) + 1u) * 1024u + a.ll_salt;

    float2 st_g = make_float2(0.f, 1.f), st_g8 = make_float2(0.f, 1.f);
    if (PRO != PRO_NONE) {
        combine_row_stats(a.stats_in, a.n_stat_tiles, a.K, a.eps, row_stats);
        __syncthreads();
        st_g = row_stats[g];
        st_g8 = row_stats[g + 8];
    }

    GM_TRACE(2);
    float acc[GEMM_TILES][4];
#pragma unroll
    for (int j = 0; j < GEMM_TILES; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[j][c] = 0.f;

    const float *x_g = a.X + (size_t)g * a.ldx + kw0 + 4 * t;
    const float *x_g8 = a.X + (size_t)(g + 8) * a.ldx + kw0 + 4 * t;

    for (int it = 0; it < iters; it++) {
        const int koff = it * KSTEP;
        float4 xa = *reinterpret_cast<const float4 *>(x_g + koff);
        float4 xb = *reinterpret_cast<const float4 *>(x_g8 + koff);
        float xs[2][4] = {{xa.x, xa.y, xa.z, xa.w}, {xb.x, xb.y, xb.z, xb.w}};
        if (PRO != PRO_NONE) {
            const int k = kw0 + 4 * t + koff;
            float gm[4] = {1.f, 1.f, 1.f, 1.f}, bt[4] = {0.f, 0.f, 0.f, 0.f};
            if (a.ln_g != nullptr) {
                float4 g4 = *reinterpret_cast<const float4 *>(a.ln_g + k);
                float4 b4 = *reinterpret_cast<const float4 *>(a.ln_b + k);
                gm[0] = g4.x; gm[1] = g4.y; gm[2] = g4.z; gm[3] = g4.w;
                bt[0] = b4.x; bt[1] = b4.y; bt[2] = b4.z; bt[3] = b4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
                xs[0][e] = (xs[0][e] - st_g.x) * st_g.y * gm[e] + bt[e];
                xs[1][e] = (xs[1][e] - st_g8.x) * st_g8.y * gm[e] + bt[e];
            }
            if (PRO == PRO_ADALN) {
                float4 sc0 = *reinterpret_cast<const float4 *>(a.mod_scale + (size_t)g * a.ld_mod + k);
                float4 sh0 = *reinterpret_cast<const float4 *>(a.mod_shift + (size_t)g * a.ld_mod + k);
                float4 sc1 = *reinterpret_cast<const float4 *>(a.mod_scale + (size_t)(g + 8) * a.ld_mod + k);
                float4 sh1 = *reinterpret_cast<const float4 *>(a.mod_shift + (size_t)(g + 8) * a.ld_mod + k);
                float s0[4] = {sc0.x, sc0.y, sc0.z, sc0.w}, h0[4] = {sh0.x, sh0.y, sh0.z, sh0.w};
                float s1[4] = {sc1.x, sc1.y, sc1.z, sc1.w}, h1[4] = {sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    xs[0][e] = xs[0][e] * (1.f + s0[e]) + h0[e];
                    xs[1][e] = xs[1][e] * (1.f + s1[e]) + h1[e];
                }
            }
        }
        uint32_t xh[2][4], xl[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int e = 0; e < 4; e++) split_tf32(xs[r][e], xh[r][e], xl[r][e]);
        // this iteration's weights have landed (at most STAGES-1 newer groups may still be in flight); pull them into
        // registers and hand the stage to iteration it + STAGES
        if (TMA) tc05::mbar_wait(bar0 + 8u * (uint32_t)(it % GEMM_STAGES), (uint32_t)(it / GEMM_STAGES) & 1u);
        else gm_cp_wait<GEMM_STAGES - 1>();
        float4 wcur[GEMM_TILES];
        {
            const uint8_t *src = gemm_smem + (warp * GEMM_STAGES + (it % GEMM_STAGES)) * GEMM_STAGE_BYTES + lane * 16;
#pragma unroll
            for (int j = 0; j < GEMM_TILES; j++) wcur[j] = *reinterpret_cast<const float4 *>(src + j * 512);
        }
        if (TMA) __syncwarp();   // every lane has read the stage before the TMA unit overwrites it
        issue(it + GEMM_STAGES);
        if (it < 4) GM_TRACE(3 + it);
#pragma unroll
        for (int j = 0; j < GEMM_TILES; j++) {
            const float wv[4] = {wcur[j].x, wcur[j].y, wcur[j].z, wcur[j].w};
            // weights: hi = trunc_tf32(w) (one LOP; the tensor core ignores the low 13 bits anyway), lo = w - hi exactly
            // (fed as raw fp32 bits, truncated to TF32 by the MMA: residual <= 2^-21 |w|)
            uint32_t wh[4], wl[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                wh[e] = __float_as_uint(wv[e]) & 0xffffe000u;
                wl[e] = __float_as_uint(wv[e] - __uint_as_float(wh[e]));
            }
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int e = 2 * half;
                // k-slot t <-> element e, k-slot t+4 <-> element e+1 (any bijection works: the sum over k is unordered)
                if (MODE == 0) {
                    mma_tf32(acc[j], xl[0][e], xl[1][e], xl[0][e + 1], xl[1][e + 1], wh[e], wh[e + 1]);
                    mma_tf32(acc[j], xh[0][e], xh[1][e], xh[0][e + 1], xh[1][e + 1], wl[e], wl[e + 1]);
                }
                if (MODE != 2) mma_tf32(acc[j], xh[0][e], xh[1][e], xh[0][e + 1], xh[1][e + 1], wh[e], wh[e + 1]);
                else acc[j][e] += __uint_as_float(wh[e]) + __uint_as_float(wh[e + 1]);
            }
        }
    }
    GM_TRACE(7);
    if (!TMA) gm_cp_wait<0>();
    __syncthreads();   // every warp is done with its ring: the reduction buffer below aliases it

    // ---- cross-warp reduction (fixed order) ----
    float *myred = red + warp * GEMM_M * GEMM_RED_LD;
#pragma unroll
    for (int j = 0; j < GEMM_TILES; j++) {
        *reinterpret_cast<float2 *>(myred + g * GEMM_RED_LD + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2 *>(myred + (g + 8) * GEMM_RED_LD + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    const int m = tid >> 4, nn = (tid & 15) * 4;  // each thread owns 4 consecutive columns of one row
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < GEMM_WARPS; w++) {
        float4 p = *reinterpret_cast<const float4 *>(red + (w * GEMM_M + m) * GEMM_RED_LD + nn);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }

    // ---- split-K, flag-carrying hand-off: the CTA of the last split sums the partials in split order ----
    if (a.splits > 1 && a.ll_salt != 0) {
        unsigned long long *wst = reinterpret_cast<unsigned long long *>(a.ws) + ((size_t)tile * a.splits) * (GEMM_M * GEMM_NT) + m * GEMM_NT + nn;
        if (split != a.splits - 1) {
            unsigned long long *dst = wst + (size_t)split * (GEMM_M * GEMM_NT);
            ll_store2(dst, ll_pack(v.x, ll_flag), ll_pack(v.y, ll_flag));
            ll_store2(dst + 2, ll_pack(v.z, ll_flag), ll_pack(v.w, ll_flag));
            GM_TRACE(8);
            return;
        }
        GM_TRACE(9);
        const float4 own = v;
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int others = a.splits - 1;
        for (int s0 = 0; s0 < others; s0 += 4) {       // four partials in flight
            unsigned long long q[4][4];
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (s0 + k < others) {
                        const unsigned long long *src = wst + (size_t)(s0 + k) * (GEMM_M * GEMM_NT);
                        ll_load2(src, q[k][0], q[k][1]);
                        ll_load2(src + 2, q[k][2], q[k][3]);
                    }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (s0 + k < others) {
#pragma unroll
                        for (int c = 0; c < 4; c++) ok = ok && ((unsigned)(q[k][c] >> 32) == ll_flag);
                    }
            } while (!ok);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (s0 + k < others) {
                    v.x += __uint_as_float((unsigned)q[k][0]); v.y += __uint_as_float((unsigned)q[k][1]);
                    v.z += __uint_as_float((unsigned)q[k][2]); v.w += __uint_as_float((unsigned)q[k][3]);
                }
        }
        v.x += own.x; v.y += own.y; v.z += own.z; v.w += own.w;
    }
    // ---- split-K, counter hand-off: last CTA of the tile to arrive reduces the partials in split order ----
    if (a.splits > 1 && a.ll_salt == 0) {
        float *wst = a.ws + ((size_t)tile * a.splits) * (GEMM_M * GEMM_NT);
        __stcg(reinterpret_cast<float4 *>(wst + (size_t)split * (GEMM_M * GEMM_NT) + m * GEMM_NT + nn), v);
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            unsigned old = atomicAdd(&a.counters[tile], 1u);
            s_is_last = (old == (unsigned)(a.splits - 1));
            if (s_is_last) a.counters[tile] = 0u;
        }
        __syncthreads();
        GM_TRACE(8 + (s_is_last ? 1 : 0));
        if (!s_is_last) return;
        __threadfence();
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < a.splits; s0 += 4) {     // four partials in flight, summed in split order
            float4 p[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
                p[k] = s0 + k < a.splits ? __ldcg(reinterpret_cast<const float4 *>(wst + (size_t)(s0 + k) * (GEMM_M * GEMM_NT) + m * GEMM_NT + nn))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (s0 + k < a.splits) { v.x += p[k].x; v.y += p[k].y; v.z += p[k].z; v.w += p[k].w; }
        }
    }

    GM_TRACE(10);
    // ---- epilogue ----
    const int n = n0 + nn;
    if (a.bias != nullptr) {
        float4 b4 = *reinterpret_cast<const float4 *>(a.bias + n);
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    }
    if (EPI == EPI_GELU) {
        v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
    }
    if (EPI == EPI_GATE_RESID) {
        float4 g4 = *reinterpret_cast<const float4 *>(a.gate + (size_t)m * a.ld_gate + n);
        v.x *= g4.x; v.y *= g4.y; v.z *= g4.z; v.w *= g4.w;
    }
    if (EPI == EPI_RESID || EPI == EPI_GATE_RESID) {
        float4 r4 = *reinterpret_cast<const float4 *>(a.resid + (size_t)m * a.ld_resid + n);
        v.x = r4.x + v.x; v.y = r4.y + v.y; v.z = r4.z + v.z; v.w = r4.w + v.w;
    }
    *reinterpret_cast<float4 *>(a.Y + (size_t)m * a.ldy + n) = v;
    if (a.stats_out != nullptr) {
        // (mean, M2) of this row's 64 columns: 16 consecutive lanes hold 4 values each
        float s = v.x + v.y + v.z + v.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / GEMM_NT);
        float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if ((tid & 15) == 0) a.stats_out[tile * GEMM_M + m] = make_float2(mean, q);
    }
    GM_TRACE(11);
}

// host-side launcher (graph-capturable)
int launch_skinny_gemm(int pro, int epi, const GemmArgs &a, cudaStream_t stream);
// number of K splits that fills the machine for an [N][K] weight (mma.sync kernel)
int pick_splits(int N, int K, int n_sms);
// split-K workspace (floats) that covers whichever kernel launch_skinny_gemm picks for this shape
size_t gemm_ws_floats(int N, int K, int splits_v0, int n_sms);

}  // namespace wmar
