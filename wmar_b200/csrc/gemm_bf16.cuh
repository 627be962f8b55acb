// Skinny GEMM with bf16 weights for the Chameleon / Anole-7B decode step:
//     Y[16][N] = bf16( pro(X)[16][K] . W[N][K]^T )  (+ residual),  W bf16, X / Y fp32 buffers holding bf16 VALUES.
//
// Replaces the bias-free nn.Linear layers of deps/chameleon/inference/transformer.py (wqkv :57-62,111; wo :63-68,158;
// w13 / w2 :184-193,217-219; output :282-286,314).  The reference model is bf16 end to end: every Linear rounds its
// output to bf16 and every activation it reads is a bf16 tensor.  The engine keeps its activations in fp32 buffers
// (one code path with the fp32 models) but rounds at exactly those points, so the products are bf16 x bf16 with fp32
// accumulation like the reference's cuBLAS calls.
//   * the 16 rows (f | i | u guided row groups of <= 5 images) are the M of mma.m16n8k16; W rows are the MMA's n, and
//     every lane fetches exactly the 16-byte pieces that are ITS B fragments: 8 consecutive k of one weight row = the
//     B fragments of two k16 MMAs (k permuted identically in the A fragments),
//   * a PERSISTENT grid (two CTAs per SM) walks the (tile, split) work items: the 7B shapes give 256..2048 items, so a
//     one-item-per-CTA launch would run 1.2-3.5 waves with a ragged tail; the weights travel through the per-lane
//     cp.async ring of gemm.cuh (three iterations of every warp in flight) and the first three iterations of a CTA's
//     NEXT item are issued before the split-K tail / epilogue of the current one, so the HBM stream does not drain
//     inside a kernel,
//   * split-K with the flag-carrying hand-off of gemm.cuh ({value, flag} words, summed in split order by the CTA that
//     owns the last split's item; the counter-based last-arriver reduction remains for ll_salt = 0, e.g. the unit test entry),
//   * prologues: RMSNorm (xformers RMSNorm, transformer.py:238-239,278) from the producer's (mean, M2) partials;
//     SwiGLU  silu(x1) * x3  over the two halves of the w13 output (transformer.py:217-218),
//   * epilogues: round to bf16; residual add (round, add, round) + LayerNorm-style (mean, M2) partials for the next RMS;
//     SwiGLU: with the rows of w13 interleaved at pack time (tile T = x1 rows 32T..32T+31 then the matching x3 rows) the
//     epilogue forms h = bf16(bf16(silu(x1)) * x3) once per element and writes [16][F] -- the w2 GEMM then reads half
//     the activation bytes and evaluates no exp (as a w2 PROLOGUE the same work was redone by each of its 64 tiles).
#pragma once
#include <cuda_bf16.h>

#include "gemm.cuh"

namespace wmar {

constexpr int BG_KI = 32;   // k per warp iteration (one LDG.128 of bf16 per lane per n8 tile)

enum Bf16Prologue { BPRO_NONE = 0, BPRO_RMS = 1, BPRO_SWIGLU = 2 };
enum Bf16Epilogue { BEPI_STORE = 0, BEPI_RESID = 1, BEPI_STORE_F32 = 2, BEPI_SWIGLU = 3 };

struct Bf16GemmArgs {
    const float *X; int ldx;
    const __nv_bfloat16 *W;
    float *Y; int ldy;
    int N, K, splits;
    const float *rms_w;           // [K] RMSNorm weight (BPRO_RMS)
    const float2 *stats_in;       // [n_stat_tiles][16] (mean, M2) of 64-wide column tiles of X
    int n_stat_tiles;
    float eps;
    int swiglu_off;               // BPRO_SWIGLU: x3 lives at X[.][k + swiglu_off]
    const float *resid; int ld_resid;
    float2 *stats_out;            // [N/64][16] or null
    float *ws;
    unsigned *counters;
    // flag-carrying split-K hand-off, see GemmArgs in gemm.cuh (ws then holds 8-byte {value, flag} words).  In this
    // persistent kernel the reducer is the CTA that owns the item of the LAST split; it only ever waits for items with
    // a lower index, which are owned by other resident CTAs or lie behind it, so the wait cannot deadlock.
    const int *ll_epoch; unsigned ll_salt;
};

__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&p);
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// per-row rsqrt(mean(x^2) + eps) from (mean, M2) partials: mean(x^2) = M2 / K + mean^2
__device__ __forceinline__ void combine_row_rms(const float2 *__restrict__ stats_in, int n_tiles, int K, float eps, float *row_rs) {
    const int tid = threadIdx.x;
    const int r = tid >> 4, sub = tid & 15;
    float n = 0.f, mean = 0.f, m2 = 0.f;
    const float w = (float)(K / n_tiles);
    for (int tl = sub; tl < n_tiles; tl += 16) {
        float2 s = __ldcg(stats_in + tl * 16 + r);   // not hoistable above griddepcontrol.wait
        chan_combine(n, mean, m2, w, s.x, s.y);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        float nb = __shfl_xor_sync(0xffffffffu, n, o);
        float mb = __shfl_xor_sync(0xffffffffu, mean, o);
        float m2b = __shfl_xor_sync(0xffffffffu, m2, o);
        if ((sub & o) == 0) chan_combine(n, mean, m2, nb, mb, m2b);
        else { float tn = nb, tm = mb, t2 = m2b; chan_combine(tn, tm, t2, n, mean, m2); n = tn; mean = tm; m2 = t2; }
    }
    if (sub == 0) row_rs[r] = 1.0f / sqrtf(m2 / (float)K + mean * mean + eps);
}

template <int PRO, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) skinny_gemm_bf16_kernel(Bf16GemmArgs a) {
    extern __shared__ __align__(16) uint8_t gemm_smem[];
    float *red = reinterpret_cast<float *>(gemm_smem);   // cross-warp reduction buffer: aliases the drained ring
    __shared__ float row_rs[GEMM_M];
    __shared__ int s_is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int tiles = a.N / GEMM_NT, n_items = tiles * a.splits;
    const int KS = a.K / a.splits;
    const int chunks = KS / BG_KI;
    const int iters = warp < chunks ? (chunks - warp + GEMM_WARPS - 1) / GEMM_WARPS : 0;
    constexpr int KSTEP = GEMM_WARPS * BG_KI;

    // weight loads of iteration `it` of `item` into ring stage it % STAGES: one commit group per call (empty groups keep
    // the group count constant).  The first STAGES iterations of the first item do not depend on the producer kernel.
    const uint32_t ring = (uint32_t)__cvta_generic_to_shared(gemm_smem) + (uint32_t)(warp * GEMM_STAGES * GEMM_STAGE_BYTES) + (uint32_t)(lane * 16);
    auto issue = [&](int item, int it) {
        if (item < n_items && it < iters) {
            const int tl = item % tiles, sp = item / tiles;
            const __nv_bfloat16 *wb = a.W + (size_t)(tl * GEMM_NT + g) * a.K + sp * KS + warp * BG_KI + 8 * t + it * KSTEP;
            const uint32_t dst = ring + (uint32_t)((it % GEMM_STAGES) * GEMM_STAGE_BYTES);
#pragma unroll
            for (int j = 0; j < GEMM_TILES; j++) gm_cp16(dst + j * 512, wb + (size_t)(8 * j) * a.K);
        }
        gm_cp_commit();
    };
#pragma unroll
    for (int s0 = 0; s0 < GEMM_STAGES; s0++) issue(blockIdx.x, s0);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    unsigned ll_flag = 0;
    if (a.ll_salt != 0) ll_flag = ((unsigned)__ldcg(a.ll_epoch) + 1u) * 1024u + a.ll_salt;

    float rs_g = 1.f, rs_g8 = 1.f;
    if (PRO == BPRO_RMS) {
        combine_row_rms(a.stats_in, a.n_stat_tiles, a.K, a.eps, row_rs);
        __syncthreads();
        rs_g = row_rs[g];
        rs_g8 = row_rs[g + 8];
    }

    float4 nxa0, nxa1, nxb0, nxb1, nw0, nw1;
    nxa0 = nxa1 = nxb0 = nxb1 = make_float4(0.f, 0.f, 0.f, 0.f);
    nw0 = nw1 = make_float4(1.f, 1.f, 1.f, 1.f);
    auto load_x = [&](int item, int it) {
        if (item < n_items && it < iters) {
            const int k = (item / tiles) * KS + warp * BG_KI + 8 * t + it * KSTEP;
            const float *p0 = a.X + (size_t)g * a.ldx + k, *p1 = a.X + (size_t)(g + 8) * a.ldx + k;
            nxa0 = __ldcg(reinterpret_cast<const float4 *>(p0)); nxa1 = __ldcg(reinterpret_cast<const float4 *>(p0 + 4));
            nxb0 = __ldcg(reinterpret_cast<const float4 *>(p1)); nxb1 = __ldcg(reinterpret_cast<const float4 *>(p1 + 4));
            if (PRO == BPRO_RMS) {
                nw0 = __ldg(reinterpret_cast<const float4 *>(a.rms_w + k)); nw1 = __ldg(reinterpret_cast<const float4 *>(a.rms_w + k + 4));
            }
        }
    };
    load_x(blockIdx.x, 0);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int tile = item % tiles, split = item / tiles;
    const int n0 = tile * GEMM_NT;
    const int kw0 = split * KS + warp * BG_KI;
    float acc[GEMM_TILES][4];
#pragma unroll
    for (int j = 0; j < GEMM_TILES; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[j][c] = 0.f;

    for (int it = 0; it < iters; it++) {
        const int koff = it * KSTEP;
        // this iteration's activations (and RMSNorm weights) were requested one iteration ago; request the next ones --
        // the next iteration's, or the first of this CTA's next item -- before touching them: an L2 round trip under
        // the weight stream costs more than the iteration's math (knock-out: 19 % of the Anole-7B pass)
        float xs[2][8];
        xs[0][0] = nxa0.x; xs[0][1] = nxa0.y; xs[0][2] = nxa0.z; xs[0][3] = nxa0.w; xs[0][4] = nxa1.x; xs[0][5] = nxa1.y; xs[0][6] = nxa1.z; xs[0][7] = nxa1.w;
        xs[1][0] = nxb0.x; xs[1][1] = nxb0.y; xs[1][2] = nxb0.z; xs[1][3] = nxb0.w; xs[1][4] = nxb1.x; xs[1][5] = nxb1.y; xs[1][6] = nxb1.z; xs[1][7] = nxb1.w;
        const float wv[8] = {nw0.x, nw0.y, nw0.z, nw0.w, nw1.x, nw1.y, nw1.z, nw1.w};
        if (it + 1 < iters) load_x(item, it + 1);
        else load_x(item + (int)gridDim.x, 0);
        if (PRO == BPRO_RMS) {
#pragma unroll
            for (int e = 0; e < 8; e++) {
                // xformers rms_norm: (x * rsqrt(mean(x^2) + eps)) * weight in fp32, stored as bf16 (the pack below rounds)
                xs[0][e] = xs[0][e] * rs_g * wv[e];
                xs[1][e] = xs[1][e] * rs_g8 * wv[e];
            }
        }
        if (PRO == BPRO_SWIGLU) {
            const float *x_g = a.X + (size_t)g * a.ldx + kw0 + 8 * t;
            const float *x_g8 = a.X + (size_t)(g + 8) * a.ldx + kw0 + 8 * t;
            const float4 c0 = *reinterpret_cast<const float4 *>(x_g + koff + a.swiglu_off), c1 = *reinterpret_cast<const float4 *>(x_g + koff + a.swiglu_off + 4);
            const float4 d0 = *reinterpret_cast<const float4 *>(x_g8 + koff + a.swiglu_off), d1 = *reinterpret_cast<const float4 *>(x_g8 + koff + a.swiglu_off + 4);
            const float x3[2][8] = {{c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w}, {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w}};
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    // F.silu on a bf16 tensor rounds its result to bf16, the product rounds again (the pack below)
                    const float s = bf16r(xs[r][e] / (1.0f + expf(-xs[r][e])));
                    xs[r][e] = s * x3[r][e];
                }
        }
        // A fragments of the two k16 MMAs of this iteration: MMA h uses this lane's elements 4h..4h+3 as its k slots
        // (2t, 2t+1, 2t+8, 2t+9); the weights below use the same bijection
        uint32_t af[2][4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            af[h][0] = pack_bf16(xs[0][4 * h], xs[0][4 * h + 1]);
            af[h][1] = pack_bf16(xs[1][4 * h], xs[1][4 * h + 1]);
            af[h][2] = pack_bf16(xs[0][4 * h + 2], xs[0][4 * h + 3]);
            af[h][3] = pack_bf16(xs[1][4 * h + 2], xs[1][4 * h + 3]);
        }
        // this iteration's weights have landed; pull them into registers and hand the stage to iteration it + STAGES
        gm_cp_wait<GEMM_STAGES - 1>();
        {
            const uint8_t *src = gemm_smem + (warp * GEMM_STAGES + (it % GEMM_STAGES)) * GEMM_STAGE_BYTES + lane * 16;
#pragma unroll
            for (int j = 0; j < GEMM_TILES; j++) {
                const uint4 w = *reinterpret_cast<const uint4 *>(src + j * 512);
                mma_bf16(acc[j], af[0][0], af[0][1], af[0][2], af[0][3], w.x, w.y);
                mma_bf16(acc[j], af[1][0], af[1][1], af[1][2], af[1][3], w.z, w.w);
            }
        }
        issue(item, it + GEMM_STAGES);   // after the reads above: the stage is this lane's own
    }
    gm_cp_wait<0>();
    __syncthreads();   // every warp is done with its ring: the reduction buffer below aliases it

    // ---- cross-warp reduction (fixed order) ----
    float *myred = red + warp * GEMM_M * GEMM_RED_LD;
#pragma unroll
    for (int j = 0; j < GEMM_TILES; j++) {
        *reinterpret_cast<float2 *>(myred + g * GEMM_RED_LD + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
        *reinterpret_cast<float2 *>(myred + (g + 8) * GEMM_RED_LD + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncthreads();
    const int m = tid >> 4, nn = (tid & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < GEMM_WARPS; w++) {
        float4 p = *reinterpret_cast<const float4 *>(red + (w * GEMM_M + m) * GEMM_RED_LD + nn);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    __syncthreads();   // the reduction buffer is consumed: the ring is free again
    // the first iterations of my next item fly during the split-K tail and the epilogue of this one
#pragma unroll
    for (int s0 = 0; s0 < GEMM_STAGES; s0++) issue(item + (int)gridDim.x, s0);

    if (a.splits > 1 && a.ll_salt != 0) {
        unsigned long long *wst = reinterpret_cast<unsigned long long *>(a.ws) + ((size_t)tile * a.splits) * (GEMM_M * GEMM_NT) + m * GEMM_NT + nn;
        if (split != a.splits - 1) {
            unsigned long long *dst = wst + (size_t)split * (GEMM_M * GEMM_NT);
            ll_store2(dst, ll_pack(v.x, ll_flag), ll_pack(v.y, ll_flag));
            ll_store2(dst + 2, ll_pack(v.z, ll_flag), ll_pack(v.w, ll_flag));
            continue;                   // CTA-uniform; `red` is protected by the barrier after the cross-warp reduction
        }
        const float4 own = v;
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int others = a.splits - 1;
        for (int s0 = 0; s0 < others; s0 += 4) {
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
        if (!s_is_last) continue;       // CTA-uniform; the barrier above also protects `red` for the next item
        __threadfence();
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < a.splits; s++) {
            float4 p = __ldcg(reinterpret_cast<const float4 *>(wst + (size_t)s * (GEMM_M * GEMM_NT) + m * GEMM_NT + nn));
            v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
    }

    // ---- epilogue ----
    const int n = n0 + nn;
    if (EPI != BEPI_STORE_F32) { v.x = bf16r(v.x); v.y = bf16r(v.y); v.z = bf16r(v.z); v.w = bf16r(v.w); }
    if (EPI == BEPI_RESID) {
        float4 r4 = *reinterpret_cast<const float4 *>(a.resid + (size_t)m * a.ld_resid + n);
        v.x = bf16r(r4.x + v.x); v.y = bf16r(r4.y + v.y); v.z = bf16r(r4.z + v.z); v.w = bf16r(r4.w + v.w);
    }
    if (EPI == BEPI_SWIGLU) {
        // lanes c (x1, columns 4c..4c+3 of the tile's first half) and c + 8 (the matching x3 columns) of the same row
        float4 o;
        o.x = __shfl_xor_sync(0xffffffffu, v.x, 8); o.y = __shfl_xor_sync(0xffffffffu, v.y, 8);
        o.z = __shfl_xor_sync(0xffffffffu, v.z, 8); o.w = __shfl_xor_sync(0xffffffffu, v.w, 8);
        if (nn < 32) {
            float4 hq;
            hq.x = bf16r(bf16r(v.x / (1.0f + expf(-v.x))) * o.x); hq.y = bf16r(bf16r(v.y / (1.0f + expf(-v.y))) * o.y);
            hq.z = bf16r(bf16r(v.z / (1.0f + expf(-v.z))) * o.z); hq.w = bf16r(bf16r(v.w / (1.0f + expf(-v.w))) * o.w);
            *reinterpret_cast<float4 *>(a.Y + (size_t)m * a.ldy + tile * 32 + nn) = hq;
        }
    } else
    *reinterpret_cast<float4 *>(a.Y + (size_t)m * a.ldy + n) = v;
    if (a.stats_out != nullptr) {
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
    __syncthreads();   // `red` / `s_is_last` are reused by the next item
  }
}

int launch_skinny_gemm_bf16(int pro, int epi, const Bf16GemmArgs &a, cudaStream_t stream);
int pick_splits_bf16(int N, int K, int n_sms);

}  // namespace wmar
