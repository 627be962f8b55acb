// tcgen05 implicit-GEMM 3x3 convolution (stride 1, pad 1) for the VQGAN conv stacks, NHWC fp32 in / out, 3xTF32.
//
// Replaces, for the layers that hold most of the FLOPs (3x3, Cin % 32 == 0, Cout % 128 == 0: 75 % of the Taming decoder,
// model.py:79-138,437-538), the mma.sync kernel of vqgan_kernels.cuh.  GEMM view: M = 128 output pixels (bh rows x bw
// columns of one image), N = 128 output channels, K = 9 taps x Cin.
//   * A (activations): per (tap, 32-channel chunk) ONE 4-D TMA box [1][bh][bw][32 ch] of the NHWC tensor at the tap's
//     spatial offset -- out-of-bounds rows / columns are zero-filled by the TMA unit, which IS the conv padding; the box
//     lands as 128 pixel rows x 128 B (SWIZZLE_128B), exactly the tile the decode kernels stream for weights,
//   * converter warps (thread <-> pixel <-> TMEM lane) split it into hi = rna_tf32(a), lo = a - hi in TENSOR MEMORY,
//   * B (weights [Cout][ky][kx][Cin], K-major): two TMA boxes of 128 rows x 32 k per chunk -- the raw fp32 weights
//     (the tensor core reads them as TF32 by truncation = w_hi) and w_lo = w - trunc_tf32(w), precomputed once --
//     stacked in shared memory as ONE 256-row SWIZZLE_128B operand [w_hi ; w_lo],
//   * per k8 step two tcgen05.mma.kind::tf32:  D[128 x 256] += a_hi . [w_hi ; w_lo]^T  and  D[:, 128:256] += a_lo . w_hi^T;
//     the epilogue adds the two accumulator halves (hi.hi) + (hi.lo + lo.hi), bias and the residual.
#pragma once
#include "gemm.cuh"
#include "tc05.cuh"

namespace wmar {

constexpr int CT_NS = 4;                      // ring stages: A raw 16 KB + w_hi 16 KB + w_lo 16 KB each
constexpr int CT_NTA = 4;                     // TMEM stages of split A
constexpr int CT_THREADS = 448;               // warps: 0 TMA, 1 MMA, 2-9 converters (2 groups), 10-13 epilogue
constexpr int CT_W_CONV = 2, CT_W_EPI = 10;
constexpr int CT_TILE_BYTES = 16384;
constexpr int CT_STAGE_BYTES = 3 * CT_TILE_BYTES;
constexpr int CT_TMEM_A0 = 256;               // cols [0,128) hi.hi, [128,256) cross terms, [256,512) 4 A stages x 64
constexpr int CT_SM_BAR = CT_NS * CT_STAGE_BYTES;
constexpr int CT_N_BARS = 2 * CT_NS + 2 * CT_NTA + 1;
constexpr int CT_SM_MISC = CT_SM_BAR + 8 * CT_N_BARS;
constexpr int CT_SM_BYTES = CT_SM_MISC + 16;
constexpr int CT_SM_ALLOC = CT_SM_BYTES + 1024;
static_assert(CT_SM_ALLOC <= 232448, "conv_tc kernel exceeds 227 KB of shared memory");

struct ConvTcArgs {
    const float *bias, *resid;
    float *out;
    int H, W, Cin, Cout;     // Cin = padded input channels (multiple of 32), Cout = real output channels (multiple of 128)
    int bw, bh;              // pixel tile: bw x bh = 128
    int tiles_x, tiles_y;    // W / bw, H / bh
    int taps;                // 9 (3x3) or 1 (1x1 conv / plain GEMM over "pixels"); 0 is read as 9
    int stride, pad;         // bf16 kernel only: 1 / 1 (same conv) or 2 / 0 (downsample conv: pad (0,1,0,1) = TMA zero fill
                             // past the right / bottom edge, model.py:69-72); H, W are the OUTPUT dims
    // bf16 kernel only: GroupNorm(32) statistics of the OUTPUT tensor from the epilogue (the consumer is a GroupNorm):
    // gn_out[(b * tiles_per_image + tile) * 32 + group] = (sum, sum of squares) over the tile's 128 pixels x the group's
    // channels; a (tile, 128-channel block) writes exactly the groups its block covers -> no atomics, fixed summation order
    double2 *gn_out;         // null: off
    int gn_cg;               // channels per group = Cout / 32 (4, 8 or 16)
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap *m, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__global__ void __launch_bounds__(CT_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapWh,
                  const __grid_constant__ CUtensorMap mapWl, const ConvTcArgs a) {
    using namespace tc05;
    extern __shared__ uint8_t ct_smem_raw[];
    const uint32_t smem_base = (smem_u32(ct_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = ct_smem_raw + (smem_base - smem_u32(ct_smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile of this CTA
    const int tpi = a.tiles_x * a.tiles_y;
    const int b = blockIdx.x / tpi, tr = blockIdx.x - b * tpi;
    const int ty = tr / a.tiles_x, tx = tr - ty * a.tiles_x;
    const int x0 = tx * a.bw, y0 = ty * a.bh, n0 = blockIdx.y * 128;
    const int taps = a.taps == 1 ? 1 : 9, off = a.taps == 1 ? 0 : 1;
    const int cchunks = a.Cin / 32, nchunks = taps * cchunks;

    const uint32_t bar0 = smem_base + CT_SM_BAR;
    auto s_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto s_empty = [&](int s) { return bar0 + 8u * (uint32_t)(CT_NS + s); };
    auto ta_full = [&](int u) { return bar0 + 8u * (uint32_t)(2 * CT_NS + u); };
    auto ta_empty = [&](int u) { return bar0 + 8u * (uint32_t)(2 * CT_NS + CT_NTA + u); };
    const uint32_t acc_full = bar0 + 8u * (uint32_t)(2 * CT_NS + 2 * CT_NTA);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + CT_SM_MISC);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapWh);
        tma_prefetch_desc(&mapWl);
        // a stage is released by the 4 converter warps that read the activation tile AND by the MMAs that read the weights
        for (int s = 0; s < CT_NS; s++) { mbar_init(s_full(s), 1); mbar_init(s_empty(s), 5); }
        for (int u = 0; u < CT_NTA; u++) { mbar_init(ta_full(u), 4); mbar_init(ta_empty(u), 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(s_tmem));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int c = 0; c < nchunks; c++) {
                const uint32_t s = c % CT_NS;
                const int tap = c / cchunks, ci0 = (c - tap * cchunks) * 32;
                const int ky = tap / 3, kx = tap - ky * 3;
                mbar_wait(s_empty(s), ((c / CT_NS) & 1) ^ 1);
                mbar_arrive_expect_tx(s_full(s), CT_STAGE_BYTES);
                const uint32_t dst = smem_base + s * CT_STAGE_BYTES;
                tma_load_4d(dst, &mapA, ci0, x0 + kx - off, y0 + ky - off, b, s_full(s));       // zero fill = padding
                tma_load_2d(dst + CT_TILE_BYTES, &mapWh, tap * a.Cin + ci0, n0, s_full(s), L2_EVICT_LAST);
                tma_load_2d(dst + 2 * CT_TILE_BYTES, &mapWl, tap * a.Cin + ci0, n0, s_full(s), L2_EVICT_LAST);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t ID256 = idesc_tf32_m128(256), ID128 = idesc_tf32_m128(128);
            for (int c = 0; c < nchunks; c++) {
                const uint32_t s = c % CT_NS, u = c % CT_NTA;
                mbar_wait(s_full(s), (c / CT_NS) & 1);          // the weight tiles of this stage have landed
                mbar_wait(ta_full(u), (c / CT_NTA) & 1);        // the split activations are in tensor memory
                fence_after_sync();
                const uint32_t a_hi = tmem + CT_TMEM_A0 + u * 64, a_lo = a_hi + 32;
                const uint32_t wb = smem_base + s * CT_STAGE_BYTES + CT_TILE_BYTES;   // [w_hi (128 rows) ; w_lo (128 rows)]
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint64_t bd = tc05::smem_desc_kmajor_sw128(wb + ks * 32);
                    mma_tf32_ts(tmem, a_hi + ks * 8, bd, ID256, (c | ks) != 0 ? 1u : 0u);   // hi.hi | hi.lo
                    mma_tf32_ts(tmem + 128, a_lo + ks * 8, bd, ID128, 1u);                  // + lo.hi
                }
                mma_commit(ta_empty(u));
                mma_commit(s_empty(s));
            }
            mma_commit(acc_full);
        }
        __syncwarp();
    } else if (warp < CT_W_EPI) {
        // ===================== activation converters (thread <-> pixel <-> TMEM lane) =====================
        const int cg = (warp - CT_W_CONV) >> 2;
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        for (int c = cg; c < nchunks; c += 2) {
            const uint32_t s = c % CT_NS, u = c % CT_NTA;
            mbar_wait_warp(s_full(s), (c / CT_NS) & 1, lane);
            const uint8_t *arow = smem + s * CT_STAGE_BYTES + r * 128;
            float4 w4[8];
#pragma unroll
            for (int k = 0; k < 8; k++) w4[k] = *reinterpret_cast<const float4 *>(arow + ((k ^ (r & 7)) << 4));
            mbar_wait_warp(ta_empty(u), ((c / CT_NTA) & 1) ^ 1, lane);
            fence_after_sync();
            const uint32_t a_hi = t_lane + CT_TMEM_A0 + u * 64, a_lo = a_hi + 32;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float wv[4] = {w4[half * 4 + k].x, w4[half * 4 + k].y, w4[half * 4 + k].z, w4[half * 4 + k].w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const uint32_t h = (__float_as_uint(wv[e]) + 0x1000u) & 0xffffe000u;
                        hi[4 * k + e] = h;
                        lo[4 * k + e] = __float_as_uint(wv[e] - __uint_as_float(h));
                    }
                }
                tmem_st16(a_hi + half * 16, hi);
                tmem_st16(a_lo + half * 16, lo);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(s));
            wait_st();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(ta_full(u));
        }
    } else {
        // ===================== epilogue (thread <-> pixel) =====================
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        const int py = y0 + r / a.bw, px = x0 + r % a.bw;
        const size_t pix = ((size_t)b * a.H + py) * a.W + px;
        float *dst = a.out + pix * a.Cout + n0;
        const float *res = a.resid != nullptr ? a.resid + pix * a.Cout + n0 : nullptr;
        mbar_wait_warp(acc_full, 0, lane);
        fence_after_sync();
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t d0[32], d1[32];
            tmem_ld32(t_lane + c0, d0);
            tmem_ld32(t_lane + 128 + c0, d1);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + c0 + k));
                float4 v;
                v.x = (__uint_as_float(d0[k]) + __uint_as_float(d1[k])) + bb.x;
                v.y = (__uint_as_float(d0[k + 1]) + __uint_as_float(d1[k + 1])) + bb.y;
                v.z = (__uint_as_float(d0[k + 2]) + __uint_as_float(d1[k + 2])) + bb.z;
                v.w = (__uint_as_float(d0[k + 3]) + __uint_as_float(d1[k + 3])) + bb.w;
                if (res != nullptr) {
                    const float4 r4 = *reinterpret_cast<const float4 *>(res + c0 + k);
                    v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
                }
                *reinterpret_cast<float4 *>(dst + c0 + k) = v;
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------------
// bf16x3 variant (wmar_vqgan_config.precision >= 2), persistent, epilogue overlapped with the next tile's MMAs.
//
// Same implicit GEMM, but every fp32 value is split into TWO bf16 terms (x = x1 + x2 + O(2^-18 |x|), both rounded to
// nearest) and the product is x1.w1 + x1.w2 + x2.w1 on tcgen05.mma.kind::f16 -- K = 16 per instruction at the K = 8 cost of
// kind::tf32, i.e. half the tensor time of the 3xTF32 kernel at ~2^-17 relative error per product (3xTF32: ~2^-21,
// one TF32 product as cuDNN's default: 2^-11).  Differences to the kernel above:
//   * a chunk is 64 channels of one tap: two 4-D TMA boxes of 32 fp32 channels (A raw, 2 x 16 KB) + the bf16 weight tiles
//     w1 and w2 (128 rows x 64 k x 2 B = 16 KB each, SWIZZLE_128B like before); 3 ring stages of 64 KB,
//   * converter warps pack (k, k+1) pairs into one 32-bit TMEM column: 32 columns x1 + 32 columns x2 per chunk,
//   * all three products accumulate into ONE 128-column fp32 accumulator (three N = 128 MMAs per k16 step: the same
//     tensor time as N = 256 + N = 128), which leaves room for TWO accumulators in tensor memory,
//   * the CTA is persistent: tiles blockIdx.x, + gridDim.x, ...; the epilogue of tile i (TMEM -> registers -> bias /
//     residual -> global) runs while the MMAs of tile i + 1 fill the other accumulator, and TMEM allocation, barrier
//     initialisation and descriptor prefetch are paid once per SM instead of once per tile.
constexpr int CB_NS = 3;                      // ring stages: A raw 32 KB + w1 16 KB + w2 16 KB each
constexpr int CB_NTA = 4;                     // TMEM stages of split A
constexpr int CB_STAGE_BYTES = 4 * CT_TILE_BYTES;
constexpr int CB_TMEM_A0 = 256;               // cols [0,128) accumulator 0, [128,256) accumulator 1, [256,512) 4 A stages x 64
constexpr int CB_SM_BAR = CB_NS * CB_STAGE_BYTES;
constexpr int CB_N_BARS = 2 * CB_NS + 2 * CB_NTA + 4;
constexpr int CB_SM_MISC = CB_SM_BAR + 8 * CB_N_BARS;
constexpr int CB_SM_GN = CB_SM_MISC + 16;         // [4 epilogue warps][32 groups][2] floats of the GroupNorm epilogue
constexpr int CB_SM_BYTES = CB_SM_GN + 4 * 32 * 2 * 4;
constexpr int CB_SM_ALLOC = CB_SM_BYTES + 1024;
static_assert(CB_SM_ALLOC <= 232448, "conv_tc bf16 kernel exceeds 227 KB of shared memory");

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_bf16_m128(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// (x0, x1) -> packed bf16 pair of the leading terms (x0 in the low half) and of the remainders
__device__ __forceinline__ void split_bf16_pair(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

struct ConvTcTiles {
    int n_tiles;     // B * tiles_x * tiles_y * (Cout / 128)
    int nblk;        // Cout / 128 (fastest-varying: the CTAs that share an activation tile run back to back)
    int dbg;         // probe only (WMAR_CB_DBG): bit 0 skips x1.w2, bit 1 skips x2.w1
};

__global__ void __launch_bounds__(CT_THREADS, 1)
conv3x3_tc_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW1,
                       const __grid_constant__ CUtensorMap mapW2, const ConvTcArgs a, const ConvTcTiles tl) {
    using namespace tc05;
    extern __shared__ uint8_t ct_smem_raw[];
    const uint32_t smem_base = (smem_u32(ct_smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = ct_smem_raw + (smem_base - smem_u32(ct_smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tpi = a.tiles_x * a.tiles_y;
    const int cchunks = a.Cin / 64, nchunks = (a.taps == 1 ? 1 : 9) * cchunks;

    const uint32_t bar0 = smem_base + CB_SM_BAR;
    auto s_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto s_empty = [&](int s) { return bar0 + 8u * (uint32_t)(CB_NS + s); };
    auto ta_full = [&](int u) { return bar0 + 8u * (uint32_t)(2 * CB_NS + u); };
    auto ta_empty = [&](int u) { return bar0 + 8u * (uint32_t)(2 * CB_NS + CB_NTA + u); };
    auto acc_full = [&](int k) { return bar0 + 8u * (uint32_t)(2 * CB_NS + 2 * CB_NTA + k); };
    auto acc_empty = [&](int k) { return bar0 + 8u * (uint32_t)(2 * CB_NS + 2 * CB_NTA + 2 + k); };
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + CB_SM_MISC);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW1);
        tma_prefetch_desc(&mapW2);
        for (int s = 0; s < CB_NS; s++) { mbar_init(s_full(s), 1); mbar_init(s_empty(s), 9); }
        for (int u = 0; u < CB_NTA; u++) { mbar_init(ta_full(u), 8); mbar_init(ta_empty(u), 1); }
        for (int k = 0; k < 2; k++) { mbar_init(acc_full(k), 1); mbar_init(acc_empty(k), 4); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(s_tmem));
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *s_tmem;

    auto tile_coords = [&](int tile, int &b, int &x0, int &y0, int &n0) {
        const int nb = tile % tl.nblk, pt = tile / tl.nblk;
        b = pt / tpi;
        const int tr = pt - b * tpi, ty = tr / a.tiles_x, tx = tr - ty * a.tiles_x;
        x0 = tx * a.bw; y0 = ty * a.bh; n0 = nb * 128;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < tl.n_tiles; tile += gridDim.x) {
                int b, x0, y0, n0;
                tile_coords(tile, b, x0, y0, n0);
                for (int c = 0; c < nchunks; c++, g++) {
                    const uint32_t s = g % CB_NS;
                    const int tap = c / cchunks, ci0 = (c - tap * cchunks) * 64;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    mbar_wait(s_empty(s), ((g / CB_NS) & 1) ^ 1);
                    mbar_arrive_expect_tx(s_full(s), CB_STAGE_BYTES);
                    const uint32_t dst = smem_base + s * CB_STAGE_BYTES;
                    const int sx = a.stride * x0 + kx - a.pad, sy = a.stride * y0 + ky - a.pad;
                    tma_load_4d(dst, &mapA, ci0, sx, sy, b, s_full(s));                                // zero fill = padding
                    tma_load_4d(dst + CT_TILE_BYTES, &mapA, ci0 + 32, sx, sy, b, s_full(s));
                    tma_load_2d(dst + 2 * CT_TILE_BYTES, &mapW1, tap * a.Cin + ci0, n0, s_full(s), L2_EVICT_LAST);
                    tma_load_2d(dst + 3 * CT_TILE_BYTES, &mapW2, tap * a.Cin + ci0, n0, s_full(s), L2_EVICT_LAST);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t ID = idesc_bf16_m128(128);
            uint32_t g = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < tl.n_tiles; tile += gridDim.x, it++) {
                const uint32_t acc = tmem + (uint32_t)(it & 1) * 128u;
                mbar_wait(acc_empty(it & 1), ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                fence_after_sync();
                for (int c = 0; c < nchunks; c++, g++) {
                    const uint32_t s = g % CB_NS, u = g % CB_NTA;
                    mbar_wait(s_full(s), (g / CB_NS) & 1);          // the weight tiles of this stage have landed
                    mbar_wait(ta_full(u), (g / CB_NTA) & 1);        // the split activations are in tensor memory
                    fence_after_sync();
                    const uint32_t a_hi = tmem + CB_TMEM_A0 + u * 64, a_lo = a_hi + 32;
                    const uint32_t wb = smem_base + s * CB_STAGE_BYTES + 2 * CT_TILE_BYTES;   // w1 tile, w2 tile 16 KB further
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint64_t d1 = tc05::smem_desc_kmajor_sw128(wb + ks * 32);
                        const uint64_t d2 = tc05::smem_desc_kmajor_sw128(wb + CT_TILE_BYTES + ks * 32);
                        mma_bf16_ts(acc, a_hi + ks * 8, d1, ID, (c | ks) != 0 ? 1u : 0u);   // x1 . w1
                        if (!(tl.dbg & 1)) mma_bf16_ts(acc, a_hi + ks * 8, d2, ID, 1u);       // x1 . w2
                        if (!(tl.dbg & 2)) mma_bf16_ts(acc, a_lo + ks * 8, d1, ID, 1u);       // x2 . w1
                    }
                    mma_commit(ta_empty(u));
                    mma_commit(s_empty(s));
                }
                mma_commit(acc_full(it & 1));
            }
        }
        __syncwarp();
    } else if (warp < CT_W_EPI) {
        // ===================== activation converters (thread <-> pixel <-> TMEM lane) =====================
        // warps 2-5 convert channels 0-31 of every chunk, warps 6-9 channels 32-63: EVERY converter warp waits for EVERY
        // phase of every ring barrier (a warp that skipped phases could be satisfied by the wrong phase: try_wait only
        // sees the parity bit -- with 3 stages and two alternating groups a stage would alternate between the groups)
        const int h = (warp - CT_W_CONV) >> 2;
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < tl.n_tiles; tile += gridDim.x) {
            for (int c = 0; c < nchunks; c++, g++) {
                const uint32_t s = g % CB_NS, u = g % CB_NTA;
                mbar_wait_warp(s_full(s), (g / CB_NS) & 1, lane);
                const uint8_t *arow = smem + s * CB_STAGE_BYTES + h * CT_TILE_BYTES + r * 128;
                float4 w4[8];
#pragma unroll
                for (int k = 0; k < 8; k++) w4[k] = *reinterpret_cast<const float4 *>(arow + ((k ^ (r & 7)) << 4));
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    split_bf16_pair(w4[k].x, w4[k].y, hi[2 * k], lo[2 * k]);
                    split_bf16_pair(w4[k].z, w4[k].w, hi[2 * k + 1], lo[2 * k + 1]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty(s));          // the raw tile is in registers
                mbar_wait_warp(ta_empty(u), ((g / CB_NTA) & 1) ^ 1, lane);
                fence_after_sync();
                const uint32_t a_hi = t_lane + CB_TMEM_A0 + u * 64 + h * 16, a_lo = a_hi + 32;
                tmem_st16(a_hi, hi);
                tmem_st16(a_lo, lo);
                wait_st();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(ta_full(u));
            }
        }
    } else {
        // ===================== epilogue (thread <-> pixel) =====================
        const int q = warp & 3, r = q * 32 + lane;
        const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < tl.n_tiles; tile += gridDim.x, it++) {
            int b, x0, y0, n0;
            tile_coords(tile, b, x0, y0, n0);
            const int py = y0 + r / a.bw, px = x0 + r % a.bw;
            const size_t pix = ((size_t)b * a.H + py) * a.W + px;
            float *dst = a.out + pix * a.Cout + n0;
            const float *res = a.resid != nullptr ? a.resid + pix * a.Cout + n0 : nullptr;
            const uint32_t acc = t_lane + (uint32_t)(it & 1) * 128u;
            mbar_wait_warp(acc_full(it & 1), (it >> 1) & 1, lane);
            fence_after_sync();
            float gsum[32], gsq[32];                  // (only live when a.gn_out != nullptr; indexed by compile-time constants)
#pragma unroll
            for (int gI = 0; gI < 32; gI++) { gsum[gI] = 0.f; gsq[gI] = 0.f; }
            const int gshift = a.gn_cg == 4 ? 0 : (a.gn_cg == 8 ? 1 : 2);   // quads per group = 1, 2, 4
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                const int c0 = cc * 32;
                uint32_t d0[32];
                tmem_ld32(acc + c0, d0);
                wait_ld();
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + c0 + k));
                    float4 v;
                    v.x = __uint_as_float(d0[k]) + bb.x;
                    v.y = __uint_as_float(d0[k + 1]) + bb.y;
                    v.z = __uint_as_float(d0[k + 2]) + bb.z;
                    v.w = __uint_as_float(d0[k + 3]) + bb.w;
                    if (res != nullptr) {
                        const float4 r4 = *reinterpret_cast<const float4 *>(res + c0 + k);
                        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
                    }
                    *reinterpret_cast<float4 *>(dst + c0 + k) = v;
                    if (a.gn_out != nullptr) {
                        const float s1 = (v.x + v.y) + (v.z + v.w), s2 = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                        const int quad = cc * 8 + k / 4;          // compile-time: 0..31
                        // group of this quad within the block: quad >> gshift; the three cases keep the index static
                        if (gshift == 0) { gsum[quad] += s1; gsq[quad] += s2; }
                        else if (gshift == 1) { gsum[quad >> 1] += s1; gsq[quad >> 1] += s2; }
                        else { gsum[quad >> 2] += s1; gsq[quad >> 2] += s2; }
                    }
                }
            }
            // every TMEM read of this accumulator has completed (wait_ld above): hand it back to the MMA issuer
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(it & 1));
            if (a.gn_out != nullptr) {
                // GroupNorm statistics of this tile: gsum / gsq hold, per lane (= pixel), the sums over the channels of each
                // group of this 128-channel block; reduce over the warp's 32 pixels, then over the 4 warps in warp order
                const int ng = 128 / a.gn_cg;                              // groups in this block: 32, 16 or 8
                float *gs = reinterpret_cast<float *>(smem + CB_SM_GN) + q * 64;
#pragma unroll
                for (int gI = 0; gI < 32; gI++) {
                    if (gI >= ng) break;                                   // (warp-uniform)
                    float s1 = gsum[gI], s2 = gsq[gI];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    }
                    if (lane == 0) { gs[2 * gI] = s1; gs[2 * gI + 1] = s2; }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps
                const int et = (warp - CT_W_EPI) * 32 + lane;              // 0..127
                if (et < ng) {
                    const float *g0 = reinterpret_cast<const float *>(smem + CB_SM_GN);
                    double d1 = 0.0, d2 = 0.0;
#pragma unroll
                    for (int w = 0; w < 4; w++) { d1 += (double)g0[w * 64 + 2 * et]; d2 += (double)g0[w * 64 + 2 * et + 1]; }
                    const int pt = tile / tl.nblk, bimg = pt / tpi, tin = pt - bimg * tpi;
                    a.gn_out[((size_t)bimg * tpi + tin) * 32 + n0 / a.gn_cg + et] = make_double2(d1, d2);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");            // gs is reused by the next tile
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 1) tmem_dealloc<512>(tmem);
}

// w -> (w1, w2) bf16 planes: w1 = bf16_rn(w), w2 = bf16_rn(w - w1)
__global__ void conv_tc_wsplit_bf16_kernel(const float *__restrict__ w, uint16_t *__restrict__ w1, uint16_t *__restrict__ w2, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = w[i];
        uint32_t hi, lo;
        split_bf16_pair(x, 0.f, hi, lo);
        w1[i] = (uint16_t)(hi & 0xffffu);
        w2[i] = (uint16_t)(lo & 0xffffu);
    }
}

// w_lo = w - trunc_tf32(w): the part of the weight the tensor core drops when it reads fp32 bits as TF32
__global__ void conv_tc_wlo_kernel(const float *__restrict__ w, float *__restrict__ wlo, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = w[i];
        wlo[i] = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    }
}

}  // namespace wmar
