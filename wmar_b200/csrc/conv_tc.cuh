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
    const int cchunks = a.Cin / 32, nchunks = 9 * cchunks;

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
                tma_load_4d(dst, &mapA, ci0, x0 + kx - 1, y0 + ky - 1, b, s_full(s));           // zero fill = padding
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

// w_lo = w - trunc_tf32(w): the part of the weight the tensor core drops when it reads fp32 bits as TF32
__global__ void conv_tc_wlo_kernel(const float *__restrict__ w, float *__restrict__ wlo, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = w[i];
        wlo[i] = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    }
}

}  // namespace wmar
