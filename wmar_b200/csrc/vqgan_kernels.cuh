// CUDA kernels of the VQGAN tokenizer (Taming / Chameleon and MaskGIT families); NHWC fp32 activations in HBM.
//   conv_igemm_kernel   3x3 / 1x1 convolution as an implicit GEMM on the tensor pipe (TF32 mma, 1x or 3x split),
//                       cp.async double-buffered smem tiles, fused bias / residual add / nearest-upsample indexing /
//                       asymmetric-pad stride-2 / NCHW+clamp output           (model.py:39-76,79-138; vqgan.py:64-73)
//   gn_partial_kernel + gn_apply_kernel   GroupNorm(32, eps 1e-6) statistics and normalise(+swish)  (model.py:30-36)
//   attn_block_kernel   single-head spatial attention of AttnBlock                                  (model.py:169-193)
//   gather / rowsumsq / argmin           codebook lookup and nearest-neighbour                      (quantize.py:272-331)
#pragma once
#include "common.cuh"
#include "conv_igemm.cuh"
#include "gemm.cuh"  // split_tf32, mma_tf32

namespace wmar {

// ------------------------------------------------------------------------------------------- GroupNorm(32)
// Stage 1: per (image, pixel-chunk) partial (sum, sum of squares) per group in fp64.
constexpr int GN_THREADS = 256;
__global__ void __launch_bounds__(GN_THREADS) gn_partial_kernel(const float *__restrict__ x, int HW, int C, int nchunk,
                                                                double2 *__restrict__ partial) {
    // grid (nchunk, B); thread -> fixed channel quad (C/4 quads), strides over pixels of the chunk
    __shared__ double2 sh[32][GN_THREADS / 32 + 1];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int quads = C >> 2;
    const int pix_per_chunk = HW / nchunk;
    const int cg = C / 32;  // channels per group (4, 8, 16)
    const float *base = x + ((size_t)b * HW + (size_t)chunk * pix_per_chunk) * C;
    const int tid = threadIdx.x;
    // threads are laid out so that a thread always sees the same channel quad: requires GN_THREADS % quads == 0 or
    // quads % GN_THREADS == 0; C in {128,256,512} -> quads in {32,64,128}
    const int q = tid % quads, prow = tid / quads, pstride = GN_THREADS / quads;
    float s = 0.f, ss = 0.f;
    double ds = 0.0, dss = 0.0;
    int cnt = 0;
    int p = prow;
    for (; p + 3 * pstride < pix_per_chunk; p += 4 * pstride) {     // four independent loads in flight, same summation order
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = *reinterpret_cast<const float4 *>(base + (size_t)(p + u * pstride) * C + q * 4);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            s += v[u].x + v[u].y + v[u].z + v[u].w;
            ss += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
            if (++cnt == 64) { ds += s; dss += ss; s = 0.f; ss = 0.f; cnt = 0; }
        }
    }
    for (; p < pix_per_chunk; p += pstride) {
        float4 v = *reinterpret_cast<const float4 *>(base + (size_t)p * C + q * 4);
        s += v.x + v.y + v.z + v.w;
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        if (++cnt == 64) { ds += s; dss += ss; s = 0.f; ss = 0.f; cnt = 0; }
    }
    ds += s; dss += ss;
    // reduce: group of this quad = (q*4)/cg; several quads and several pixel rows map to one group
    const int grp = (q * 4) / cg;
    // shared accumulation in a fixed order: first write per-thread values, then one thread per group sums them
    __shared__ double2 vals[GN_THREADS];
    vals[tid] = make_double2(ds, dss);
    __syncthreads();
    if (tid < 32) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < GN_THREADS; i++) {
            int qi = i % quads;
            if ((qi * 4) / cg == tid) { a0 += vals[i].x; a1 += vals[i].y; }
        }
        partial[((size_t)b * nchunk + chunk) * 32 + tid] = make_double2(a0, a1);
    }
    (void)sh; (void)grp;
}

// Stage 1': the statistics came from the producing conv's epilogue as per-tile partials [B][tiles][32] (conv_tc.cuh): sum
// them in a fixed order into the [B][1][32] layout stage 2 reads with nchunk = 1.  grid B, 256 threads.
__global__ void __launch_bounds__(256) gn_finalize_kernel(const double2 *__restrict__ tile_partial, int tiles,
                                                          double2 *__restrict__ partial) {
    __shared__ double2 sh[8][32];
    const int b = blockIdx.x, g = threadIdx.x & 31, part = threadIdx.x >> 5;
    double a0 = 0.0, a1 = 0.0;
    for (int t = part; t < tiles; t += 8) {
        const double2 p = tile_partial[((size_t)b * tiles + t) * 32 + g];
        a0 += p.x; a1 += p.y;
    }
    sh[part][g] = make_double2(a0, a1);
    __syncthreads();
    if (part == 0) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) { s0 += sh[k][g].x; s1 += sh[k][g].y; }
        partial[(size_t)b * 32 + g] = make_double2(s0, s1);
    }
}

// Stage 2: y = act((x - mean_g) * rstd_g * gamma_c + beta_c); act = swish (x*sigmoid(x)) or identity.
__global__ void __launch_bounds__(256) gn_apply_kernel(const float *__restrict__ x, float *__restrict__ y, int HW, int C,
                                                       int nchunk, const double2 *__restrict__ partial,
                                                       const float *__restrict__ gamma, const float *__restrict__ beta,
                                                       float eps, int swish) {
    __shared__ float s_mean[32], s_rstd[32];
    const int b = blockIdx.y;
    if (threadIdx.x < 32) {
        double a0 = 0.0, a1 = 0.0;
        for (int c = 0; c < nchunk; c++) {
            double2 p = partial[((size_t)b * nchunk + c) * 32 + threadIdx.x];
            a0 += p.x; a1 += p.y;
        }
        const double n = (double)HW * (C / 32);
        const double mean = a0 / n;
        double var = a1 / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int cg = C / 32;
    const size_t total4 = (size_t)HW * C / 4;
    const float4 *xb = reinterpret_cast<const float4 *>(x + (size_t)b * HW * C);
    float4 *yb = reinterpret_cast<float4 *>(y + (size_t)b * HW * C);
    // a CTA owns a CONTIGUOUS range of the image (multiple of 1024 quads) and walks it 256 quads at a time: the walk stride
    // (256 quads) is a multiple of C / 4 (C / 4 in {32, 64, 128}), so a thread sees the same channel quad in every iteration
    // and gamma / beta / statistics are loop invariants; four independent 16-byte loads per iteration, 16 KB contiguous per
    // CTA iteration (the grid-strided version touched four pages 4 MB apart per iteration and ran at 2.1-3.3 TB/s)
    const size_t per = ((total4 + gridDim.x - 1) / gridDim.x + 1023) / 1024 * 1024;
    const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < total4 ? lo + per : total4;
    const size_t i0 = lo + threadIdx.x, stride = blockDim.x;
    const int c = (int)((i0 * 4) % C);
    const float mean = s_mean[c / cg], rstd = s_rstd[c / cg];
    const float4 g4 = *reinterpret_cast<const float4 *>(gamma + c);
    const float4 b4 = *reinterpret_cast<const float4 *>(beta + c);
    auto act = [&](float4 v) {
        float o[4] = {(v.x - mean) * rstd * g4.x + b4.x, (v.y - mean) * rstd * g4.y + b4.y,
                      (v.z - mean) * rstd * g4.z + b4.z, (v.w - mean) * rstd * g4.w + b4.w};
        if (swish) {
#pragma unroll
            for (int e = 0; e < 4; e++) o[e] = o[e] * (1.0f / (1.0f + expf(-o[e])));
        }
        return make_float4(o[0], o[1], o[2], o[3]);
    };
    size_t i = i0;
    for (; i + 3 * stride < hi; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldcs(xb + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; u++) yb[i + u * stride] = act(v[u]);
    }
    for (; i < hi; i += stride) yb[i] = act(__ldcs(xb + i));
}

// ------------------------------------------------------------------------------------------- decoder tail
// norm_out -> swish -> conv_out (C -> 3, 3x3, pad 1) -> clamp -> NCHW image, one kernel (model.py:533-538,
// maskgit_vqgan.py decoder tail).  As an implicit GEMM this layer wastes the tensor tile (N = 3 of 64 columns: 3.4 ms per
// 16 images at 256^2); it is a 3-row GEMV per pixel, so it runs on the fp32 FMA pipe: a CTA owns a 16 x 16 pixel tile,
// stages the 18 x 18 halo of 32 channels at a time in shared memory (GroupNorm + swish applied while staging, zeros outside
// the image = the conv's padding of the ACTIVATED tensor) and every thread accumulates its pixel's 3 outputs in plain
// fp32 (no TF32 split needed).  Pixel stride 36 floats: the 8 lanes of an LDS.128 phase hit 8 different bank quads.
constexpr int CO_T = 16, CO_HALO = CO_T + 2, CO_CH = 32, CO_LD = CO_CH + 4;
constexpr int CO_SMEM_BYTES = (CO_HALO * CO_HALO * CO_LD + 3 * 9 * CO_CH) * 4;
struct ConvOutArgs {
    const float *x;          // NHWC [B][H][W][C], input of the GroupNorm
    const float *w;          // [>= 3][3][3][C]
    const float *bias;
    float *out;              // NCHW [B][3][H][W]
    int H, W, C;
    const double2 *gn_partial; int nchunk;     // statistics of gn_partial_kernel over x
    const float *gamma, *beta; float eps;
    float clamp_lo, clamp_hi, out_scale, out_shift;
};
__global__ void __launch_bounds__(CO_T * CO_T) conv_out3_kernel(const ConvOutArgs a) {
    extern __shared__ __align__(16) float co_smem[];
    float *tile = co_smem;                                   // [18 * 18][CO_LD]
    float *wsm = co_smem + CO_HALO * CO_HALO * CO_LD;        // [3][9][CO_CH]
    __shared__ float s_mean[32], s_rstd[32];
    const int tid = threadIdx.x, tx = tid % CO_T, ty = tid / CO_T;
    const int b = blockIdx.z, x0 = blockIdx.x * CO_T, y0 = blockIdx.y * CO_T;
    const int HW = a.H * a.W, cg = a.C / 32;
    if (tid < 32) {
        double a0 = 0.0, a1 = 0.0;
        for (int c = 0; c < a.nchunk; c++) {
            const double2 p = a.gn_partial[((size_t)b * a.nchunk + c) * 32 + tid];
            a0 += p.x; a1 += p.y;
        }
        const double n = (double)HW * cg;
        const double mean = a0 / n;
        double var = a1 / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[tid] = (float)mean;
        s_rstd[tid] = (float)(1.0 / sqrt(var + (double)a.eps));
    }
    float acc[3] = {0.f, 0.f, 0.f};
    const float *xb = a.x + (size_t)b * HW * a.C;
    for (int c0 = 0; c0 < a.C; c0 += CO_CH) {
        __syncthreads();   // statistics visible (first pass) / previous chunk consumed
        // stage: 324 halo pixels x 8 channel quads
        for (int i = tid; i < CO_HALO * CO_HALO * (CO_CH / 4); i += CO_T * CO_T) {
            const int q = i % (CO_CH / 4), p = i / (CO_CH / 4);
            const int hy = p / CO_HALO, hx = p - hy * CO_HALO;
            const int gy = y0 + hy - 1, gx = x0 + hx - 1;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
                const int c = c0 + 4 * q;
                const float4 v = *reinterpret_cast<const float4 *>(xb + ((size_t)gy * a.W + gx) * a.C + c);
                const float4 g4 = *reinterpret_cast<const float4 *>(a.gamma + c);
                const float4 b4 = *reinterpret_cast<const float4 *>(a.beta + c);
                const float mean = s_mean[c / cg], rstd = s_rstd[c / cg];   // a quad never straddles a group (cg >= 4)
                float t[4] = {(v.x - mean) * rstd * g4.x + b4.x, (v.y - mean) * rstd * g4.y + b4.y,
                              (v.z - mean) * rstd * g4.z + b4.z, (v.w - mean) * rstd * g4.w + b4.w};
#pragma unroll
                for (int e = 0; e < 4; e++) t[e] = t[e] * (1.0f / (1.0f + expf(-t[e])));
                o = make_float4(t[0], t[1], t[2], t[3]);
            }
            *reinterpret_cast<float4 *>(tile + p * CO_LD + 4 * q) = o;
        }
        for (int i = tid; i < 3 * 9 * (CO_CH / 4); i += CO_T * CO_T) {
            const int q = i % (CO_CH / 4), ot = i / (CO_CH / 4);      // ot = o * 9 + tap
            *reinterpret_cast<float4 *>(wsm + ot * CO_CH + 4 * q) =
                *reinterpret_cast<const float4 *>(a.w + (size_t)ot * a.C + c0 + 4 * q);
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < 9; tap++) {
            const float *px = tile + ((ty + tap / 3) * CO_HALO + tx + tap % 3) * CO_LD;
#pragma unroll
            for (int q = 0; q < CO_CH / 4; q++) {
                const float4 v = *reinterpret_cast<const float4 *>(px + 4 * q);
#pragma unroll
                for (int o = 0; o < 3; o++) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(wsm + (o * 9 + tap) * CO_CH + 4 * q);
                    acc[o] = fmaf(v.x, w4.x, acc[o]); acc[o] = fmaf(v.y, w4.y, acc[o]);
                    acc[o] = fmaf(v.z, w4.z, acc[o]); acc[o] = fmaf(v.w, w4.w, acc[o]);
                }
            }
        }
    }
    const int gy = y0 + ty, gx = x0 + tx;
    if (gy < a.H && gx < a.W) {
#pragma unroll
        for (int o = 0; o < 3; o++) {
            float v = acc[o] + a.bias[o];
            v = fminf(fmaxf(v, a.clamp_lo), a.clamp_hi);
            v = v * a.out_scale + a.out_shift;
            a.out[(((size_t)b * 3 + o) * a.H + gy) * a.W + gx] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------- encoder head
// conv_in (3 -> Cout, 3x3, pad 1; model.py:368-372, maskgit_vqgan.py encoder head) straight from the caller's NCHW image:
// K = 27, so as an implicit GEMM (input padded to 32 channels) 29/32 of the tensor work multiplies zeros (0.66 ms + 0.13 ms
// for the NHWC repack per 16 images).  Here a CTA owns 32 x 2 output pixels: the 34 x 4 x 3 input halo and the
// [27][Cout] weights sit in shared memory, a thread accumulates 8 pixels x 4 output channels in fp32 (exact products).
constexpr int CI_TW = 32, CI_TH = 8;   // 256 pixels per CTA: the 13.8 KB weight stage is amortised over four row pairs
struct ConvInArgs {
    const float *img;        // NCHW [B][3][H][W]
    const float *w;          // [Cout][3][3][Cin_pad] (first 3 input channels real)
    const float *bias;
    float *out;              // NHWC [B][H][W][Cout]
    int H, W, Cout, Cin_pad;
    float scale, shift;      // v = x * scale + shift before the conv (RAR: (x+1)/2), padding stays zero
};
__global__ void __launch_bounds__(256, 3) conv_in3_kernel(const ConvInArgs a) {
    extern __shared__ __align__(16) float ci_smem[];
    float *wsm = ci_smem;                                   // [27][Cout]
    float *xin = ci_smem + 27 * a.Cout;                     // [3][CI_TH + 2][CI_TW + 2]
    const int tid = threadIdx.x, b = blockIdx.z, x0 = blockIdx.x * CI_TW, y0 = blockIdx.y * CI_TH;
    for (int i = tid; i < 27 * a.Cout; i += 256) {
        const int n = i % a.Cout, j = i / a.Cout, tap = j / 3, c = j - tap * 3;      // j = tap * 3 + c
        wsm[i] = a.w[((size_t)n * 9 + tap) * a.Cin_pad + c];
    }
    for (int i = tid; i < 3 * (CI_TH + 2) * (CI_TW + 2); i += 256) {
        const int hx = i % (CI_TW + 2), r = i / (CI_TW + 2), hy = r % (CI_TH + 2), c = r / (CI_TH + 2);
        const int gy = y0 + hy - 1, gx = x0 + hx - 1;
        xin[i] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? a.img[(((size_t)b * 3 + c) * a.H + gy) * a.W + gx] * a.scale + a.shift : 0.f;
    }
    __syncthreads();
    const int pg = tid >> 5, pxb = (pg & 3) * 8;                       // 8 pixels of one row per warp, two rows per round
#pragma unroll 1
    for (int rr = 0; rr < CI_TH / 2; rr++) {
        const int py = 2 * rr + (pg >> 2);
        for (int n4 = (tid & 31) * 4; n4 < a.Cout; n4 += 128) {
            float acc[8][4];
#pragma unroll
            for (int p = 0; p < 8; p++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[p][e] = 0.f;
#pragma unroll 3
            for (int tap = 0; tap < 9; tap++) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(wsm + (tap * 3 + c) * a.Cout + n4);
                    const float *xr = xin + (c * (CI_TH + 2) + py + tap / 3) * (CI_TW + 2) + pxb + tap % 3;
#pragma unroll
                    for (int p = 0; p < 8; p++) {
                        const float xv = xr[p];
                        acc[p][0] = fmaf(xv, w4.x, acc[p][0]); acc[p][1] = fmaf(xv, w4.y, acc[p][1]);
                        acc[p][2] = fmaf(xv, w4.z, acc[p][2]); acc[p][3] = fmaf(xv, w4.w, acc[p][3]);
                    }
                }
            }
            const float4 b4 = *reinterpret_cast<const float4 *>(a.bias + n4);
            const int gy = y0 + py;
#pragma unroll
            for (int p = 0; p < 8; p++) {
                const int gx = x0 + pxb + p;
                if (gy < a.H && gx < a.W)
                    *reinterpret_cast<float4 *>(a.out + (((size_t)b * a.H + gy) * a.W + gx) * a.Cout + n4) =
                        make_float4(acc[p][0] + b4.x, acc[p][1] + b4.y, acc[p][2] + b4.z, acc[p][3] + b4.w);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- AttnBlock core
// out[b][i][:] = sum_j softmax_j(q_i . k_j * C^-0.5) v_j ; q,k,v,out NHWC [B][N][C].  grid (N/16, B), 256 threads.
// Round 2: register tiles that halve / quarter the shared-memory reads per FMA (the first version issued one broadcast
// LDS per 4 FMAs and ran at 126 GB/s, 200 us per launch): q.k^T as 2 keys x 8 queries per thread, the probabilities
// stored key-major [N][16] so that p.v reads them as four LDS.128 per key for 2 channels x 16 queries of FMAs.
constexpr int AB_Q = 16;
__global__ void __launch_bounds__(256) attn_block_kernel(const float *__restrict__ q, const float *__restrict__ k,
                                                         const float *__restrict__ v, float *__restrict__ out, int N,
                                                         int C) {
    extern __shared__ __align__(16) float sm[];
    float *qs = sm;               // [AB_Q][C]
    float *ps = sm + AB_Q * C;    // [N][AB_Q]  (key-major)
    const int b = blockIdx.y, i0 = blockIdx.x * AB_Q, tid = threadIdx.x;
    const float *qb = q + ((size_t)b * N + i0) * C;
    for (int i = tid; i < AB_Q * C / 4; i += 256) reinterpret_cast<float4 *>(qs)[i] = reinterpret_cast<const float4 *>(qb)[i];
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)C);  // int(c)**(-0.5)
    // ---- scores: thread = (key pair, query half); keys j and j + 128 * ceil, queries qh * 8 .. qh * 8 + 7 ----
    const int qh = tid & 1, kp = tid >> 1;         // 128 key pairs per sweep, 2 query halves
    for (int j0 = kp; j0 < N; j0 += 256) {
        const int j1 = j0 + 128;
        const bool has1 = j1 < N;
        const float *ka = k + ((size_t)b * N + j0) * C;
        const float *kb = k + ((size_t)b * N + (has1 ? j1 : j0)) * C;
        float acc0[8], acc1[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { acc0[i] = 0.f; acc1[i] = 0.f; }
        for (int c = 0; c < C; c += 4) {
            const float4 a4 = *reinterpret_cast<const float4 *>(ka + c);
            const float4 b4 = *reinterpret_cast<const float4 *>(kb + c);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float4 q4 = *reinterpret_cast<const float4 *>(qs + (qh * 8 + i) * C + c);
                acc0[i] += q4.x * a4.x + q4.y * a4.y + q4.z * a4.z + q4.w * a4.w;
                acc1[i] += q4.x * b4.x + q4.y * b4.y + q4.z * b4.z + q4.w * b4.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ps[j0 * AB_Q + qh * 8 + i] = acc0[i] * scale;
            if (has1) ps[j1 * AB_Q + qh * 8 + i] = acc1[i] * scale;
        }
    }
    __syncthreads();
    // softmax per query row: 8 warps, 2 rows each
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < AB_Q; i += 8) {
        float m = -INFINITY;
        for (int j = lane; j < N; j += 32) m = fmaxf(m, ps[j * AB_Q + i]);
        m = warp_max(m);
        float s = 0.f;
        for (int j = lane; j < N; j += 32) { float e = expf(ps[j * AB_Q + i] - m); ps[j * AB_Q + i] = e; s += e; }
        s = warp_sum(s);
        const float inv = 1.0f / s;
        for (int j = lane; j < N; j += 32) ps[j * AB_Q + i] *= inv;
    }
    __syncthreads();
    // ---- out = P V: thread = channels c and c + 256 (same summation order over j as before) ----
    for (int c = tid; c < C; c += 512) {
        const bool has1 = c + 256 < C;
        float acc0[AB_Q], acc1[AB_Q];
#pragma unroll
        for (int i = 0; i < AB_Q; i++) { acc0[i] = 0.f; acc1[i] = 0.f; }
        const float *vb = v + (size_t)b * N * C;
        for (int j = 0; j < N; j++) {
            const float va = vb[(size_t)j * C + c];
            const float vc = has1 ? vb[(size_t)j * C + c + 256] : 0.f;
#pragma unroll
            for (int i4 = 0; i4 < AB_Q / 4; i4++) {
                const float4 p4 = *reinterpret_cast<const float4 *>(ps + j * AB_Q + 4 * i4);
                acc0[4 * i4] += p4.x * va; acc0[4 * i4 + 1] += p4.y * va; acc0[4 * i4 + 2] += p4.z * va; acc0[4 * i4 + 3] += p4.w * va;
                acc1[4 * i4] += p4.x * vc; acc1[4 * i4 + 1] += p4.y * vc; acc1[4 * i4 + 2] += p4.z * vc; acc1[4 * i4 + 3] += p4.w * vc;
            }
        }
#pragma unroll
        for (int i = 0; i < AB_Q; i++) {
            out[((size_t)b * N + i0 + i) * C + c] = acc0[i];
            if (has1) out[((size_t)b * N + i0 + i) * C + c + 256] = acc1[i];
        }
    }
}

// ------------------------------------------------------------------------------------------- small kernels
// NCHW image (3 channels) -> NHWC with Cpad channels (zeros beyond 3); v = x*scale + shift
__global__ void nchw_to_nhwc_pad_kernel(const float *__restrict__ img, float *__restrict__ out, int B, int HW, int Cpad,
                                        float scale, float shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * HW * Cpad;
    if (i >= total) return;
    int c = (int)(i % Cpad);
    size_t p = i / Cpad;
    int b = (int)(p / HW);
    int r = (int)(p - (size_t)b * HW);
    out[i] = c < 3 ? img[((size_t)b * 3 + c) * HW + r] * scale + shift : 0.f;
}

// zq[b][p][:] = codebook[codes[b][p]][:]
__global__ void codebook_gather_kernel(const int64_t *__restrict__ codes, const float *__restrict__ emb, float *__restrict__ out,
                                       size_t n_tok, int D, int n_e) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tok * (size_t)(D / 4)) return;
    size_t tok = i / (D / 4);
    int c4 = (int)(i - tok * (D / 4));
    long long id = codes[tok];
    if (id < 0 || id >= n_e) id = 0;
    reinterpret_cast<float4 *>(out)[i] = reinterpret_cast<const float4 *>(emb + (size_t)id * D)[c4];
}

// out[r] = sum_c x[r][c]^2   (one warp per row)
__global__ void row_sumsq_kernel(const float *__restrict__ x, float *__restrict__ out, int rows, int D) {
    int r = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (r >= rows) return;
    int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { float v = x[(size_t)r * D + c]; s += v * v; }
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
}

// codes[r] = argmin_j (zz[r] + ee[j]) - 2 dots[r][j]   (first index on ties)   quantize.py:281-285
__global__ void __launch_bounds__(256) vq_argmin_kernel(const float *__restrict__ dots, const float *__restrict__ zz,
                                                        const float *__restrict__ ee, int64_t *__restrict__ codes, int n_e) {
    __shared__ float sv[8];
    __shared__ int si[8];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float z = zz[r];
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int j = tid; j < n_e; j += 256) {
        float d = (z + ee[j]) - 2.0f * dots[(size_t)r * n_e + j];
        if (d < best) { best = d; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = best; si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; w++)
            if (sv[w] < best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        codes[r] = bi;
    }
}

// 2x2 average pool, NHWC
__global__ void avgpool2_kernel(const float *__restrict__ x, float *__restrict__ y, int B, int Ho, int Wo, int C) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)B * Ho * Wo * (C / 4);
    if (i >= total) return;
    int c4 = (int)(i % (C / 4));
    size_t p = i / (C / 4);
    int ox = (int)(p % Wo);
    size_t q = p / Wo;
    int oy = (int)(q % Ho);
    int b = (int)(q / Ho);
    const int Wi = Wo * 2, Hi = Ho * 2;
    const float4 *src = reinterpret_cast<const float4 *>(x);
    size_t base = (((size_t)b * Hi + 2 * oy) * Wi + 2 * ox) * (C / 4) + c4;
    float4 a = src[base], bq = src[base + (C / 4)], c = src[base + (size_t)Wi * (C / 4)], d = src[base + (size_t)Wi * (C / 4) + (C / 4)];
    reinterpret_cast<float4 *>(y)[i] = make_float4((a.x + bq.x + c.x + d.x) * 0.25f, (a.y + bq.y + c.y + d.y) * 0.25f,
                                                   (a.z + bq.z + c.z + d.z) * 0.25f, (a.w + bq.w + c.w + d.w) * 0.25f);
}

}  // namespace wmar
