// Implicit-GEMM convolution on the tensor pipe (TF32 mma.sync, 1x or 3x split), NHWC fp32: the 128 x 64 x 32 tile kernel
// behind the VQGAN's 1x1 / strided / first / last convolutions (model.py:39-76,79-138; vqgan.py:64-73), the codebook
// dot products (quantize.py:272-331) and -- as a plain [M][K] x [N][K]^T GEMM (ks = 1, B = Hs = 1, Ws = M) -- the hoisted
// adaLN tables of the RAR engine (rar.py:180).  Templates only: safe to include from several translation units.
#pragma once
#include "common.cuh"
#include "gemm.cuh"  // split_tf32, mma_tf32

namespace wmar {

constexpr int CV_THREADS = 256;
constexpr int CV_BM = 128, CV_BN = 64, CV_BK = 32, CV_LD = 36;

struct ConvArgs {
    const float *in, *w, *bias, *resid;
    float *out;
    int B, Hs, Ws, Cin, Ho, Wo, Cout, Cout_pad;
    int ks, stride, pad, up;
    int nchw_out, do_clamp;
    float clamp_lo, clamp_hi;
    float out_scale, out_shift;  // applied before the clamp when nchw_out (RAR: clamp(0,1) then *2-1 handled by caller)
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int PREC3>
__global__ void __launch_bounds__(CV_THREADS) conv_igemm_kernel(ConvArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *As = smem;                               // [2][CV_BM][CV_LD]
    float *Bs = smem + 2 * CV_BM * CV_LD;           // [2][CV_BN][CV_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.x * CV_BM, n0 = blockIdx.y * CV_BN;
    const int Hl = a.up ? a.Hs * 2 : a.Hs, Wl = a.up ? a.Ws * 2 : a.Ws;
    const int cchunks = a.Cin / CV_BK;
    const int nchunks = a.ks * a.ks * cchunks;

    // the 4 A rows and 2 B rows this thread copies every chunk
    int ab[4], aoy[4], aox[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int row = (tid + i * CV_THREADS) >> 3;
        int m = m0 + row;
        int b = m / (a.Ho * a.Wo);
        int r = m - b * (a.Ho * a.Wo);
        ab[i] = b; aoy[i] = r / a.Wo; aox[i] = r - aoy[i] * a.Wo;
    }
    const int seg = tid & 7;

    auto load_chunk = [&](int c, int buf) {
        const int tap = c / cchunks, ci0 = (c - tap * cchunks) * CV_BK;
        const int ky = tap / a.ks, kx = tap - ky * a.ks;
        float *Ab = As + buf * CV_BM * CV_LD, *Bb = Bs + buf * CV_BN * CV_LD;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int row = (tid + i * CV_THREADS) >> 3;
            int iy = aoy[i] * a.stride + ky - a.pad, ix = aox[i] * a.stride + kx - a.pad;
            bool valid = iy >= 0 && iy < Hl && ix >= 0 && ix < Wl;
            int sy = a.up ? iy >> 1 : iy, sx = a.up ? ix >> 1 : ix;
            const float *src = valid ? a.in + (((size_t)ab[i] * a.Hs + sy) * a.Ws + sx) * a.Cin + ci0 + seg * 4 : a.in;
            cp_async16(Ab + row * CV_LD + seg * 4, src, valid);
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
            int row = (tid + i * CV_THREADS) >> 3;
            const float *src = a.w + ((size_t)(n0 + row) * (a.ks * a.ks) + tap) * a.Cin + ci0 + seg * 4;
            cp_async16(Bb + row * CV_LD + seg * 4, src, true);
        }
        cp_async_commit();
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[i][j][c] = 0.f;

    load_chunk(0, 0);
    for (int c = 0; c < nchunks; c++) {
        const int buf = c & 1;
        if (c + 1 < nchunks) {
            load_chunk(c + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float *Ab = As + buf * CV_BM * CV_LD + (wm * 32) * CV_LD;
        const float *Bb = Bs + buf * CV_BN * CV_LD + (wn * 32) * CV_LD;
#pragma unroll
        for (int k8 = 0; k8 < CV_BK / 8; k8++) {
            float af[2][4], bf[4][2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const float *p = Ab + (i * 16 + g) * CV_LD + k8 * 8 + t;
                af[i][0] = p[0]; af[i][1] = p[8 * CV_LD]; af[i][2] = p[4]; af[i][3] = p[8 * CV_LD + 4];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float *p = Bb + (j * 8 + g) * CV_LD + k8 * 8 + t;
                bf[j][0] = p[0]; bf[j][1] = p[4];
            }
            if (PREC3) {
                uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int e = 0; e < 4; e++) split_tf32(af[i][e], ah[i][e], al[i][e]);
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) split_tf32(bf[j][e], bh[j][e], bl[j][e]);
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        mma_tf32(acc[i][j], al[i][0], al[i][1], al[i][2], al[i][3], bh[j][0], bh[j][1]);
                        mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bl[j][0], bl[j][1]);
                        mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh[j][0], bh[j][1]);
                    }
            } else {
                uint32_t ah[2][4], bh[4][2];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int e = 0; e < 4; e++) ah[i][e] = tf32_hi(af[i][e]);
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) bh[j][e] = tf32_hi(bf[j][e]);
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        mma_tf32(acc[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh[j][0], bh[j][1]);
            }
        }
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int m = m0 + wm * 32 + i * 16 + g + half * 8;
            const int b = m / (a.Ho * a.Wo);
            const int r = m - b * (a.Ho * a.Wo);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = n0 + wn * 32 + j * 8 + 2 * t;
                float v0 = acc[i][j][half * 2 + 0], v1 = acc[i][j][half * 2 + 1];
                if (a.bias != nullptr) { v0 += a.bias[n]; v1 += a.bias[n + 1]; }
                if (a.nchw_out) {
                    const int oy = r / a.Wo, ox = r - oy * a.Wo;
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        float v = e ? v1 : v0;
                        if (n + e < a.Cout) {
                            if (a.do_clamp) v = fminf(fmaxf(v, a.clamp_lo), a.clamp_hi);
                            v = v * a.out_scale + a.out_shift;
                            a.out[(((size_t)b * a.Cout + n + e) * a.Ho + oy) * a.Wo + ox] = v;
                        }
                    }
                } else if (n + 1 < a.Cout) {
                    const size_t o = (size_t)m * a.Cout + n;
                    if (a.resid != nullptr) {
                        float2 rr = *reinterpret_cast<const float2 *>(a.resid + o);
                        v0 += rr.x; v1 += rr.y;
                    }
                    *reinterpret_cast<float2 *>(a.out + o) = make_float2(v0, v1);
                } else if (n < a.Cout) {
                    const size_t o = (size_t)m * a.Cout + n;
                    if (a.resid != nullptr) v0 += a.resid[o];
                    a.out[o] = v0;
                }
            }
        }
}

}  // namespace wmar
