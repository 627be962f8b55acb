// Evaluation augmentations of the round-trip loop: the transforms generate.py:142-163 applies
// between codes_to_images and images_to_codes, as fused image kernels on NCHW fp32 images in [0,1] so the evaluation
// stays on the device.  Reference: wmar/augmentations/valuemetric.py:76-137 (GaussianBlur, Brightness, GaussianNoise),
// geometric.py:26-117 (Rotate, UpperLeftCropWithResizeBack, UpperLeftCropWithPadBack, HorizontalFlip); the arithmetic
// below those classes is torchvision.transforms.functional (third party, restated in oracle/augment.py):
//   gaussian_blur   = reflect pad + depthwise conv2d with the normalised 2-D Gaussian (sigma = 0.3((k-1)/2-1)+0.8), clamp
//   adjust_brightness = (factor * img).clamp(0,1)
//   rotate          = inverse affine grid (align_corners=False conventions) + grid_sample nearest, zeros outside
//   resize (upscale)= bilinear, half-pixel centres, edge clamp
// JPEG stays on the host (PIL), as in the reference (valuemetric.py:18-40).
#include "common.cuh"

using namespace wmar;

namespace {

__global__ void aug_brightness_kernel(const float *__restrict__ in, float *__restrict__ out, size_t n, float factor) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fminf(fmaxf(__fmul_rn(factor, in[i]), 0.f), 1.f);
}

__global__ void aug_noise_kernel(const float *__restrict__ in, const float *__restrict__ noise, float *__restrict__ out, size_t n,
                                 float std) {
    // noise * std, then the add, each rounded like the two torch ops (no FMA contraction)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fminf(fmaxf(__fadd_rn(in[i], __fmul_rn(noise[i], std)), 0.f), 1.f);
}

__global__ void aug_hflip_kernel(const float *__restrict__ in, float *__restrict__ out, int planes, int H, int W) {
    const size_t n = (size_t)planes * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        out[i] = in[i - x + (W - 1 - x)];
    }
}

// reflect padding of k/2, depthwise k x k correlation with the given weights, clamp(0,1)
__global__ void aug_blur_kernel(const float *__restrict__ in, float *__restrict__ out, int planes, int H, int W,
                                const float *__restrict__ k2d, int ks) {
    extern __shared__ float kw[];
    for (int i = threadIdx.x; i < ks * ks; i += blockDim.x) kw[i] = k2d[i];
    __syncthreads();
    const int r = ks / 2;
    const size_t n = (size_t)planes * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const float *pl = in + (i - (size_t)y * W - x);
        float acc = 0.f;
        for (int dy = 0; dy < ks; dy++) {
            int yy = y + dy - r;
            yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
            for (int dx = 0; dx < ks; dx++) {
                int xx = x + dx - r;
                xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
                acc = fmaf(kw[dy * ks + dx], pl[(size_t)yy * W + xx], acc);
            }
        }
        out[i] = fminf(fmaxf(acc, 0.f), 1.f);
    }
}

// torchvision _gen_affine_grid + grid_sample(nearest, zeros, align_corners=False): theta = inverse affine [2][3]
__global__ void aug_affine_nearest_kernel(const float *__restrict__ in, float *__restrict__ out, int planes, int H, int W, int OH,
                                          int OW, float t00, float t01, float t02, float t10, float t11, float t12) {
    const size_t n = (size_t)planes * OH * OW;
    // rescaled theta^T / (0.5 w, 0.5 h)
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    const float a0 = t00 / hw, a1 = t01 / hw, a2 = t02 / hw, b0 = t10 / hh, b1 = t11 / hh, b2 = t12 / hh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW), oy = (int)((i / OW) % OH);
        const size_t p = i / ((size_t)OW * OH);
        const float bx = -(float)OW * 0.5f + 0.5f + (float)ox, by = -(float)OH * 0.5f + 0.5f + (float)oy;
        const float gx = __fadd_rn(__fmaf_rn(by, a1, __fmul_rn(bx, a0)), a2);
        const float gy = __fadd_rn(__fmaf_rn(by, b1, __fmul_rn(bx, b0)), b2);
        const float ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
        const int xi = (int)nearbyintf(ix), yi = (int)nearbyintf(iy);
        float v = 0.f;
        if (xi >= 0 && xi < W && yi >= 0 && yi < H) v = in[(p * H + yi) * W + xi];
        out[i] = v;
    }
}

// bilinear up-scaling of the upper-left h2 x w2 crop back to H x W (half-pixel centres, edge clamp)
__global__ void aug_crop_resize_kernel(const float *__restrict__ in, float *__restrict__ out, int planes, int H, int W, int h2,
                                       int w2) {
    const size_t n = (size_t)planes * H * W;
    const float sy = (float)h2 / (float)H, sx = (float)w2 / (float)W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const float *pl = in + (i - (size_t)y * W - x);
        float fy = ((float)y + 0.5f) * sy - 0.5f, fx = ((float)x + 0.5f) * sx - 0.5f;
        fy = fy < 0.f ? 0.f : fy;
        fx = fx < 0.f ? 0.f : fx;
        int y0 = (int)fy, x0 = (int)fx;
        y0 = y0 > h2 - 1 ? h2 - 1 : y0;
        x0 = x0 > w2 - 1 ? w2 - 1 : x0;
        const int y1 = y0 + 1 < h2 ? y0 + 1 : h2 - 1, x1 = x0 + 1 < w2 ? x0 + 1 : w2 - 1;
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        const float top = pl[(size_t)y0 * W + x0] * (1.f - lx) + pl[(size_t)y0 * W + x1] * lx;
        const float bot = pl[(size_t)y1 * W + x0] * (1.f - lx) + pl[(size_t)y1 * W + x1] * lx;
        out[i] = top * (1.f - ly) + bot * ly;
    }
}

__global__ void aug_crop_pad_kernel(const float *__restrict__ in, float *__restrict__ out, int planes, int H, int W, int h2, int w2) {
    const size_t n = (size_t)planes * H * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        out[i] = (y < h2 && x < w2) ? in[i] : 0.f;
    }
}

inline unsigned grid_for(size_t n) {
    size_t g = (n + 255) / 256;
    return (unsigned)(g > 148 * 16 ? 148 * 16 : (g ? g : 1));
}

}  // namespace

extern "C" {

/* op codes of wmar_augment */
int wmar_augment(int op, const float *d_in, float *d_out, int64_t B, int64_t H, int64_t W, int64_t OH, int64_t OW,
                 const float *params, int n_params, const float *d_aux, void *stream) {
    WMAR_REQUIRE(d_in != nullptr && d_out != nullptr && d_in != d_out && B > 0 && H > 0 && W > 0, "bad arguments");
    const int planes = (int)B * 3;
    const size_t n = (size_t)planes * H * W;
    cudaStream_t s = as_stream(stream);
    switch (op) {
    case WMAR_AUG_BRIGHTNESS:
        WMAR_REQUIRE(n_params == 1, "brightness: params = {factor}");
        aug_brightness_kernel<<<grid_for(n), 256, 0, s>>>(d_in, d_out, n, params[0]);
        break;
    case WMAR_AUG_GAUSSIAN_NOISE:
        WMAR_REQUIRE(n_params == 1 && d_aux != nullptr, "noise: params = {std}, aux = N(0,1) noise of the image's shape");
        aug_noise_kernel<<<grid_for(n), 256, 0, s>>>(d_in, d_aux, d_out, n, params[0]);
        break;
    case WMAR_AUG_HFLIP:
        aug_hflip_kernel<<<grid_for(n), 256, 0, s>>>(d_in, d_out, planes, (int)H, (int)W);
        break;
    case WMAR_AUG_GAUSSIAN_BLUR: {
        WMAR_REQUIRE(n_params == 1 && d_aux != nullptr, "blur: params = {kernel_size}, aux = k x k weights");
        const int ks = (int)params[0];
        WMAR_REQUIRE(ks >= 1 && (ks & 1) && ks / 2 < H && ks / 2 < W && ks <= 63, "kernel size must be odd, <= 63 and smaller than the image");
        aug_blur_kernel<<<grid_for(n), 256, sizeof(float) * ks * ks, s>>>(d_in, d_out, planes, (int)H, (int)W, d_aux, ks);
        break;
    }
    case WMAR_AUG_AFFINE_NEAREST: {
        WMAR_REQUIRE(n_params == 6 && OH > 0 && OW > 0, "affine: params = inverse affine matrix [2][3], OH, OW = output size");
        const size_t no = (size_t)planes * OH * OW;
        aug_affine_nearest_kernel<<<grid_for(no), 256, 0, s>>>(d_in, d_out, planes, (int)H, (int)W, (int)OH, (int)OW, params[0],
                                                                params[1], params[2], params[3], params[4], params[5]);
        break;
    }
    case WMAR_AUG_CROP_RESIZE:
    case WMAR_AUG_CROP_PAD: {
        WMAR_REQUIRE(n_params == 2, "crop: params = {h2, w2}");
        const int h2 = (int)params[0], w2 = (int)params[1];
        WMAR_REQUIRE(h2 >= 1 && h2 <= H && w2 >= 1 && w2 <= W, "crop size out of range");
        if (op == WMAR_AUG_CROP_RESIZE) aug_crop_resize_kernel<<<grid_for(n), 256, 0, s>>>(d_in, d_out, planes, (int)H, (int)W, h2, w2);
        else aug_crop_pad_kernel<<<grid_for(n), 256, 0, s>>>(d_in, d_out, planes, (int)H, (int)W, h2, w2);
        break;
    }
    default:
        return set_error(WMAR_ERR_INVALID, "unknown augmentation op%s%s");
    }
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

}  // extern "C"
