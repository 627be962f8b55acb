// VQGAN tokenizer engine: codes_to_images / images_to_codes for the Taming (Chameleon) and MaskGIT (RAR) families.
// A plan of kernel launches is built once from the config + weight table and replayed per call.
//   Taming   deps/taming/modules/diffusionmodules/model.py:343-538, models/vqgan.py:64-73, vqvae/quantize.py:272-331,
//            models/cond_transformer.py:169-192, wmar/models/taming_wrapper.py:79-92
//   MaskGIT  deps/rar/modeling/modules/maskgit_vqgan.py:54-321, modeling/titok.py:75-85, wmar/models/rar_wrapper.py:109-128
// Weight table order (all fp32 device pointers; conv weights packed [Cout_pad64][ky][kx][Cin_pad32], biases [Cout_pad64],
// bias-free convs get a zero bias) -- mirrored by wmar_b200/models/vqgan_pack.py:
//   [0] codebook [n_e][D]
//   encoder: conv_in, then per level: per block: ResBlock (+ AttnBlock if Taming and res == attn_resolution),
//            Taming: downsample.conv when not last level;  mid (Taming: block_1, attn_1, block_2; MaskGIT: nrb blocks);
//            norm_out, conv_out;  Taming only: quant_conv
//   decoder: Taming only: post_quant_conv;  conv_in;  mid;  per level (high to low): nrb(+1 for Taming) ResBlocks
//            (+ AttnBlock), upsample conv when level != 0;  norm_out, conv_out
//   ResBlock  = norm1.g, norm1.b, conv1.w, conv1.b, norm2.g, norm2.b, conv2.w, conv2.b [, nin_shortcut.w, nin_shortcut.b]
//   AttnBlock = norm.g, norm.b, q.w, q.b, k.w, k.b, v.w, v.b, proj_out.w, proj_out.b
#include <vector>

#include <functional>

#include "conv_tc.cuh"
#include "gemm_tc.cuh"
#include "vqgan_kernels.cuh"

using namespace wmar;

namespace {

enum OpKind { OP_CONV, OP_GN, OP_ATTN, OP_POOL };

struct Op {
    OpKind kind;
    int src, dst, res;  // buffer indices (res = residual buffer or -1)
    // conv
    const float *w, *b;
    int Hs, Ws, Cin, Ho, Wo, Cout, Cout_pad, ks, stride, pad, up;
    int final_out;  // decoder conv_out: NCHW + clamp into the caller's image
    float *wlo;     // tcgen05 path (conv_tc.cuh): w - trunc_tf32(w), same layout as w; null = mma.sync kernel
    uint16_t *wb1, *wb2;   // tcgen05 bf16x3 path: bf16 planes w1 = rn(w), w2 = rn(w - w1); null = 3xTF32 kernel
    int tmp;        // up convs on the tcgen05 path: buffer that receives the materialised nearest x2 input
    // gn (conv: gamma / beta of the GroupNorm fused into the decoder-tail kernel, fused_gn = 1)
    const float *gamma, *beta;
    int swish, H, W, C;
    int direct_in;    // conv: encoder conv_in on conv_in3_kernel, straight from the caller's NCHW image
    int gn_emit;      // conv (bf16 tcgen05 path): the epilogue emits the GroupNorm statistics of its output (per-tile partials)
    int from_conv;    // gn: statistics come from the producing conv's epilogue (gn_finalize_kernel instead of gn_partial_kernel)
    int stats_only;   // gn: only the statistics pass runs, the consumer (conv_out3_kernel) normalises while staging
    int fused_gn;
    // attn: q,k,v buffers
    int bq, bk, bv;
};

struct Cursor {
    const void *const *tab;
    int n, pos;
    const float *next() { return pos < n ? reinterpret_cast<const float *>(tab[pos++]) : (pos++, nullptr); }
};

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct wmar_vqgan {
    wmar_vqgan_config cfg;
    const float *codebook;
    std::vector<Op> enc, dec;
    const float *quant_w = nullptr, *quant_b = nullptr;
    float *buf[6];
    size_t buf_floats;
    float *dots, *zz, *ee;
    double2 *gn_partial;
    float *cb_lo = nullptr, *zero_bias = nullptr;   // codebook - trunc_tf32(codebook), zeros[n_embed]: distance GEMM on the tcgen05 kernel
    double2 *gn_tile = nullptr;   // [max_batch][R * R / 128][32] per-tile GroupNorm partials written by the conv epilogues
    double flops_enc, flops_dec;
    int enc_out_buf, latent;  // buffer holding the encoder output (pre-quant z), latent side
    const float *enc_images = nullptr;   // NCHW images of the running encode call (conv_in3_kernel reads them)
    // CUDA graphs of the two directions (round 2): ~150 / ~110 host launches + tensor-map encodes per call became ONE
    // launch -- a descheduled host thread showed up as 56 ms decodes among 20.3 ms ones.  Graphs need fixed addresses: the
    // caller's codes / images are copied into (out of) engine-owned staging buffers around the graph launch.
    bool use_graph = true;
    int64_t *g_codes = nullptr;          // [max_batch][latent^2]
    float *g_images = nullptr;           // [max_batch][3][R][R]
    std::vector<cudaGraphExec_t> dec_exec, enc_exec;   // one per batch size, captured on first use
};

namespace {

struct Builder {
    wmar_vqgan *v;
    Cursor cur;
    std::vector<Op> *ops;
    int H, W, C;    // current activation shape
    int x;          // buffer index of the current activation
    double flops = 0.0;

    int free_buf(std::initializer_list<int> used) {
        for (int i = 0; i < 6; i++) {
            bool u = false;
            for (int k : used) u |= (k == i);
            if (!u) return i;
        }
        return -1;
    }
    void conv(int src, int dst, int res, int cout, int ks, int stride, int pad, int up, bool final_out = false) {
        Op o{};
        o.kind = OP_CONV; o.src = src; o.dst = dst; o.res = res;
        o.w = cur.next(); o.b = cur.next();
        o.Hs = H; o.Ws = W; o.Cin = round_up(C, 32); o.C = C;
        const int Hl = up ? H * 2 : H, Wl = up ? W * 2 : W;
        o.Ho = stride == 2 ? Hl / 2 : Hl;
        o.Wo = stride == 2 ? Wl / 2 : Wl;
        o.Cout = cout; o.Cout_pad = round_up(cout, 64);
        o.ks = ks; o.stride = stride; o.pad = pad; o.up = up; o.final_out = final_out ? 1 : 0; o.tmp = -1; o.wlo = nullptr; o.wb1 = o.wb2 = nullptr;
        ops->push_back(o);
        flops += 2.0 * o.Ho * o.Wo * (double)cout * (double)(ks * ks) * (double)C;
        H = o.Ho; W = o.Wo; C = cout;
    }
    void gn(int src, int dst, int swish) {
        Op o{};
        o.kind = OP_GN; o.src = src; o.dst = dst; o.res = -1;
        o.gamma = cur.next(); o.beta = cur.next();
        o.swish = swish; o.H = H; o.W = W; o.C = C;
        ops->push_back(o);
    }
    // Taming: x + h or nin(x) + h ; MaskGIT: h + x or h + nin(h)
    void resblock(int cout, bool maskgit) {
        const int cin = C, h0 = H, w0 = W;
        int t1 = free_buf({x}), t2 = free_buf({x, t1});
        gn(x, t1, 1);
        conv(t1, t2, -1, cout, 3, 1, 1, 0);
        gn(t2, t1, 1);
        if (cin == cout) {
            conv(t1, t2, x, cout, 3, 1, 1, 0);  // conv2 + residual x
            x = t2;
        } else if (!maskgit) {
            int t3 = free_buf({x, t1, t2});
            conv(t1, t3, -1, cout, 3, 1, 1, 0);            // h = conv2
            H = h0; W = w0; C = cin;
            conv(x, t2, t3, cout, 1, 1, 0, 0);             // nin_shortcut(x) + h
            x = t2;
        } else {
            int t3 = free_buf({x, t1, t2});
            conv(t1, t3, -1, cout, 3, 1, 1, 0);            // h = conv2
            conv(t3, t2, t3, cout, 1, 1, 0, 0);            // nin_shortcut(h) + h   (maskgit_vqgan.py:87-90)
            x = t2;
        }
    }
    void attnblock() {
        int t1 = free_buf({x}), bq = free_buf({x, t1}), bk = free_buf({x, t1, bq}), bv = free_buf({x, t1, bq, bk});
        gn(x, t1, 0);
        const int c = C;
        conv(t1, bq, -1, c, 1, 1, 0, 0);
        conv(t1, bk, -1, c, 1, 1, 0, 0);
        conv(t1, bv, -1, c, 1, 1, 0, 0);
        Op o{};
        o.kind = OP_ATTN; o.bq = bq; o.bk = bk; o.bv = bv; o.dst = t1; o.H = H; o.W = W; o.C = C;
        ops->push_back(o);
        flops += 4.0 * (double)(H * W) * (double)(H * W) * (double)C;
        conv(t1, bq, x, c, 1, 1, 0, 0);  // proj_out + x
        x = bq;
    }
    void pool() {
        int t1 = free_buf({x});
        Op o{};
        o.kind = OP_POOL; o.src = x; o.dst = t1; o.H = H / 2; o.W = W / 2; o.C = C;
        ops->push_back(o);
        H /= 2; W /= 2; x = t1;
    }
};

int build_plans(wmar_vqgan *v, const void *const *tab, int n) {
    const wmar_vqgan_config &c = v->cfg;
    const bool mg = c.family == 1;
    Builder b{};
    b.v = v;
    b.cur = Cursor{tab, n, 0};
    v->codebook = b.cur.next();
    // ---------------- encoder: input = NHWC image padded to 32 channels in buffer 0
    b.ops = &v->enc;
    b.H = b.W = c.resolution; b.C = 32; b.x = 0; b.flops = 0.0;
    {
        int t = b.free_buf({b.x});
        b.conv(b.x, t, -1, c.ch, 3, 1, 1, 0);
        b.flops -= 2.0 * c.resolution * c.resolution * (double)c.ch * 9.0 * 29.0;  // padded input channels are zeros
        const char *e = getenv("WMAR_CONVIN");
        if (!(e && e[0] == 'v') && c.ch % 128 == 0) b.ops->back().direct_in = 1;
        b.x = t;
    }
    int res = c.resolution;
    for (int l = 0; l < c.n_levels; l++) {
        const int cout = c.ch * c.ch_mult[l];
        for (int k = 0; k < c.num_res_blocks; k++) {
            b.resblock(cout, mg);
            if (!mg && c.attn_resolution > 0 && res == c.attn_resolution) b.attnblock();
        }
        if (l != c.n_levels - 1) {
            if (mg) b.pool();
            else {
                int t = b.free_buf({b.x});
                b.conv(b.x, t, -1, b.C, 3, 2, 0, 0);  // pad (0,1,0,1) + stride 2 (model.py:69-72)
                b.x = t;
            }
            res /= 2;
        }
    }
    if (mg) {
        for (int k = 0; k < c.num_res_blocks; k++) b.resblock(b.C, true);
    } else {
        b.resblock(b.C, false);
        b.attnblock();
        b.resblock(b.C, false);
    }
    {
        int t1 = b.free_buf({b.x}), t2 = b.free_buf({b.x, t1});
        b.gn(b.x, t1, 1);
        b.conv(t1, t2, -1, c.z_channels, mg ? 1 : 3, 1, mg ? 0 : 1, 0);
        b.x = t2;
        if (!mg) {
            int t3 = b.free_buf({b.x});
            b.conv(b.x, t3, -1, c.embed_dim, 1, 1, 0, 0);  // quant_conv
            b.x = t3;
        }
    }
    v->enc_out_buf = b.x;
    v->latent = b.H;
    v->flops_enc = b.flops + 2.0 * (double)(b.H * b.W) * (double)c.n_embed * (double)c.embed_dim;
    // ---------------- decoder: input = gathered codebook vectors (NHWC) in buffer 0
    b.ops = &v->dec;
    b.H = b.W = v->latent; b.C = c.embed_dim; b.x = 0; b.flops = 0.0;
    if (!mg) {
        int t = b.free_buf({b.x});
        b.conv(b.x, t, -1, c.z_channels, 1, 1, 0, 0);  // post_quant_conv
        b.x = t;
    }
    {
        int t = b.free_buf({b.x});
        b.conv(b.x, t, -1, c.ch * c.ch_mult[c.n_levels - 1], 3, 1, 1, 0);
        b.x = t;
    }
    if (mg) {
        for (int k = 0; k < c.num_res_blocks; k++) b.resblock(b.C, true);
    } else {
        b.resblock(b.C, false);
        b.attnblock();
        b.resblock(b.C, false);
    }
    res = v->latent;
    for (int l = c.n_levels - 1; l >= 0; l--) {
        const int cout = c.ch * c.ch_mult[l];
        const int nblk = mg ? c.num_res_blocks : c.num_res_blocks + 1;
        for (int k = 0; k < nblk; k++) {
            b.resblock(cout, mg);
            if (!mg && c.attn_resolution > 0 && res == c.attn_resolution) b.attnblock();
        }
        if (l != 0) {
            int t = b.free_buf({b.x});
            b.conv(b.x, t, -1, b.C, 3, 1, 1, 1);  // nearest x2 folded into the conv's input indexing (mma.sync kernel)
            b.ops->back().tmp = b.free_buf({b.x, t});   // ... or materialised there for the tcgen05 kernel
            b.x = t;
            res *= 2;
        }
    }
    {
        int t1 = b.free_buf({b.x});
        b.gn(b.x, t1, 1);
        b.conv(t1, -1, -1, 3, 3, 1, 1, 0, true);
        // decoder tail as one fp32 kernel (conv_out3_kernel): GroupNorm + swish applied while the conv stages its input
        const char *e = getenv("WMAR_CONVOUT");
        if (!(e && e[0] == 'v') && b.ops->back().C % 32 == 0) {
            Op &cv = b.ops->back(), &g = (*b.ops)[b.ops->size() - 2];
            g.stats_only = 1;
            cv.fused_gn = 1; cv.src = g.src; cv.gamma = g.gamma; cv.beta = g.beta;
        }
    }
    v->flops_dec = b.flops;
    WMAR_REQUIRE(b.cur.pos == n, "weight table length does not match the architecture");
    return WMAR_OK;
}

// 3x3 / stride 1 / pad 1 convs whose shapes tile into 128 pixels x 128 channels go to the tcgen05 kernel
bool conv_tc_eligible(const Op &o) {
    if (o.ks != 3 || o.stride != 1 || o.pad != 1 || o.final_out) return false;
    if (o.up && o.tmp < 0) return false;
    if (o.C != o.Cin || o.Cout % 128 != 0 || o.Cout != o.Cout_pad) return false;   // no channel padding on either side
    const int Hi = o.up ? 2 * o.Hs : o.Hs, Wi = o.up ? 2 * o.Ws : o.Ws;            // conv input = (upsampled) source
    if (o.Ho != Hi || o.Wo != Wi || Wi < 8) return false;
    if (Wi >= 128) return Wi % 128 == 0;
    return 128 % Wi == 0 && Hi % (128 / Wi) == 0 && 128 / Wi <= 256;
}

// stride-2 downsample convs (3x3, pad (0,1,0,1), model.py:57-76) on the bf16x3 tcgen05 kernel: TMA element stride 2
bool conv_tc_s2_eligible(const Op &o) {
    if (o.ks != 3 || o.stride != 2 || o.pad != 0 || o.final_out || o.up) return false;
    if (o.C != o.Cin || o.Cin % 64 != 0 || o.Cout % 128 != 0 || o.Cout != o.Cout_pad) return false;
    if (o.Hs != 2 * o.Ho || o.Ws != 2 * o.Wo || o.Wo < 8) return false;
    if (o.Wo >= 128) return o.Wo % 128 == 0;
    return 128 % o.Wo == 0 && o.Ho % (128 / o.Wo) == 0;
}

// 1x1 convs (attention q / k / v / proj_out, nin_shortcut, quant / post_quant convs) on the bf16x3 tcgen05 kernel: one tap
bool conv_tc_1x1_eligible(const Op &o) {
    if (o.ks != 1 || o.stride != 1 || o.pad != 0 || o.final_out || o.up) return false;
    if (o.C != o.Cin || o.Cin % 64 != 0 || o.Cout % 128 != 0 || o.Cout != o.Cout_pad) return false;
    if (o.Ho != o.Hs || o.Wo != o.Ws || o.Wo < 8) return false;
    if (o.Wo >= 128) return o.Wo % 128 == 0;
    return 128 % o.Wo == 0 && o.Ho % (128 / o.Wo) == 0;
}

// nearest x2 (taming model.py:39-54 Upsample: F.interpolate(scale_factor=2, mode="nearest")), NHWC, 16-byte vectors
__global__ void upsample2x_nhwc_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, int B, int H, int W, int C4) {
    const size_t n = (size_t)B * 2 * H * 2 * W * C4;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        size_t p = i / C4;
        const int x = (int)(p % (2 * W));
        p /= 2 * W;
        const int y = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        out[i] = in[(((size_t)b * H + (y >> 1)) * W + (x >> 1)) * C4 + c];
    }
}

int run_conv_tc(const wmar_vqgan *v, const Op &o, int B, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SM_ALLOC));
        configured = true;
    }
    const float *src = v->buf[o.src];
    if (o.up) {
        const size_t n4 = (size_t)B * o.Ho * o.Wo * (o.Cin / 4);
        size_t gx = (n4 + 255) / 256;
        upsample2x_nhwc_kernel<<<(unsigned)(gx > 148 * 32 ? 148 * 32 : gx), 256, 0, s>>>(
            reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(v->buf[o.tmp]), B, o.Hs, o.Ws, o.Cin / 4);
        WMAR_LAUNCH_CHECK();
        src = v->buf[o.tmp];
    }
    ConvTcArgs a{};
    a.bias = o.b; a.resid = o.res >= 0 ? v->buf[o.res] : nullptr; a.out = v->buf[o.dst];
    a.H = o.Ho; a.W = o.Wo; a.Cin = o.Cin; a.Cout = o.Cout;
    a.bw = o.Wo >= 128 ? 128 : o.Wo; a.bh = 128 / a.bw;
    a.tiles_x = o.Wo / a.bw; a.tiles_y = o.Ho / a.bh;
    CUtensorMap mA, mWh, mWl;
    int rc;
    a.taps = o.ks == 1 ? 1 : 9;
    a.stride = o.stride; a.pad = (o.ks == 3 && o.stride == 1) ? 1 : 0;
    const int Kw = o.ks * o.ks * o.Cin;      // row length of the [Cout][ky][kx][Cin] weights
    a.gn_out = (o.gn_emit && o.wb1 != nullptr) ? v->gn_tile : nullptr;
    a.gn_cg = o.Cout / 32;
    if ((rc = tc_nhwc_map(src, B, o.stride * o.Ho, o.stride * o.Wo, o.Cin, a.bw, a.bh, &mA, o.stride))) return rc;
    if (o.wb1 != nullptr) {
        // bf16x3: persistent kernel, one CTA per SM walking the (pixel tile, 128-channel block) list
        static bool configured_b = false;
        if (!configured_b) {
            WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SM_ALLOC));
            configured_b = true;
        }
        if ((rc = tc_weight_map_bf16(o.wb1, o.Cout_pad, Kw, &mWh))) return rc;
        if ((rc = tc_weight_map_bf16(o.wb2, o.Cout_pad, Kw, &mWl))) return rc;
        ConvTcTiles tl{};
        tl.nblk = o.Cout / 128;
        tl.n_tiles = B * a.tiles_x * a.tiles_y * tl.nblk;
        { const char *e = getenv("WMAR_CB_DBG"); tl.dbg = e ? atoi(e) : 0; }
        static int sms = 0;
        if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        const int grid = tl.n_tiles < sms ? tl.n_tiles : sms;
        conv3x3_tc_bf16_kernel<<<grid, CT_THREADS, CB_SM_ALLOC, s>>>(mA, mWh, mWl, a, tl);
        WMAR_LAUNCH_CHECK();
        return WMAR_OK;
    }
    if ((rc = tc_weight_map(o.w, o.Cout_pad, Kw, &mWh))) return rc;
    if ((rc = tc_weight_map(o.wlo, o.Cout_pad, Kw, &mWl))) return rc;
    dim3 grid((unsigned)(B * a.tiles_x * a.tiles_y), (unsigned)(o.Cout / 128));
    conv3x3_tc_kernel<<<grid, CT_THREADS, CT_SM_ALLOC, s>>>(mA, mWh, mWl, a);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

int gn_chunks(int HW, int C);

int run_conv(const wmar_vqgan *v, const Op &o, int B, float *final_out, cudaStream_t s) {
    if (o.direct_in) {
        ConvInArgs a{};
        a.img = v->enc_images; a.w = o.w; a.bias = o.b; a.out = v->buf[o.dst];
        a.H = o.Ho; a.W = o.Wo; a.Cout = o.Cout; a.Cin_pad = o.Cin;
        a.scale = v->cfg.family == 1 ? 0.5f : 1.f; a.shift = v->cfg.family == 1 ? 0.5f : 0.f;   // rar_wrapper.py:124
        dim3 grid((unsigned)((o.Wo + CI_TW - 1) / CI_TW), (unsigned)((o.Ho + CI_TH - 1) / CI_TH), (unsigned)B);
        const size_t smem = sizeof(float) * (27 * (size_t)o.Cout + 3 * (CI_TH + 2) * (CI_TW + 2));
        conv_in3_kernel<<<grid, 256, smem, s>>>(a);
        WMAR_LAUNCH_CHECK();
        return WMAR_OK;
    }
    if (o.wlo != nullptr) return run_conv_tc(v, o, B, s);
    if (o.fused_gn) {
        ConvOutArgs a{};
        a.x = v->buf[o.src]; a.w = o.w; a.bias = o.b; a.out = final_out;
        a.H = o.Ho; a.W = o.Wo; a.C = o.C;
        a.gn_partial = v->gn_partial; a.nchunk = o.from_conv ? 1 : gn_chunks(o.Ho * o.Wo, o.C);   // (from_conv copied from its GroupNorm)
        a.gamma = o.gamma; a.beta = o.beta; a.eps = 1e-6f;
        a.out_scale = 1.f; a.out_shift = 0.f; a.clamp_lo = -1.f; a.clamp_hi = 1.f;
        if (v->cfg.family == 1) { a.clamp_lo = 0.f; a.clamp_hi = 1.f; a.out_scale = 2.f; a.out_shift = -1.f; }
        dim3 grid((unsigned)((o.Wo + CO_T - 1) / CO_T), (unsigned)((o.Ho + CO_T - 1) / CO_T), (unsigned)B);
        conv_out3_kernel<<<grid, CO_T * CO_T, CO_SMEM_BYTES, s>>>(a);
        WMAR_LAUNCH_CHECK();
        return WMAR_OK;
    }
    ConvArgs a{};
    a.in = v->buf[o.src]; a.w = o.w; a.bias = o.b;
    a.resid = o.res >= 0 ? v->buf[o.res] : nullptr;
    a.out = o.final_out ? final_out : v->buf[o.dst];
    a.B = B; a.Hs = o.Hs; a.Ws = o.Ws; a.Cin = o.Cin; a.Ho = o.Ho; a.Wo = o.Wo; a.Cout = o.Cout; a.Cout_pad = o.Cout_pad;
    a.ks = o.ks; a.stride = o.stride; a.pad = o.pad; a.up = o.up;
    a.out_scale = 1.f; a.out_shift = 0.f;
    if (o.final_out) {
        a.nchw_out = 1; a.do_clamp = 1;
        if (v->cfg.family == 1) { a.clamp_lo = 0.f; a.clamp_hi = 1.f; a.out_scale = 2.f; a.out_shift = -1.f; }
        else { a.clamp_lo = -1.f; a.clamp_hi = 1.f; }
    }
    const long long M = (long long)B * o.Ho * o.Wo;
    WMAR_REQUIRE(M % CV_BM == 0, "B*Ho*Wo must be a multiple of 128");
    dim3 grid((unsigned)(M / CV_BM), (unsigned)(o.Cout_pad / CV_BN));
    const size_t smem = sizeof(float) * 2 * (CV_BM + CV_BN) * CV_LD;
    if (v->cfg.precision != 1) conv_igemm_kernel<1><<<grid, CV_THREADS, smem, s>>>(a);
    else conv_igemm_kernel<0><<<grid, CV_THREADS, smem, s>>>(a);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

int gn_chunks(int HW, int C) {
    int n = 1;
    while (n < 64 && (long long)HW * C / (n * 2) >= 16384 && (HW / (n * 2)) >= 1) n *= 2;
    return n;
}

int run_ops(const wmar_vqgan *v, const std::vector<Op> &ops, int B, float *final_out, cudaStream_t s) {
    int rc;
    for (const Op &o : ops) {
        switch (o.kind) {
            case OP_CONV:
                if ((rc = run_conv(v, o, B, final_out, s))) return rc;
                break;
            case OP_GN: {
                const int HW = o.H * o.W, nchunk = o.from_conv ? 1 : gn_chunks(HW, o.C);
                if (o.from_conv) gn_finalize_kernel<<<B, 256, 0, s>>>(v->gn_tile, HW / 128, v->gn_partial);
                else gn_partial_kernel<<<dim3(nchunk, B), GN_THREADS, 0, s>>>(v->buf[o.src], HW, o.C, nchunk, v->gn_partial);
                WMAR_LAUNCH_CHECK();
                if (o.stats_only) break;
                long long total4 = (long long)HW * o.C / 4;
                int gx = (int)((total4 + 255) / 256);
                if (gx > 1024) gx = 1024;
                gn_apply_kernel<<<dim3(gx, B), 256, 0, s>>>(v->buf[o.src], v->buf[o.dst], HW, o.C, nchunk, v->gn_partial,
                                                         o.gamma, o.beta, 1e-6f, o.swish);
                WMAR_LAUNCH_CHECK();
                break;
            }
            case OP_ATTN: {
                const int N = o.H * o.W;
                const size_t smem = sizeof(float) * (size_t)AB_Q * (o.C + N);
                attn_block_kernel<<<dim3(N / AB_Q, B), 256, smem, s>>>(v->buf[o.bq], v->buf[o.bk], v->buf[o.bv],
                                                                      v->buf[o.dst], N, o.C);
                WMAR_LAUNCH_CHECK();
                break;
            }
            case OP_POOL: {
                size_t total = (size_t)B * o.H * o.W * (o.C / 4);
                avgpool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(v->buf[o.src], v->buf[o.dst], B, o.H, o.W, o.C);
                WMAR_LAUNCH_CHECK();
                break;
            }
        }
    }
    return WMAR_OK;
}

}  // namespace

extern "C" {

int wmar_vqgan_create(const wmar_vqgan_config *cfg, const void *const *d_weights, int n_weights, wmar_vqgan **out) {
    WMAR_REQUIRE(cfg && d_weights && out, "NULL argument");
    WMAR_REQUIRE(cfg->family == 0 || cfg->family == 1, "family must be 0 (Taming) or 1 (MaskGIT)");
    WMAR_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= 8, "n_levels out of range");
    WMAR_REQUIRE(cfg->ch % 128 == 0, "base channel count must be a multiple of 128 (GroupNorm kernel: >= 4 channels/group)");
    WMAR_REQUIRE(cfg->z_channels % 32 == 0 && cfg->embed_dim % 32 == 0 && cfg->n_embed % 64 == 0, "bad latent dims");
    WMAR_REQUIRE(cfg->max_batch >= 1, "max_batch must be >= 1");
    WMAR_REQUIRE(cfg->precision >= 0 && cfg->precision <= 3, "precision must be 0 (3xTF32), 1 (TF32), 2 (bf16x3) or 3 (bf16x3 decoder, 3xTF32 encoder)");
    for (int i = 0; i < n_weights; i++) WMAR_REQUIRE(d_weights[i] != nullptr, "NULL weight pointer");
    wmar_vqgan *v = new (std::nothrow) wmar_vqgan();
    if (!v) return set_error(WMAR_ERR_NOMEM, "out of host memory%s%s");
    v->cfg = *cfg;
    int rc = build_plans(v, d_weights, n_weights);
    if (rc) { delete v; return rc; }
    const int R = cfg->resolution;
    size_t per_img = (size_t)R * R * (size_t)(cfg->ch * cfg->ch_mult[0] > 32 ? cfg->ch * cfg->ch_mult[0] : 32);
    // the widest activation is at full resolution with ch*ch_mult[0] channels; upsampled tensors at lower levels are
    // (R/2)^2 * ch*ch_mult[1] <= that as long as ch_mult[1] <= 4*ch_mult[0]
    for (int l = 0; l < cfg->n_levels; l++) {
        size_t side = (size_t)R >> l;
        size_t n = side * side * (size_t)cfg->ch * cfg->ch_mult[l];
        if (n > per_img) per_img = n;
        if (l + 1 < cfg->n_levels) {  // decoder: level l+1 channels upsampled to level l resolution
            size_t n2 = side * side * (size_t)cfg->ch * cfg->ch_mult[l + 1];
            if (n2 > per_img) per_img = n2;
        }
    }
    v->buf_floats = per_img * (size_t)cfg->max_batch;
    for (int i = 0; i < 6; i++) WMAR_CUDA_CHECK(cudaMalloc(&v->buf[i], sizeof(float) * v->buf_floats));
    const size_t tokens = (size_t)cfg->max_batch * v->latent * v->latent;
    WMAR_CUDA_CHECK(cudaMalloc(&v->dots, sizeof(float) * tokens * cfg->n_embed));
    WMAR_CUDA_CHECK(cudaMalloc(&v->zz, sizeof(float) * tokens));
    WMAR_CUDA_CHECK(cudaMalloc(&v->ee, sizeof(float) * cfg->n_embed));
    WMAR_CUDA_CHECK(cudaMalloc(&v->gn_partial, sizeof(double2) * 64 * 32 * (size_t)cfg->max_batch));
    row_sumsq_kernel<<<(cfg->n_embed + 7) / 8, 256>>>(v->codebook, v->ee, cfg->n_embed, cfg->embed_dim);
    WMAR_LAUNCH_CHECK();
    // tcgen05 path for the 3x3 convs that tile (WMAR_CONV=v0 keeps everything on the mma.sync kernel): precompute w_lo
    {
        const char *e = getenv("WMAR_CONV");
        const bool want_tc = !(e && e[0] == 'v') && tc_available();
        if (want_tc) {
            const size_t ncb = (size_t)cfg->n_embed * cfg->embed_dim;
            WMAR_CUDA_CHECK(cudaMalloc(&v->cb_lo, sizeof(float) * ncb));
            conv_tc_wlo_kernel<<<1024, 256>>>(v->codebook, v->cb_lo, ncb);
            WMAR_LAUNCH_CHECK();
            WMAR_CUDA_CHECK(cudaMalloc(&v->zero_bias, sizeof(float) * cfg->n_embed));
            WMAR_CUDA_CHECK(cudaMemset(v->zero_bias, 0, sizeof(float) * cfg->n_embed));
        }
        for (auto *ops : {&v->enc, &v->dec})
            for (Op &o : *ops) {
                o.wlo = nullptr;
                const bool bf = cfg->precision == 2 || (cfg->precision == 3 && ops == &v->dec);
                const bool s2 = bf && o.kind == OP_CONV && (conv_tc_s2_eligible(o) || conv_tc_1x1_eligible(o));   // bf16 kernel only
                if (!want_tc || o.kind != OP_CONV || o.direct_in || !(conv_tc_eligible(o) || s2)) continue;
                const size_t nw = (size_t)o.Cout_pad * o.ks * o.ks * o.Cin;
                WMAR_CUDA_CHECK(cudaMalloc(&o.wlo, sizeof(float) * nw));   // (also the "tcgen05 path" marker of run_conv)
                conv_tc_wlo_kernel<<<1024, 256>>>(o.w, o.wlo, nw);
                WMAR_LAUNCH_CHECK();
                if (bf && o.Cin % 64 == 0) {
                    WMAR_CUDA_CHECK(cudaMalloc(&o.wb1, sizeof(uint16_t) * nw));
                    WMAR_CUDA_CHECK(cudaMalloc(&o.wb2, sizeof(uint16_t) * nw));
                    conv_tc_wsplit_bf16_kernel<<<1024, 256>>>(o.w, o.wb1, o.wb2, nw);
                    WMAR_LAUNCH_CHECK();
                }
            }
    }
    // GroupNorm statistics from the producing conv's epilogue (WMAR_GN_FUSE=0: always the separate statistics pass): a
    // GroupNorm whose input was written by a conv on the persistent bf16 tcgen05 kernel (whole output = that conv's tiles)
    {
        const char *e = getenv("WMAR_GN_FUSE");
        const bool fuse = !(e && e[0] == '0');
        WMAR_CUDA_CHECK(cudaMalloc(&v->gn_tile, sizeof(double2) * 32 * (size_t)cfg->max_batch * ((size_t)R * R / 128)));
        for (auto *ops : {&v->enc, &v->dec})
            for (size_t i = 0; fuse && i < ops->size(); i++) {
                Op &gn = (*ops)[i];
                if (gn.kind != OP_GN) continue;
                int p = -1;                      // the most recent writer of the GroupNorm's input buffer
                for (int k = (int)i - 1; k >= 0 && p < 0; k--) {
                    const Op &w = (*ops)[k];
                    const bool writes = (w.kind == OP_CONV && !w.final_out && w.dst == gn.src) || (w.kind != OP_CONV && w.dst == gn.src);
                    if (writes) p = k;
                }
                if (p < 0) continue;
                Op &cv = (*ops)[p];
                if (cv.kind != OP_CONV || cv.wb1 == nullptr || cv.Cout != gn.C || cv.Ho != gn.H || cv.Wo != gn.W) continue;
                if ((cv.Ho * cv.Wo) % 128 != 0) continue;
                bool clash = false;              // no other emitting conv between producer and consumer (one statistics buffer)
                for (size_t k = (size_t)p + 1; k < i; k++) clash |= (*ops)[k].kind == OP_CONV && (*ops)[k].gn_emit;
                if (clash) continue;
                cv.gn_emit = 1;
                gn.from_conv = 1;
                if (gn.stats_only && i + 1 < ops->size()) (*ops)[i + 1].from_conv = 1;   // the fused decoder-tail conv reads nchunk = 1
            }
    }
    const size_t smem = sizeof(float) * 2 * (CV_BM + CV_BN) * CV_LD;
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(attn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv_out3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CO_SMEM_BYTES));
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SM_ALLOC));
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_tc_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SM_ALLOC));
    {
        const char *e = getenv("WMAR_VQ_GRAPH");
        v->use_graph = !(e && e[0] == '0');
        v->dec_exec.assign((size_t)cfg->max_batch + 1, nullptr);
        v->enc_exec.assign((size_t)cfg->max_batch + 1, nullptr);
        WMAR_CUDA_CHECK(cudaMalloc(&v->g_codes, sizeof(int64_t) * tokens));
        WMAR_CUDA_CHECK(cudaMalloc(&v->g_images, sizeof(float) * (size_t)cfg->max_batch * 3 * R * R));
    }
    WMAR_CUDA_CHECK(cudaDeviceSynchronize());
    *out = v;
    return WMAR_OK;
}

void wmar_vqgan_destroy(wmar_vqgan *v) {
    if (!v) return;
    cudaDeviceSynchronize();
    for (auto *ops : {&v->enc, &v->dec})
        for (Op &o : *ops) { cudaFree(o.wlo); cudaFree(o.wb1); cudaFree(o.wb2); }
    for (int i = 0; i < 6; i++) cudaFree(v->buf[i]);
    cudaFree(v->dots); cudaFree(v->zz); cudaFree(v->ee); cudaFree(v->gn_partial);
    for (cudaGraphExec_t e : v->dec_exec) if (e) cudaGraphExecDestroy(e);
    for (cudaGraphExec_t e : v->enc_exec) if (e) cudaGraphExecDestroy(e);
    cudaFree(v->g_codes); cudaFree(v->g_images); cudaFree(v->gn_tile); cudaFree(v->cb_lo); cudaFree(v->zero_bias);
    delete v;
}

static int vqgan_decode_eager(wmar_vqgan *v, const int64_t *d_codes, int64_t B, float *d_images, void *stream) {
    WMAR_REQUIRE(v && d_codes && d_images, "NULL argument");
    WMAR_REQUIRE(B >= 1 && B <= v->cfg.max_batch, "batch exceeds max_batch");
    cudaStream_t s = as_stream(stream);
    const size_t n_tok = (size_t)B * v->latent * v->latent;
    const int D = v->cfg.embed_dim;
    size_t n4 = n_tok * (size_t)(D / 4);
    codebook_gather_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(d_codes, v->codebook, v->buf[0], n_tok, D, v->cfg.n_embed);
    WMAR_LAUNCH_CHECK();
    return run_ops(v, v->dec, (int)B, d_images, s);
}

static int vqgan_encode_eager(wmar_vqgan *v, const float *d_images, int64_t B, int64_t *d_codes, void *stream) {
    WMAR_REQUIRE(v && d_codes && d_images, "NULL argument");
    WMAR_REQUIRE(B >= 1 && B <= v->cfg.max_batch, "batch exceeds max_batch");
    cudaStream_t s = as_stream(stream);
    const int R = v->cfg.resolution;
    size_t total = (size_t)B * R * R * 32;
    // RAR feeds (x+1)/2 to its encoder (rar_wrapper.py:124)
    const float scale = v->cfg.family == 1 ? 0.5f : 1.f, shift = v->cfg.family == 1 ? 0.5f : 0.f;
    v->enc_images = d_images;
    if (!v->enc.front().direct_in) {
        nchw_to_nhwc_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_images, v->buf[0], (int)B, R * R, 32, scale, shift);
        WMAR_LAUNCH_CHECK();
    }
    int rc = run_ops(v, v->enc, (int)B, nullptr, s);
    if (rc) return rc;
    // nearest codebook entry: dots = z . e^T (always 3xTF32), d = (|z|^2 + |e|^2) - 2 dots, first arg-min
    const int tokens = (int)B * v->latent * v->latent, D = v->cfg.embed_dim, NE = v->cfg.n_embed;
    const float *z = v->buf[v->enc_out_buf];
    row_sumsq_kernel<<<(tokens + 7) / 8, 256, 0, s>>>(z, v->zz, tokens, D);
    WMAR_LAUNCH_CHECK();
    WMAR_REQUIRE(tokens % CV_BM == 0, "token count must be a multiple of 128");
    if (v->cb_lo != nullptr && D % 32 == 0 && NE % 128 == 0) {
        // tcgen05 3xTF32 kernel as a plain GEMM: the tokens are a 1 x tokens "image", one tap (always fp32-faithful 3xTF32:
        // the arg-min decides ties at fp32 rounding distance, whatever the precision of the conv stacks)
        ConvTcArgs t{};
        t.bias = v->zero_bias; t.resid = nullptr; t.out = v->dots;
        t.H = 1; t.W = tokens; t.Cin = D; t.Cout = NE; t.bw = 128; t.bh = 1; t.tiles_x = tokens / 128; t.tiles_y = 1; t.taps = 1;
        CUtensorMap mA, mWh, mWl;
        int rc2;
        if ((rc2 = tc_nhwc_map(z, 1, 1, tokens, D, 128, 1, &mA))) return rc2;
        if ((rc2 = tc_weight_map(v->codebook, NE, D, &mWh))) return rc2;
        if ((rc2 = tc_weight_map(v->cb_lo, NE, D, &mWl))) return rc2;
        conv3x3_tc_kernel<<<dim3((unsigned)(tokens / 128), (unsigned)(NE / 128)), CT_THREADS, CT_SM_ALLOC, s>>>(mA, mWh, mWl, t);
        WMAR_LAUNCH_CHECK();
    } else {
        ConvArgs a{};
        a.in = z; a.w = v->codebook; a.bias = nullptr; a.resid = nullptr; a.out = v->dots;
        a.B = 1; a.Hs = 1; a.Ws = tokens; a.Cin = D; a.Ho = 1; a.Wo = tokens; a.Cout = NE; a.Cout_pad = NE;
        a.ks = 1; a.stride = 1; a.pad = 0; a.up = 0; a.out_scale = 1.f;
        const size_t smem = sizeof(float) * 2 * (CV_BM + CV_BN) * CV_LD;
        conv_igemm_kernel<1><<<dim3(tokens / CV_BM, NE / CV_BN), CV_THREADS, smem, s>>>(a);
        WMAR_LAUNCH_CHECK();
    }
    vq_argmin_kernel<<<tokens, 256, 0, s>>>(v->dots, v->zz, v->ee, d_codes, NE);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

// one direction as a cached CUDA graph over the engine's staging buffers (re-captured when the batch size changes)
static int vqgan_capture(cudaGraphExec_t *exec, const std::function<int(cudaStream_t)> &body) {
    if (*exec) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
    cudaStream_t cs;
    WMAR_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    WMAR_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const int rc = body(cs);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cs, &graph);
    cudaStreamDestroy(cs);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return set_error(WMAR_ERR_CUDA, "cudaStreamEndCapture (vqgan): %s%s", cudaGetErrorString(e));
    const cudaError_t ei = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) return set_error(WMAR_ERR_CUDA, "cudaGraphInstantiate (vqgan): %s%s", cudaGetErrorString(ei));
    return WMAR_OK;
}

int wmar_vqgan_decode(wmar_vqgan *v, const int64_t *d_codes, int64_t B, float *d_images, void *stream) {
    WMAR_REQUIRE(v && d_codes && d_images, "NULL argument");
    WMAR_REQUIRE(B >= 1 && B <= v->cfg.max_batch, "batch exceeds max_batch");
    if (!v->use_graph) return vqgan_decode_eager(v, d_codes, B, d_images, stream);
    cudaStream_t s = as_stream(stream);
    const size_t n_tok = (size_t)B * v->latent * v->latent, n_img = (size_t)B * 3 * v->cfg.resolution * v->cfg.resolution;
    if (v->dec_exec[B] == nullptr) {
        int rc = vqgan_capture(&v->dec_exec[B], [&](cudaStream_t cs) { return vqgan_decode_eager(v, v->g_codes, B, v->g_images, cs); });
        if (rc) return rc;
    }
    WMAR_CUDA_CHECK(cudaMemcpyAsync(v->g_codes, d_codes, sizeof(int64_t) * n_tok, cudaMemcpyDeviceToDevice, s));
    WMAR_CUDA_CHECK(cudaGraphLaunch(v->dec_exec[B], s));
    WMAR_CUDA_CHECK(cudaMemcpyAsync(d_images, v->g_images, sizeof(float) * n_img, cudaMemcpyDeviceToDevice, s));
    return WMAR_OK;
}

int wmar_vqgan_encode(wmar_vqgan *v, const float *d_images, int64_t B, int64_t *d_codes, void *stream) {
    WMAR_REQUIRE(v && d_codes && d_images, "NULL argument");
    WMAR_REQUIRE(B >= 1 && B <= v->cfg.max_batch, "batch exceeds max_batch");
    if (!v->use_graph) return vqgan_encode_eager(v, d_images, B, d_codes, stream);
    cudaStream_t s = as_stream(stream);
    const size_t n_tok = (size_t)B * v->latent * v->latent, n_img = (size_t)B * 3 * v->cfg.resolution * v->cfg.resolution;
    if (v->enc_exec[B] == nullptr) {
        int rc = vqgan_capture(&v->enc_exec[B], [&](cudaStream_t cs) { return vqgan_encode_eager(v, v->g_images, B, v->g_codes, cs); });
        if (rc) return rc;
    }
    WMAR_CUDA_CHECK(cudaMemcpyAsync(v->g_images, d_images, sizeof(float) * n_img, cudaMemcpyDeviceToDevice, s));
    WMAR_CUDA_CHECK(cudaGraphLaunch(v->enc_exec[B], s));
    WMAR_CUDA_CHECK(cudaMemcpyAsync(d_codes, v->g_codes, sizeof(int64_t) * n_tok, cudaMemcpyDeviceToDevice, s));
    return WMAR_OK;
}

double wmar_vqgan_flops(const wmar_vqgan *v, int decode) { return v ? (decode ? v->flops_dec : v->flops_enc) : 0.0; }

}  // extern "C"
