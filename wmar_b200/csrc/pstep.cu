// Host side of the persistent decode-step kernel: plan, pack, launch.  See pstep.cuh for the design.
#include <stdlib.h>

#include <vector>

#include "pstep.cuh"
#include "pstep_host.h"

using namespace wmar;
using namespace wmar::ps;

namespace wmar {

struct PstepState {
    wmar_gpt_config cfg;
    int G, NG;
    PsPlan plan;
    PsArgs args;                 // everything but B / err
    uint8_t *wpack, *head_pack;
    PsProg *d_prog;
    PsLayer *d_layers;
    uint8_t *pool;               // all {value, flag} buffers + the abort flag: one memset per generation
    size_t pool_bytes;
    unsigned long long *d_trace;
};

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool pstep_eligible(const wmar_gpt_config &c, int n_sms) {
    if (c.n_embd % 64 != 0 || c.n_embd / c.n_head != 64 || c.vocab_size % 64 != 0) return false;
    if (c.block_size > 1024 || c.n_layer > 120 || n_sms < 8) return false;
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    return max_smem >= PS_SMEM_BYTES;
}

template <int NG>
static int launch_t(const PsArgs &a, int G, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        WMAR_CUDA_CHECK(cudaFuncSetAttribute(pstep_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_BYTES));
        configured = true;
    }
    pstep_kernel<NG><<<G, (NG * 4 + 1) * 32, PS_SMEM_BYTES, s>>>(a);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
}

int pstep_create(const wmar_gpt_config &cfg, int n_sms, const PstepWeights &w, float *kcache, float *vcache, float *logits,
                 const int *d_step, const int64_t *d_seq, int seq_ld, PstepState **out) {
    PstepState *s = new (std::nothrow) PstepState();
    if (!s) return set_error(WMAR_ERR_NOMEM, "out of host memory%s%s");
    *s = PstepState{};
    s->cfg = cfg;
    const int d = cfg.n_embd, V = cfg.vocab_size, L = cfg.n_layer;
    // one CTA per SM; 144 = 24 * 6 = 72 * 2 tiles the d = 1536 shapes exactly (WMAR_PSTEP_G overrides)
    s->G = env_int("WMAR_PSTEP_G", n_sms >= 144 ? 144 : n_sms);
    if (s->G > n_sms) s->G = n_sms;
    s->NG = env_int("WMAR_PSTEP_NG", 4) == 2 ? 2 : 4;
    std::vector<PsProg> progs((size_t)s->G);
    const int prc = ps_make_plan(s->G, d, cfg.n_head, V, progs.data(), &s->plan);
    if (prc != 0) { delete s; return set_error(WMAR_ERR_INVALID, "the persistent step plan does not fit this model%s%s"); }

    const size_t dd = (size_t)d * d * sizeof(float);
    const size_t layer_bytes = 12 * dd;
    WMAR_CUDA_CHECK(cudaMalloc(&s->wpack, layer_bytes * (size_t)L));
    WMAR_CUDA_CHECK(cudaMalloc(&s->head_pack, (size_t)V * d * sizeof(float)));
    WMAR_CUDA_CHECK(cudaMalloc(&s->d_prog, sizeof(PsProg) * (size_t)s->G));
    WMAR_CUDA_CHECK(cudaMalloc(&s->d_layers, sizeof(PsLayer) * (size_t)L));
    WMAR_CUDA_CHECK(cudaMemcpy(s->d_prog, progs.data(), sizeof(PsProg) * (size_t)s->G, cudaMemcpyHostToDevice));
    std::vector<PsLayer> hl((size_t)L);
    const size_t ph_off[4] = {0, 3 * dd, 4 * dd, 8 * dd};
    for (int l = 0; l < L; l++) {
        const PstepWeights::Layer &W = w.layers[l];
        hl[l] = PsLayer{W.ln1_g, W.ln1_b, W.bqkv, W.bproj, W.ln2_g, W.ln2_b, W.b1, W.b2};
        uint8_t *base = s->wpack + (size_t)l * layer_bytes;
        pack_weight_kernel<<<1024, 256>>>(W.wqkv, 3 * d, d, reinterpret_cast<float4 *>(base + ph_off[0]));
        pack_weight_kernel<<<1024, 256>>>(W.wproj, d, d, reinterpret_cast<float4 *>(base + ph_off[1]));
        pack_weight_kernel<<<1024, 256>>>(W.w1, 4 * d, d, reinterpret_cast<float4 *>(base + ph_off[2]));
        pack_weight_kernel<<<1024, 256>>>(W.w2, d, 4 * d, reinterpret_cast<float4 *>(base + ph_off[3]));
    }
    pack_weight_kernel<<<1024, 256>>>(w.head, V, d, reinterpret_cast<float4 *>(s->head_pack));
    WMAR_CUDA_CHECK(cudaGetLastError());
    WMAR_CUDA_CHECK(cudaMemcpy(s->d_layers, hl.data(), sizeof(PsLayer) * (size_t)L, cudaMemcpyHostToDevice));

    // {value, flag} pool
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_xa = take(16 * (size_t)d * 8), o_xb = take(16 * (size_t)d * 8), o_y = take(16 * (size_t)d * 8);
    const size_t o_qkv = take(16 * (size_t)3 * d * 8), o_h = take(16 * (size_t)4 * d * 8);
    const size_t o_sta = take((size_t)(d / 64) * 16 * 16), o_stb = take((size_t)(d / 64) * 16 * 16);
    size_t o_ws[PH_N];
    for (int ph = 0; ph < PH_N; ph++) o_ws[ph] = take((size_t)(s->plan.n_slots[ph] > 0 ? s->plan.n_slots[ph] : 1) * 1024 * 8);
    const size_t o_abort = take(256);
    s->pool_bytes = off;
    WMAR_CUDA_CHECK(cudaMalloc(&s->pool, s->pool_bytes));
    WMAR_CUDA_CHECK(cudaMemset(s->pool, 0, s->pool_bytes));

    PsArgs &a = s->args;
    a.prog = s->d_prog; a.layers = s->d_layers; a.wpack = s->wpack; a.layer_bytes = layer_bytes;
    for (int ph = 0; ph < 4; ph++) a.ph_off16[ph] = (uint32_t)(ph_off[ph] / 16);
    a.head_pack = s->head_pack;
    a.tok_emb = w.tok_emb; a.pos_emb = w.pos_emb; a.lnf_g = w.lnf_g; a.lnf_b = w.lnf_b;
    a.d = d; a.H = cfg.n_head; a.V = V; a.L = L; a.T = cfg.block_size; a.B = 0;
    a.step = d_step; a.seq = d_seq; a.seq_ld = seq_ld;
    auto U64 = [&](size_t o) { return reinterpret_cast<unsigned long long *>(s->pool + o); };
    a.xa = U64(o_xa); a.xb = U64(o_xb); a.y = U64(o_y); a.qkv = U64(o_qkv); a.h = U64(o_h);
    a.sta = reinterpret_cast<ulonglong2 *>(s->pool + o_sta); a.stb = reinterpret_cast<ulonglong2 *>(s->pool + o_stb);
    for (int ph = 0; ph < PH_N; ph++) a.ws[ph] = U64(o_ws[ph]);
    a.kcache = kcache; a.vcache = vcache; a.logits = logits;
    a.abort_flag = reinterpret_cast<int *>(s->pool + o_abort);
    a.err = nullptr; a.trace = nullptr; a.dbg = env_int("WMAR_PSTEP_DBG", 0);
    if (getenv("WMAR_PSTEP_TRACE")) {
        WMAR_CUDA_CHECK(cudaMalloc(&s->d_trace, sizeof(unsigned long long) * (size_t)s->G * PS_TRACE_EV));
        WMAR_CUDA_CHECK(cudaMemset(s->d_trace, 0, sizeof(unsigned long long) * (size_t)s->G * PS_TRACE_EV));
        a.trace = s->d_trace;
    }
    WMAR_CUDA_CHECK(cudaDeviceSynchronize());   // packing done before the caller may free / patch the originals
    *out = s;
    return WMAR_OK;
}

void pstep_destroy(PstepState *s) {
    if (!s) return;
    cudaFree(s->wpack); cudaFree(s->head_pack); cudaFree(s->d_prog); cudaFree(s->d_layers); cudaFree(s->pool);
    cudaFree(s->d_trace);
    delete s;
}

int pstep_reset(PstepState *s, cudaStream_t stream) {
    WMAR_CUDA_CHECK(cudaMemsetAsync(s->pool, 0, s->pool_bytes, stream));
    return WMAR_OK;
}

int pstep_enqueue(PstepState *s, int B, int *d_err, cudaStream_t stream) {
    PsArgs a = s->args;
    a.B = B;
    a.err = d_err;
    return s->NG == 2 ? launch_t<2>(a, s->G, stream) : launch_t<4>(a, s->G, stream);
}

int pstep_trace(PstepState *s, unsigned long long *out, size_t cap) {
    if (!s || !s->d_trace) return -1;
    const size_t n = (size_t)s->G * PS_TRACE_EV;
    if (cap < n) return -1;
    cudaDeviceSynchronize();
    cudaMemcpy(out, s->d_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return s->G;
}

}  // namespace wmar

/* CPU-testable view of the static plan: h_progs_out receives G records of wmar_pstep_prog_bytes() bytes
 * (layout: pstep_plan.h PsProg), h_slots_out the partial-slot count of the five GEMM phases.  No CUDA calls. */
extern "C" int wmar_pstep_prog_bytes(void) { return (int)sizeof(PsProg); }
extern "C" int wmar_pstep_plan_debug(int G, int d, int H, int V, void *h_progs_out, int *h_slots_out) {
    if (!h_progs_out || !h_slots_out) return WMAR_ERR_INVALID;
    PsPlan plan;
    const int rc = ps_make_plan(G, d, H, V, reinterpret_cast<PsProg *>(h_progs_out), &plan);
    if (rc != 0) return set_error(WMAR_ERR_INVALID, "plan does not fit%s%s");
    for (int ph = 0; ph < PH_N; ph++) h_slots_out[ph] = plan.n_slots[ph];
    return WMAR_OK;
}
