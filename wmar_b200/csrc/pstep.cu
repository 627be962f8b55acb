// Host side of the persistent decode-step kernel: plan, pack, launch.  See pstep.cuh for the design.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "pstep.cuh"
#include "pstep_host.h"

using namespace wmar;
using namespace wmar::ps;

namespace wmar {

struct PstepState {
    wmar_gpt_config cfg;
    int G;
    PsPlan plan;
    PsArgs args;                 // everything but B / err
    uint8_t *wpack, *head_pack;
    PsProg *d_prog;
    PsLayer *d_layers;
    uint8_t *pool;               // activations, fc2 partials, flags, abort flag
    size_t pool_bytes, flags_off, flags_bytes;
    unsigned long long *d_trace;
};

static int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

bool pstep_eligible(const wmar_gpt_config &c, int n_sms) {
    if (c.n_embd % 64 != 0 || c.n_embd / c.n_head != 64 || c.n_embd % c.n_head != 0 || c.vocab_size % 16 != 0) return false;
    if (c.n_embd > PS_DMAX || c.block_size > 1024 || c.n_layer > 200 || n_sms < 8 || n_sms > PS_RED_NQ * PS_RED_MAXC) return false;
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return false;
    if (max_smem < PS_SMEM_BYTES) return false;
    const int G = std::min(env_int("WMAR_PSTEP_G", n_sms), n_sms);
    std::vector<PsProg> progs((size_t)std::max(G, 1));
    PsPlan plan;
    return ps_make_plan(G, c.n_embd, c.n_head, c.vocab_size, progs.data(), &plan) == 0;
}

int pstep_create(const wmar_gpt_config &cfg, int n_sms, const PstepWeights &w, float *kcache, float *vcache, float *logits,
                 const int *d_step, const int64_t *d_seq, int seq_ld, PstepState **out) {
    PstepState *s = new (std::nothrow) PstepState();
    if (!s) return set_error(WMAR_ERR_NOMEM, "out of host memory%s%s");
    *s = PstepState{};
    s->cfg = cfg;
    const int d = cfg.n_embd, V = cfg.vocab_size, L = cfg.n_layer;
    s->G = std::min(env_int("WMAR_PSTEP_G", n_sms), n_sms);   // one CTA per SM
    std::vector<PsProg> progs((size_t)s->G);
    const int prc = ps_make_plan(s->G, d, cfg.n_head, V, progs.data(), &s->plan);
    if (prc != 0) { delete s; return set_error(WMAR_ERR_INVALID, "the persistent step plan does not fit this model%s%s"); }
    const PsPlan &pl = s->plan;

    const size_t layer_bytes = (size_t)pl.layer_stages * PS_STAGE_BYTES;
    const size_t head_bytes = (size_t)pl.head_stages * PS_STAGE_BYTES;
    WMAR_CUDA_CHECK(cudaMalloc(&s->wpack, layer_bytes * (size_t)L));
    WMAR_CUDA_CHECK(cudaMalloc(&s->head_pack, head_bytes));
    WMAR_CUDA_CHECK(cudaMalloc(&s->d_prog, sizeof(PsProg) * (size_t)s->G));
    WMAR_CUDA_CHECK(cudaMalloc(&s->d_layers, sizeof(PsLayer) * (size_t)L));
    WMAR_CUDA_CHECK(cudaMemcpy(s->d_prog, progs.data(), sizeof(PsProg) * (size_t)s->G, cudaMemcpyHostToDevice));
    unsigned max_ls = 1, max_hs = 1;
    for (int c = 0; c < s->G; c++) { max_ls = std::max(max_ls, progs[c].layer_stages); max_hs = std::max(max_hs, progs[c].head_stages); }
    std::vector<PsLayer> hl((size_t)L);
    for (int l = 0; l < L; l++) {
        const PstepWeights::Layer &W = w.layers[l];
        hl[l] = PsLayer{W.ln1_g, W.ln1_b, W.bqkv, W.bproj, W.ln2_g, W.ln2_b, W.b1, W.b2};
        PackArgs pa{};
        pa.prog = s->d_prog; pa.w[0] = W.wqkv; pa.w[1] = W.wproj; pa.w[2] = W.w1; pa.w2 = W.w2;
        pa.dst = s->wpack + (size_t)l * layer_bytes;
        pa.d = d; pa.KC = pl.KC; pa.NBn = pl.NBn; pa.head = 0;
        pa.Nrows[0] = 3 * d; pa.Nrows[1] = d; pa.Nrows[2] = 4 * d;
        pack_stage_kernel<<<dim3((unsigned)s->G, max_ls), 256>>>(pa);
    }
    {
        PackArgs pa{};
        pa.prog = s->d_prog; pa.w[0] = w.head; pa.w[1] = nullptr; pa.w[2] = nullptr; pa.w2 = nullptr;
        pa.dst = s->head_pack;
        pa.d = d; pa.KC = pl.KC; pa.NBn = pl.NBn; pa.head = 1;
        pa.Nrows[0] = V; pa.Nrows[1] = 0; pa.Nrows[2] = 0;
        pack_stage_kernel<<<dim3((unsigned)s->G, max_hs), 256>>>(pa);
    }
    WMAR_CUDA_CHECK(cudaGetLastError());
    WMAR_CUDA_CHECK(cudaMemcpy(s->d_layers, hl.data(), sizeof(PsLayer) * (size_t)L, cudaMemcpyHostToDevice));

    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const int GP = (s->G + 31) / 32 * 32;
    const size_t o_x = take(16 * (size_t)d * 4), o_xb = take(16 * (size_t)d * 4), o_y = take(16 * (size_t)d * 4);
    const size_t o_qkv = take(16 * (size_t)3 * d * 4);
    const size_t o_part = take((size_t)s->G * 16 * d * 4);
    const size_t o_flags = take((size_t)FL_N * GP * 4 + 256);      // + the abort flag
    s->pool_bytes = off;
    s->flags_off = o_flags;
    s->flags_bytes = (size_t)FL_N * GP * 4 + 256;
    WMAR_CUDA_CHECK(cudaMalloc(&s->pool, s->pool_bytes));
    WMAR_CUDA_CHECK(cudaMemset(s->pool, 0, s->pool_bytes));    // partials of CTAs without fc1 tiles stay zero forever

    PsArgs &a = s->args;
    a.prog = s->d_prog; a.layers = s->d_layers; a.wpack = s->wpack; a.layer_bytes = layer_bytes;
    a.head_pack = s->head_pack;
    a.tok_emb = w.tok_emb; a.pos_emb = w.pos_emb; a.lnf_g = w.lnf_g; a.lnf_b = w.lnf_b;
    a.d = d; a.H = cfg.n_head; a.V = V; a.L = L; a.T = cfg.block_size; a.B = 0; a.G = s->G; a.GP = GP;
    a.Kp = pl.Kp; a.KC = pl.KC; a.NBn = pl.NBn;
    a.pf_dist = (unsigned)std::max(0, env_int("WMAR_PSTEP_PF_KB", 0)) * 1024u;
    a.wait_hint_ns = (unsigned)std::max(0, env_int("WMAR_PSTEP_HINT_NS", 20000));
    a.dbg = env_int("WMAR_PSTEP_DBG", 0);
    a.step = d_step; a.seq = d_seq; a.seq_ld = seq_ld;
    auto F32 = [&](size_t o) { return reinterpret_cast<float *>(s->pool + o); };
    a.x = F32(o_x); a.xb = F32(o_xb); a.y = F32(o_y); a.qkv = F32(o_qkv); a.part = F32(o_part);
    a.flags = reinterpret_cast<unsigned *>(s->pool + o_flags);
    a.abort_flag = reinterpret_cast<int *>(s->pool + o_flags + (size_t)FL_N * GP * 4);
    a.kcache = kcache; a.vcache = vcache; a.logits = logits;
    a.err = nullptr; a.trace = nullptr;
    if (getenv("WMAR_PSTEP_TRACE")) {
        WMAR_CUDA_CHECK(cudaMalloc(&s->d_trace, sizeof(unsigned long long) * (size_t)s->G * PS_TRACE_EV));
        WMAR_CUDA_CHECK(cudaMemset(s->d_trace, 0, sizeof(unsigned long long) * (size_t)s->G * PS_TRACE_EV));
        a.trace = s->d_trace;
    }
    WMAR_CUDA_CHECK(cudaFuncSetAttribute(pstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_BYTES));
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
    // epochs restart with the token-step counter: clear the flags (and the abort flag) once per generation
    WMAR_CUDA_CHECK(cudaMemsetAsync(s->pool + s->flags_off, 0, s->flags_bytes, stream));
    return WMAR_OK;
}

int pstep_enqueue(PstepState *s, int B, int *d_err, cudaStream_t stream) {
    PsArgs a = s->args;
    a.B = B;
    a.err = d_err;
    pstep_kernel<<<s->G, PS_THREADS, PS_SMEM_BYTES, stream>>>(a);
    WMAR_LAUNCH_CHECK();
    return WMAR_OK;
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

/* CPU-testable view of the static plan: h_progs_out receives G records of wmar_pstep_prog_bytes() bytes (layout:
 * pstep_plan.h PsProg); h_info_out[8] = {Kp, KC, NBn, layer_stages, head_stages, max_load, min_load, 0}.  No CUDA calls. */
extern "C" int wmar_pstep_prog_bytes(void) { return (int)sizeof(PsProg); }
extern "C" int wmar_pstep_plan_debug(int G, int d, int H, int V, void *h_progs_out, long long *h_info_out) {
    if (!h_progs_out || !h_info_out) return WMAR_ERR_INVALID;
    PsPlan plan;
    const int rc = ps_make_plan(G, d, H, V, reinterpret_cast<PsProg *>(h_progs_out), &plan);
    if (rc != 0) return set_error(WMAR_ERR_INVALID, "plan does not fit%s%s");
    h_info_out[0] = plan.Kp; h_info_out[1] = plan.KC; h_info_out[2] = plan.NBn; h_info_out[3] = plan.layer_stages;
    h_info_out[4] = plan.head_stages; h_info_out[5] = plan.max_load; h_info_out[6] = plan.min_load; h_info_out[7] = 0;
    return WMAR_OK;
}
/* where stage `s` of CTA `cta`'s layer block (head == 0) or head block comes from: out[4] = {phase (-1 = fc2), n16 tile
 * (fc2: the fc1 tile whose 16 columns are contracted), k chunk, fc2 n-block}.  Host logic only. */
extern "C" int wmar_pstep_stage_src(int G, int d, int H, int V, int cta, int s, int head, int *out4) {
    if (!out4 || cta < 0 || cta >= G) return WMAR_ERR_INVALID;
    std::vector<PsProg> progs((size_t)G);
    PsPlan plan;
    if (ps_make_plan(G, d, H, V, progs.data(), &plan) != 0) return set_error(WMAR_ERR_INVALID, "plan does not fit%s%s");
    const PsStageSrc r = ps_stage_src(plan, progs[cta], s, head);
    out4[0] = r.ph; out4[1] = r.tile; out4[2] = r.kc; out4[3] = r.nb;
    return WMAR_OK;
}
