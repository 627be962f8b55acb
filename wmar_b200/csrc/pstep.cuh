// Persistent decode-step kernel for the Taming minGPT engine: ONE launch runs a whole token step
// (embedding -> 48 x [LN1+QKV, attention over the KV cache, proj+residual, LN2+fc1+GELU, fc2+residual] -> LN_f + head)
// replacing the 244 dependent launches per token of the per-GEMM graph (mingpt.py:183-214, 98-122, 42-95).
//
// Design (B200 first):
//   * one CTA per SM (grid = PS_G <= #SMs, all co-resident), 1 producer warp + NG x 4 consumer warps;
//   * the step is HBM-bound on the weights (5.5 GB / token) and on the K/V cache.  Neither depends on the
//     activations, so ONE elected producer thread per CTA streams the CTA's static share of every weight matrix
//     and of the K/V rows of earlier tokens with cp.async.bulk (TMA bulk copies, 16 KB each) into a 10-deep
//     mbarrier ring (160 KB / SM in flight) and simply keeps going across GEMM and layer boundaries: the stream
//     never drains while the consumers wait for a dependency;
//   * weights are re-tiled once at create time (wmar_gpt_create) into 16 KB stages [n64 tile][k64 stage] whose
//     bytes are already in mma.m16n8k8 B-fragment order for 4 warps x 32 lanes: a stage is one contiguous
//     bulk copy, and every lane reads its fragments back with conflict-free LDS.128;
//   * a GEMM phase is cut into (tile, k-range) items, a static share per CTA (pstep_plan.h); the 16 batch rows
//     are the M of the MMA, products are 3xTF32 with fp32 accumulation (fp32-faithful, as the reference's
//     TF32-off Linear layers need for greedy token parity), warps split K and reduce through shared memory
//     in a fixed order, CTAs that share a tile hand their partial to the CTA owning the tile's last k-range;
//   * every hand-off between CTAs (split-K partials, activations, LayerNorm statistics, q/k/v, attention
//     output) travels as self-validating {value, flag} 8-byte words (flag = token step and phase): no grid
//     barrier, no fence, no atomic; a consumer polls exactly the words it needs;
//   * attention: (head, row) items run on 128-thread warp groups, up to NG at a time per CTA, K then V of the
//     cached tokens arrive through the same ring; two-pass softmax in shared memory.
// All waits are bounded: a wait that exceeds ~1 s raises the device error flag (bit 2) and the kernel drains.
#pragma once
#include "gemm.cuh"
#include "pstep_plan.h"
#include "tc05.cuh"

namespace wmar {
namespace ps {

using namespace tc05;

constexpr int PS_NS = 10;                       // ring depth (stages)
constexpr int PS_XBUF_BYTES = 65536;            // X slice of an item, later its cross-warp reduction buffer
constexpr int PS_RING_BYTES = PS_NS * PS_STAGE_BYTES;
constexpr int PS_OFF_XBUF = PS_RING_BYTES;
constexpr int PS_OFF_BARS = PS_OFF_XBUF + PS_XBUF_BYTES;     // full[NS], empty[NS]
constexpr int PS_OFF_STATS = PS_OFF_BARS + 2 * PS_NS * 8;    // float2 row_stats[16]
constexpr int PS_OFF_DEAD = PS_OFF_STATS + 16 * 8;           // int
constexpr int PS_SMEM_BYTES = PS_OFF_DEAD + 16;
constexpr int PS_RED_LD = 72;
constexpr int PS_ATT_SCRATCH = 8192;            // per warp group: scores[1024], q/k/v[192], part[8][64], red[8]
constexpr unsigned PS_SPIN_LIMIT = 1u << 24;    // LL polls (~40 ns apart)
constexpr unsigned PS_MBAR_LIMIT = 1u << 16;    // try_wait suspends up to 20 us each

struct PsLayer {
    const float *ln1_g, *ln1_b, *bqkv, *bproj, *ln2_g, *ln2_b, *b1, *b2;
};

struct PsArgs {
    const PsProg *prog;          // [G]
    const PsLayer *layers;       // [L]
    const uint8_t *wpack;        // [L][layer_bytes] packed qkv | proj | fc1 | fc2
    unsigned long long layer_bytes;
    uint32_t ph_off16[4];        // phase base inside a layer block, in 16-byte units
    const uint8_t *head_pack;
    const float *tok_emb, *pos_emb, *lnf_g, *lnf_b;
    int d, H, V, L, T, B;
    const int *step;
    const int64_t *seq; int seq_ld;
    unsigned long long *xa, *xb, *qkv, *y, *h;     // {value, flag} activations [16][ld]
    ulonglong2 *sta, *stb;                         // LN statistics of xa / xb: [d/64][16] {mean|flag, M2|flag}
    unsigned long long *ws[PH_N];                  // split-K partial slots [slot][16*64] words
    float *kcache, *vcache;
    float *logits;                                 // plain fp32 [16][V]
    int *abort_flag;                               // global: non-zero = a wait timed out somewhere, drain
    int *err;                                      // device error flag (bit 2 = pstep timeout)
    unsigned long long *trace;                     // probe only: [G][PS_TRACE_EV] globaltimer stamps
    int dbg;
};
constexpr int PS_TRACE_EV = 512;

// ---------------------------------------------------------------------------------------------------------
struct Ctx {
    const PsArgs &a;
    uint8_t *smem;
    int *s_dead;
    unsigned long long *tr_item;   // probe only: stamps inside the GEMM items of one layer
    __device__ __forceinline__ bool dead() const { return *reinterpret_cast<volatile int *>(s_dead) != 0; }
    __device__ __forceinline__ void timeout(int code) const {
        atomicExch(a.abort_flag, code);
        atomicOr(a.err, 4 | (code << 8));
        *reinterpret_cast<volatile int *>(s_dead) = 1;
    }
    // periodic check of the global abort flag from inside spin loops
    __device__ __forceinline__ bool poll_abort(unsigned n, unsigned limit, int code) const {
        if ((n & 255u) == 0u) {
            if (dead()) return true;
            if (*reinterpret_cast<volatile int *>(a.abort_flag) != 0) { *reinterpret_cast<volatile int *>(s_dead) = 1; return true; }
            if (n > limit) { timeout(code); return true; }
        }
        return false;
    }
    __device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity, int code) const {
        if (mbar_try_wait(bar, parity)) return;
        unsigned n = 0;
        while (!mbar_try_wait(bar, parity)) {
            ++n;
            if (dead()) return;
            if ((n & 15u) == 0u && *reinterpret_cast<volatile int *>(a.abort_flag) != 0) { *reinterpret_cast<volatile int *>(s_dead) = 1; return; }
            if (n > PS_MBAR_LIMIT) { timeout(code); return; }
        }
    }
};

template <int NT>
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory"); }
__device__ __forceinline__ void bar_group(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

__device__ __forceinline__ void ll_load1(const unsigned long long *p, unsigned long long &a) {
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar, uint64_t hint) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(hint) : "memory");
}

// 4 consecutive {value, flag} words -> float4 once all four carry `flag`
__device__ __forceinline__ float4 ll_wait4(const Ctx &c, const unsigned long long *p, uint32_t flag, int code) {
    unsigned long long q0, q1, q2, q3;
    unsigned n = 0;
    while (true) {
        ll_load2(p, q0, q1);
        ll_load2(p + 2, q2, q3);
        if ((uint32_t)(q0 >> 32) == flag && (uint32_t)(q1 >> 32) == flag && (uint32_t)(q2 >> 32) == flag && (uint32_t)(q3 >> 32) == flag) break;
        if (c.poll_abort(++n, PS_SPIN_LIMIT, code)) break;
        __nanosleep(32);
    }
    return make_float4(__uint_as_float((uint32_t)q0), __uint_as_float((uint32_t)q1), __uint_as_float((uint32_t)q2), __uint_as_float((uint32_t)q3));
}
__device__ __forceinline__ void ll_store4(unsigned long long *p, float4 v, uint32_t flag) {
    ll_store2(p, ll_pack(v.x, flag), ll_pack(v.y, flag));
    ll_store2(p + 2, ll_pack(v.z, flag), ll_pack(v.w, flag));
}

// ---------------------------------------------------------------------------------------------------------
// Resolved description of a GEMM phase of layer l at token step t, split so that only the input half is live during
// the main loop and only the output half during the epilogue.
struct PhaseIn {
    const unsigned long long *x; int ldx; uint32_t x_flag;
    const ulonglong2 *stats_in; const float *ln_g, *ln_b; int K;
};
struct PhaseRt {
    const float *bias; int epi;                                   // 0 store, 1 GELU, 2 residual
    const unsigned long long *resid; uint32_t resid_flag;
    unsigned long long *out; int ldo; uint32_t out_flag;         // out == nullptr -> plain logits
    ulonglong2 *stats_out;
};

__device__ __forceinline__ uint32_t mkflag(int t, int fid) { return ((uint32_t)(t + 1) << 10) | (uint32_t)fid; }
// flag ids: 1 = embedding; layer l: 8l+2 qkv, 8l+3 attention, 8l+4 proj, 8l+5 fc1, 8l+6 fc2
__device__ __forceinline__ int fid_in(int l) { return l == 0 ? 1 : 8 * (l - 1) + 6; }

__device__ __forceinline__ PhaseIn resolve_in(const PsArgs &a, int phase, int l, int t) {
    PhaseIn r;
    const int d = a.d;
    const PsLayer *L = (phase == PH_HEAD) ? nullptr : a.layers + l;
    r.stats_in = nullptr; r.ln_g = r.ln_b = nullptr; r.K = d; r.ldx = d;
    switch (phase) {
    case PH_QKV: r.x = a.xa; r.x_flag = mkflag(t, fid_in(l)); r.stats_in = a.sta; r.ln_g = L->ln1_g; r.ln_b = L->ln1_b; break;
    case PH_PROJ: r.x = a.y; r.x_flag = mkflag(t, 8 * l + 3); break;
    case PH_FC1: r.x = a.xb; r.x_flag = mkflag(t, 8 * l + 4); r.stats_in = a.stb; r.ln_g = L->ln2_g; r.ln_b = L->ln2_b; break;
    case PH_FC2: r.x = a.h; r.ldx = 4 * d; r.K = 4 * d; r.x_flag = mkflag(t, 8 * l + 5); break;
    default: r.x = a.xa; r.x_flag = mkflag(t, fid_in(l)); r.stats_in = a.sta; r.ln_g = a.lnf_g; r.ln_b = a.lnf_b; break;
    }
    return r;
}
__device__ __forceinline__ PhaseRt resolve_out(const PsArgs &a, int phase, int l, int t) {
    PhaseRt r;
    const int d = a.d;
    const PsLayer *L = (phase == PH_HEAD) ? nullptr : a.layers + l;
    r.resid = nullptr; r.resid_flag = 0; r.stats_out = nullptr; r.bias = nullptr; r.epi = 0;
    switch (phase) {
    case PH_QKV: r.bias = L->bqkv; r.out = a.qkv; r.ldo = 3 * d; r.out_flag = mkflag(t, 8 * l + 2); break;
    case PH_PROJ:
        r.bias = L->bproj; r.epi = 2; r.resid = a.xa; r.resid_flag = mkflag(t, fid_in(l));
        r.out = a.xb; r.ldo = d; r.out_flag = mkflag(t, 8 * l + 4); r.stats_out = a.stb;
        break;
    case PH_FC1: r.bias = L->b1; r.epi = 1; r.out = a.h; r.ldo = 4 * d; r.out_flag = mkflag(t, 8 * l + 5); break;
    case PH_FC2:
        r.bias = L->b2; r.epi = 2; r.resid = a.xb; r.resid_flag = mkflag(t, 8 * l + 4);
        r.out = a.xa; r.ldo = d; r.out_flag = mkflag(t, 8 * l + 6); r.stats_out = a.sta;
        break;
    default: r.out = nullptr; r.ldo = a.V; r.out_flag = 0; break;   // PH_HEAD
    }
    return r;
}

// Per-row (mean, rstd) from the producers' per-tile (mean, M2) words; one warp, result in row_stats[16].
__device__ __forceinline__ void combine_row_stats_ll(const Ctx &c, const ulonglong2 *stats_in, uint32_t flag, int n_tiles, int K,
                                                     float eps, float2 *row_stats, int lane) {
    const int r = lane & 15, half = lane >> 4;
    constexpr int MAXT = 12;
    float n = 0.f, mean = 0.f, m2 = 0.f;
    const float w = (float)(K / n_tiles);
    for (int t0 = 0; t0 < n_tiles; t0 += 2 * MAXT) {
        float2 sv[MAXT];
        unsigned spins = 0;
        while (true) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < MAXT; k++) {
                const int tl = t0 + half + 2 * k;
                sv[k] = make_float2(0.f, 0.f);
                if (tl < n_tiles) {
                    unsigned long long q0, q1;
                    ll_load2(reinterpret_cast<const unsigned long long *>(stats_in + tl * 16 + r), q0, q1);
                    ok = ok && (uint32_t)(q0 >> 32) == flag && (uint32_t)(q1 >> 32) == flag;
                    sv[k] = make_float2(__uint_as_float((uint32_t)q0), __uint_as_float((uint32_t)q1));
                }
            }
            if (ok) break;
            if (c.poll_abort(++spins, PS_SPIN_LIMIT, 11)) break;
            __nanosleep(32);
        }
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
    const float nb = __shfl_xor_sync(0xffffffffu, n, 16);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, 16);
    const float m2b = __shfl_xor_sync(0xffffffffu, m2, 16);
    if (half == 0) {
        chan_combine(n, mean, m2, nb, mb, m2b);
        row_stats[r] = make_float2(mean, 1.0f / sqrtf(m2 / (float)K + eps));
    }
}

// Epilogue of a finished 16 x 64 output tile: thread (m, nn..nn+3) holds v.
__device__ __forceinline__ void tile_epilogue(const Ctx &c, const PhaseRt &rt, int tile, int m, int nn, float4 v, int ctid) {
    const int n = tile * 64 + nn;
    if (rt.bias != nullptr) {
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(rt.bias + n));
        v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    }
    if (rt.epi == 1) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
    if (rt.epi == 2) {
        const float4 r4 = ll_wait4(c, rt.resid + (size_t)m * rt.ldo + n, rt.resid_flag, 12);
        v.x = r4.x + v.x; v.y = r4.y + v.y; v.z = r4.z + v.z; v.w = r4.w + v.w;
    }
    if (rt.out != nullptr) ll_store4(rt.out + (size_t)m * rt.ldo + n, v, rt.out_flag);
    else *reinterpret_cast<float4 *>(c.a.logits + (size_t)m * rt.ldo + n) = v;
    if (rt.stats_out != nullptr) {
        float s = v.x + v.y + v.z + v.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 64.0f);
        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        float q = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if ((ctid & 15) == 0)
            ll_store2(reinterpret_cast<unsigned long long *>(rt.stats_out + tile * 16 + m), ll_pack(mean, rt.out_flag), ll_pack(q, rt.out_flag));
    }
}

// ---------------------------------------------------------------------------------------------------------
// One GEMM item: Y[16][64 of tile] (+)= X[16][k-range] . Wtile^T over nst ring stages starting at ring index gi.
template <int NG>
__device__ __forceinline__ void gemm_item(const Ctx &c, const int phase, const int l, const int t, const PsItem it, uint32_t gi, int ctid) {
    constexpr int NW = NG * 4, NT = NW * 32;
    const int lane = ctid & 31, cw = ctid >> 5, group = cw >> 2, wg = cw & 3;
    const int g = lane >> 2;
    const int nst = it.nst, k0 = it.k0st * 64;
    uint8_t *xbuf = c.smem + PS_OFF_XBUF;
    float2 *row_stats = reinterpret_cast<float2 *>(c.smem + PS_OFF_STATS);
    const uint32_t full0 = smem_u32(c.smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;

    // ---- 1. stage the X slice (16 x nst*64 fp32) into shared memory in A-fragment order, LayerNorm applied
    bar_consumers<NT>();                     // the previous user of xbuf (reduction buffer / attention scratch) is done
    {
    const PhaseIn rt = resolve_in(c.a, phase, l, t);
    // 1a. LIGHT wait for the producers of this K slice.  Polling the slice itself from every thread saturated L2
    // (measured: 17-19 us per LN phase); instead one word per source is polled -- the LayerNorm statistics (written
    // after the tile's values, needed anyway) or the last word of each source (tile, row) -- and the slice is then
    // loaded once, every word still verified by its own flag (a miss only costs a retry).
    if (rt.stats_in != nullptr) {
        if (cw == 0) combine_row_stats_ll(c, rt.stats_in, rt.x_flag, rt.K / 64, rt.K, 1e-5f, row_stats, lane);
    } else {
        // attention output rows come from 16 different CTAs per tile; a GEMM-produced tile from one CTA (row 15 = its last warp)
        const int rows_mode = (phase == PH_PROJ) ? 16 : 1;
        if (ctid < nst * rows_mode) {
            const int ti = ctid / rows_mode, row = rows_mode == 16 ? (ctid & 15) : 15;
            const unsigned long long *p = rt.x + (size_t)row * rt.ldx + k0 + ti * 64 + 63;
            unsigned long long q;
            unsigned n = 0;
            while (true) {
                ll_load1(p, q);
                if ((uint32_t)(q >> 32) == rt.x_flag) break;
                if (c.poll_abort(++n, PS_SPIN_LIMIT, 19)) break;
                __nanosleep(64);
            }
        }
    }
    bar_consumers<NT>();
    {
        const int per_row = nst * 16;        // float4 chunks per row
        const int total = per_row * 16;
        for (int q0 = ctid; q0 < total; q0 += 4 * NT) {
            float4 xv[4];
            int rr[4], kk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int q = q0 + u * NT;
                rr[u] = q / per_row; kk[u] = (q - rr[u] * per_row) * 4;
            }
            // all loads of the batch in flight, then verify; on a miss the whole batch is re-read
            unsigned spins = 0;
            while (true) {
                unsigned long long w[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (q0 + u * NT < total) {
                        const unsigned long long *p = rt.x + (size_t)rr[u] * rt.ldx + k0 + kk[u];
                        ll_load2(p, w[u][0], w[u][1]);
                        ll_load2(p + 2, w[u][2], w[u][3]);
                    }
                bool ok = true;
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (q0 + u * NT < total) {
                        ok = ok && (uint32_t)(w[u][0] >> 32) == rt.x_flag && (uint32_t)(w[u][1] >> 32) == rt.x_flag &&
                             (uint32_t)(w[u][2] >> 32) == rt.x_flag && (uint32_t)(w[u][3] >> 32) == rt.x_flag;
                        xv[u] = make_float4(__uint_as_float((uint32_t)w[u][0]), __uint_as_float((uint32_t)w[u][1]),
                                            __uint_as_float((uint32_t)w[u][2]), __uint_as_float((uint32_t)w[u][3]));
                    }
                if (ok) break;
                if (c.poll_abort(++spins, PS_SPIN_LIMIT, 13)) break;
                __nanosleep(64);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (q0 + u * NT < total) {
                    float4 v = xv[u];
                    if (rt.stats_in != nullptr) {
                        const float2 st = row_stats[rr[u]];
                        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(rt.ln_g + k0 + kk[u]));
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(rt.ln_b + k0 + kk[u]));
                        v.x = (v.x - st.x) * st.y * g4.x + b4.x; v.y = (v.y - st.x) * st.y * g4.y + b4.y;
                        v.z = (v.z - st.x) * st.y * g4.z + b4.z; v.w = (v.w - st.x) * st.y * g4.w + b4.w;
                    }
                    // chunk (k16) c16, quad tq inside it; rows 0-7 in the first 512 bytes of the chunk, 8-15 in the second
                    const int c16 = kk[u] >> 4, tq = (kk[u] >> 2) & 3;
                    *reinterpret_cast<float4 *>(xbuf + c16 * 1024 + (rr[u] >> 3) * 512 + ((rr[u] & 7) * 4 + tq) * 16) = v;
                }
        }
    }
    }
    bar_consumers<NT>();
    if (c.tr_item && ctid == 0) c.tr_item[phase * 4 + 0] = timer_ns();

    // ---- 2. main loop: this warp group takes every NG-th stage, each warp one k16 chunk of it
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][e] = 0.f;
    for (int s = group; s < nst; s += NG) {
        const uint32_t gs = gi + (uint32_t)s, slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
        c.mbar_wait_b(full0 + slot * 8, parity, 14);
        const uint8_t *src = c.smem + slot * PS_STAGE_BYTES + wg * 4096 + lane * 16;
        const float4 xa = *reinterpret_cast<const float4 *>(xbuf + (s * 4 + wg) * 1024 + lane * 16);
        const float4 xb = *reinterpret_cast<const float4 *>(xbuf + (s * 4 + wg) * 1024 + 512 + lane * 16);
        const float xs[2][4] = {{xa.x, xa.y, xa.z, xa.w}, {xb.x, xb.y, xb.z, xb.w}};
        uint32_t xh[2][4], xl[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int e = 0; e < 4; e++) split_tf32(xs[r][e], xh[r][e], xl[r][e]);
#pragma unroll
        for (int jh = 0; jh < 2; jh++) {     // two halves of four n8 tiles: 16 weight registers live at a time
            float4 wcur[4];
#pragma unroll
            for (int j = 0; j < 4; j++) wcur[j] = *reinterpret_cast<const float4 *>(src + (jh * 4 + j) * 512);
            if (jh == 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + slot * 8);   // the stage's bytes are in registers
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float wv[4] = {wcur[j].x, wcur[j].y, wcur[j].z, wcur[j].w};
                uint32_t wh[4], wl[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    wh[e] = __float_as_uint(wv[e]) & 0xffffe000u;
                    wl[e] = __float_as_uint(wv[e] - __uint_as_float(wh[e]));
                }
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int e = 2 * half;
                    mma_tf32(acc[jh * 4 + j], xl[0][e], xl[1][e], xl[0][e + 1], xl[1][e + 1], wh[e], wh[e + 1]);
                    mma_tf32(acc[jh * 4 + j], xh[0][e], xh[1][e], xh[0][e + 1], xh[1][e + 1], wl[e], wl[e + 1]);
                    mma_tf32(acc[jh * 4 + j], xh[0][e], xh[1][e], xh[0][e + 1], xh[1][e + 1], wh[e], wh[e + 1]);
                }
            }
        }
    }

    // ---- 3. cross-warp reduction in a fixed order (the buffer aliases the X slice)
    if (c.tr_item && ctid == 0) c.tr_item[phase * 4 + 1] = timer_ns();
    bar_consumers<NT>();
    float *red = reinterpret_cast<float *>(xbuf);
    const int tq = lane & 3;
    auto red_write = [&](int w) {
        float *my = red + w * 16 * PS_RED_LD;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            *reinterpret_cast<float2 *>(my + g * PS_RED_LD + 8 * j + 2 * tq) = make_float2(acc[j][0], acc[j][1]);
            *reinterpret_cast<float2 *>(my + (g + 8) * PS_RED_LD + 8 * j + 2 * tq) = make_float2(acc[j][2], acc[j][3]);
        }
    };
    if (NG == 4) {
        if (cw >= 8) red_write(cw - 8);
        bar_consumers<NT>();
        if (cw < 8) {
            const float *my = red + cw * 16 * PS_RED_LD;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float2 p0 = *reinterpret_cast<const float2 *>(my + g * PS_RED_LD + 8 * j + 2 * tq);
                const float2 p1 = *reinterpret_cast<const float2 *>(my + (g + 8) * PS_RED_LD + 8 * j + 2 * tq);
                acc[j][0] += p0.x; acc[j][1] += p0.y; acc[j][2] += p1.x; acc[j][3] += p1.y;
            }
            red_write(cw);
        }
    } else {
        red_write(cw);
    }
    bar_consumers<NT>();
    if (ctid >= 256) return;                 // warps 8.. go on to the next item's first barrier
    const int m = ctid >> 4, nn = (ctid & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const float4 p = *reinterpret_cast<const float4 *>(red + (w * 16 + m) * PS_RED_LD + nn);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }

    // ---- 4. hand-off
    if (c.tr_item && ctid == 0) c.tr_item[phase * 4 + 2] = timer_ns();
    unsigned long long *ws = c.a.ws[phase];
    const uint32_t pflag = resolve_in(c.a, phase, l, t).x_flag;      // unique per (step, layer) inside this phase's slots
    if (!it.reducer) {
        ll_store4(ws + (size_t)it.slot * 1024 + m * 64 + nn, v, pflag);
        return;
    }
    if (it.nparts > 0) {
        const float4 own = v;
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < it.nparts; s0 += 4) {          // four partials in flight, summed in k order
            unsigned long long q[4][4];
            unsigned spins = 0;
            while (true) {
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (s0 + k < it.nparts) {
                        const unsigned long long *src = ws + (size_t)(it.slot + s0 + k) * 1024 + m * 64 + nn;
                        ll_load2(src, q[k][0], q[k][1]);
                        ll_load2(src + 2, q[k][2], q[k][3]);
                    }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (s0 + k < it.nparts) {
#pragma unroll
                        for (int e = 0; e < 4; e++) ok = ok && ((uint32_t)(q[k][e] >> 32) == pflag);
                    }
                if (ok) break;
                if (c.poll_abort(++spins, PS_SPIN_LIMIT, 15)) break;
                __nanosleep(64);
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (s0 + k < it.nparts) {
                    v.x += __uint_as_float((uint32_t)q[k][0]); v.y += __uint_as_float((uint32_t)q[k][1]);
                    v.z += __uint_as_float((uint32_t)q[k][2]); v.w += __uint_as_float((uint32_t)q[k][3]);
                }
        }
        v.x += own.x; v.y += own.y; v.z += own.z; v.w += own.w;
    }
    tile_epilogue(c, resolve_out(c.a, phase, l, t), it.tile, m, nn, v, ctid);
}

// ---------------------------------------------------------------------------------------------------------
// Attention batch: up to NG (head,row) items of this CTA, one per warp group.  K stages then V stages of the cached
// tokens come through the ring (order: stage-major, item-minor; rows >= B are skipped by producer and consumers alike).
template <int NG>
__device__ __forceinline__ uint32_t attn_stage_count(const PsArgs &a, const PsProg &pg, int batch, int t) {
    int nact = 0;
    for (int i = 0; i < NG; i++) {
        const int r = batch * NG + i;
        if (r < pg.n_attn && (pg.attn[r] & 15) < a.B) nact++;
    }
    return (uint32_t)(2 * ((t + 63) >> 6) * nact);
}

template <int NG>
__device__ __forceinline__ void attn_batch(const Ctx &c, const PsProg &pg, int batch, int l, int t, uint32_t gi, int ctid) {
    constexpr int NW = NG * 4, NT = NW * 32;
    const PsArgs &a = c.a;
    const int lane = ctid & 31, cw = ctid >> 5, group = cw >> 2, wg = cw & 3;
    const int tid128 = ctid & 127, sub = lane & 15, hf = lane >> 4;
    const uint32_t full0 = smem_u32(c.smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;
    bar_consumers<NT>();                     // xbuf is free
    const int r = batch * NG + group;
    if (r >= pg.n_attn) return;
    const int item = pg.attn[r], h = item >> 4, b = item & 15;
    const int d = a.d;
    const uint32_t out_flag = mkflag(t, 8 * l + 3);
    unsigned long long *yout = a.y + (size_t)b * d + h * 64;
    if (b >= a.B) {                           // inactive row: zeros keep the proj GEMM's rows finite
        if (tid128 < 16) ll_store4(yout + 4 * tid128, make_float4(0.f, 0.f, 0.f, 0.f), out_flag);
        return;
    }
    int nact = 0, ai = 0;
    for (int i = 0; i < NG; i++) {
        const int rr = batch * NG + i;
        if (rr < pg.n_attn && (pg.attn[rr] & 15) < a.B) { if (i < group) ai++; nact++; }
    }
    const int nK = (t + 63) >> 6;
    float *scr = reinterpret_cast<float *>(c.smem + PS_OFF_XBUF + group * PS_ATT_SCRATCH);
    float *sc = scr, *qs = scr + 1024, *kn = scr + 1088, *vn = scr + 1152, *part = scr + 1216, *redv = scr + 1728;
    const size_t cbase = (((size_t)l * 16 + b) * a.H + h) * (size_t)a.T * 64;

    // 1. this step's q, k, v (written by the qkv reducers); k, v are appended to the cache for later steps
    if (tid128 < 48) {
        const int which = tid128 >> 4, cq = tid128 & 15;
        const float4 v4 = ll_wait4(c, a.qkv + (size_t)b * 3 * d + which * d + h * 64 + 4 * cq, mkflag(t, 8 * l + 2), 16);
        *reinterpret_cast<float4 *>(scr + 1024 + which * 64 + 4 * cq) = v4;
        if (which == 1) *reinterpret_cast<float4 *>(a.kcache + cbase + (size_t)t * 64 + 4 * cq) = v4;
        if (which == 2) *reinterpret_cast<float4 *>(a.vcache + cbase + (size_t)t * 64 + 4 * cq) = v4;
    }
    bar_group(group);
    const float4 q4 = *reinterpret_cast<const float4 *>(qs + 4 * sub);
    const float scale = 0.125f;               // 1 / sqrt(64)

    // 2. scores: half-warp per key, 16 keys of every stage per warp
    for (int st = 0; st < nK; st++) {
        const uint32_t gs = gi + (uint32_t)(st * nact + ai), slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
        c.mbar_wait_b(full0 + slot * 8, parity, 17);
        const uint8_t *base = c.smem + slot * PS_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int jl = wg * 16 + i * 2 + hf, j = st * 64 + jl;
            float s = 0.f;
            if (j < t) {
                const float4 k4 = *reinterpret_cast<const float4 *>(base + jl * 256 + sub * 16);
                s = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (sub == 0 && j < t) sc[j] = s * scale;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + slot * 8);
    }
    if (wg == 0) {                             // this step's own key
        const float4 k4 = *reinterpret_cast<const float4 *>(kn + 4 * sub);
        float s = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sc[t] = s * scale;
    }
    bar_group(group);

    // 3. softmax over keys 0..t
    const int nk = t + 1;
    float mx = -INFINITY;
    for (int j = tid128; j < nk; j += 128) mx = fmaxf(mx, sc[j]);
    mx = warp_max(mx);
    if (lane == 0) redv[wg] = mx;
    bar_group(group);
    mx = fmaxf(fmaxf(redv[0], redv[1]), fmaxf(redv[2], redv[3]));
    float sum = 0.f;
    for (int j = tid128; j < nk; j += 128) {
        const float e = expf(sc[j] - mx);
        sc[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) redv[4 + wg] = sum;
    bar_group(group);
    sum = (redv[4] + redv[5]) + (redv[6] + redv[7]);
    const float inv = 1.0f / sum;

    // 4. P.V
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int st = 0; st < nK; st++) {
        const uint32_t gs = gi + (uint32_t)((nK + st) * nact + ai), slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
        c.mbar_wait_b(full0 + slot * 8, parity, 18);
        const uint8_t *base = c.smem + slot * PS_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int jl = wg * 16 + i * 2 + hf, j = st * 64 + jl;
            if (j < t) {
                const float4 v4 = *reinterpret_cast<const float4 *>(base + jl * 256 + sub * 16);
                const float p = sc[j];
                acc.x += p * v4.x; acc.y += p * v4.y; acc.z += p * v4.z; acc.w += p * v4.w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + slot * 8);
    }
    if (wg == 0 && hf == 0) {
        const float4 v4 = *reinterpret_cast<const float4 *>(vn + 4 * sub);
        const float p = sc[t];
        acc.x += p * v4.x; acc.y += p * v4.y; acc.z += p * v4.z; acc.w += p * v4.w;
    }
    *reinterpret_cast<float4 *>(part + (wg * 2 + hf) * 64 + 4 * sub) = acc;
    bar_group(group);
    if (tid128 < 16) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float4 p = *reinterpret_cast<const float4 *>(part + i * 64 + 4 * tid128);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        o.x *= inv; o.y *= inv; o.z *= inv; o.w *= inv;
        ll_store4(yout + 4 * tid128, o, out_flag);
    }
}

// ---------------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__((NG * 4 + 1) * 32, 1) pstep_kernel(const __grid_constant__ PsArgs a) {
    constexpr int NW = NG * 4, NT = NW * 32;
    extern __shared__ __align__(1024) uint8_t ps_smem[];
    const int tid = threadIdx.x, cta = blockIdx.x;
    const int t = *a.step;
    if (t >= a.T) return;
    const PsProg &pg = a.prog[cta];
    Ctx c{a, ps_smem, reinterpret_cast<int *>(ps_smem + PS_OFF_DEAD), nullptr};
    const uint32_t full0 = smem_u32(ps_smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;
    if (tid == 0) {
        for (int s = 0; s < PS_NS; s++) { mbar_init(full0 + s * 8, 1); mbar_init(empty0 + s * 8, 4); }
        *c.s_dead = 0;
        mbar_fence_init();
    }
    __syncthreads();
    const int n_batches = (pg.n_attn + NG - 1) / NG;

    if (tid < 32) {
        // =========================== producer: one thread streams this CTA's weights and K/V ===========================
        if (tid != 0) return;
        uint32_t gi = 0;
        auto issue = [&](const void *src, uint32_t bytes) {
            const uint32_t slot = gi % PS_NS;
            if (gi >= PS_NS) c.mbar_wait_b(empty0 + slot * 8, ((gi / PS_NS) - 1u) & 1u, 20);
            mbar_arrive_expect_tx(full0 + slot * 8, bytes);
            bulk_load_hint(smem_u32(ps_smem + slot * PS_STAGE_BYTES), src, bytes, full0 + slot * 8, L2_EVICT_FIRST);
            gi++;
        };
        auto issue_items = [&](int phase, const uint8_t *base) {
            for (int i = 0; i < pg.n_items[phase]; i++) {
                const PsItem it = pg.items[pg.first[phase] + i];
                const uint8_t *src = base + (size_t)it.w_off16 * 16;
                for (int s = 0; s < it.nst; s++) issue(src + (size_t)s * PS_STAGE_BYTES, PS_STAGE_BYTES);
            }
        };
        const int nK = (t + 63) >> 6;
        for (int l = 0; l < a.L; l++) {
            const uint8_t *wl = a.wpack + (size_t)l * a.layer_bytes;
            issue_items(PH_QKV, wl + (size_t)a.ph_off16[PH_QKV] * 16);
            for (int bt = 0; bt < n_batches; bt++) {
                for (int kv = 0; kv < 2; kv++)
                    for (int st = 0; st < nK; st++)
                        for (int i = 0; i < NG; i++) {
                            const int r = bt * NG + i;
                            if (r >= pg.n_attn) break;
                            const int item = pg.attn[r], h = item >> 4, b = item & 15;
                            if (b >= a.B) continue;
                            const size_t cbase = (((size_t)l * 16 + b) * a.H + h) * (size_t)a.T * 64 + (size_t)st * 64 * 64;
                            const int rows = min(64, t - st * 64);
                            issue((kv == 0 ? a.kcache : a.vcache) + cbase, (uint32_t)rows * 256u);
                        }
            }
            issue_items(PH_PROJ, wl + (size_t)a.ph_off16[PH_PROJ] * 16);
            issue_items(PH_FC1, wl + (size_t)a.ph_off16[PH_FC1] * 16);
            issue_items(PH_FC2, wl + (size_t)a.ph_off16[PH_FC2] * 16);
        }
        issue_items(PH_HEAD, a.head_pack);
        return;
    }

    // ================================================ consumers ================================================
    const int ctid = tid - 32;
    unsigned long long *tr = a.trace ? a.trace + (size_t)cta * PS_TRACE_EV : nullptr;
    if (tr && ctid == 0) tr[0] = timer_ns();
    // embedding: x = tok_emb[id] + pos_emb[t] for the tiles of this CTA (mingpt.py:186-200), rows >= B are zero
    for (int tile = cta; tile < a.d / 64; tile += gridDim.x) {
        if (ctid < 256) {
            const int m = ctid >> 4, nn = (ctid & 15) * 4, n = tile * 64 + nn;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < a.B) {
                long long id = a.seq[(size_t)m * a.seq_ld + t];
                if (id < 0 || id >= a.V) id = 0;
                const float4 e = __ldg(reinterpret_cast<const float4 *>(a.tok_emb + (size_t)id * a.d + n));
                const float4 p = __ldg(reinterpret_cast<const float4 *>(a.pos_emb + (size_t)t * a.d + n));
                v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
            }
            PhaseRt rt{};
            rt.epi = 0; rt.out = a.xa; rt.ldo = a.d; rt.out_flag = mkflag(t, 1); rt.stats_out = a.sta;
            tile_epilogue(c, rt, tile, m, nn, v, ctid);
        }
    }
    uint32_t gi = 0;
    // layers 0..L-1 run phases qkv [attention] proj fc1 fc2; "layer" L is the head.  One call site for the GEMM item
    // and one for the attention batch keep the kernel small (every phase shares the same code).
#pragma unroll 1
    for (int l = 0; l <= a.L; l++) {
        const int ph0 = l < a.L ? PH_QKV : PH_HEAD, ph1 = l < a.L ? PH_FC2 : PH_HEAD;
        c.tr_item = (tr && l == a.L / 2) ? tr + 300 : nullptr;
#pragma unroll 1
        for (int ph = ph0; ph <= ph1; ph++) {
#pragma unroll 1
            for (int i = 0; i < pg.n_items[ph]; i++) {
                const PsItem it = pg.items[pg.first[ph] + i];
                gemm_item<NG>(c, ph, l, t, it, gi, ctid);
                gi += it.nst;
            }
            if (ph == PH_QKV) {
                if (c.tr_item && ctid == 0) c.tr_item[3] = timer_ns();
#pragma unroll 1
                for (int bt = 0; bt < n_batches; bt++) {
                    attn_batch<NG>(c, pg, bt, l, t, gi, ctid);
                    gi += attn_stage_count<NG>(a, pg, bt, t);
                }
            }
            if (tr && ctid == 0 && l < 48) tr[1 + l * 5 + ph] = timer_ns();
        }
    }
    if (tr && ctid == 0) tr[1 + 48 * 5] = timer_ns();
}

// Re-tiles a row-major weight W[N][K] into ring stages: dst[tile][kstage][warp 4][n8 tile 8][lane 32][4 floats] with
// lane (g = lane / 4, tq = lane % 4) holding W[tile*64 + 8 j + g][kstage*64 + 16 warp + 4 tq .. + 3].
__global__ void pack_weight_kernel(const float *__restrict__ W, int N, int K, float4 *__restrict__ dst) {
    const size_t total = (size_t)N * K / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int lane = (int)(i & 31), j = (int)((i >> 5) & 7), w = (int)((i >> 8) & 3);
        const size_t stage = i >> 10;
        const int KSt = K / 64;
        const int tile = (int)(stage / KSt), ks = (int)(stage % KSt);
        const int g = lane >> 2, tq = lane & 3;
        dst[i] = *reinterpret_cast<const float4 *>(W + (size_t)(tile * 64 + 8 * j + g) * K + ks * 64 + 16 * w + 4 * tq);
    }
}

}  // namespace ps
}  // namespace wmar
