// Persistent decode-step kernel for the Taming minGPT engine: ONE launch runs a whole token step
// (embedding -> L x [LN1+QKV, attention over the KV cache, proj+residual, LN2+fc1+GELU, fc2+residual] -> LN_f + head)
// replacing the 244 dependent launches per token of the per-GEMM graph (mingpt.py:183-214, 98-122, 42-95).
//
// Design (B200 first; the plan is in pstep_plan.h):
//   * one CTA per SM (grid G <= #SMs, all co-resident), 1 producer warp + 16 consumer warps;
//   * the step is HBM-bound on the weights (5.5 GB / token) and on the K/V cache.  Neither depends on the activations,
//     so ONE elected producer thread per CTA walks the CTA's packed weight stream (contiguous per layer, re-tiled at
//     create time) and the K/V rows of earlier tokens with cp.async.bulk (TMA bulk copies, 16 KB each) into a 7-deep
//     mbarrier ring, and asks L2 for the bytes a few hundred KB further down the stream (cp.async.bulk.prefetch.L2):
//     HBM keeps streaming while the consumers sit in a dependency, the ring then only has to cover the L2 latency.
//     At the end of a step the prefetch cursor has wrapped into layer 0 of the next token step;
//   * every GEMM but fc2 is split along N only (a CTA owns n16 tiles with the full K): no split-K partials cross CTAs.
//     A 16 KB stage is 16 units of n16 x k16 (one per consumer warp, bytes already in mma.m16n8k8 B-fragment order,
//     read back with two conflict-free LDS.128 per lane); the 16 batch rows are the M of the MMA, products are 3xTF32
//     with fp32 accumulation (fp32-faithful: the reference's Linear layers run with TF32 off), the 16 warps split K and
//     are summed in warp order through shared memory (deterministic);
//   * fc2 is split along K: the CTA keeps its 16 f columns of gelu(fc1) in shared memory and multiplies them with the
//     matching columns of W2; the [16][d] partials are reduce-scattered through L2 in CTA order, fused with bias +
//     residual.  fc1 -> fc2 needs no exchange and the 4d-wide hidden activations never leave the SM;
//   * five exchanges per layer (x -> qkv -> y -> xb -> fc2 partials -> x), each a release/acquire epoch flag per CTA:
//     writers store plain fp32, bar.sync, one thread fences and stores the flag; one warp of every reader polls the G
//     flags, fences, bar.sync, then all threads load through L2 (ld.global.cg).  LayerNorm is recomputed by every
//     consumer CTA from the full row (two-pass, exactly like the reference) -- no statistics exchange;
//   * attention: (head, row) items on 128-thread warp groups, up to 4 at a time per CTA; K then V of the cached tokens
//     arrive through the same ring; two-pass softmax in shared memory.
// All waits are bounded: a wait that exceeds ~1 s raises the device error flag (bit 2) and the kernel drains.
#pragma once
#include "gemm.cuh"
#include "pstep_plan.h"
#include "tc05.cuh"

namespace wmar {
namespace ps {

using namespace tc05;

constexpr int PS_NW = 16;                       // consumer warps
constexpr int PS_NT = PS_NW * 32;               // consumer threads
constexpr int PS_THREADS = PS_NT + 32;          // + the producer warp
constexpr int PS_NS = 7;                        // ring depth (stages)
constexpr int PS_DMAX = 1536;                   // widest residual stream whose activations fit beside the ring
constexpr int PS_XPAD = 16;                     // row stride of X = Kp + 16 floats (== 16 mod 32: conflict-free LDS.128)
constexpr int PS_HLD = 16 * PS_MAX_F + 16;      // row stride of the gelu(fc1) slice
constexpr int PS_RING_BYTES = PS_NS * PS_STAGE_BYTES;
constexpr int PS_OFF_X = PS_RING_BYTES;
constexpr int PS_X_BYTES = 16 * (PS_DMAX + PS_XPAD) * 4;
constexpr int PS_OFF_H = PS_OFF_X + PS_X_BYTES;
constexpr int PS_H_BYTES = 16 * PS_HLD * 4;
constexpr int PS_OFF_BARS = PS_OFF_H + PS_H_BYTES;          // full[NS], empty[NS]
constexpr int PS_OFF_DEAD = PS_OFF_BARS + 2 * PS_NS * 8;
constexpr int PS_SMEM_BYTES = PS_OFF_DEAD + 16;
constexpr int PS_ATT_SCRATCH = 8192;            // per warp group: scores[1024], q/k/v[192], part[8][64], red[8]
constexpr unsigned PS_SPIN_LIMIT = 1u << 23;    // flag polls (~100 ns apart)
constexpr unsigned PS_MBAR_LIMIT = 1u << 16;    // try_wait suspends up to 20 us each
constexpr int PS_TRACE_EV = 1024;                // probe only: 1 + 48 layers x 20 events + head
static_assert(PS_SMEM_BYTES <= 232448, "shared memory budget");
static_assert(4 * PS_ATT_SCRATCH <= PS_X_BYTES && PS_PASS_TILES * 16 * 1024 <= PS_X_BYTES, "scratch aliases X");

enum { FL_X = 0, FL_QKV = 1, FL_ATT = 2, FL_XB = 3, FL_P = 4, FL_N = 5 };

struct PsLayer {
    const float *ln1_g, *ln1_b, *bqkv, *bproj, *ln2_g, *ln2_b, *b1, *b2;
};

struct PsArgs {
    const PsProg *prog;          // [G]
    const PsLayer *layers;       // [L]
    const uint8_t *wpack;        // [L][layer_bytes]: per CTA qkv | proj | fc1 | fc2 stages
    unsigned long long layer_bytes;
    const uint8_t *head_pack;
    const float *tok_emb, *pos_emb, *lnf_g, *lnf_b;
    int d, H, V, L, T, B, G, GP, Kp, KC, NBn;
    unsigned pf_dist;            // L2 prefetch distance of the weight stream, bytes per CTA
    unsigned wait_hint_ns;       // mbarrier.try_wait suspend-time hint (0 = plain try_wait)
    int dbg;                     // probe only (WMAR_PSTEP_DBG): bit 0 = skip the MMAs, bit 1 = skip the flag waits (results are garbage)
    const int *step;
    const int64_t *seq; int seq_ld;
    float *x, *xb, *y, *qkv;     // plain fp32 activations [16][d], [16][d], [16][d], [16][3d]
    float *part;                 // fc2 partials [G][16][d]
    unsigned *flags;             // [FL_N][GP] epoch flags
    float *kcache, *vcache;
    float *logits;               // plain fp32 [16][V]
    int *abort_flag;             // global: non-zero = a wait timed out somewhere, drain
    int *err;                    // device error flag (bit 2 = pstep timeout)
    unsigned long long *trace;   // probe only: [G][PS_TRACE_EV] globaltimer stamps
};

// ---------------------------------------------------------------------------------------------------------
// barrier.sync without .aligned: the threads of a warp need not arrive convergently (a lane may still be in a poll loop
// or a trace stamp when its warp mates reach the barrier)
__device__ __forceinline__ void bar_consumers() { __syncwarp(); asm volatile("barrier.sync 8, %0;" ::"n"(PS_NT) : "memory"); }
__device__ __forceinline__ void bar_group(int group) { __syncwarp(); asm volatile("barrier.sync %0, 128;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cg4(const float *p) {     // L2 only: the line may have been written by another SM
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
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

struct Ctx {
    const PsArgs &a;
    uint8_t *smem;
    int *s_dead;
    unsigned long long *tr;        // probe only
    int cta;
    __device__ __forceinline__ bool dead() const { return *reinterpret_cast<volatile int *>(s_dead) != 0; }
    __device__ __forceinline__ void timeout(int code) const {
        atomicExch(a.abort_flag, code);
        atomicOr(a.err, 4 | (code << 8));
        *reinterpret_cast<volatile int *>(s_dead) = 1;
    }
    __device__ __forceinline__ bool poll_abort(unsigned n, unsigned limit, int code) const {
        if ((n & 63u) == 0u) {
            if (dead()) return true;
            if (*reinterpret_cast<volatile int *>(a.abort_flag) != 0) { *reinterpret_cast<volatile int *>(s_dead) = 1; return true; }
            if (n > limit) { timeout(code); return true; }
        }
        return false;
    }
    __device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) const {
        uint32_t ok;
        if (a.wait_hint_ns != 0u)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(parity), "r"(a.wait_hint_ns) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        return ok != 0;
    }
    __device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity, int code) const {
        if (try_wait(bar, parity)) return;
        unsigned n = 0;
        while (!try_wait(bar, parity)) {
            ++n;
            if (dead()) return;
            if ((n & 15u) == 0u && *reinterpret_cast<volatile int *>(a.abort_flag) != 0) { *reinterpret_cast<volatile int *>(s_dead) = 1; return; }
            if (n > (a.wait_hint_ns != 0u ? PS_MBAR_LIMIT : (PS_MBAR_LIMIT << 8))) { timeout(code); return; }
        }
    }
    // probe only; tr is set for the whole of consumer warp 0, lane 0 stores, the warp reconverges
    __device__ __forceinline__ void stamp(int ev) const {
        if (tr) {
            if ((threadIdx.x & 31) == 0) tr[ev] = timer_ns();
            __syncwarp();
        }
    }
};

// Release: every consumer thread's global stores of this phase are ordered before the CTA's epoch flag.
__device__ __forceinline__ void signal_flag(const Ctx &c, int which, unsigned epoch, int ctid) {
    bar_consumers();
    if (ctid == 0) {
        __threadfence();
        st_relaxed_u32(c.a.flags + which * c.a.GP + c.cta, epoch);
    }
}
// Acquire: consumer warp 0 polls the G flags of exchange `which` until all carry `epoch` (flags only grow inside a
// generation), then the whole consumer side synchronises.  Returns with the data of all CTAs visible to ld.global.cg.
__device__ __forceinline__ void wait_flags(const Ctx &c, int which, unsigned epoch, int ctid, int code) {
    if (ctid < 32 && !(c.a.dbg & 2)) {
        const unsigned *f = c.a.flags + which * c.a.GP;
        unsigned spins = 0;
        while (true) {
            unsigned ok = 1u;
#pragma unroll 5
            for (int i = ctid; i < c.a.G; i += 32) ok &= (unsigned)(ld_relaxed_u32(f + i) >= epoch);
            if (__all_sync(0xffffffffu, ok != 0u)) break;
            if (c.poll_abort(++spins, PS_SPIN_LIMIT, code)) break;
            __nanosleep(40);
        }
        __threadfence();
    }
    bar_consumers();
}

// ---------------------------------------------------------------------------------------------------------
// X[16][Kp] (row stride Kp + 16) <- LayerNorm(src[16][d]) or src itself; warp = row, lane = 4 columns of every 128.
// mode 0: plain copy, 1: LayerNorm(g, b), 2: embedding tok_emb[id] + pos_emb[t] then LayerNorm (layer 0)
template <int MODE>
__device__ __forceinline__ void load_x(const Ctx &c, const float *src, const float *g, const float *b, float eps, int t, int ctid) {
    const PsArgs &a = c.a;
    const int row = ctid >> 5, lane = ctid & 31, d = a.d, ldx = a.Kp + PS_XPAD;
    float *X = reinterpret_cast<float *>(c.smem + PS_OFF_X) + row * ldx;
    constexpr int MAXI = PS_DMAX / 128;
    float4 v[MAXI];
    const float *tok = nullptr, *pos = nullptr;
    bool active = row < a.B;
    if (MODE == 2) {
        long long id = active ? a.seq[(size_t)row * a.seq_ld + t] : 0;
        if (id < 0 || id >= a.V) id = 0;
        tok = a.tok_emb + (size_t)id * d;
        pos = a.pos_emb + (size_t)t * d;
    }
#pragma unroll
    for (int i = 0; i < MAXI; i++) {
        const int col = i * 128 + lane * 4;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < d && active) {
            if (MODE == 2) {
                const float4 e = __ldg(reinterpret_cast<const float4 *>(tok + col));
                const float4 p = __ldg(reinterpret_cast<const float4 *>(pos + col));
                v[i] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
            } else {
                v[i] = ld_cg4(src + (size_t)row * d + col);
            }
        }
    }
    if (MODE != 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXI; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < MAXI; i++) {
            const int col = i * 128 + lane * 4;
            if (col < d) {
                const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
                q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
        for (int i = 0; i < MAXI; i++) {
            const int col = i * 128 + lane * 4;
            if (col < d) {
                const float4 g4 = __ldg(reinterpret_cast<const float4 *>(g + col));
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(b + col));
                v[i].x = (v[i].x - mean) * rstd * g4.x + b4.x; v[i].y = (v[i].y - mean) * rstd * g4.y + b4.y;
                v[i].z = (v[i].z - mean) * rstd * g4.z + b4.z; v[i].w = (v[i].w - mean) * rstd * g4.w + b4.w;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < MAXI; i++) {
        const int col = i * 128 + lane * 4;
        if (col < a.Kp) *reinterpret_cast<float4 *>(X + col) = (col < d) ? v[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// A fragments of one k16 step: rows g and g + 8, four consecutive k per lane (k slot tq <-> k 4tq, tq+4 <-> 4tq+1 in the
// first MMA, 4tq+2 / 4tq+3 in the second: any bijection works as long as the packed weights use the same one).
struct XFrag { uint32_t h[2][4], l[2][4]; };
__device__ __forceinline__ void make_xfrag(XFrag &f, const float4 xa, const float4 xb) {
    const float r0[4] = {xa.x, xa.y, xa.z, xa.w}, r1[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
    for (int e = 0; e < 4; e++) { split_tf32(r0[e], f.h[0][e], f.l[0][e]); split_tf32(r1[e], f.h[1][e], f.l[1][e]); }
}
// acc[u] (16 x 8, n8 tile u of the n16 tile) += X(16 x 16) . W(16 x 16)^T, 3xTF32
__device__ __forceinline__ void unit_mma(float (&acc)[2][4], const XFrag &x, const uint8_t *wunit, int lane) {
    const float4 w0 = *reinterpret_cast<const float4 *>(wunit + lane * 16);
    const float4 w1 = *reinterpret_cast<const float4 *>(wunit + 512 + lane * 16);
    const float wv[2][4] = {{w0.x, w0.y, w0.z, w0.w}, {w1.x, w1.y, w1.z, w1.w}};
#pragma unroll
    for (int u = 0; u < 2; u++) {
        uint32_t wh[4], wl[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            wh[e] = __float_as_uint(wv[u][e]) & 0xffffe000u;
            wl[e] = __float_as_uint(wv[u][e] - __uint_as_float(wh[e]));
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int e = 2 * half;
            mma_tf32(acc[u], x.l[0][e], x.l[1][e], x.l[0][e + 1], x.l[1][e + 1], wh[e], wh[e + 1]);
            mma_tf32(acc[u], x.h[0][e], x.h[1][e], x.h[0][e + 1], x.h[1][e + 1], wl[e], wl[e + 1]);
            mma_tf32(acc[u], x.h[0][e], x.h[1][e], x.h[0][e + 1], x.h[1][e + 1], wh[e], wh[e + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// One K-type GEMM phase of this CTA: Y[16][its n16 tiles] = X[16][Kp] . W^T (+ epilogue), in passes of <= 4 tiles.
// Stage order inside a pass: k chunk major, tile minor (the X fragments of a chunk are split once and reused).
// EPI 0: + bias -> qkv (global)   1: + bias + residual -> xb (global)   2: + bias, GELU -> gelu slice (shared)   3: -> logits
template <int EPI>
__device__ __forceinline__ void gemm_phase(const Ctx &c, const PsProg &pg, int ph, int l, int t, uint32_t &gi, int ctid,
                                           const float *bias, bool reload_ln, int trace_ev) {
    const PsArgs &a = c.a;
    const int lane = ctid & 31, cw = ctid >> 5, g = lane >> 2, tq = lane & 3;
    const int nt = pg.n_tiles[ph], ldx = a.Kp + PS_XPAD;
    const uint32_t full0 = smem_u32(c.smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;
    const float *X = reinterpret_cast<const float *>(c.smem + PS_OFF_X);
    float *red = reinterpret_cast<float *>(c.smem + PS_OFF_X);
    for (int t0 = 0; t0 < nt; t0 += PS_PASS_TILES) {
        const int np = min(PS_PASS_TILES, nt - t0);
        if (t0 > 0) {
            // the reduction scratch of the previous pass overwrote X: build it again (head only at full size)
            bar_consumers();
            if (reload_ln) load_x<1>(c, a.x, a.lnf_g, a.lnf_b, 1e-5f, t, ctid);
            bar_consumers();
        }
        float acc[PS_PASS_TILES][2][4];
#pragma unroll
        for (int i = 0; i < PS_PASS_TILES; i++)
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[i][u][e] = 0.f;
        for (int kc = 0; kc < a.KC; kc++) {
            XFrag xf;
            {
                const float *xp = X + g * ldx + kc * PS_KS + cw * 16 + 4 * tq;
                make_xfrag(xf, *reinterpret_cast<const float4 *>(xp), *reinterpret_cast<const float4 *>(xp + 8 * ldx));
            }
#pragma unroll
            for (int i = 0; i < PS_PASS_TILES; i++) {
                if (i < np) {
                    const uint32_t gs = gi + (uint32_t)(kc * np + i), slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
                    c.mbar_wait_b(full0 + slot * 8, parity, 14);
                    __syncwarp();          // mma.sync.aligned needs the warp converged after the per-lane wait loop
                    if (!(a.dbg & 1)) unit_mma(acc[i], xf, c.smem + slot * PS_STAGE_BYTES + cw * 1024, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty0 + slot * 8);
                }
            }
        }
        gi += (uint32_t)(a.KC * np);
        // cross-warp reduction in warp order; the scratch aliases X
        bar_consumers();
        if (trace_ev >= 0) c.stamp(trace_ev);
#pragma unroll
        for (int i = 0; i < PS_PASS_TILES; i++) {
            if (i < np) {
                float *my = red + ((i * PS_NW + cw) * 16) * 16;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    *reinterpret_cast<float2 *>(my + g * 16 + 8 * u + 2 * tq) = make_float2(acc[i][u][0], acc[i][u][1]);
                    *reinterpret_cast<float2 *>(my + (g + 8) * 16 + 8 * u + 2 * tq) = make_float2(acc[i][u][2], acc[i][u][3]);
                }
            }
        }
        bar_consumers();
        if (ctid < np * 64) {
            const int i = ctid >> 6, m = (ctid >> 2) & 15, n4 = (ctid & 3) * 4;
            const int tile = pg.tiles[pg.first[ph] + t0 + i], n = tile * 16 + n4;
            // global operands of the epilogue first: their latency overlaps the shared-memory reduction
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr) b4 = __ldg(reinterpret_cast<const float4 *>(bias + n));
            if (EPI == 1) {
                if (l == 0) {                 // layer 0: the residual is the embedding itself (mingpt.py:186-200)
                    if (m < a.B) {
                        long long id = a.seq[(size_t)m * a.seq_ld + t];
                        if (id < 0 || id >= a.V) id = 0;
                        const float4 e = __ldg(reinterpret_cast<const float4 *>(a.tok_emb + (size_t)id * a.d + n));
                        const float4 p = __ldg(reinterpret_cast<const float4 *>(a.pos_emb + (size_t)t * a.d + n));
                        r4 = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
                    }
                } else {
                    r4 = ld_cg4(a.x + (size_t)m * a.d + n);
                }
            }
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < PS_NW; w++) {
                const float4 p = *reinterpret_cast<const float4 *>(red + ((i * PS_NW + w) * 16 + m) * 16 + n4);
                v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
            }
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            if (EPI == 0) {
                *reinterpret_cast<float4 *>(a.qkv + (size_t)m * 3 * a.d + n) = v;
            } else if (EPI == 1) {
                v.x = r4.x + v.x; v.y = r4.y + v.y; v.z = r4.z + v.z; v.w = r4.w + v.w;
                *reinterpret_cast<float4 *>(a.xb + (size_t)m * a.d + n) = v;
            } else if (EPI == 2) {
                v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w);
                float *hb = reinterpret_cast<float *>(c.smem + PS_OFF_H);
                *reinterpret_cast<float4 *>(hb + m * PS_HLD + (t0 + i) * 16 + n4) = v;
            } else {
                if (m < a.B) *reinterpret_cast<float4 *>(a.logits + (size_t)m * a.V + n) = v;
            }
        }
        if (trace_ev >= 0) c.stamp(trace_ev + 1);
    }
}

// fc2 of this CTA: part[cta][16][d] = gelu slice [16][16 f] . W2[:, its 16 f columns]^T.  Stage (nb, j): warp w owns
// n16 tile nb * 16 + w, k16 step j; a tile is complete after f stages and goes straight from registers to L2.
__device__ __forceinline__ void fc2_phase(const Ctx &c, const PsProg &pg, uint32_t &gi, int ctid) {
    const PsArgs &a = c.a;
    const int lane = ctid & 31, cw = ctid >> 5, g = lane >> 2, tq = lane & 3;
    const int f = pg.n_tiles[PH_FC1];
    if (f == 0) return;
    const uint32_t full0 = smem_u32(c.smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;
    const float *hb = reinterpret_cast<const float *>(c.smem + PS_OFF_H);
    float *out = a.part + (size_t)c.cta * 16 * a.d;
    for (int nb = 0; nb < a.NBn; nb++) {
        float acc[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[u][e] = 0.f;
        for (int j = 0; j < f; j++) {
            XFrag xf;
            const float *xp = hb + g * PS_HLD + j * 16 + 4 * tq;
            make_xfrag(xf, *reinterpret_cast<const float4 *>(xp), *reinterpret_cast<const float4 *>(xp + 8 * PS_HLD));
            const uint32_t gs = gi + (uint32_t)(nb * f + j), slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
            c.mbar_wait_b(full0 + slot * 8, parity, 21);
            __syncwarp();
            if (!(a.dbg & 1)) unit_mma(acc, xf, c.smem + slot * PS_STAGE_BYTES + cw * 1024, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + slot * 8);
        }
        const int n0 = nb * PS_NB + cw * 16;
        if (n0 < a.d) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                *reinterpret_cast<float2 *>(out + (size_t)g * a.d + n0 + 8 * u + 2 * tq) = make_float2(acc[u][0], acc[u][1]);
                *reinterpret_cast<float2 *>(out + (size_t)(g + 8) * a.d + n0 + 8 * u + 2 * tq) = make_float2(acc[u][2], acc[u][3]);
            }
        }
    }
    gi += (uint32_t)(a.NBn * f);
}

// x[slice] = xb[slice] + b2 + sum over CTAs (fixed order) of their fc2 partials: the reduce-scatter half of fc2.
// thread = (element e, group q): group q sums the partials of CTAs q, q + NQ, ... (all its loads in flight at once),
// the groups are then added in q order through shared memory.
constexpr int PS_RED_NQ = 12, PS_RED_PER = PS_NT / PS_RED_NQ, PS_RED_MAXC = 16;   // up to 12 x 16 = 192 CTAs
__device__ __forceinline__ void reduce_phase(const Ctx &c, const PsProg &pg, const float *b2, int ctid) {
    const PsArgs &a = c.a;
    const int n = pg.red_hi - pg.red_lo;             // float4 elements of the flattened [16][d]
    float4 *scr = reinterpret_cast<float4 *>(c.smem + PS_OFF_X);
    for (int e0 = 0; e0 < n; e0 += PS_RED_PER) {
        const int e = e0 + ctid % PS_RED_PER, q = ctid / PS_RED_PER;
        const bool mine = e < n && q < PS_RED_NQ;
        const size_t idx = (size_t)(pg.red_lo + (mine ? e : 0)) * 4;
        float4 xb4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = xb4;
        if (mine && q == 0) {
            xb4 = ld_cg4(a.xb + idx);
            b4 = __ldg(reinterpret_cast<const float4 *>(b2 + (int)(idx % (size_t)a.d)));
        }
        if (mine) {
            // CTAs without fc1 tiles never write their partial: it stays zero from create time
            float4 p[PS_RED_MAXC];
#pragma unroll
            for (int k = 0; k < PS_RED_MAXC; k++) {
                const int cc = q + k * PS_RED_NQ;
                p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cc < a.G) p[k] = ld_cg4(a.part + (size_t)cc * 16 * a.d + idx);
            }
            float4 s = p[0];
#pragma unroll
            for (int k = 1; k < PS_RED_MAXC; k++) { s.x += p[k].x; s.y += p[k].y; s.z += p[k].z; s.w += p[k].w; }
            scr[q * PS_RED_PER + (e - e0)] = s;
        }
        bar_consumers();
        if (mine && q == 0) {
            float4 sum = scr[e - e0];
#pragma unroll
            for (int qq = 1; qq < PS_RED_NQ; qq++) {
                const float4 pp = scr[qq * PS_RED_PER + (e - e0)];
                sum.x += pp.x; sum.y += pp.y; sum.z += pp.z; sum.w += pp.w;
            }
            float4 v;
            v.x = xb4.x + (sum.x + b4.x); v.y = xb4.y + (sum.y + b4.y); v.z = xb4.z + (sum.z + b4.z); v.w = xb4.w + (sum.w + b4.w);
            *reinterpret_cast<float4 *>(a.x + idx) = v;
        }
        bar_consumers();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Attention batch: up to 4 (head,row) items of this CTA, one per warp group.  K stages then V stages of the cached
// tokens come through the ring (order: K/V major, stage, item minor; rows >= B are skipped by producer and consumers).
__device__ __forceinline__ uint32_t attn_stage_count(const PsArgs &a, const PsProg &pg, int batch, int t) {
    int nact = 0;
    for (int i = 0; i < 4; i++) {
        const int r = batch * 4 + i;
        if (r < pg.n_attn && (pg.attn[r] & 15) < a.B) nact++;
    }
    return (uint32_t)(2 * ((t + 63) >> 6) * nact);
}

__device__ __forceinline__ void attn_batch(const Ctx &c, const PsProg &pg, int batch, int l, int t, uint32_t gi, int ctid) {
    const PsArgs &a = c.a;
    const int lane = ctid & 31, cw = ctid >> 5, group = cw >> 2, wg = cw & 3;
    const int tid128 = ctid & 127, sub = lane & 15, hf = lane >> 4;
    const uint32_t full0 = smem_u32(c.smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;
    const int r = batch * 4 + group;
    const int item = r < pg.n_attn ? pg.attn[r] : 0, h = item >> 4, b = item & 15;
    // inactive groups (no item, or a row >= B whose y keeps its finite old contents) still take part in the ring
    // protocol below: EVERY consumer warp waits for and releases EVERY stage, in ring order.  (If only the owning group
    // touched a stage, a group could wait for ring index s while the previous fill of that slot, index s - 7 of another
    // group, is still in flight; mbarrier.try_wait.parity only sees the phase bit and would pass on the wrong data.)
    const bool mine = r < pg.n_attn && b < a.B;
    const int d = a.d;
    int nact = 0, ai = 0;
    for (int i = 0; i < 4; i++) {
        const int rr = batch * 4 + i;
        if (rr < pg.n_attn && (pg.attn[rr] & 15) < a.B) { if (i < group) ai++; nact++; }
    }
    if (nact == 0) return;
    const int nK = (t + 63) >> 6;
    float *scr = reinterpret_cast<float *>(c.smem + PS_OFF_X + group * PS_ATT_SCRATCH);
    float *sc = scr, *qs = scr + 1024, *kn = scr + 1088, *vn = scr + 1152, *part = scr + 1216, *redv = scr + 1728;
    const size_t cbase = (((size_t)l * 16 + b) * a.H + h) * (size_t)a.T * 64;

    // 1. this step's q, k, v (written by the qkv tiles' owners); k, v are appended to the cache for later steps
    float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mine) {
        if (tid128 < 48) {
            const int which = tid128 >> 4, cq = tid128 & 15;
            const float4 v4 = ld_cg4(a.qkv + (size_t)b * 3 * d + which * d + h * 64 + 4 * cq);
            *reinterpret_cast<float4 *>(scr + 1024 + which * 64 + 4 * cq) = v4;
            if (which == 1) *reinterpret_cast<float4 *>(a.kcache + cbase + (size_t)t * 64 + 4 * cq) = v4;
            if (which == 2) *reinterpret_cast<float4 *>(a.vcache + cbase + (size_t)t * 64 + 4 * cq) = v4;
        }
        bar_group(group);
        q4 = *reinterpret_cast<const float4 *>(qs + 4 * sub);
    }
    const float scale = 0.125f;               // 1 / sqrt(64)

    // 2. scores: half-warp per key, 16 keys of every stage per warp
    for (int st = 0; st < nK; st++) {
        for (int j = 0; j < nact; j++) {
            const uint32_t gs = gi + (uint32_t)(st * nact + j), slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
            c.mbar_wait_b(full0 + slot * 8, parity, 17);
            __syncwarp();
            if (mine && j == ai) {
                const uint8_t *base = c.smem + slot * PS_STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int jl = wg * 16 + i * 2 + hf, jj = st * 64 + jl;
                    float sv = 0.f;
                    if (jj < t) {
                        const float4 k4 = *reinterpret_cast<const float4 *>(base + jl * 256 + sub * 16);
                        sv = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
                    }
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
                    if (sub == 0 && jj < t) sc[jj] = sv * scale;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + slot * 8);
        }
    }
    float inv = 0.f;
    if (mine) {
        if (wg == 0) {                             // this step's own key
            const float4 k4 = *reinterpret_cast<const float4 *>(kn + 4 * sub);
            float sv = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
            if (lane == 0) sc[t] = sv * scale;
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
        inv = 1.0f / sum;
    }

    // 4. P.V
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int st = 0; st < nK; st++) {
        for (int j = 0; j < nact; j++) {
            const uint32_t gs = gi + (uint32_t)((nK + st) * nact + j), slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
            c.mbar_wait_b(full0 + slot * 8, parity, 18);
            __syncwarp();
            if (mine && j == ai) {
                const uint8_t *base = c.smem + slot * PS_STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int jl = wg * 16 + i * 2 + hf, jj = st * 64 + jl;
                    if (jj < t) {
                        const float4 v4 = *reinterpret_cast<const float4 *>(base + jl * 256 + sub * 16);
                        const float p = sc[jj];
                        acc.x += p * v4.x; acc.y += p * v4.y; acc.z += p * v4.z; acc.w += p * v4.w;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + slot * 8);
        }
    }
    if (!mine) return;
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
        *reinterpret_cast<float4 *>(a.y + (size_t)b * d + h * 64 + 4 * tid128) = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PS_THREADS, 1) pstep_kernel(const __grid_constant__ PsArgs a) {
    extern __shared__ __align__(1024) uint8_t ps_smem[];
    const int tid = threadIdx.x, cta = blockIdx.x;
    const int t = *a.step;
    if (t >= a.T) return;
    const PsProg &pg = a.prog[cta];
    Ctx c{a, ps_smem, reinterpret_cast<int *>(ps_smem + PS_OFF_DEAD), nullptr, cta};
    const uint32_t full0 = smem_u32(ps_smem + PS_OFF_BARS), empty0 = full0 + PS_NS * 8;
    if (tid == 0) {
        for (int s = 0; s < PS_NS; s++) { mbar_init(full0 + s * 8, 1); mbar_init(empty0 + s * 8, PS_NW); }
        *c.s_dead = 0;
        mbar_fence_init();
    }
    __syncthreads();
    const int n_batches = (pg.n_attn + 3) / 4;
    const unsigned ep0 = (unsigned)t * (unsigned)(a.L + 1) + 1u;      // epoch of (step t, layer 0)

    if (tid < 32) {
        // =========================== producer: one thread streams this CTA's weights and K/V ===========================
        if (tid != 0) return;
        uint32_t gi = 0;
        // L2 prefetch cursor: runs pf_dist bytes ahead of the loads along [layer 0 block | ... | layer L-1 block | head
        // block] and wraps into layer 0 (= the next token step) at the end
        int pf_l = 0;
        unsigned long long pf_off = 0;
        auto blk_base = [&](int li) -> const uint8_t * {
            return li < a.L ? a.wpack + (size_t)li * a.layer_bytes + (size_t)pg.layer_off16 * 16 : a.head_pack + (size_t)pg.head_off16 * 16;
        };
        auto blk_bytes = [&](int li) -> unsigned long long {
            return (unsigned long long)(li < a.L ? pg.layer_stages : pg.head_stages) * PS_STAGE_BYTES;
        };
        auto prefetch = [&](unsigned long long n) {
            int guard = 0;
            while (n > 0 && guard++ < 64) {
                const unsigned long long left = blk_bytes(pf_l) - pf_off;
                if (left == 0) { pf_l = (pf_l + 1) % (a.L + 1); pf_off = 0; continue; }
                const unsigned long long chunk = n < left ? n : left;
                prefetch_l2_bulk(blk_base(pf_l) + pf_off, (uint32_t)chunk);
                pf_off += chunk; n -= chunk;
            }
        };
        if (a.pf_dist > 0) for (unsigned long long p = 0; p < a.pf_dist; p += 65536) prefetch(65536);
        auto issue = [&](const void *src, uint32_t bytes) {
            const uint32_t slot = gi % PS_NS;
            if (gi >= PS_NS) c.mbar_wait_b(empty0 + slot * 8, ((gi / PS_NS) - 1u) & 1u, 20);
            if (c.dead()) return;
            mbar_arrive_expect_tx(full0 + slot * 8, bytes);
            bulk_load_hint(smem_u32(ps_smem + slot * PS_STAGE_BYTES), src, bytes, full0 + slot * 8, L2_EVICT_FIRST);
            gi++;
        };
        auto issue_weights = [&](const uint8_t *&cur, int n) {
            for (int s = 0; s < n && !c.dead(); s++) {
                issue(cur, PS_STAGE_BYTES);
                cur += PS_STAGE_BYTES;
                if (a.pf_dist > 0) prefetch(PS_STAGE_BYTES);
            }
        };
        const int nK = (t + 63) >> 6;
        for (int l = 0; l < a.L && !c.dead(); l++) {
            const uint8_t *cur = blk_base(l);
            issue_weights(cur, pg.n_tiles[PH_QKV] * a.KC);
            for (int bt = 0; bt < n_batches; bt++) {
                for (int kv = 0; kv < 2; kv++)
                    for (int st = 0; st < nK; st++)
                        for (int i = 0; i < 4; i++) {
                            const int r = bt * 4 + i;
                            if (r >= pg.n_attn) break;
                            const int item = pg.attn[r], h = item >> 4, b = item & 15;
                            if (b >= a.B) continue;
                            const size_t cbase = (((size_t)l * 16 + b) * a.H + h) * (size_t)a.T * 64 + (size_t)st * 64 * 64;
                            const int rows = min(64, t - st * 64);
                            issue((kv == 0 ? a.kcache : a.vcache) + cbase, (uint32_t)rows * 256u);
                        }
            }
            issue_weights(cur, pg.n_tiles[PH_PROJ] * a.KC);
            issue_weights(cur, pg.n_tiles[PH_FC1] * a.KC);
            issue_weights(cur, pg.n_tiles[PH_FC1] * a.NBn);
        }
        {
            const uint8_t *cur = blk_base(a.L);
            issue_weights(cur, pg.n_tiles[PH_HEAD] * a.KC);
        }
        // drain: every issued copy has landed before the CTA may retire (matters only when the consumers bailed out)
        if (c.dead()) {
            const uint32_t n_out = gi < PS_NS ? gi : PS_NS;
            for (uint32_t k = 0; k < n_out; k++) {
                const uint32_t gs = gi - 1 - k, slot = gs % PS_NS, parity = (gs / PS_NS) & 1u;
                unsigned n = 0;
                while (!mbar_try_wait(full0 + slot * 8, parity) && ++n < 64) {}
            }
        }
        return;
    }

    // ================================================ consumers ================================================
    const int ctid = tid - 32;
    c.tr = (a.trace && ctid < 32) ? a.trace + (size_t)cta * PS_TRACE_EV : nullptr;
    c.stamp(0);
    uint32_t gi = 0;
    // trace events of a layer (probe only): 0 x flags seen, 1 x loaded, 2 qkv loop done, 3 qkv stored, 4 qkv signalled,
    // 5 qkv flags seen, 6 attention done, 7 attention signalled, 8 y flags seen, 9 y loaded, 10 proj loop done,
    // 11 proj stored, 12 xb signalled, 13 xb flags seen, 14 xb loaded, 15 fc1 done, 16 fc2 done, 17 fc2 signalled,
    // 18 partial flags seen, 19 x reduced (+ signalled)
#pragma unroll 1
    for (int l = 0; l < a.L; l++) {
        const PsLayer &Ly = a.layers[l];
        const unsigned ep = ep0 + (unsigned)l;
        const int tb = (l < 48) ? 1 + l * 20 : 1000;      // trace slot base of this layer (layers >= 48 share a spill slot)
        // ---- x -> LN1 -> qkv
        if (pg.n_tiles[PH_QKV] > 0) {
            if (l == 0) { bar_consumers(); c.stamp(tb + 0); load_x<2>(c, nullptr, Ly.ln1_g, Ly.ln1_b, 1e-5f, t, ctid); }
            else { wait_flags(c, FL_X, ep, ctid, 31); c.stamp(tb + 0); load_x<1>(c, a.x, Ly.ln1_g, Ly.ln1_b, 1e-5f, t, ctid); }
            bar_consumers();
            c.stamp(tb + 1);
            gemm_phase<0>(c, pg, PH_QKV, l, t, gi, ctid, Ly.bqkv, false, tb + 2);
        }
        signal_flag(c, FL_QKV, ep, ctid);
        c.stamp(tb + 4);
        // ---- attention
        {
            bool any = false;
            for (int r = 0; r < pg.n_attn; r++) any = any || ((pg.attn[r] & 15) < a.B);
            if (any) {
                wait_flags(c, FL_QKV, ep, ctid, 32);
                c.stamp(tb + 5);
#pragma unroll 1
                for (int bt = 0; bt < n_batches; bt++) {
                    attn_batch(c, pg, bt, l, t, gi, ctid);
                    gi += attn_stage_count(a, pg, bt, t);
                    bar_consumers();
                }
                c.stamp(tb + 6);
            }
        }
        signal_flag(c, FL_ATT, ep, ctid);
        c.stamp(tb + 7);
        // ---- y -> proj + residual -> xb
        if (pg.n_tiles[PH_PROJ] > 0) {
            wait_flags(c, FL_ATT, ep, ctid, 33);
            c.stamp(tb + 8);
            load_x<0>(c, a.y, nullptr, nullptr, 0.f, t, ctid);
            bar_consumers();
            c.stamp(tb + 9);
            gemm_phase<1>(c, pg, PH_PROJ, l, t, gi, ctid, Ly.bproj, false, tb + 10);
        }
        signal_flag(c, FL_XB, ep, ctid);
        c.stamp(tb + 12);
        // ---- xb -> LN2 -> fc1 -> GELU -> fc2 partial
        if (pg.n_tiles[PH_FC1] > 0) {
            wait_flags(c, FL_XB, ep, ctid, 34);
            c.stamp(tb + 13);
            load_x<1>(c, a.xb, Ly.ln2_g, Ly.ln2_b, 1e-5f, t, ctid);
            bar_consumers();
            c.stamp(tb + 14);
            gemm_phase<2>(c, pg, PH_FC1, l, t, gi, ctid, Ly.b1, false, -1);
            bar_consumers();
            c.stamp(tb + 15);
            fc2_phase(c, pg, gi, ctid);
            c.stamp(tb + 16);
        }
        signal_flag(c, FL_P, ep, ctid);
        c.stamp(tb + 17);
        // ---- reduce-scatter of the fc2 partials + bias + residual -> x
        wait_flags(c, FL_P, ep, ctid, 35);
        c.stamp(tb + 18);
        reduce_phase(c, pg, Ly.b2, ctid);
        signal_flag(c, FL_X, ep + 1, ctid);
        c.stamp(tb + 19);
        if (c.dead()) break;
    }
    // ---- LN_f + head
    if (pg.n_tiles[PH_HEAD] > 0 && !c.dead()) {
        wait_flags(c, FL_X, ep0 + (unsigned)a.L, ctid, 36);
        load_x<1>(c, a.x, a.lnf_g, a.lnf_b, 1e-5f, t, ctid);
        bar_consumers();
        c.stamp(1 + 48 * 20);
        gemm_phase<3>(c, pg, PH_HEAD, a.L, t, gi, ctid, nullptr, true, -1);
    }
    c.stamp(2 + 48 * 20);
}

// Re-tiles the weights of one layer (or the head) into the ring stages of every CTA, in consumption order.
// grid = (G, stages per CTA rounded up), block = 256: one block per (CTA, stage).
// Unit w of a K-type stage = [u 2][lane 32][4 floats]: lane (g = lane / 4, tq = lane % 4) holds
// W[16 tile + 8 u + g][256 kc + 16 w + 4 tq .. + 3]; of an fc2-type stage: W2[256 nb + 16 w + 8 u + g][16 tile + 4 tq .. + 3].
struct PackArgs {
    const PsProg *prog;
    const float *w[3];           // qkv [3d][d], proj [d][d], fc1 [4d][d]  (or head [V][d] in slot 0)
    const float *w2;             // fc2 [d][4d]
    uint8_t *dst;                // packed layer (or head) base
    int d, KC, NBn, head, Nrows[3];
};
__global__ void pack_stage_kernel(const PackArgs p) {
    const PsProg &pg = p.prog[blockIdx.x];
    const int s = blockIdx.y;
    const int n_stages = p.head ? (int)pg.head_stages : (int)pg.layer_stages;
    if (s >= n_stages) return;
    // locate the stage (same walk as ps_stage_src)
    int ph = 0, tile = 0, kc = 0, nb = 0, rem = s;
    bool found = false;
    const int ph0 = p.head ? PH_HEAD : PH_QKV, ph1 = p.head ? PH_HEAD : PH_FC1;
    for (int q = ph0; q <= ph1 && !found; q++) {
        const int nt = pg.n_tiles[q], n = nt * p.KC;
        if (rem < n) {
            const int per_pass = PS_PASS_TILES * p.KC, pass = rem / per_pass, r2 = rem % per_pass;
            const int nt_pass = min(PS_PASS_TILES, nt - pass * PS_PASS_TILES);
            ph = q; kc = r2 / nt_pass; tile = pg.tiles[pg.first[q] + pass * PS_PASS_TILES + r2 % nt_pass];
            found = true;
        } else rem -= n;
    }
    if (!found) {
        const int f = pg.n_tiles[PH_FC1];
        ph = -1; nb = rem / f; tile = pg.tiles[pg.first[PH_FC1] + rem % f];
    }
    float4 *dst = reinterpret_cast<float4 *>(p.dst + ((size_t)(p.head ? pg.head_off16 : pg.layer_off16)) * 16 + (size_t)s * PS_STAGE_BYTES);
    for (int i = threadIdx.x; i < PS_STAGE_BYTES / 16; i += blockDim.x) {
        const int lane = i & 31, u = (i >> 5) & 1, w = i >> 6;
        const int g = lane >> 2, tq = lane & 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ph >= 0) {
            const int src = p.head ? 0 : ph;
            const int n = 16 * tile + 8 * u + g, k = PS_KS * kc + 16 * w + 4 * tq;
            if (n < p.Nrows[src] && k < p.d) v = *reinterpret_cast<const float4 *>(p.w[src] + (size_t)n * p.d + k);
        } else {
            const int n = PS_NB * nb + 16 * w + 8 * u + g, k = 16 * tile + 4 * tq;
            if (n < p.d) v = *reinterpret_cast<const float4 *>(p.w2 + (size_t)n * 4 * p.d + k);
        }
        dst[i] = v;
    }
}

}  // namespace ps
}  // namespace wmar
