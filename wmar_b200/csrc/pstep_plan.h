// Static work plan of the persistent decode-step kernel (pstep.cuh): which n16 output tiles of every GEMM, which
// (head, row) attention items and which slice of the residual stream each CTA owns, and where its packed weights live.
// Pure host logic (no CUDA calls) so it is unit-tested on the CPU through wmar_pstep_plan_debug (tests/test_pstep_plan.py).
//
// Decomposition (one CTA per SM, G CTAs):
//   * qkv / proj / fc1 / head are split along N only: a CTA owns whole n16 tiles with the FULL K, so no partial sums
//     ever cross CTAs inside these GEMMs (the per-GEMM graph path pays a split-K hand-off at every one of them);
//   * fc2 is split along K instead: the CTA that produced 16 f columns of gelu(fc1) keeps them in shared memory and
//     multiplies them with the matching 16 f columns of W2 -> a [16][d] partial per CTA, reduce-scattered through L2
//     (fixed order): fc1 -> fc2 needs no exchange at all and nobody ever gathers the 4d-wide hidden activations;
//   * tiles are dealt greedily to the least-loaded CTA, heaviest class first (an fc1 tile carries its fc2 share, so it
//     weighs twice a qkv / proj tile; an attention item about half): every CTA streams the same bytes per layer.
// Weights are re-tiled once into 16 KB ring stages, stored per (layer, CTA) in exactly the order the CTA consumes them,
// so its producer thread walks ONE contiguous stream per layer.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace wmar {
namespace ps {

constexpr int PS_STAGE_BYTES = 16384;   // one ring stage: 16 units (one per consumer warp) of n16 x k16 fp32
constexpr int PS_KS = 256;              // k per K-type stage (16 warps x k16)
constexpr int PS_NB = 256;              // n per fc2-type stage (16 warps x n16)
constexpr int PS_PASS_TILES = 4;        // n16 tiles whose accumulators a warp holds at once
constexpr int PS_MAX_TILES = 48;        // n16 tiles per CTA over qkv + proj + fc1 + head
constexpr int PS_MAX_F = 4;             // fc1 tiles per CTA (columns of gelu(fc1) kept in shared memory: 16 f <= 64)
constexpr int PS_MAX_ATTN = 16;         // attention items per CTA
enum { PH_QKV = 0, PH_PROJ = 1, PH_FC1 = 2, PH_HEAD = 3, PH_NK = 4 };

struct PsProg {
    int n_tiles[PH_NK];
    int first[PH_NK];
    uint16_t tiles[PS_MAX_TILES];   // n16 tile indices, per phase in ascending order
    int n_attn;
    uint16_t attn[PS_MAX_ATTN];     // head * 16 + row
    int red_lo, red_hi;             // float4 range [lo, hi) of the flattened [16][d] residual stream this CTA reduces
    uint32_t layer_off16;           // this CTA's block inside a packed layer, in 16-byte units: qkv | proj | fc1 | fc2
    uint32_t layer_stages;          // its length in 16 KB stages
    uint32_t head_off16;            // its block inside the packed head
    uint32_t head_stages;
};

struct PsPlan {
    int G, d, H, V;
    int Kp;                         // d rounded up to PS_KS (zero padded), KC = Kp / PS_KS chunks per tile
    int KC;
    int NBn;                        // fc2 n-blocks: ceil(d / PS_NB)
    long long layer_stages;         // 16 KB stages of one packed layer (all CTAs)
    long long head_stages;
    int max_load, min_load;         // greedy loads (half units) for the balance test
};

inline int ps_stages_of(const PsPlan &pl, const PsProg &p, int ph) { return p.n_tiles[ph] * pl.KC; }
inline int ps_fc2_stages_of(const PsPlan &pl, const PsProg &p) { return pl.NBn * p.n_tiles[PH_FC1]; }

// Returns 0 on success; negative when the model does not fit the static limits (the caller keeps the per-GEMM graph path).
inline int ps_make_plan(int G, int d, int H, int V, PsProg *progs, PsPlan *plan) {
    if (G < 1 || G > 1024 || d < 64 || d % 64 != 0 || V % 16 != 0 || H < 1 || d / H != 64 || d % H != 0) return -1;
    memset(progs, 0, sizeof(PsProg) * (size_t)G);
    PsPlan &pl = *plan;
    pl.G = G; pl.d = d; pl.H = H; pl.V = V;
    pl.Kp = (d + PS_KS - 1) / PS_KS * PS_KS;
    pl.KC = pl.Kp / PS_KS;
    pl.NBn = (d + PS_NB - 1) / PS_NB;
    const int NT[PH_NK] = {3 * d / 16, d / 16, 4 * d / 16, V / 16};
    std::vector<int> load((size_t)G, 0);
    std::vector<uint16_t> tmp_store((size_t)G * PH_NK * PS_MAX_TILES, 0);
    auto tmp_at = [&](int c, int ph, int i) -> uint16_t & { return tmp_store[((size_t)c * PH_NK + ph) * PS_MAX_TILES + i]; };
    auto least = [&]() {
        int best = 0;
        for (int c = 1; c < G; c++)
            if (load[c] < load[best]) best = c;
        return best;
    };
    // heaviest class first: fc1 (+ its fc2 share) = 4 half units, qkv = proj = 2, attention item = 1
    const int order[3] = {PH_FC1, PH_QKV, PH_PROJ};
    const int weight[3] = {4, 2, 2};
    for (int oi = 0; oi < 3; oi++) {
        const int ph = order[oi];
        for (int t = 0; t < NT[ph]; t++) {
            const int c = least();
            PsProg &p = progs[c];
            if (p.n_tiles[ph] >= PS_MAX_TILES) return -2;
            tmp_at(c, ph, p.n_tiles[ph]++) = (uint16_t)t;
            load[c] += weight[oi];
        }
    }
    for (int i = 0; i < 16 * H; i++) {
        const int c = least();
        PsProg &p = progs[c];
        if (p.n_attn >= PS_MAX_ATTN) return -3;
        p.attn[p.n_attn++] = (uint16_t)i;
        load[c] += 1;
    }
    pl.max_load = pl.min_load = load[0];
    for (int c = 1; c < G; c++) {
        if (load[c] > pl.max_load) pl.max_load = load[c];
        if (load[c] < pl.min_load) pl.min_load = load[c];
    }
    // the head runs alone at the end of the step: plain round robin
    for (int t = 0; t < NT[PH_HEAD]; t++) {
        PsProg &p = progs[t % G];
        if (p.n_tiles[PH_HEAD] >= PS_MAX_TILES) return -2;
        tmp_at(t % G, PH_HEAD, p.n_tiles[PH_HEAD]++) = (uint16_t)t;
    }
    const long long nf4 = 16ll * d / 4;
    long long layer_off = 0, head_off = 0;
    for (int c = 0; c < G; c++) {
        PsProg &p = progs[c];
        int cur = 0;
        for (int ph = 0; ph < PH_NK; ph++) {
            p.first[ph] = cur;
            if (cur + p.n_tiles[ph] > PS_MAX_TILES) return -2;
            for (int i = 0; i < p.n_tiles[ph]; i++) p.tiles[cur + i] = tmp_at(c, ph, i);
            cur += p.n_tiles[ph];
        }
        if (p.n_tiles[PH_FC1] > PS_MAX_F) return -4;
        // only the head may take several passes (the reduction scratch of a pass overwrites the activations)
        for (int ph = PH_QKV; ph <= PH_FC1; ph++)
            if (p.n_tiles[ph] > PS_PASS_TILES) return -6;
        p.red_lo = (int)(c * nf4 / G);
        p.red_hi = (int)((c + 1) * nf4 / G);
        const long long st = (long long)(p.n_tiles[PH_QKV] + p.n_tiles[PH_PROJ] + p.n_tiles[PH_FC1]) * pl.KC +
                             (long long)pl.NBn * p.n_tiles[PH_FC1];
        const long long hs = (long long)p.n_tiles[PH_HEAD] * pl.KC;
        if (layer_off * (PS_STAGE_BYTES / 16) > 0xffffffffll || head_off * (PS_STAGE_BYTES / 16) > 0xffffffffll) return -5;
        p.layer_off16 = (uint32_t)(layer_off * (PS_STAGE_BYTES / 16));
        p.layer_stages = (uint32_t)st;
        p.head_off16 = (uint32_t)(head_off * (PS_STAGE_BYTES / 16));
        p.head_stages = (uint32_t)hs;
        layer_off += st;
        head_off += hs;
    }
    pl.layer_stages = layer_off;
    pl.head_stages = head_off;
    return 0;
}

// Where stage `s` (0-based inside the CTA's layer block, or inside its head block when head != 0) comes from.
// K-type stage: 16 units, unit w = k16 step w of chunk kc of n16 tile `tile` of phase `ph`.
// fc2-type stage (ph == -1): unit w = n16 tile nb*16 + w, contracted with fc1 tile `tile` (k = 16 tile .. 16 tile + 15).
struct PsStageSrc { int ph, tile, kc, nb; };
inline PsStageSrc ps_stage_src(const PsPlan &pl, const PsProg &p, int s, int head) {
    PsStageSrc r{0, 0, 0, 0};
    if (head) {
        // passes of PS_PASS_TILES tiles; inside a pass: kc major, tile minor
        const int nt = p.n_tiles[PH_HEAD];
        const int per_pass = PS_PASS_TILES * pl.KC;
        const int pass = s / per_pass, rem = s % per_pass;
        const int nt_pass = (nt - pass * PS_PASS_TILES) < PS_PASS_TILES ? (nt - pass * PS_PASS_TILES) : PS_PASS_TILES;
        r.ph = PH_HEAD; r.kc = rem / nt_pass; r.tile = p.tiles[p.first[PH_HEAD] + pass * PS_PASS_TILES + rem % nt_pass];
        return r;
    }
    for (int ph = PH_QKV; ph <= PH_FC1; ph++) {
        const int nt = p.n_tiles[ph], n = nt * pl.KC;
        if (s < n) {
            const int per_pass = PS_PASS_TILES * pl.KC;
            const int pass = s / per_pass, rem = s % per_pass;
            const int nt_pass = (nt - pass * PS_PASS_TILES) < PS_PASS_TILES ? (nt - pass * PS_PASS_TILES) : PS_PASS_TILES;
            r.ph = ph; r.kc = rem / nt_pass; r.tile = p.tiles[p.first[ph] + pass * PS_PASS_TILES + rem % nt_pass];
            return r;
        }
        s -= n;
    }
    const int f = p.n_tiles[PH_FC1];
    r.ph = -1; r.nb = s / f; r.tile = p.tiles[p.first[PH_FC1] + s % f];
    return r;
}

}  // namespace ps
}  // namespace wmar
