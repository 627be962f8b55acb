// Static work plan of the persistent decode-step kernel (pstep.cuh): which (tile, k-range) items of every GEMM phase
// and which (head, row) attention items each CTA owns.  Pure host logic (no CUDA calls) so it is unit-tested on the CPU
// through wmar_pstep_plan_debug (tests/test_host_logic.py).
//
// A GEMM phase Y[16][N] = X[16][K] W[N][K]^T is a grid of units (n64 tile, k64 stage); one unit = one 16 KB ring stage
// of packed weights, stored in HBM in unit order [tile][kstage] (so the packed format does not depend on the plan).
// CTA c owns the contiguous unit range [c*U/G, (c+1)*U/G), cut at tile boundaries and at PS_MAX_ST stages into items.
// The items of a tile, in k order, are its split-K parts: every part but the last writes its partial into a slot
// (slots of a tile are consecutive), the owner of the last part sums them in k order and runs the epilogue.
#pragma once
#include <stdint.h>

#include <string.h>

namespace wmar {
namespace ps {

constexpr int PS_STAGE_BYTES = 16384;   // 64 W rows x 64 k x fp32
constexpr int PS_MAX_ST = 16;           // stages per item (X slice <= 1024 columns)
constexpr int PS_MAX_ITEMS = 40;        // GEMM items per CTA (all phases of a layer + head)
constexpr int PS_MAX_ATTN = 16;         // attention items per CTA
enum { PH_QKV = 0, PH_PROJ = 1, PH_FC1 = 2, PH_FC2 = 3, PH_HEAD = 4, PH_N = 5 };

struct PsItem {
    uint32_t w_off16;   // offset of the first stage inside the phase's packed block, in 16-byte units
    uint16_t tile;      // n64 tile
    uint16_t k0st;      // first k64 stage
    uint16_t nst;       // stages (1..PS_MAX_ST)
    uint16_t slot;      // own partial slot (non-reducer) / first foreign slot (reducer)
    uint16_t nparts;    // reducer: foreign partials to sum before its own
    uint8_t phase;
    uint8_t reducer;
};
static_assert(sizeof(PsItem) == 16, "PsItem layout");

struct PsProg {
    int n_items[PH_N];
    int first[PH_N];
    int n_attn;
    int pad_;
    PsItem items[PS_MAX_ITEMS];
    uint16_t attn[PS_MAX_ATTN];   // head * 16 + row
};

struct PsPlan {
    int G;
    int n_slots[PH_N];            // partial slots per phase
    long long units[PH_N];
};

// Returns 0 on success; negative when the model does not fit the static limits (caller falls back to the graph path).
inline int ps_make_plan(int G, int d, int H, int V, PsProg *progs, PsPlan *plan) {
    if (G < 1 || d % 64 != 0 || V % 64 != 0 || H < 1) return -1;
    const int N[PH_N] = {3 * d, d, 4 * d, d, V};
    const int K[PH_N] = {d, d, d, 4 * d, d};
    memset(progs, 0, sizeof(PsProg) * (size_t)G);
    plan->G = G;
    int cursor[4096];
    if (G > 4096) return -1;
    for (int c = 0; c < G; c++) cursor[c] = 0;
    for (int ph = 0; ph < PH_N; ph++) {
        const long long NT = N[ph] / 64, KSt = K[ph] / 64, U = NT * KSt;
        plan->units[ph] = U;
        if (NT > 65535 || KSt > 65535) return -1;
        int slots = 0;
        // first pass: items in global unit order; remember where the current tile's parts started
        struct Ref { int cta, idx; };
        Ref parts[4096];
        int n_parts = 0;
        int cur_tile = -1;
        auto close_tile = [&]() {
            if (n_parts == 0) return 0;
            // parts[0..n-2] get consecutive slots, parts[n-1] reduces
            const int first_slot = slots;
            for (int i = 0; i + 1 < n_parts; i++) {
                PsItem &it = progs[parts[i].cta].items[parts[i].idx];
                it.reducer = 0; it.slot = (uint16_t)slots++; it.nparts = 0;
            }
            PsItem &last = progs[parts[n_parts - 1].cta].items[parts[n_parts - 1].idx];
            last.reducer = 1; last.slot = (uint16_t)first_slot; last.nparts = (uint16_t)(n_parts - 1);
            n_parts = 0;
            return 0;
        };
        for (int c = 0; c < G; c++) {
            PsProg &p = progs[c];
            p.first[ph] = cursor[c];
            long long u = (long long)c * U / G;
            const long long u1 = (long long)(c + 1) * U / G;
            while (u < u1) {
                const int tile = (int)(u / KSt), ks = (int)(u % KSt);
                long long n = u1 - u;
                if (n > KSt - ks) n = KSt - ks;
                if (n > PS_MAX_ST) n = PS_MAX_ST;
                if (tile != cur_tile) { close_tile(); cur_tile = tile; }
                if (cursor[c] >= PS_MAX_ITEMS || n_parts >= 4096) return -2;
                PsItem &it = p.items[cursor[c]];
                it.w_off16 = (uint32_t)((u * PS_STAGE_BYTES) / 16);
                if ((u * PS_STAGE_BYTES) / 16 > 0xffffffffll) return -3;
                it.tile = (uint16_t)tile; it.k0st = (uint16_t)ks; it.nst = (uint16_t)n; it.phase = (uint8_t)ph;
                parts[n_parts++] = Ref{c, cursor[c]};
                cursor[c]++;
                p.n_items[ph]++;
                u += n;
            }
        }
        close_tile();
        if (slots > 65535) return -4;
        plan->n_slots[ph] = slots;
    }
    // attention items (head, row) round-robin over the CTAs
    const int n_att = H * 16;
    for (int i = 0; i < n_att; i++) {
        PsProg &p = progs[i % G];
        if (p.n_attn >= PS_MAX_ATTN) return -5;
        p.attn[p.n_attn++] = (uint16_t)i;
    }
    return 0;
}

}  // namespace ps
}  // namespace wmar
