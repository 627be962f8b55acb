"""Persistent decode-step kernel (csrc/pstep.cuh): invariants of its static work plan and of the packed weight format,
checked on the CPU through the C-ABI (wmar_pstep_plan_debug is host logic only, no CUDA call)."""
import ctypes

import numpy as np
import pytest

PS_MAX_ITEMS, PS_MAX_ATTN, PS_MAX_ST = 40, 16, 16


class PsItem(ctypes.Structure):
    _fields_ = [("w_off16", ctypes.c_uint32), ("tile", ctypes.c_uint16), ("k0st", ctypes.c_uint16),
                ("nst", ctypes.c_uint16), ("slot", ctypes.c_uint16), ("nparts", ctypes.c_uint16),
                ("phase", ctypes.c_uint8), ("reducer", ctypes.c_uint8)]


class PsProg(ctypes.Structure):
    _fields_ = [("n_items", ctypes.c_int * 5), ("first", ctypes.c_int * 5), ("n_attn", ctypes.c_int),
                ("pad_", ctypes.c_int), ("items", PsItem * PS_MAX_ITEMS), ("attn", ctypes.c_uint16 * PS_MAX_ATTN)]


def _plan(G, d, H, V):
    from wmar_b200 import _lib
    L = _lib.lib()
    assert L.wmar_pstep_prog_bytes() == ctypes.sizeof(PsProg)
    progs = (PsProg * G)()
    slots = (ctypes.c_int * 5)()
    _lib.check(L.wmar_pstep_plan_debug(G, d, H, V, ctypes.cast(progs, ctypes.c_void_p), slots))
    return progs, list(slots)


@pytest.mark.parametrize("G,d,H,V", [(144, 1536, 24, 16384), (148, 1536, 24, 16384), (144, 128, 2, 16384),
                                     (144, 384, 6, 16384), (132, 1024, 16, 1024), (7, 64, 1, 64)])
def test_plan_covers_every_unit_once_and_reduces_in_k_order(G, d, H, V):
    progs, slots = _plan(G, d, H, V)
    N = [3 * d, d, 4 * d, d, V]
    K = [d, d, d, 4 * d, d]
    for ph in range(5):
        NT, KSt = N[ph] // 64, K[ph] // 64
        cover = np.zeros((NT, KSt), dtype=np.int32)
        parts = {}                                  # tile -> [(k0, cta, order, item)]
        per_cta = []
        for c in range(G):
            p = progs[c]
            tot = 0
            for i in range(p.n_items[ph]):
                it = p.items[p.first[ph] + i]
                assert it.phase == ph and 1 <= it.nst <= PS_MAX_ST
                assert it.w_off16 == (it.tile * KSt + it.k0st) * 1024       # packed stream is in unit order
                cover[it.tile, it.k0st:it.k0st + it.nst] += 1
                parts.setdefault(it.tile, []).append((it.k0st, c, i, it))
                tot += it.nst
            per_cta.append(tot)
        assert (cover == 1).all(), f"phase {ph}: a (tile, k-stage) unit is missed or duplicated"
        assert max(per_cta) - min(per_cta) <= 1, f"phase {ph}: unbalanced {min(per_cta)}..{max(per_cta)}"
        n_slots = 0
        for tile, lst in parts.items():
            lst.sort(key=lambda x: x[0])
            *others, last = lst
            assert last[3].reducer == 1 and last[3].nparts == len(others)
            for j, o in enumerate(others):            # partials of a tile: consecutive slots, k order
                assert o[3].reducer == 0 and o[3].slot == last[3].slot + j
                # a reducer never waits on a CTA that could wait on it: parts come from lower-or-equal CTA ids
                assert (o[1], o[2]) < (last[1], last[2])
            n_slots += len(others)
        assert n_slots == slots[ph]
    seen = sorted(progs[c].attn[i] for c in range(G) for i in range(progs[c].n_attn))
    assert seen == list(range(16 * H))


def test_full_size_plan_is_the_exact_decomposition():
    """Taming C2 on 144 CTAs: every GEMM phase gives every CTA the same number of 16 KB stages (12 / 4 / 16 / 16)."""
    progs, slots = _plan(144, 1536, 24, 16384)
    for ph, want in enumerate([12, 4, 16, 16]):
        for c in range(144):
            p = progs[c]
            assert sum(p.items[p.first[ph] + i].nst for i in range(p.n_items[ph])) == want
    assert all(progs[c].n_attn in (2, 3) for c in range(144))


def test_packed_layout_matches_the_fragment_order_the_kernel_reads():
    """numpy model of pack_weight_kernel + the consumer's addressing: lane (g, tq) of warp w, n8 tile j of a stage
    must receive W[tile*64 + 8j + g][kstage*64 + 16w + 4tq .. +3] -- the mma.m16n8k8 B fragments of its k16 chunk."""
    rng = np.random.default_rng(0)
    N, K = 128, 192
    W = rng.standard_normal((N, K)).astype(np.float32)
    KSt = K // 64
    i = np.arange(N * K // 4)
    lane, j, w, stage = i & 31, (i >> 5) & 7, (i >> 8) & 3, i >> 10
    tile, ks = stage // KSt, stage % KSt
    g, tq = lane >> 2, lane & 3
    rows = tile * 64 + 8 * j + g
    cols = ks * 64 + 16 * w + 4 * tq
    packed = np.stack([W[rows, cols + e] for e in range(4)], axis=1).reshape(-1)      # what the pack kernel writes
    # consumer: stage s of the phase = 4096 floats; warp w, tile j, lane -> float4 at w*1024 + j*128 + lane*4
    x = rng.standard_normal((16, K)).astype(np.float32)
    y = np.zeros((16, N), dtype=np.float64)
    for st in range(N // 64 * KSt):
        t_, k_ = st // KSt, st % KSt
        blk = packed[st * 4096:(st + 1) * 4096]
        for w_ in range(4):
            for j_ in range(8):
                for ln in range(32):
                    g_, tq_ = ln >> 2, ln & 3
                    w4 = blk[w_ * 1024 + j_ * 128 + ln * 4: w_ * 1024 + j_ * 128 + ln * 4 + 4]
                    kk = k_ * 64 + 16 * w_ + 4 * tq_
                    y[:, t_ * 64 + 8 * j_ + g_] += x[:, kk:kk + 4].astype(np.float64) @ w4.astype(np.float64)
    np.testing.assert_allclose(y, x.astype(np.float64) @ W.astype(np.float64).T, rtol=1e-12, atol=1e-12)
