"""Persistent decode-step kernel (csrc/pstep.cuh): invariants of its static work plan and of the packed weight format,
checked on the CPU through the C-ABI (wmar_pstep_plan_debug / wmar_pstep_stage_src are host logic only, no CUDA call)."""
import ctypes

import numpy as np
import pytest

PS_MAX_TILES, PS_MAX_ATTN, PS_PASS_TILES, PS_MAX_F = 48, 16, 4, 4
PH_QKV, PH_PROJ, PH_FC1, PH_HEAD = 0, 1, 2, 3


class PsProg(ctypes.Structure):
    _fields_ = [("n_tiles", ctypes.c_int * 4), ("first", ctypes.c_int * 4), ("tiles", ctypes.c_uint16 * PS_MAX_TILES),
                ("n_attn", ctypes.c_int), ("attn", ctypes.c_uint16 * PS_MAX_ATTN), ("red_lo", ctypes.c_int),
                ("red_hi", ctypes.c_int), ("layer_off16", ctypes.c_uint32), ("layer_stages", ctypes.c_uint32),
                ("head_off16", ctypes.c_uint32), ("head_stages", ctypes.c_uint32)]


def _plan(G, d, H, V):
    from wmar_b200 import _lib
    L = _lib.lib()
    assert L.wmar_pstep_prog_bytes() == ctypes.sizeof(PsProg)
    progs = (PsProg * G)()
    info = (ctypes.c_longlong * 8)()
    _lib.check(L.wmar_pstep_plan_debug(G, d, H, V, ctypes.cast(progs, ctypes.c_void_p), info))
    keys = ("Kp", "KC", "NBn", "layer_stages", "head_stages", "max_load", "min_load")
    return progs, dict(zip(keys, list(info)))


def _tiles(p, ph):
    return [p.tiles[p.first[ph] + i] for i in range(p.n_tiles[ph])]


@pytest.mark.parametrize("G,d,H,V", [(148, 1536, 24, 16384), (144, 1536, 24, 16384), (132, 1536, 24, 16384),
                                     (148, 384, 6, 16384), (148, 128, 2, 16384), (148, 1024, 16, 1024)])
def test_plan_deals_every_tile_item_and_slice_exactly_once(G, d, H, V):
    progs, info = _plan(G, d, H, V)
    NT = [3 * d // 16, d // 16, 4 * d // 16, V // 16]
    for ph in range(4):
        seen = sorted(t for c in range(G) for t in _tiles(progs[c], ph))
        assert seen == list(range(NT[ph])), f"phase {ph}: a tile is missed or duplicated"
        for c in range(G):
            ts = _tiles(progs[c], ph)
            assert ts == sorted(ts)
            if ph != PH_HEAD:
                assert len(ts) <= PS_PASS_TILES           # single pass: the reduction scratch may alias the activations
    assert sorted(progs[c].attn[i] for c in range(G) for i in range(progs[c].n_attn)) == list(range(16 * H))
    # residual-stream slices: a partition of the 16 * d / 4 float4 words, in CTA order
    assert progs[0].red_lo == 0 and progs[G - 1].red_hi == 16 * d // 4
    assert all(progs[c].red_hi == progs[c + 1].red_lo for c in range(G - 1))
    # packed stream: per CTA contiguous, in CTA order, nothing in between
    off, hoff = 0, 0
    for c in range(G):
        p = progs[c]
        f = p.n_tiles[PH_FC1]
        assert f <= PS_MAX_F
        assert p.layer_off16 == off * 1024 and p.head_off16 == hoff * 1024
        assert p.layer_stages == (p.n_tiles[PH_QKV] + p.n_tiles[PH_PROJ] + f) * info["KC"] + info["NBn"] * f
        assert p.head_stages == p.n_tiles[PH_HEAD] * info["KC"]
        off += p.layer_stages
        hoff += p.head_stages
    assert off == info["layer_stages"] and hoff == info["head_stages"]
    assert info["Kp"] % 256 == 0 and info["Kp"] >= d and info["KC"] * 256 == info["Kp"]


def test_full_size_plan_is_balanced():
    """Taming C2 (d = 1536, 24 heads) on 148 CTAs: every CTA streams nearly the same number of bytes per layer (greedy
    loads in half tiles: fc1 tile + its fc2 share = 4, qkv / proj tile = 2, attention item = 1) and the whole packed
    layer is exactly the 12 d^2 fp32 weights (no padding at this width)."""
    progs, info = _plan(148, 1536, 24, 16384)
    assert info["layer_stages"] * 16384 == 12 * 1536 * 1536 * 4
    assert info["head_stages"] * 16384 == 16384 * 1536 * 4
    assert info["max_load"] - info["min_load"] <= 2, info
    st = [progs[c].layer_stages for c in range(148)]
    assert max(st) - min(st) <= 12, (min(st), max(st))          # at most two K-type tiles of 6 stages apart
    assert all(2 <= progs[c].n_attn <= 4 for c in range(148))        # one batch of at most four warp groups


def _stage_src(G, d, H, V, cta, s, head):
    from wmar_b200 import _lib
    out = (ctypes.c_int * 4)()
    _lib.check(_lib.lib().wmar_pstep_stage_src(G, d, H, V, cta, s, head, out))
    return tuple(out)


@pytest.mark.parametrize("G,d,H,V", [(12, 128, 2, 1024), (40, 384, 6, 96)])
def test_packed_stream_reproduces_every_gemm(G, d, H, V):
    """numpy model of pack_stage_kernel + the consumers' addressing for a whole layer on a few CTAs: walking every CTA's
    stage stream with the kernel's own index arithmetic (unit w of a K-type stage = k16 step w of chunk kc of an n16
    tile; unit w of an fc2-type stage = n16 tile 16 nb + w contracted with one fc1 tile's 16 columns; lane (g, tq) holds
    four consecutive k) must give X W^T for qkv / proj / fc1 / head and sum_c partial_c = H W2^T for fc2."""
    progs, info = _plan(G, d, H, V)
    KC, NBn = info["KC"], info["NBn"]
    rng = np.random.default_rng(d)
    W = {PH_QKV: rng.standard_normal((3 * d, d)), PH_PROJ: rng.standard_normal((d, d)),
         PH_FC1: rng.standard_normal((4 * d, d)), PH_HEAD: rng.standard_normal((V, d))}
    W2 = rng.standard_normal((d, 4 * d))
    x = rng.standard_normal((16, d))
    hid = rng.standard_normal((16, 4 * d))

    def pack_stage(src):
        """[w 16][u 2][lane 32][4] exactly as pack_stage_kernel writes it"""
        ph, tile, kc, nb = src
        out = np.zeros((16, 2, 32, 4))
        for w in range(16):
            for u in range(2):
                for lane in range(32):
                    g, tq = lane >> 2, lane & 3
                    if ph >= 0:
                        n, k = 16 * tile + 8 * u + g, 256 * kc + 16 * w + 4 * tq
                        if n < W[ph].shape[0] and k < d:
                            out[w, u, lane] = W[ph][n, k:k + 4]
                    else:
                        n, k = 256 * nb + 16 * w + 8 * u + g, 16 * tile + 4 * tq
                        if n < d:
                            out[w, u, lane] = W2[n, k:k + 4]
        return out

    def unit(acc, xrows, wunit):
        """acc[16][16] += X[16][16] . Wunit^T with the kernel's lane mapping (xrows = X[:, k0:k0+16])"""
        for u in range(2):
            for lane in range(32):
                g, tq = lane >> 2, lane & 3
                acc[:, 8 * u + g] += xrows[:, 4 * tq:4 * tq + 4] @ wunit[u, lane]

    xp = np.zeros((16, info["Kp"]))
    xp[:, :d] = x
    got = {ph: np.zeros((16, W[ph].shape[0])) for ph in W}
    fc2 = np.zeros((16, d))
    for c in range(G):
        p = progs[c]
        s = 0
        for ph in (PH_QKV, PH_PROJ, PH_FC1):
            ts = _tiles(p, ph)
            acc = {t: np.zeros((16, 16)) for t in ts}
            for kc in range(KC):                        # kernel order: k chunk major, tile minor (single pass)
                for t in ts:
                    src = _stage_src(G, d, H, V, c, s, 0)
                    assert src[:3] == (ph, t, kc), (src, ph, t, kc)
                    st = pack_stage(src)
                    for w in range(16):
                        unit(acc[t], xp[:, 256 * kc + 16 * w:256 * kc + 16 * w + 16], st[w])
                    s += 1
            for t in ts:
                got[ph][:, 16 * t:16 * t + 16] = acc[t]
        ts = _tiles(p, PH_FC1)
        part = np.zeros((16, 256 * NBn))
        for nb in range(NBn):
            for j, t in enumerate(ts):
                src = _stage_src(G, d, H, V, c, s, 0)
                assert src == (-1, t, 0, nb), (src, t, nb)
                st = pack_stage(src)
                for w in range(16):
                    acc = np.zeros((16, 16))
                    unit(acc, hid[:, 16 * t:16 * t + 16], st[w])
                    part[:, 256 * nb + 16 * w:256 * nb + 16 * w + 16] += acc
                s += 1
        assert s == p.layer_stages
        fc2 += part[:, :d]
        # head: passes of four tiles
        ts = _tiles(p, PH_HEAD)
        s = 0
        for t0 in range(0, len(ts), PS_PASS_TILES):
            tp = ts[t0:t0 + PS_PASS_TILES]
            acc = {t: np.zeros((16, 16)) for t in tp}
            for kc in range(KC):
                for t in tp:
                    src = _stage_src(G, d, H, V, c, s, 1)
                    assert src[:3] == (PH_HEAD, t, kc)
                    st = pack_stage(src)
                    for w in range(16):
                        unit(acc[t], xp[:, 256 * kc + 16 * w:256 * kc + 16 * w + 16], st[w])
                    s += 1
            for t in tp:
                got[PH_HEAD][:, 16 * t:16 * t + 16] = acc[t]
        assert s == p.head_stages
    for ph in W:
        np.testing.assert_allclose(got[ph], x @ W[ph].T, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(fc2, hid @ W2.T, rtol=1e-11, atol=1e-11)
