"""ctypes binding for oracle/wm_oracle.c (TEST INFRASTRUCTURE ONLY) + detector statistics.

Follows wmar/watermarking/gentime_watermark.py (reference): greenlist :161-226, detect :285-344.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libwm_oracle.so")

SPLIT = {"rand": 0, "stratifiedrand": 1}
SEED = {"fixed": 0, "linear": 1, "spatial": 2}
SALT = 15485863


def build():
    src = os.path.join(_HERE, "wm_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_randperm.argtypes = [ctypes.c_uint64, ctypes.c_int64, i64p]
        L.oracle_context_seed.restype = ctypes.c_uint64
        L.oracle_context_seed.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.oracle_greenlist_ids.restype = ctypes.c_int64
        L.oracle_greenlist_ids.argtypes = [ctypes.c_int64, ctypes.c_double, ctypes.c_int, i64p, ctypes.c_int64, i64p,
                                           ctypes.c_int64, ctypes.c_uint64, i64p]
        L.oracle_greenlist_bitmask.argtypes = [ctypes.c_int64, ctypes.c_double, ctypes.c_int, i64p, ctypes.c_int64,
                                               i64p, ctypes.c_int64, ctypes.c_uint64,
                                               ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_detect_one.restype = ctypes.c_int
        L.oracle_detect_one.argtypes = [i64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_uint64, i64p, ctypes.c_int64, i64p,
                                        ctypes.c_int64, i64p, i64p, ctypes.POINTER(ctypes.c_int32), i64p]
        _lib = L
    return _lib


def _p64(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def randperm(seed, n):
    out = np.empty(n, dtype=np.int64)
    lib().oracle_randperm(seed, n, _p64(out))
    return out


def context_seed(ctx_sum, salt=SALT):
    return int(lib().oracle_context_seed(salt, int(ctx_sum)))


def load_ids(path):
    """armm_wrapper.py:42-55 -- ids in FILE ORDER (comma separated, possibly multi-line)."""
    ids = []
    with open(path) as f:
        for line in f:
            ids.extend(int(t) for t in line.split(",") if t.strip())
    return ids


def alive_dead(alive_list, n_e):
    """armm_wrapper.py:53: dead = list(set(range(n_e)) - set(alive)) -- CPython small-int set order == sorted."""
    alive = np.asarray(alive_list, dtype=np.int64)
    dead = np.asarray(sorted(set(range(n_e)) - set(alive_list)), dtype=np.int64)
    return alive, dead


def greenlist_ids(V, gamma, split, alive, dead, seed):
    alive = np.ascontiguousarray(alive, dtype=np.int64)
    dead = np.ascontiguousarray(dead, dtype=np.int64)
    out = np.empty(int(V * gamma) + 1, dtype=np.int64)
    n = lib().oracle_greenlist_ids(V, float(gamma), SPLIT[split], _p64(alive), len(alive), _p64(dead), len(dead),
                                   seed, _p64(out))
    return out[:n].copy()


def greenlist_bitmask(V, gamma, split, alive, dead, seed):
    alive = np.ascontiguousarray(alive, dtype=np.int64)
    dead = np.ascontiguousarray(dead, dtype=np.int64)
    out = np.zeros((V + 31) // 32, dtype=np.uint32)
    lib().oracle_greenlist_bitmask(V, float(gamma), SPLIT[split], _p64(alive), len(alive), _p64(dead), len(dead),
                                   seed, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    return out


def bitmask_table(V, gamma, split, alive, dead, n_ctx_sums, salt=SALT):
    """[n_ctx_sums, ceil(V/32)] u32: row s = greenlist for context sum s (LINEAR/SPATIAL seeding)."""
    return np.stack([greenlist_bitmask(V, gamma, split, alive, dead, context_seed(s, salt))
                     for s in range(n_ctx_sums)])


def detect_counts(codes, V, gamma, split, seed_strategy, h, alive, dead, salt=SALT, return_mask=False):
    """codes int64[B, L] -> (n_green[B], n_scored[B], masks) exactly as gentime_watermark.py:285-344 counts them."""
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    alive = np.ascontiguousarray(alive, dtype=np.int64)
    dead = np.ascontiguousarray(dead, dtype=np.int64)
    B, L = codes.shape
    ng = np.zeros(B, dtype=np.int64)
    ns = np.zeros(B, dtype=np.int64)
    masks = []
    for b in range(B):
        g = ctypes.c_int64(0)
        s = ctypes.c_int64(0)
        ml = ctypes.c_int64(0)
        m = np.full(L + h + 1, -2, dtype=np.int32)
        rc = lib().oracle_detect_one(_p64(codes[b]), L, V, float(gamma), SPLIT[split], SEED[seed_strategy], h, salt,
                                     _p64(alive), len(alive), _p64(dead), len(dead), ctypes.byref(g),
                                     ctypes.byref(s), m.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                     ctypes.byref(ml))
        if rc == -1:
            raise ValueError("Must have at least 1 token to score after the context")  # :287-291
        if rc != 0:
            raise ValueError("bad spatial n-gram configuration")
        ng[b], ns[b] = g.value, s.value
        masks.append(m[:ml.value].tolist())
    if return_mask:
        return ng, ns, masks
    return ng, ns


def pvalue(n_green, n_scored, gamma):
    """gentime_watermark.py:338 -- the reference's own function (scipy is third-party to the reference)."""
    from scipy import special
    return special.betainc(np.asarray(n_green, dtype=np.float64),
                           1.0 + np.asarray(n_scored, dtype=np.float64) - np.asarray(n_green, dtype=np.float64),
                           float(gamma))


def zscore(n_green, n_scored, gamma):
    """z := (n_green - gamma T) / sqrt(T gamma (1-gamma)); not in the reference, defined in SURVEY.md A.6."""
    n_green = np.asarray(n_green, dtype=np.float64)
    T = np.asarray(n_scored, dtype=np.float64)
    return (n_green - gamma * T) / np.sqrt(T * gamma * (1.0 - gamma))


def context_sum_for_row(past, seed_strategy, h, spatial_dim=16):
    """Context selection of _process_logits (gentime_watermark.py:233-263) for one row's history (list/array).

    Returns the context SUM (only the sum enters the seed, :225), 0 for FIXED, or None when the reference skips
    the row (ValueError swallowed at :268-270)."""
    n = len(past)
    if seed_strategy == "fixed":
        return 0
    if seed_strategy == "linear":
        if n < h:
            return None
        return int(sum(int(t) for t in past[n - h:])) if h > 0 else 0
    if seed_strategy == "spatial":
        assert h in (1, 3)
        if h == 3:
            if n < spatial_dim + 1:
                return None
            return int(past[n - spatial_dim - 1]) + int(past[n - spatial_dim]) + int(past[n - 1])
        if n < h:
            return None
        if n % spatial_dim == 0:
            # past[-spatial_dim : -spatial_dim + 1]; for spatial_dim == 1 that slice is empty (sum 0)
            sl = past[n - spatial_dim: n - spatial_dim + 1] if spatial_dim <= n else past[0:max(n - spatial_dim + 1, 0)]
            return int(sum(int(t) for t in sl))
        return int(past[n - 1])
    raise ValueError(seed_strategy)


class GreenRows:
    """Callable past_ids int64[B,t] -> list of bitmask rows (None = row skipped), with a per-seed cache."""

    def __init__(self, V, gamma, split, seed_strategy, h, alive, dead, salt=SALT, spatial_dim=16):
        self.V, self.gamma, self.split, self.seed_strategy, self.h = V, gamma, split, seed_strategy, h
        self.alive, self.dead, self.salt, self.spatial_dim = alive, dead, salt, spatial_dim
        self.cache = {}

    def row_for_sum(self, s):
        seed = 0 if self.seed_strategy == "fixed" else context_seed(s, self.salt)
        if seed not in self.cache:
            self.cache[seed] = greenlist_bitmask(self.V, self.gamma, self.split, self.alive, self.dead, seed)
        return self.cache[seed]

    def __call__(self, past_ids):
        rows = []
        for b in range(past_ids.shape[0]):
            s = context_sum_for_row([int(t) for t in past_ids[b]], self.seed_strategy, self.h, self.spatial_dim)
            rows.append(None if s is None else self.row_for_sum(s))
        return rows
