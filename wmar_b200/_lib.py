"""ctypes binding of libwmar_b200.so (the C-ABI declared in include/wmar_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing the product path raises.
torch is used only for device memory, streams and tensors whose ``data_ptr()`` is handed to the C side.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libwmar_b200.so")

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_voidp = ctypes.c_void_p


class WmParams(ctypes.Structure):
    _fields_ = [("d_table", c_voidp), ("n_rows", ctypes.c_int64), ("vocab_size", ctypes.c_int64),
                ("seed_strategy", ctypes.c_int), ("context_size", ctypes.c_int), ("spatial_dim", ctypes.c_int),
                ("delta", ctypes.c_float), ("gamma", ctypes.c_double)]


class SampleParams(ctypes.Structure):
    _fields_ = [("temperature", ctypes.c_float), ("top_k", ctypes.c_int), ("top_p", ctypes.c_double),
                ("greedy", ctypes.c_int), ("seed", ctypes.c_uint64), ("rng_mode", ctypes.c_int),
                ("torch_threads", ctypes.c_int), ("torch_offset", ctypes.c_uint64), ("torch_numel", ctypes.c_int64),
                ("torch_rowlen", ctypes.c_int64)]


class GptConfig(ctypes.Structure):
    _fields_ = [("vocab_size", ctypes.c_int), ("block_size", ctypes.c_int), ("n_layer", ctypes.c_int),
                ("n_head", ctypes.c_int), ("n_embd", ctypes.c_int), ("max_batch", ctypes.c_int)]


class RarConfig(ctypes.Structure):
    _fields_ = [("codebook_size", ctypes.c_int), ("n_classes", ctypes.c_int), ("image_seq_len", ctypes.c_int),
                ("n_layer", ctypes.c_int), ("n_head", ctypes.c_int), ("hidden", ctypes.c_int), ("mlp", ctypes.c_int),
                ("max_batch", ctypes.c_int)]


class ChamConfig(ctypes.Structure):
    _fields_ = [("vocab_size", ctypes.c_int), ("dim", ctypes.c_int), ("n_layer", ctypes.c_int), ("n_head", ctypes.c_int),
                ("n_kv_head", ctypes.c_int), ("ffn_hidden", ctypes.c_int), ("max_seq", ctypes.c_int),
                ("max_batch", ctypes.c_int), ("image_token_lo", ctypes.c_int), ("image_token_hi", ctypes.c_int),
                ("norm_eps", ctypes.c_float), ("rope_theta", ctypes.c_float), ("qk_norm", ctypes.c_int)]


class VqganConfig(ctypes.Structure):
    _fields_ = [("family", ctypes.c_int), ("ch", ctypes.c_int), ("n_levels", ctypes.c_int),
                ("ch_mult", ctypes.c_int * 8), ("num_res_blocks", ctypes.c_int), ("attn_resolution", ctypes.c_int),
                ("resolution", ctypes.c_int), ("z_channels", ctypes.c_int), ("embed_dim", ctypes.c_int),
                ("n_embed", ctypes.c_int), ("max_batch", ctypes.c_int), ("precision", ctypes.c_int)]


EXPORTS = {
    # name: (restype, argtypes)
    "wmar_version": (ctypes.c_int, []),
    "wmar_last_error": (ctypes.c_char_p, []),
    "wmar_launch_count": (ctypes.c_uint64, []),
    "wmar_greenlist_build_host": (ctypes.c_int, [ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_uint64, c_voidp, ctypes.c_int64, c_voidp, ctypes.c_int64,
                                                 ctypes.c_int64, c_voidp, ctypes.c_int]),
    "wmar_greenlist_build_device": (ctypes.c_int, [ctypes.c_int64, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_uint64, c_voidp, ctypes.c_int64, c_voidp, ctypes.c_int64,
                                                   ctypes.c_int64, c_voidp, c_voidp]),
    "wmar_wm_process_logits": (ctypes.c_int, [ctypes.POINTER(WmParams), c_voidp, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_int64, c_voidp, c_voidp]),
    "wmar_wm_sample": (ctypes.c_int, [ctypes.POINTER(WmParams), ctypes.POINTER(SampleParams), c_voidp, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.c_int64, c_voidp, c_voidp, c_voidp, c_voidp]),
    "wmar_check_device_flag": (ctypes.c_int, [c_voidp]),
    "wmar_detect": (ctypes.c_int, [ctypes.POINTER(WmParams), c_voidp, ctypes.c_int64, ctypes.c_int64, c_voidp, c_voidp,
                                   c_voidp, c_voidp, c_voidp, ctypes.c_int64, c_voidp, c_voidp]),
    "wmar_gpt_create": (ctypes.c_int, [ctypes.POINTER(GptConfig), ctypes.POINTER(c_voidp), ctypes.c_int,
                                       ctypes.POINTER(c_voidp)]),
    "wmar_gpt_destroy": (None, [c_voidp]),
    "wmar_gpt_sample": (ctypes.c_int, [c_voidp, ctypes.POINTER(WmParams), ctypes.POINTER(SampleParams), c_voidp,
                                       ctypes.c_int64, ctypes.c_int64, c_voidp, c_voidp, c_voidp, c_voidp]),
    "wmar_gpt_algorithmic_bytes": (ctypes.c_double, [c_voidp, ctypes.c_int64, ctypes.c_int64]),
    "wmar_gpt_launches_per_step": (ctypes.c_int, [c_voidp]),
    "wmar_debug_torch_exponential": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64,
                                                    ctypes.c_int, c_voidp, c_voidp]),
    "wmar_pstep_prog_bytes": (ctypes.c_int, []),
    "wmar_pstep_plan_debug": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_voidp,
                                             ctypes.POINTER(ctypes.c_longlong)]),
    "wmar_pstep_stage_src": (ctypes.c_int, [ctypes.c_int] * 7 + [ctypes.POINTER(ctypes.c_int)]),
    "wmar_skinny_gemm": (ctypes.c_int, [c_voidp, c_voidp, c_voidp, c_voidp, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_int, c_voidp]),
    "wmar_rar_create": (ctypes.c_int, [ctypes.POINTER(RarConfig), ctypes.POINTER(c_voidp), ctypes.c_int,
                                       ctypes.POINTER(c_voidp)]),
    "wmar_rar_destroy": (None, [c_voidp]),
    "wmar_rar_sample": (ctypes.c_int, [c_voidp, ctypes.POINTER(WmParams), ctypes.POINTER(SampleParams), c_voidp,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_float, c_voidp, c_voidp, c_voidp,
                                       c_voidp]),
    "wmar_rar_algorithmic_bytes": (ctypes.c_double, [c_voidp, ctypes.c_int64, ctypes.c_int64]),
    "wmar_skinny_gemm_bf16": (ctypes.c_int, [c_voidp, c_voidp, c_voidp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                             c_voidp]),
    "wmar_cham_create": (ctypes.c_int, [ctypes.POINTER(ChamConfig), ctypes.POINTER(c_voidp), ctypes.c_int,
                                        ctypes.POINTER(c_voidp)]),
    "wmar_cham_destroy": (None, [c_voidp]),
    "wmar_cham_sample": (ctypes.c_int, [c_voidp, ctypes.POINTER(WmParams), ctypes.POINTER(SampleParams), c_voidp, c_voidp,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                        ctypes.c_int64, c_voidp, c_voidp, c_voidp, c_voidp]),
    "wmar_cham_algorithmic_bytes": (ctypes.c_double, [c_voidp, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int64]),
    "wmar_cham_launches_per_pass": (ctypes.c_int, [c_voidp]),
    "wmar_cham_select": (ctypes.c_int, [ctypes.POINTER(WmParams), ctypes.POINTER(SampleParams), c_voidp, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, ctypes.c_float,
                                        c_voidp, ctypes.c_int64, ctypes.c_int64, c_voidp, c_voidp, c_voidp, c_voidp]),
    "wmar_augment": (ctypes.c_int, [ctypes.c_int, c_voidp, c_voidp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.c_int, c_voidp, c_voidp]),
    "wmar_vqgan_create": (ctypes.c_int, [ctypes.POINTER(VqganConfig), ctypes.POINTER(c_voidp), ctypes.c_int,
                                         ctypes.POINTER(c_voidp)]),
    "wmar_vqgan_destroy": (None, [c_voidp]),
    "wmar_vqgan_decode": (ctypes.c_int, [c_voidp, c_voidp, ctypes.c_int64, c_voidp, c_voidp]),
    "wmar_vqgan_encode": (ctypes.c_int, [c_voidp, c_voidp, ctypes.c_int64, c_voidp, c_voidp]),
    "wmar_vqgan_flops": (ctypes.c_double, [c_voidp, ctypes.c_int]),
}

_lib = None
MISSING = []


class WmarError(RuntimeError):
    pass


def lib():
    """Loads the CUDA library; raises if it has not been built (python -m wmar_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise WmarError(f"{SO_PATH} is missing: the CUDA extension was not built "
                            "(run `python -m wmar_b200.build`); wmar_b200 has no CPU fallback")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in EXPORTS.items():
            try:
                fn = getattr(L, name)
            except AttributeError:
                MISSING.append(name)  # tests/test_abi.py asserts this list is empty
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    """Maps a wmar_status to the exception the reference would raise for that condition."""
    if rc == 0:
        return
    msg = lib().wmar_last_error().decode("utf-8", "replace")
    if rc in (-1, -5):      # WMAR_ERR_INVALID / WMAR_ERR_SHORT -> the reference raises ValueError / AssertionError
        raise ValueError(msg)
    if rc == -3:
        raise IndexError(msg)
    if rc == -4:
        raise MemoryError(msg)
    raise WmarError(f"wmar_b200 CUDA error ({rc}): {msg}")


def check_device_flag():
    """Reads and clears the current device's error flag (one stream synchronisation).  The product entry points call it
    once per sample() / detect() so that an out-of-range context id or a top-p overflow raises like the reference's
    indexing would, instead of yielding silently un-watermarked samples or wrong p-values."""
    check(lib().wmar_check_device_flag(current_stream()))


def ptr(t):
    """Device (or host) pointer of a contiguous torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous(), "tensor must be contiguous"
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pointer_table(tensors):
    arr = (c_voidp * len(tensors))()
    for i, t in enumerate(tensors):
        assert t.is_cuda and t.is_contiguous(), f"weight {i} must be a contiguous CUDA tensor"
        arr[i] = t.data_ptr()
    return arr
