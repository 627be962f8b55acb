"""The evaluation round trip of the reference's generate.py kept on the device:

    fill_batch_log                          generate.py:111-164   decode -> [round trips | augment] -> re-encode
    compute_metric("pvalue" | "l0" | "psnr") wmar/utils/metrics.py:19-45
    compute_metrics_and_save_from_batch_log generate.py:37-108    (the metrics part; files are written by wmar_b200.generate)

The reference copies every intermediate tensor to the host (`.cpu().numpy()`), calls `watermarker.detect` once per
image and converts images to PIL for the PSNR.  Here the log holds device tensors, the detector scores a whole
[B, L] batch per transform / parameter in one launch and the metrics are batched tensor expressions; `to_numpy=True`
reproduces the reference's log layout (tuples of numpy arrays) for callers that need it.
"""
import torch

from .augmentations import default_augmentations


@torch.no_grad()
def fill_batch_log(batch_log, key, model, codes, eval_params, to_numpy=False):
    """batch_log[key][transform] = [(param, codes [B,L], imgs [B,3,S,S] in [-1,1], None), ...]   (generate.py:111-164)"""
    conv = (lambda t: t.cpu().numpy()) if to_numpy else (lambda t: t)
    imgs = model.codes_to_images(codes)
    log = batch_log[key] = {}
    log["roundtrips"] = [(0, conv(codes), conv(imgs), None)]
    curr = imgs
    for T in range(1, eval_params.get("max_roundtrips", 1) + 1):
        curr_codes = model.images_to_codes(curr)
        curr = model.codes_to_images(curr_codes)
        log["roundtrips"].append((T, conv(curr_codes), conv(curr), None))
    for aug_name, aug_fn, aug_params in eval_params.get("augmentations", default_augmentations()):
        log[aug_name] = []
        for p in aug_params:
            x01 = imgs / 2.0 + 0.5                                   # augmentations expect [0, 1]
            aug = aug_fn(x01, p).clamp(0, 1) * 2.0 - 1.0
            aug_codes = model.images_to_codes(aug)
            log[aug_name].append((p, conv(aug_codes), conv(aug), None))
    return batch_log


def to_uint8(x):
    """wmar/utils/utils.py:69-80 chw_to_pillow on a batch: 255 * (x + 1) / 2 in fp32, clip to [0, 255], round half to even."""
    return torch.round((255 * ((x.float() + 1.0) / 2.0)).clamp(0, 255)).to(torch.uint8)


def psnr_uint8(a, b):
    """compute_psnr (metrics.py:19-21) on the 8-bit images the reference converts to (chw_to_pillow); a, b in [-1,1]."""
    qa, qb = to_uint8(a).double(), to_uint8(b).double()
    mse = (qa - qb).pow(2).flatten(1).mean(dim=1)
    return 10.0 * torch.log10(255.0 ** 2 / mse)


@torch.no_grad()
def compute_metrics(batch_log, key, watermarker, metric_names=("pvalue", "l0", "psnr")):
    """{transform: [(param, {metric: tensor[B]})]} for one method of the log (metrics.py:19-45, batched)."""
    log = batch_log[key]
    as_t = lambda x, dev: x if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=dev)
    dev = watermarker.device if watermarker is not None else None
    orig_codes = as_t(log["roundtrips"][0][1], dev)
    orig_imgs = as_t(log["roundtrips"][0][2], dev)
    out = {}
    for transform, entries in log.items():
        out[transform] = []
        for param, codes, imgs, _ in entries:
            codes, imgs = as_t(codes, dev), as_t(imgs, dev)
            m = {}
            for name in metric_names:
                if name == "l0":
                    m[name] = (orig_codes != codes).sum(dim=1).double() / orig_codes.shape[1]
                elif name == "psnr":
                    m[name] = psnr_uint8(imgs, orig_imgs)
                elif name == "pvalue":
                    m[name] = watermarker.detect(codes) if watermarker is not None else None
                else:
                    raise ValueError(f"Metric {name} not found")
            out[transform].append((param, m))
    return out


class _HostDetectState:
    """Copy stream, two device staging buffers and the pinned result buffer of detect_host_batches, kept on the model
    between calls (page-locking a fresh result buffer and creating a stream per call cost up to tens of ms)."""

    def __init__(self, dev, shape, n_batches, n_keys):
        self.shape, self.n_batches, self.n_keys = tuple(shape), n_batches, n_keys
        self.copy = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.out = torch.empty((n_batches, shape[0], n_keys), dtype=torch.float64).pin_memory()


@torch.no_grad()
def detect_host_batches(model, watermarker, host_batches, keys=("n_green", "n_scored", "z", "pvalue")):
    """Detection-only job (BASELINE configs[4]) over images that live in (pinned) HOST memory: for every batch
    images_to_codes -> watermarker.detect_stats, with the host->device copy of batch i + 1 running on a copy stream while
    batch i is encoded (two device staging buffers).  Returns a float64 pinned host tensor [n_batches, B, len(keys)]
    (filled asynchronously, synchronised before returning; the buffer is reused by the next call with the same shapes)."""
    host_batches = list(host_batches)
    if not host_batches:
        return torch.empty((0, 0, len(keys)), dtype=torch.float64)
    dev = model.device
    main = torch.cuda.current_stream(dev)
    shape = host_batches[0].shape
    stt = getattr(model, "_host_detect_state", None)
    if stt is None or stt.shape != tuple(shape) or stt.n_batches < len(host_batches) or stt.n_keys != len(keys):
        stt = model._host_detect_state = _HostDetectState(dev, shape, len(host_batches), len(keys))
    copy, stage, out = stt.copy, stt.stage, stt.out[: len(host_batches)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [None, None]
    copy.wait_stream(main)

    def upload(i):
        k = i % 2
        with torch.cuda.stream(copy):
            if free[k] is not None:
                copy.wait_event(free[k])                # batch i - 2 has been consumed
            stage[k][: host_batches[i].shape[0]].copy_(host_batches[i], non_blocking=True)
            ready[k].record(copy)

    upload(0)
    for i, hb in enumerate(host_batches):
        if i + 1 < len(host_batches):
            upload(i + 1)
        k = i % 2
        main.wait_event(ready[k])
        st = watermarker.detect_stats(model.images_to_codes(stage[k][: hb.shape[0]]))
        free[k] = torch.cuda.Event()
        free[k].record(main)
        res = torch.stack([st[key].to(torch.float64) for key in keys], dim=1)      # [B, n_keys] on the device: ONE copy
        out[i, : hb.shape[0]].copy_(res, non_blocking=True)
    main.synchronize()
    return out
