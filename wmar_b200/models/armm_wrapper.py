"""Host-side mirror of ``wmar.models.armm_wrapper.AutoregressiveMultimodalModelWrapper`` (armm_wrapper.py:22-89).

Same method names, argument meaning and assertion behaviour as the reference; the arithmetic is done by the CUDA
engines of libwmar_b200.so.  ``get_image_tokenizer()`` keeps returning a real ``nn.Module`` tree (``StateModule``)
whose ``encoder`` / ``decoder`` / ``quantize`` sub-modules carry the reference's state-dict keys, so the delta
checkpoints of ``generate.py:327-332`` still apply; call ``sync_weights()`` after patching so the engines re-pack.
"""
import torch

from .state import load_ids


class AutoregressiveMultimodalModelWrapper:
    def __init__(self):
        self.watermarker = None

    def set_watermarker(self, watermarker=None):
        self.watermarker = watermarker

    def get_image_tokenizer(self):
        raise NotImplementedError("Subclass should implement this")

    def get_vq(self):
        return self.get_image_tokenizer().quantize

    def get_total_vocab_size(self):
        raise NotImplementedError("Subclass should implement this")

    @property
    def device(self):
        return self._device

    def init_alivecodes(self, alive_ids_path):
        """armm_wrapper.py:42-55: alive ids in file order; dead = set(range(n)) - alive (ascending)."""
        vq = self.get_vq()
        vocab_sz = vq.n_e if hasattr(vq, "n_e") else vq.num_embeddings
        alive_ids = load_ids(alive_ids_path)
        dead_ids = list(set(range(vocab_sz)) - set(alive_ids))
        vq.alive_ids = torch.tensor(alive_ids, dtype=torch.long)
        vq.dead_ids = torch.tensor(dead_ids, dtype=torch.long)

    def sample(self, conditioning, gen_params, apply_watermark=False):
        raise NotImplementedError("Subclass should implement this")

    def codes_to_images(self, codes):
        raise NotImplementedError("Subclass should implement this")

    def images_to_codes(self, images):
        raise NotImplementedError("Subclass should implement this")

    def sync_weights(self):
        raise NotImplementedError("Subclass should implement this")

    def is_codes_shaped(self, codes):
        return (isinstance(codes, torch.Tensor) and codes.ndim == 2
                and codes.shape[1] == self.codes_size * self.codes_size)

    def is_images_shaped(self, images):
        return (isinstance(images, torch.Tensor) and images.ndim == 4 and images.shape[1] == 3
                and images.shape[2] == self.image_size and images.shape[3] == self.image_size)

    # -- shared by the subclasses ---------------------------------------------------------------------------
    def _torch_stream(self, steps, rows, rowlen):
        """Parameters that let the sampler kernel draw, element by element, the very Exp(1) tensors that `steps`
        consecutive `torch.multinomial(probs[rows, rowlen], 1)` calls would draw from torch's CUDA generator on this device
        (ATen DistributionTemplates.h: calc_execution_policy + distribution_elementwise_grid_stride_kernel), and advance
        the generator past them -- a run seeded like the reference consumes the identical Philox stream, without the
        [steps, rows, rowlen] noise buffer and its `steps` host-issued exponential_ launches."""
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        gen = torch.cuda.default_generators[idx]
        props = torch.cuda.get_device_properties(idx)
        numel = rows * rowlen
        grid = min(props.multi_processor_count * (props.max_threads_per_multi_processor // 256), (numel + 255) // 256)
        threads = 256 * grid
        iters = (numel - 1) // (threads * 4) + 1
        seed, offset = gen.initial_seed(), gen.get_offset()
        gen.set_offset(offset + 4 * iters * steps)
        return dict(seed=seed, torch_offset=offset, torch_threads=threads, torch_numel=numel, torch_rowlen=rowlen)

    def _draw_noise(self, steps, rows, vocab):
        """q ~ Exp(1) for every step, drawn from torch's CUDA generator with the SAME sequence of calls the reference
        makes (`torch.multinomial(probs, 1)` == `argmax(probs / empty_like(probs).exponential_(1))`, one call per
        step on a [rows, V] tensor), so a run seeded like the reference consumes the identical Philox stream."""
        noise = torch.empty((steps, rows, vocab), dtype=torch.float32, device=self.device)
        for t in range(steps):
            noise[t].exponential_(1)
        return noise
