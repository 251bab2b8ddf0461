"""The caller side of the boundary: ``GenerationManager.generate`` (``bem/GenerationManager.py:8-63``) with its
post-processing fused into the sampler.

The reference calls ``method.sample(...)``, then -- three more passes and a blocking copy -- clamps the samples to +-1
(images) / +-6 (2-D data), moves them to the CPU and maps images by (x+1)/2 (``:50-63``).  Here the clamp / affine map
(/ optional uint8 quantisation for PNG writers) is written by the LAST reverse-step kernel (``FusedPost`` ->
``dlpm_b200_*_step_post``, include/dlpm_b200.h) and the device -> host copy is issued asynchronously into pinned memory
inside ``generate``; only the plotting / animation helpers of the reference class are out of scope (viz).
"""
import copy

import torch

from . import _lib


class FusedPost:
    """Post-processing of the final sample requested from ``GenerativeLevyProcess.sample(postprocess=...)``.

    ``clamp``: +-1 for images, +-6 for 2-D data (bem/GenerationManager.py:50); ``is_image``: apply (x+1)/2 (:58-63);
    ``uint8``: quantise like ``torchvision.utils.save_image`` (x*255 + 0.5, clamp, truncate) into NHWC bytes (images only).
    After ``sample`` returns, ``out`` is the device tensor holding the processed samples (fp32 of the sample shape, or
    uint8 ``[B, H, W, C]``)."""

    def __init__(self, clamp, is_image, uint8=False):
        assert clamp > 0
        assert not (uint8 and not is_image), "uint8 output is the image form"
        self.clamp, self.is_image, self.uint8 = float(clamp), bool(is_image), bool(uint8)
        self.out = None
        self._keep = None

    @property
    def mode(self):
        if self.uint8:
            return _lib.POST_U8_NHWC
        return _lib.POST_F32_IMAGE if self.is_image else _lib.POST_F32

    def bind(self, shape, device):
        """Allocate ``out`` for a batch of ``shape`` and return the ctypes ``dlpm_b200_post_t*`` the step kernels take."""
        shape = [int(s) for s in shape]
        if self.uint8:
            assert len(shape) == 4, "uint8 NHWC output needs (B, C, H, W) samples"
            self.out = torch.empty((shape[0], shape[2], shape[3], shape[1]), device=device, dtype=torch.uint8)
            channels = shape[1]
        else:
            self.out = torch.empty(shape, device=device, dtype=torch.float32)
            channels = 1
        self._keep = _lib.Post(_lib.ptr(self.out), self.clamp, self.mode, channels)
        import ctypes
        return ctypes.byref(self._keep)

    def run_standalone(self, x):
        """Same result from a finished x_0 (paths whose last step cannot carry the fusion): one extra launch."""
        self.bind(list(x.shape), x.device)
        B = x.shape[0]
        D = x[0].numel()
        xc = x.contiguous()
        if self.uint8:
            # a zero-noise DLIM-free way to reuse the step kernel's store is not worth it: clamp/affine kernel + torch cast
            tmp = torch.empty_like(xc)
            _lib.call("dlpm_b200_postprocess", _lib.ptr(tmp), _lib.ptr(xc), B * D, self.clamp, 1, _lib.stream_ptr())
            self.out.copy_((tmp * 255.0 + 0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8))
        else:
            _lib.call("dlpm_b200_postprocess", _lib.ptr(self.out), _lib.ptr(xc), B * D, self.clamp, 1 if self.is_image else 0,
                      _lib.stream_ptr())
        return self.out


def inverse_affine_transform(x):
    """bem/datasets/__init__.py:108-109."""
    return (x + 1) / 2


class GenerationManager:
    """``bem/GenerationManager.py:8-63``: same constructor and ``generate`` signature, same ``samples`` / ``history``
    attributes afterwards (CPU tensors).  ``dataloader`` is only used to learn the sample shape (:40-42); a plain shape
    (list / tuple / torch.Size) is accepted as well."""

    def __init__(self, method, dataloader, is_image, **kwargs):
        self.method = method
        self.original_data = dataloader
        self.is_image = is_image
        self.kwargs = kwargs
        self.samples = []
        self.history = []
        self.device_samples = None  # the post-processed samples on the device (multi-GPU callers gather these)
        self._pinned = None

    def _data_shape(self):
        od = self.original_data
        if isinstance(od, (list, tuple, torch.Size)) and all(isinstance(v, int) for v in od):
            return list(od)
        _, (data, y) = next(enumerate(od))
        return list(data.size())

    def generate(self, models, nsamples, get_sample_history=False, print_progression=False, **kwargs):
        assert nsamples > 0, 'nsamples must be greater than 0, got {}'.format(nsamples)
        tmp_kwargs = copy.deepcopy(self.kwargs)
        tmp_kwargs.update(kwargs)
        data_shape = self._data_shape()
        size = list(data_shape)
        size[0] = nsamples
        clamp = 1. if self.is_image else 6.
        last = data_shape[-1]
        if get_sample_history:
            # the history is a (T, B, ...) device tensor: the reference's own post-processing order (:51-63)
            x = self.method.sample(shape=size, models=models, print_progression=print_progression, get_sample_history=True,
                                   **tmp_kwargs)
            samples, hist = x
            self.samples = hist[-1, ..., :last]
            self.history = hist[..., :last].clamp(-clamp, clamp).cpu()
            self.samples = self.samples.clamp(-clamp, clamp).cpu()
            if self.is_image:
                self.samples = inverse_affine_transform(self.samples)
                if len(self.history) != 0:
                    self.history = torch.stack([inverse_affine_transform(h) for h in self.history])
            return
        post = FusedPost(clamp, self.is_image)
        x = self.method.sample(shape=size, models=models, print_progression=print_progression, get_sample_history=False,
                               postprocess=post, **tmp_kwargs)
        out = post.out if post.out is not None else post.run_standalone(x)
        self.device_samples = out
        out = out[..., :last]  # "select positions in case of pdmp" (:56); a no-op for the DLPM / LIM methods
        if not out.is_contiguous():
            out = out.contiguous()
        if self._pinned is None or self._pinned.shape != out.shape:
            self._pinned = torch.empty(out.shape, dtype=out.dtype).pin_memory()
        self._pinned.copy_(out, non_blocking=True)   # async D2H of the already post-processed samples
        torch.cuda.current_stream(out.device).synchronize()
        self.samples = self._pinned.clone()
        self.history = []
