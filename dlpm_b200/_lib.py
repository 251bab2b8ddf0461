"""ctypes binding of libdlpm_b200.so (the C ABI in ``include/dlpm_b200.h``).

There is NO fallback: if the shared library is missing, or a CUDA device is required and absent,
the call raises.  PyTorch is used only for device memory and streams; every pointer handed to the
library is ``tensor.data_ptr()`` of a CUDA tensor and every launch goes to torch's current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdlpm_b200.so")

c_i64, c_u64, c_int, c_f32, c_vp = ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

# name -> argtypes, exactly the declarations of include/dlpm_b200.h / dlpm_b200_unet.h
SIGNATURES = {
    "dlpm_b200_stable_A": [c_vp, c_i64, c_i64, c_int, c_f32, c_f32, c_u64, c_u64, c_i64, c_vp],
    "dlpm_b200_sas": [c_vp, c_vp, c_i64, c_i64, c_int, c_f32, c_f32, c_f32, c_u64, c_u64, c_i64, c_vp],
    "dlpm_b200_normal": [c_vp, c_i64, c_i64, c_u64, c_u64, c_i64, c_vp],
    "dlpm_b200_sigma_scan": [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_f32, c_f32, c_u64, c_u64, c_i64, c_vp],
    "dlpm_b200_reverse_step": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_u64, c_u64,
                               c_i64, c_vp, c_vp],
    "dlpm_b200_dlim_step": [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_vp],
    "dlpm_b200_lim_step": [c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_u64,
                           c_u64, c_i64, c_vp, c_vp],
    "dlpm_b200_reverse_step_post": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_u64, c_u64,
                                    c_i64, c_vp, c_vp, c_vp],
    "dlpm_b200_dlim_step_post": [c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_int, c_vp, c_vp, c_vp],
    "dlpm_b200_lim_step_post": [c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_i64, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_u64,
                                c_u64, c_i64, c_vp, c_vp, c_int, c_vp],
    "dlpm_b200_advance_counter": [c_vp, c_int, c_vp],
    "dlpm_b200_set_counter": [c_vp, c_int, c_vp],
    "dlpm_b200_training_elements": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_f32, c_f32, c_u64,
                                    c_u64, c_i64, c_vp],
    "dlpm_b200_scale_by_step": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_i64, c_vp],
    "dlpm_b200_lim_training_elements": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_int, c_f32, c_u64, c_u64, c_i64,
                                        c_vp],
    "dlpm_b200_loss_terms": [c_vp, c_vp, c_vp, c_i64, c_i64, c_f32, c_int, c_vp],
    "dlpm_b200_postprocess": [c_vp, c_vp, c_i64, c_f32, c_int, c_vp],
    "dlpm_b200_mlp_forward": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp],
    "dlpm_b200_mlp_sample_chain": [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp,
                                   c_vp, c_u64, c_u64, c_i64, c_vp],
}

class Post(ctypes.Structure):
    """``dlpm_b200_post_t``: post-processing of the final sample fused into the last step (bem/GenerationManager.py:50-63)."""
    _fields_ = [("out", c_vp), ("clamp", c_f32), ("mode", c_int), ("channels", c_int)]


POST_NONE, POST_F32, POST_F32_IMAGE, POST_U8_NHWC = 0, 1, 2, 3


def make_post(out, clamp, mode, channels=1):
    """ctypes pointer to a ``dlpm_b200_post_t`` (or None): ``out`` is a CUDA tensor that receives the processed x_0."""
    if out is None or mode == POST_NONE:
        return None
    return ctypes.byref(Post(ptr(out), float(clamp), int(mode), int(channels)))


A_COMPACT, A_ISOTROPIC, A_FULL = 0, 1, 2
STEP_CLIP_DENOISED, STEP_EPS_BF16, STEP_SIGMA_FULL = 1, 2, 4

_lib = None


class DlpmB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DlpmB200Error(
            "libdlpm_b200.so not found at %s -- build it with `python -m dlpm_b200.build` "
            "(there is no CPU / PyTorch fallback for the DLPM hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.dlpm_b200_abi_version.restype = c_int
    lib.dlpm_b200_philox_rounds.restype = c_int
    lib.dlpm_b200_last_error.restype = ctypes.c_char_p
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    try:
        from . import _unet_lib  # noqa: F401  (declares the UNet entry points on the same handle)
        _unet_lib.declare(lib)
    except ImportError:
        pass
    _lib = lib
    return lib


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a CUDA tensor (or NULL for None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DlpmB200Error("dlpm_b200 kernels need CUDA tensors (got device %s); there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise DlpmB200Error("dlpm_b200 kernels need contiguous tensors")
    return ctypes.c_void_p(t.data_ptr())


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.dlpm_b200_last_error().decode("utf-8", "replace")
        raise DlpmB200Error("%s failed (%d): %s" % (name, rc, msg))


def require_cuda(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise DlpmB200Error("dlpm_b200 runs on CUDA devices only (got %r); there is no CPU fallback" % (device,))
    if not torch.cuda.is_available():
        raise DlpmB200Error("CUDA device requested but torch.cuda.is_available() is False")
    return dev
