"""alpha-stable generators with the reference signatures (``bem/datasets/Distributions.py:9-73``),
backed by the Philox/CMS CUDA kernels (``csrc/process.cu`` K1a/K1b) instead of host scipy + H2D.

Same names, argument meaning and error behaviour as the reference; tensors are created directly
on ``device`` (which must be a CUDA device -- there is no CPU path).
"""
import math

import torch

from .. import _lib, rng


def match_last_dims(data, size):
    """Distributions.py:9-28: expand a (B,) tensor to ``size`` (contiguous)."""
    assert data.dim() == 1, f"Data must be 1-dimensional, got {data.size()}"
    for _ in range(len(size) - 1):
        data = data.unsqueeze(-1)
    return data.expand(*size).contiguous()


def _outer_inner(size):
    size = [int(s) for s in size]
    if len(size) == 0:
        raise Exception("size must have at least one dimension")
    return size, size[0], int(math.prod(size[1:]))


def _clamp_arg(c):
    return -1.0 if c is None else float(c)


def gen_skewed_levy(alpha, size, device=None, isotropic=True, clamp_a=None, compact=False, state=None):
    """A ~ S(alpha/2, 1, 0, 2cos(pi alpha/4)^(2/alpha)); Distributions.py:33-51.

    ``compact=True`` (extension) returns the (B,) per-sample draws without the broadcast copy."""
    if alpha > 2.0 or alpha <= 0.0:
        raise Exception("Wrong value of alpha ({}) for skewed levy r.v generation".format(alpha))
    dev = _lib.require_cuda(device if device is not None else "cuda")
    size, n_outer, inner = _outer_inner(size)
    st = state or rng.default_state()
    if compact:
        assert isotropic, "compact layout only exists for isotropic noise"
        out = torch.empty(n_outer, device=dev, dtype=torch.float32)
        mode = _lib.A_COMPACT
    else:
        out = torch.empty(size, device=dev, dtype=torch.float32)
        mode = _lib.A_ISOTROPIC if isotropic else _lib.A_FULL
    with torch.cuda.device(dev):
        _lib.call("dlpm_b200_stable_A", _lib.ptr(out), n_outer, max(inner, 1), mode, float(alpha), _clamp_arg(clamp_a),
                  st.seed, st.reserve(1), st.sample_base, _lib.stream_ptr())
    return out


def gen_sas(alpha, size, a=None, device=None, isotropic=True, clamp_eps=None, scale=1.0, state=None):
    """eps = sqrt(A) * G, clamped to +-clamp_eps; Distributions.py:57-73 (clamp_a is NOT applied to the
    internal A draw, like the reference :64).  ``a`` may be a full-size tensor (reference behaviour) or,
    for isotropic noise, a compact (B,) tensor.  ``scale`` (extension) multiplies the result in-kernel."""
    if alpha > 2.0 or alpha <= 0.0:
        raise Exception("Wrong value of alpha ({}) for skewed levy r.v generation".format(alpha))
    dev = _lib.require_cuda(device if device is not None else (a.device if a is not None else "cuda"))
    size, n_outer, inner = _outer_inner(size)
    st = state or rng.default_state()
    out = torch.empty(size, device=dev, dtype=torch.float32)
    a_iso = isotropic
    if a is not None:
        a = a.to(dev, torch.float32)
        if a.numel() == n_outer and (inner != 1 or a.dim() == 1):
            a, a_iso = a.reshape(n_outer).contiguous(), True
        else:
            assert list(a.shape) == size, "a must have shape `size` (or (B,) for isotropic noise)"
            a, a_iso = a.contiguous(), False
    with torch.cuda.device(dev):
        _lib.call("dlpm_b200_sas", _lib.ptr(out), _lib.ptr(a), n_outer, max(inner, 1), 1 if a_iso else 0, float(alpha),
                  _clamp_arg(clamp_eps), float(scale), st.seed, st.reserve(1), st.sample_base, _lib.stream_ptr())
    return out


def gen_normal(size, device=None, state=None):
    """N(0, I) from the same Philox stream family (stands in for ``torch.randn`` on the hot path)."""
    dev = _lib.require_cuda(device if device is not None else "cuda")
    size, n_outer, inner = _outer_inner(size)
    st = state or rng.default_state()
    out = torch.empty(size, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.call("dlpm_b200_normal", _lib.ptr(out), n_outer, max(inner, 1), st.seed, st.reserve(1), st.sample_base,
                  _lib.stream_ptr())
    return out
