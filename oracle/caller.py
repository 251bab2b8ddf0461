"""Oracle: the CALLER of the drop-in boundary, restated.  TEST INFRASTRUCTURE ONLY.

``bem/GenerationManager.py:28-63`` is the code that calls ``method.sample(...)`` in the reference and
post-processes what it returns.  ``/root/reference`` does not travel to the GPU box, so the GPU tests drive
``dlpm_b200.GenerativeLevyProcess`` through this restatement; ``tests/test_oracle_golden.py`` checks it against the real
``GenerationManager`` (same kwargs forwarded to ``sample``, same samples / history out) whenever the reference tree is present.
"""
import copy

import torch


def inverse_affine_transform(x):
    """bem/datasets/__init__.py:108-109."""
    return (x + 1) / 2


def generation_manager_generate(method, models, data_shape, nsamples, is_image, manager_kwargs=None, get_sample_history=False,
                                print_progression=False, **kwargs):
    """GenerationManager.generate (bem/GenerationManager.py:28-63) for a data loader whose batches have shape ``data_shape``.
    Returns (samples, history) as the manager stores them (history is [] when not requested)."""
    assert nsamples > 0, 'nsamples must be greater than 0, got {}'.format(nsamples)
    tmp_kwargs = copy.deepcopy(manager_kwargs or {})  # :37-38
    tmp_kwargs.update(kwargs)
    size = list(data_shape)                          # :40-42
    size[0] = nsamples
    x = method.sample(shape=size, models=models, print_progression=print_progression, get_sample_history=get_sample_history,
                      **tmp_kwargs)                  # :43-47  <- the boundary
    clamp = 1. if is_image else 6.                   # :50
    history = []
    last = data_shape[-1]
    if get_sample_history:                           # :51-54
        samples, hist = x
        samples = hist[-1, ..., :last]
        history = hist[..., :last].clamp(-clamp, clamp).cpu()
    else:
        samples = x[..., :last]                      # :56
    samples = samples.clamp(-clamp, clamp).cpu()     # :57
    if is_image:                                     # :58-63
        samples = inverse_affine_transform(samples)
        if len(history) != 0:
            history = torch.stack([inverse_affine_transform(h) for h in history])
    return samples, history
