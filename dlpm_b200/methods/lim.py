"""LIM continuous-time sampler: ``VPSDE`` (``dlpm/methods/LIM/functions/sde.py:5-49``) and the
``LIM_sampler`` loop (``dlpm/methods/LIM/functions/sampler.py:15-259``) on the fused step kernel.

The per-step coefficients are batch-constant (sampler.py:229 builds ``ones(B) * timesteps[i]``), so
they are computed once on the host -- CPU fp32, same op order as the reference -- into a
``(steps, 4)`` device table (score_scale, a, c_score, c_noise) that the kernel indexes by step.
"""
import math

import torch

from .. import _lib, rng


class VPSDE:
    def __init__(self, alpha, schedule="cosine", T=0.9946):
        self.beta_0 = 0
        self.beta_1 = 20
        self.alpha = alpha
        self.cosine_s = 0.008
        self.schedule = schedule
        self.cosine_beta_max = 0.999
        self.cosine_t_max = math.atan(self.cosine_beta_max * (1.0 + self.cosine_s) / math.pi) * 2.0 \
            * (1.0 + self.cosine_s) / math.pi - self.cosine_s
        self.T = T if schedule == "cosine" else 1.0
        self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1.0 + self.cosine_s) * math.pi / 2.0))

    def beta(self, t):
        if self.schedule == "linear":
            return (self.beta_1 - self.beta_0) * t + self.beta_0
        return math.pi / 2 * self.alpha / (self.cosine_s + 1) * torch.tan(
            (t + self.cosine_s) / (1 + self.cosine_s) * math.pi / 2)

    def marginal_log_mean_coeff(self, t):
        if self.schedule == "linear":
            return -1 / (2 * self.alpha) * (t ** 2) * (self.beta_1 - self.beta_0) - 1 / self.alpha * t * self.beta_0
        return torch.log(torch.cos((t + self.cosine_s) / (1.0 + self.cosine_s) * math.pi / 2.0)) - self.cosine_log_alpha_0

    def diffusion_coeff(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.pow(1.0 - torch.exp(self.marginal_log_mean_coeff(t) * self.alpha), 1 / self.alpha)

    def inverse_a(self, a):
        return 2 / math.pi * (1 + self.cosine_s) * torch.acos(a) - self.cosine_s


def lim_step_table(sde, steps, ode):
    """(timesteps[steps+1], coef[steps,4]) on the CPU.  sampler.py:218 (grid), :86-111 (ODE), :120-155 (SDE)."""
    timesteps = torch.linspace(sde.T, 1e-5, steps + 1)
    s, t = timesteps[:-1], timesteps[1:]
    score_scale = torch.pow(sde.marginal_std(s), -(sde.alpha - 1))
    if ode:
        a = sde.diffusion_coeff(t) * torch.pow(sde.diffusion_coeff(s), -1)
        c_score = -sde.alpha * (1 - a)
        c_noise = torch.zeros_like(a)
    else:
        a = torch.exp(sde.marginal_log_mean_coeff(t) - sde.marginal_log_mean_coeff(s))
        c_score = sde.alpha ** 2 * (-1 + a)
        c_noise = torch.pow(-1 + torch.pow(a, sde.alpha), 1 / sde.alpha)
    return timesteps, torch.stack([score_scale, a, c_score, c_noise], dim=1).contiguous()


def LIM_sampler(ddim, x, y, model, sde, levy, isotropic, steps, gen_a, gen_eps, sde_clamp=None, masked_data=None,
                mask=None, t0=None, device="cuda", get_sample_history=False, injected_noise=None, net_call=None,
                state=None, postprocess=None):
    """Same signature as sampler.py:15-33 (+ ``injected_noise`` (steps, *x.shape) for parity tests).
    Heavy-tailed branch only (alpha != 2), 'sde' / 'ode' methods (imputation is not on the hot path)."""
    if sde.alpha == 2:
        raise NotImplementedError("LIM with alpha == 2 (plain Gaussian VPSDE branch) is out of scope")
    dev = _lib.require_cuda(device)
    timesteps, coef = lim_step_table(sde, steps, ode=bool(ddim))
    coef_d = coef.to(dev)
    x = x.to(dev, torch.float32).contiguous().clone()
    B = x.shape[0]
    if B == 0:  # the reference's LIM loop runs over the empty batch and returns empty tensors (sampler.py:218-258)
        return (x, torch.empty((steps + 1, *x.shape), device=dev, dtype=torch.float32)) if get_sample_history else x
    D = x[0].numel()
    clamp_eps = gen_eps.kwargs.get("clamp_eps", None)
    st = state or rng.default_state()
    offset = st.reserve(steps)
    hist = None
    if get_sample_history:
        hist = torch.empty((steps + 1, *x.shape), device=dev, dtype=torch.float32)
        hist[0].copy_(x)
    call = net_call or (lambda xx, tt: model(xx, tt))
    post = None if postprocess is None else postprocess.bind(list(x.shape), dev)
    if getattr(model, "native_kind", None) == "unet" and injected_noise is None and x.dim() == 4:
        # image nets: one CUDA graph per step, replayed `steps` times (times come from a device table)
        from .. import _unet_lib
        with torch.no_grad(), torch.cuda.device(dev):
            _unet_lib.run_lim_loop(model, x, coef_d, timesteps[:-1].contiguous().to(dev), steps, bool(ddim), bool(isotropic),
                                   float(sde.alpha), clamp_eps, hist, st.seed, offset, st.sample_base, post=post)
        return (x, hist) if get_sample_history else x
    with torch.no_grad(), torch.cuda.device(dev):
        for i in range(steps):
            vec_s = torch.full((B,), float(timesteps[i]), device=dev, dtype=torch.float32)
            out = call(x, vec_s)
            flags = _lib.STEP_EPS_BF16 if out.dtype == torch.bfloat16 else 0
            out = out.contiguous() if out.dtype == torch.bfloat16 else out.to(torch.float32).contiguous()
            e_L = None if injected_noise is None else injected_noise[i].to(dev, torch.float32).contiguous()
            _lib.call("dlpm_b200_lim_step_post", _lib.ptr(x), _lib.ptr(out), _lib.ptr(coef_d), i, None, B, D, flags,
                      1 if ddim else 0, 1 if isotropic else 0, float(sde.alpha),
                      -1.0 if clamp_eps is None else float(clamp_eps), _lib.ptr(e_L), st.seed, offset, st.sample_base,
                      _lib.ptr(hist[i + 1]) if hist is not None else None, post, steps - 1, _lib.stream_ptr())
    if get_sample_history:
        return x, hist
    return x
