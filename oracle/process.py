"""Oracle: DLPM / DLIM / LIM reverse processes (torch CPU fp32).  TEST INFRASTRUCTURE ONLY.

Restates, with explicit injected noise, the arithmetic of
  * ``dlpm/methods/dlpm.py``  (schedules :103-156, Sigma recursion :230-239,
    Gamma / posterior :250-278, DLIM :281-297, predict_xstart/eps :191-202,
    one-r.v. training elements :384-401),
  * ``dlpm/methods/GenerativeLevyProcess.py``  (p_mean_variance :154-219, p_sample
    :225-239, p_sample_loop_progressive :291-330, ddim_* :332-452, lim_sample :454-506,
    training_losses_dlpm :612-677, compute_loss_terms :19-31),
  * ``dlpm/methods/LIM/functions/sampler.py`` (:81-181, :218-258) and ``sde.py`` (:5-49).
Every tensor op is written in the same order as the reference so fp32 results
agree to rounding; pinned by ``tests/test_oracle_golden.py``.
"""
import math

import torch


# ----------------------------------------------------------------------------------------
# schedules  (dlpm.py:103-156)
# ----------------------------------------------------------------------------------------
def get_timesteps(steps, time_spacing="linear"):
    """dlpm.py:103-110."""
    if time_spacing == "linear":
        return torch.tensor(range(0, steps), dtype=torch.float32)
    if time_spacing == "quadratic":
        return steps * (torch.tensor(range(0, steps), dtype=torch.float32) / steps) ** 2
    raise NotImplementedError(time_spacing)


def gen_noise_schedule(alpha, diffusion_steps, time_spacing="linear", scale="scale_preserving"):
    """dlpm.py:114-156.  Returns (gammas, bargammas, sigmas, barsigmas), each (T,) fp32."""
    timesteps = get_timesteps(diffusion_steps, time_spacing)
    if scale == "scale_preserving":
        s = 0.008
        schedule = torch.cos((timesteps / diffusion_steps + s) / (1 + s) * torch.pi / 2) ** 2
        baralphas = schedule / schedule[0]
        betas = 1 - baralphas / torch.concatenate([baralphas[0:1], baralphas[0:-1]])
        alphas = 1 - betas
        gammas = alphas ** (1 / alpha)
        bargammas = torch.cumprod(gammas, dim=0)
        sigmas = (1 - gammas ** alpha) ** (1 / alpha)
        barsigmas = (1 - bargammas ** alpha) ** (1 / alpha)
    elif scale == "scale_exploding":
        sigma_min, sigma_max, rho = 0.002, 80, 7
        gammas = torch.ones_like(timesteps)
        bargammas = torch.ones_like(timesteps)
        barsigmas = (sigma_min ** (1 / rho) + (timesteps / (diffusion_steps - 1))
                     * (sigma_max ** (1 / rho) - sigma_min ** (1 / rho))) ** rho
        barsigmas_alpha = barsigmas ** alpha
        sigmas_alpha = torch.ones_like(barsigmas) * barsigmas_alpha[0]
        for i in range(1, len(barsigmas)):
            sigmas_alpha[i] = barsigmas_alpha[i] - torch.sum(sigmas_alpha[:i])
        sigmas = sigmas_alpha ** (1 / alpha)
    else:
        raise AssertionError("Unknown scale")
    return gammas, bargammas, sigmas, barsigmas


def _bc(v, x):
    """per-sample (B,) vector -> broadcastable against x (B, ...)."""
    return v.view(-1, *([1] * (x.dim() - 1)))


# ----------------------------------------------------------------------------------------
# Sigma recursion and posterior  (dlpm.py:230-278)
# ----------------------------------------------------------------------------------------
def compute_Sigmas(A, gammas, sigmas):
    """dlpm.py:230-239.  A: (T, B, ...) -> Sigmas (T, B, ...);  Sigma_0 = s_0^2 A_0,
    Sigma_t = s_t^2 A_t + g_t^2 Sigma_{t-1}."""
    out = [sigmas[0] ** 2 * A[0]]
    for t in range(1, A.shape[0]):
        out.append(sigmas[t] ** 2 * A[t] + gammas[t] ** 2 * out[-1])
    return torch.stack(out)


def dlpm_posterior(x_t, eps, t, Sigmas, gammas, barsigmas):
    """dlpm.py:250-278 (anterior_mean_variance_dlpm with scalar t)."""
    Gamma_t = 1 - (gammas[t] ** 2 * Sigmas[t - 1]) / Sigmas[t]
    mean = (x_t - barsigmas[t] * Gamma_t * eps) / gammas[t]
    var = Gamma_t * Sigmas[t - 1]
    return mean, var


def clip_eps(x_t, eps, t, bargammas, barsigmas):
    """clip_denoised path, GenerativeLevyProcess.py:186-207 + dlpm.py:191-202."""
    xstart = ((x_t - eps * barsigmas[t]) / bargammas[t]).clamp(-1, 1)
    return (x_t - xstart * bargammas[t]) / barsigmas[t]


def dlpm_step(x_t, eps, z, t, Sigmas, sched, clip_denoised=False):
    """One stochastic reverse step (p_sample, GenerativeLevyProcess.py:225-239)."""
    g, bg, s, bs = sched
    if clip_denoised:
        eps = clip_eps(x_t, eps, t, bg, bs)
    mean, var = dlpm_posterior(x_t, eps, t, Sigmas, g, bs)
    nonzero = 0.0 if t == 1 else 1.0
    return mean + nonzero * torch.sqrt(var) * z


def dlim_step(x_t, eps, t, sched, clip_denoised=False):
    """eta = 0 deterministic step, dlpm.py:281-287."""
    g, bg, s, bs = sched
    if clip_denoised:
        eps = clip_eps(x_t, eps, t, bg, bs)
    return (x_t - bs[t] * eps) / g[t] + bs[t - 1] * eps


def dlpm_sample_loop(model, x_init, A, z, alpha, T, time_spacing="linear",
                     clip_denoised=False, deterministic=False, rescale_timesteps=True, scale="scale_preserving",
                     input_scaling=False):
    """p_sample_loop_progressive (GenerativeLevyProcess.py:291-330) / ddim loop (:413-452)
    with injected noise.

    x_init: the already-scaled x_{T-1} = barsigma_{T-1} * eps_init (:313);  A: (T, B, ...)
    full-shape subordinators (dlpm.py:226-227);  z: (T-1, B, ...) Gaussians, z[k] is used at
    the k-th loop iteration (t = T-1-k).  ``input_scaling`` with the 'scale_exploding' schedule feeds the network
    x / (1 + barsigma_t) (GenerativeLevyProcess.py:177-180).  Returns (final, history[T, B, ...])."""
    sched = gen_noise_schedule(alpha, T, time_spacing, scale)
    g, bg, s, bs = sched
    Sigmas = compute_Sigmas(A, g, s)
    x = x_init
    hist = [x]
    B = x.shape[0]
    for k, t in enumerate(range(T - 1, 0, -1)):
        tt = torch.tensor([t] * B)
        tin = tt.float() * (1.0 / T) if rescale_timesteps else tt
        x_in = x
        if input_scaling and scale == "scale_exploding":
            x_in = x * _bc(1 / (1 + bs[tt]), x)
        eps = model(x_in, tin)
        if deterministic:
            x = dlim_step(x, eps, t, sched, clip_denoised)
        else:
            x = dlpm_step(x, eps, z[k], t, Sigmas, sched, clip_denoised)
        hist.append(x)
    return x, torch.stack(hist)


# ----------------------------------------------------------------------------------------
# LIM  (sde.py:5-49, sampler.py:81-181,218-258)
# ----------------------------------------------------------------------------------------
class VPSDE:
    """Cosine VPSDE, sde.py:5-49."""

    def __init__(self, alpha, T=0.9946):
        self.alpha = alpha
        self.cosine_s = 0.008
        self.T = T
        self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1.0 + self.cosine_s) * math.pi / 2.0))

    def marginal_log_mean_coeff(self, t):
        return torch.log(torch.cos((t + self.cosine_s) / (1.0 + self.cosine_s) * math.pi / 2.0)) \
            - self.cosine_log_alpha_0

    def diffusion_coeff(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.pow(1.0 - torch.exp(self.marginal_log_mean_coeff(t) * self.alpha), 1 / self.alpha)

    def beta(self, t):
        return math.pi / 2 * self.alpha / (self.cosine_s + 1) * torch.tan(
            (t + self.cosine_s) / (1 + self.cosine_s) * math.pi / 2)


def lim_coefficients(sde, s, t, ode):
    """Per-sample coefficient vectors of one LIM step (alpha != 2 branches).

    SDE (sampler.py:120-155):  x <- a x + alpha^2 (a-1) score + (a^alpha - 1)^(1/alpha) e_L
    ODE (sampler.py:86-111):   x <- a x - alpha (1-a) score,    a via diffusion_coeff ratio
    with score = model(x, s) * marginal_std(s)^-(alpha-1)."""
    score_scale = torch.pow(sde.marginal_std(s), -(sde.alpha - 1))
    if ode:
        a = sde.diffusion_coeff(t) * torch.pow(sde.diffusion_coeff(s), -1)
        return score_scale, a, -sde.alpha * (1 - a), None
    a = torch.exp(sde.marginal_log_mean_coeff(t) - sde.marginal_log_mean_coeff(s))
    noise_coeff = torch.pow(-1 + torch.pow(a, sde.alpha), 1 / sde.alpha)
    return score_scale, a, sde.alpha ** 2 * (-1 + a), noise_coeff


def lim_sample_loop(model, x_init, e_L, alpha, steps, ode=False):
    """LIM_sampler (sampler.py:218-258), alpha != 2.  e_L: (steps, B, ...) SaS noises
    (ignored for the ODE).  Returns (final, history[steps+1, B, ...])."""
    assert alpha != 2.0, "oracle covers the heavy-tailed branch only"
    sde = VPSDE(alpha)
    timesteps = torch.linspace(sde.T, 1e-5, steps + 1)
    x = x_init
    hist = [x]
    B = x.shape[0]
    for i in range(steps):
        vec_s = torch.ones((B,)) * timesteps[i]
        vec_t = torch.ones((B,)) * timesteps[i + 1]
        sc, a, c_score, c_noise = lim_coefficients(sde, vec_s, vec_t, ode)
        score = model(x, vec_s) * _bc(sc, x)
        x_new = _bc(a, x) * x + _bc(c_score, x) * score
        if not ode:
            x_new = x_new + _bc(c_noise, x) * e_L[i]
        x = x_new
        hist.append(x)
    return x, torch.stack(hist)


# ----------------------------------------------------------------------------------------
# training loss, Proposition (9)  (GenerativeLevyProcess.py:612-677, dlpm.py:384-401)
# ----------------------------------------------------------------------------------------
def compute_loss_terms(x, y, lploss):
    """GenerativeLevyProcess.py:19-31."""
    dims = list(range(1, x.dim()))
    if lploss == 2.0:
        return torch.sqrt(torch.nn.functional.mse_loss(x, y, reduction="none").mean(dim=dims))
    if lploss == 1.0:
        return torch.nn.functional.smooth_l1_loss(x, y, beta=1, reduction="none").mean(dim=dims)
    if lploss == -1:
        return torch.nn.functional.mse_loss(x, y, reduction="none").mean(dim=dims)
    return torch.pow(torch.linalg.norm(x - y, ord=lploss, dim=dims), 1 / lploss)


def one_rv_loss_elements(x0, t, A, z, bargammas, barsigmas):
    """dlpm.py:384-401: Sigma' = A bs_t^2; x_t = bg_t x0 + sqrt(Sigma') z; eps_t = (x_t - bg_t x0)/bs_t."""
    bg = _bc(bargammas[t], x0)
    bs = _bc(barsigmas[t], x0)
    Sigma = A * bs ** 2
    x_t = bg * x0 + Sigma ** (1 / 2) * z
    eps_t = (x_t - x0 * bg) / bs
    return x_t, eps_t


def training_loss_dlpm(model, x0, t, A, z, alpha, T, lploss=2.0, rescale_timesteps=True, scale="scale_preserving",
                       input_scaling=False):
    """training_losses_dlpm with mean aggregation and M=1 (GenerativeLevyProcess.py:612-677)."""
    g, bg, s, bs = gen_noise_schedule(alpha, T, scale=scale)
    x_t, eps_t = one_rv_loss_elements(x0, t, A, z, bg, bs)
    tin = t.float() * (1.0 / T) if rescale_timesteps else t
    x_in = x_t
    if input_scaling and scale == "scale_exploding":  # :651-654
        x_in = x_t * _bc(1 / (1 + bs[t]), x_t)
    model_eps = model(x_in, tin)
    return compute_loss_terms(model_eps, eps_t, lploss).mean()


def lim_training_elements(x0, t, e, alpha):
    """LIM/functions/loss.py:21-31: x_t = x0 * diffusion_coeff(t) + e * marginal_std(t);  score = -e / alpha (alpha != 2)."""
    sde = VPSDE(alpha)
    x_t = x0 * _bc(sde.diffusion_coeff(t), x0) + e * _bc(sde.marginal_std(t), x0)
    return x_t, -e / alpha


def training_loss_lim(model, x0, u, e, alpha):
    """training_losses_lim (GenerativeLevyProcess.py:680-709) + loss_fn (LIM/functions/loss.py:13-39): t = u (T - 1e-5) + 1e-5,
    loss = mean smooth-L1(model(x_t, t), score), beta = 1."""
    sde = VPSDE(alpha)
    t = u * (sde.T - 1e-5) + 1e-5
    x_t, score = lim_training_elements(x0, t, e, alpha)
    return torch.nn.functional.smooth_l1_loss(model(x_t, t), score, beta=1, reduction="mean")
