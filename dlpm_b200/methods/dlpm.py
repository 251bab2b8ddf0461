"""``DLPM``: noise schedules, the A_{1:T} / Sigma_t chains and the posterior, with the surface of
the reference's ``dlpm/methods/dlpm.py`` (class ``DLPM`` :56-415) on top of the CUDA kernels.

Differences in *representation* (not in results), all motivated by SURVEY.md section 3.1 "memory quirk":
  * isotropic A and Sigma are stored compact as ``(T, B)`` instead of ``(T, B, C, H, W)`` (50 GB at
    the CIFAR config); ``A_full()`` / ``Sigmas_full()`` expand on request;
  * the four schedule vectors are also kept as one packed ``(T, 4)`` device table ``sched`` --
    the ``(T, C, H, W)`` tables of ``get_schedule`` (:158-174) are never built;
  * ``sample_A`` + ``compute_Sigmas`` (:226-239) run as ONE scan kernel (K2) that draws A_t from the
    Philox/CMS generator and applies the recursion.
"""
import torch

from .. import _lib, rng
from ..datasets import Data


class ModelMeanType:
    """dlpm.py:10-19."""
    PREVIOUS_X = "PREVIOUS_X"
    START_X = "START_X"
    EPSILON = "EPSILON"
    Z = "Z"
    SQRT_GAMMA_EPSILON = "SQRT_GAMMA_EPSILON"


class ModelVarType:
    """dlpm.py:21-32."""
    FIXED = "FIXED"


class LossType:
    """dlpm.py:34-42."""
    LP_LOSS = "LP_LOSS"
    MEAN_LOSS = "MEAN_LOSS"
    EPS_LOSS = "EPS_LOSS"
    LAMBDA_LOSS = "LAMBDA_LOSS"
    VAR_KL = "VAR_KL"
    VAR_LP_SUM = "VAR_LP_SUM"


def match_last_dims(data, size):
    """dlpm.py:47-51 (repeat a (B,) tensor over the trailing dims of ``size``)."""
    assert len(data.size()) == 1
    for _ in range(len(size) - 1):
        data = data.unsqueeze(-1)
    return data.repeat(1, *(size[1:]))


def _bc(v, like):
    return v.view(-1, *([1] * (like.dim() - 1)))


class DLPM:
    def __init__(self, alpha, device, diffusion_steps, time_spacing="linear", isotropic=True, clamp_a=None,
                 clamp_eps=None, scale="scale_preserving"):
        self.alpha = alpha
        self.device = device
        self.time_spacing = time_spacing
        self.isotropic = isotropic
        self.use_single_a_chain = True
        self.scale = scale
        self._set_schedule(diffusion_steps, scale)
        self.constants = None
        self.gen_a = Data.Generator("skewed_levy", alpha=self.alpha, device=self.device, isotropic=isotropic,
                                    clamp_a=clamp_a)
        self.gen_eps = Data.Generator("sas", alpha=self.alpha, device=self.device, isotropic=isotropic,
                                      clamp_eps=clamp_eps)
        self.A = None        # compact (T, B) when isotropic, (T, B, *shape[1:]) otherwise
        self.Sigmas = None   # same layout as A
        self._shape = None
        self._sigma_src = None

    # ------------------------------------------------------------------ schedules (dlpm.py:103-185)
    def get_timesteps(self, steps):
        if self.time_spacing == "linear":
            return torch.tensor(range(0, steps), dtype=torch.float32)
        elif self.time_spacing == "quadratic":
            return steps * (torch.tensor(range(0, steps), dtype=torch.float32) / steps) ** 2
        raise NotImplementedError(self.time_spacing)

    def gen_noise_schedule(self, diffusion_steps, scale="scale_preserving"):
        """Host-side (CPU fp32, same op order as dlpm.py:114-156 so the tables are bit-identical)."""
        a = self.alpha
        ts = self.get_timesteps(diffusion_steps)
        if scale == "scale_preserving":
            s = 0.008
            f = torch.cos((ts / diffusion_steps + s) / (1 + s) * torch.pi / 2) ** 2
            baralphas = f / f[0]
            betas = 1 - baralphas / torch.concatenate([baralphas[0:1], baralphas[0:-1]])
            alphas = 1 - betas
            gammas = alphas ** (1 / a)
            bargammas = torch.cumprod(gammas, dim=0)
            sigmas = (1 - gammas ** a) ** (1 / a)
            barsigmas = (1 - bargammas ** a) ** (1 / a)
        elif scale == "scale_exploding":
            smin, smax, rho = 0.002, 80, 7
            gammas = torch.ones_like(ts)
            bargammas = torch.ones_like(ts)
            barsigmas = (smin ** (1 / rho) + (ts / (diffusion_steps - 1)) * (smax ** (1 / rho) - smin ** (1 / rho))) ** rho
            bsa = barsigmas ** a
            sa = torch.ones_like(barsigmas) * bsa[0]
            for i in range(1, len(barsigmas)):
                sa[i] = bsa[i] - torch.sum(sa[:i])
            sigmas = sa ** (1 / a)
        else:
            assert False, "Unknown scale"
        return gammas, bargammas, sigmas, barsigmas

    def _set_schedule(self, diffusion_steps, scale):
        g, bg, s, bs = self.gen_noise_schedule(diffusion_steps, scale=scale)
        self._sched_host = torch.stack([g, bg, s, bs], dim=1).contiguous()  # (T, 4) CPU
        self.sched = self._sched_host.to(self.device)
        self.gammas, self.bargammas, self.sigmas, self.barsigmas = (x.to(self.device) for x in (g, bg, s, bs))
        self.diffusion_steps = int(diffusion_steps)

    def rescale_diffusion(self, diffusion_steps, time_spacing=None):
        """dlpm.py:176-185 (like the reference, the rebuilt schedule is always 'scale_preserving')."""
        assert isinstance(diffusion_steps, int), "Diffusion steps must be an integer"
        if time_spacing is not None:
            self.time_spacing = time_spacing
        self._set_schedule(diffusion_steps, "scale_preserving")
        self.constants = None

    def snapshot_schedule(self):
        """Everything ``rescale_diffusion`` overwrites, so that ``sample(reverse_steps != train steps)`` can put the
        TRAINING schedule back afterwards -- including a 'scale_exploding' one, which ``rescale_diffusion`` (like the
        reference, dlpm.py:183) would silently replace by the cosine schedule."""
        return (self._sched_host, self.sched, self.gammas, self.bargammas, self.sigmas, self.barsigmas, self.diffusion_steps,
                self.time_spacing, self.constants)

    def restore_schedule(self, snap):
        (self._sched_host, self.sched, self.gammas, self.bargammas, self.sigmas, self.barsigmas, self.diffusion_steps,
         self.time_spacing, self.constants) = snap

    def get_schedule(self, shape):
        return tuple(match_last_dims(v, shape) for v in (self.gammas, self.bargammas, self.sigmas, self.barsigmas))

    def update_constants(self, shape):
        """dlpm.py:170-174 (API parity).  The kernels read the (T, 4) ``sched`` table; nothing is cached per shape."""
        self.constants = self.get_schedule(shape)
        return self.constants

    def get_t_to_batch_size(self, x_t, t):
        if isinstance(t, int):
            return torch.full([x_t.shape[0]], t).to(self.device)
        return t

    # ------------------------------------------------------------------ simple dynamics (dlpm.py:191-219)
    def _rows(self, x, t):
        t = self.get_t_to_batch_size(x, t)
        return (_bc(v[t], x) for v in (self.gammas, self.bargammas, self.sigmas, self.barsigmas))

    def predict_xstart(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        g, bg, s, bs = self._rows(x_t, t)
        return (x_t - eps * bs) / bg

    def predict_eps(self, x_t, t, xstart):
        g, bg, s, bs = self._rows(x_t, t)
        return (x_t - xstart * bg) / bs

    def sample_x_t_from_xstart(self, xstart, t, eps=None):
        g, bg, s, bs = self._rows(xstart, t)
        if eps is None:
            eps = self.gen_eps.generate(size=xstart.size())
        return bg * xstart + bs * eps, eps

    # API-parity helpers of the conditioned dynamics (dlpm.py:199-202, :243-270, :377-382): torch expressions on device
    # tensors, not used by the fused loops
    def _sigma_rows(self, like, t):
        tb = self.get_t_to_batch_size(like, t)
        S1, St = self.Sigmas[tb - 1], self.Sigmas[tb]
        if S1.dim() != like.dim():  # compact (T, B) tables: the batched index gives (B, B); take the matching rows
            idx = torch.arange(like.shape[0], device=S1.device)
            S1, St = _bc(self.Sigmas[tb - 1, idx], like), _bc(self.Sigmas[tb, idx], like)
        return S1, St

    def predict_eps_from_m_tilde(self, x_t, t, m_tilde_t_1):
        g, bg, s, bs = self._rows(m_tilde_t_1, t)
        S1, St = self._sigma_rows(m_tilde_t_1, t)
        Gamma_t = 1 - (g ** 2 * S1) / St
        return (x_t - m_tilde_t_1 * g) / (bs * Gamma_t)

    def sample_x_t_from_xstart_given_Sigma(self, xstart, t, Sigma_t, z_t=None):
        g, bg, s, bs = self._rows(xstart, t)
        if z_t is None:
            from ..datasets.Distributions import gen_normal
            z_t = gen_normal(xstart.shape, device=self.device)
        return bg * xstart + Sigma_t ** (1 / 2) * z_t

    def compute_m_tilde_t_1(self, x_t, t, Gamma_t, eps_t):
        g, bg, s, bs = self._rows(x_t, t)
        return (x_t - bs * Gamma_t * eps_t) / g

    def compute_one_rv_Sigma_prime_t(self, t, a_t):
        g, bg, s, bs = self._rows(a_t, t)
        return a_t * bs ** 2

    # ------------------------------------------------------------------ A / Sigma chains (dlpm.py:226-239)
    def _clamp_a(self):
        c = self.gen_a.kwargs.get("clamp_a", None)
        return -1.0 if c is None else float(c)

    def sample_A(self, shape, diffusion_steps, state=None):
        """Draw A_{0:T-1} and (fused) the Sigma chain.  Compact (T, B) when isotropic."""
        dev = _lib.require_cuda(self.device)
        shape = [int(s) for s in shape]
        B = shape[0]
        inner = 1
        for s in shape[1:]:
            inner *= s
        T = int(diffusion_steps)
        assert T == self.diffusion_steps, "diffusion_steps must match the current schedule"
        st = state or rng.default_state()
        n = B if self.isotropic else B * inner
        out_shape = (T, B) if self.isotropic else (T, *shape)
        self.A = torch.empty(out_shape, device=dev, dtype=torch.float32)
        self.Sigmas = torch.empty(out_shape, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.call("dlpm_b200_sigma_scan", _lib.ptr(self.Sigmas), None, _lib.ptr(self.A), _lib.ptr(self.sched), T, n,
                      max(inner, 1), 0 if self.isotropic else 1, float(self.alpha), self._clamp_a(), st.seed, st.reserve(T),
                      st.sample_base, _lib.stream_ptr())
        self._shape = shape
        self._sigma_src = self.A

    def compute_Sigmas(self, force=False):
        """Sigma_t = s_t^2 A_t + g_t^2 Sigma_{t-1}.  No-op if ``sample_A`` just produced them; re-runs
        the scan (K2) when ``self.A`` was replaced (e.g. injected for parity tests); pass ``force=True`` after editing
        ``self.A`` in place."""
        assert self.A is not None, "sample_A must be called first"
        if not force and self.Sigmas is not None and self._sigma_src is self.A and self.Sigmas.shape == self.A.shape:
            return
        dev = _lib.require_cuda(self.device)
        A = self.A.to(dev, torch.float32).contiguous()
        T = A.shape[0]
        n = A[0].numel()
        self.Sigmas = torch.empty_like(A)
        with torch.cuda.device(dev):
            _lib.call("dlpm_b200_sigma_scan", _lib.ptr(self.Sigmas), _lib.ptr(A), None, _lib.ptr(self.sched), T, n, 1, 0,
                      float(self.alpha), -1.0, 0, 0, 0, _lib.stream_ptr())
        self.A = A
        self._sigma_src = self.A

    def _full(self, v):
        if v is None or self._shape is None or v.dim() != 2:
            return v
        T, B = v.shape
        return v.view(T, B, *([1] * (len(self._shape) - 1))).expand(T, *self._shape)

    def A_full(self):
        """(T, *shape) view of A, as the reference stores it (dlpm.py:227)."""
        return self._full(self.A)

    def Sigmas_full(self):
        return self._full(self.Sigmas)

    # ------------------------------------------------------------------ posterior (dlpm.py:250-297)
    def compute_Gamma_t(self, t, Sigma_t_1, Sigma_t):
        g = self.gammas[t] if isinstance(t, int) else _bc(self.gammas[t], Sigma_t_1)
        return 1 - (g ** 2 * Sigma_t_1) / Sigma_t

    def compute_Sigma_tilde_t_1(self, Gamma_t, Sigma_t_1):
        return Gamma_t * Sigma_t_1

    def anterior_mean_variance_dlpm(self, x_t, t, eps):
        """API-parity helper (torch expressions on device tensors); the sampling loop itself uses the
        fused kernel K3 (``dlpm_b200_reverse_step``)."""
        t = int(t)
        S1, St = self.Sigmas[t - 1], self.Sigmas[t]
        if S1.dim() == 1:
            S1, St = _bc(S1, x_t), _bc(St, x_t)
        Gamma_t = 1 - (self.gammas[t] ** 2 * S1) / St
        x_t_1 = (x_t - self.barsigmas[t] * Gamma_t * eps) / self.gammas[t]
        return x_t_1, Gamma_t * S1

    def anterior_mean_variance_dlim(self, x_t, t, eps, eta=0.0):
        if eta != 0.0:
            raise NotImplementedError("dlim_eta != 0 is broken in the reference (dlpm.py:295 indexes A with a "
                                      "batched t -> (B,B,...) output); only eta = 0 is implemented")
        g, bg, s, bs = self._rows(x_t, t)
        tb = self.get_t_to_batch_size(x_t, t)
        return (x_t - bs * eps) / g + _bc(self.barsigmas[tb - 1], x_t) * eps, 0

    # ------------------------------------------------------------------ training, Prop. (9) (dlpm.py:384-401)
    def get_one_rv_faster_sampling(self, shape):
        return self.gen_a.generate(size=shape)

    def get_one_rv_loss_elements(self, t, x_0, a_t=None, z_t=None, state=None):
        """x_t = bg_t x_0 + sqrt(a_t bs_t^2) z_t ;  eps_t = (x_t - bg_t x_0) / bs_t  -- one fused kernel.
        ``a_t`` may be compact (B,) or full-shape (isotropic: constant per sample); None -> in-kernel."""
        if not self.isotropic:
            raise NotImplementedError("the fused training-elements kernel covers isotropic noise (all shipped configs)")
        dev = _lib.require_cuda(self.device)
        x0 = x_0.to(dev, torch.float32).contiguous()
        B = x0.shape[0]
        D = x0[0].numel()
        t = t.to(dev, torch.int64).contiguous()
        if a_t is not None:
            a_t = a_t.to(dev, torch.float32)
            a_t = a_t.reshape(B, -1)[:, 0].contiguous() if a_t.numel() != B else a_t.reshape(B).contiguous()
        if z_t is not None:
            z_t = z_t.to(dev, torch.float32).contiguous()
        st = state or rng.default_state()
        x_t = torch.empty_like(x0)
        eps_t = torch.empty_like(x0)
        with torch.cuda.device(dev):
            _lib.call("dlpm_b200_training_elements", _lib.ptr(x_t), _lib.ptr(eps_t), _lib.ptr(x0), _lib.ptr(t),
                      _lib.ptr(a_t), _lib.ptr(z_t), _lib.ptr(self.sched), self.diffusion_steps, B, D, float(self.alpha),
                      self._clamp_a(), st.seed, st.reserve(1), st.sample_base, _lib.stream_ptr())
        return x_t, eps_t
