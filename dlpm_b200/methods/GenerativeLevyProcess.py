"""``GenerativeLevyProcess``: the drop-in boundary.

Same constructor, ``sample`` / ``training_losses`` / ``get_timesteps`` entry points, kwargs and
attributes as ``dlpm/methods/GenerativeLevyProcess.py:35-709`` of the reference; the work is done by
libdlpm_b200.so:

  sample()  ->  K2 Sigma scan  ->  per step [score net (K4 / K5-K7)  ->  K3 fused update]
                (2-D nets: the whole chain is ONE persistent launch; image nets: one CUDA graph
                 replayed T-1 times with a device-side step counter)

Deliberate deviations from reference *quirks* (all documented in DESIGN.md):
  B.3  ``deterministic=True`` requires ``dlim_eta == 0`` (the eta != 0 branch of the reference is broken);
  B.4  after ``sample(reverse_steps != train steps)`` the original schedule IS restored.
Everything else (T-1 network evaluations, x_T scaled by barsigma_{T-1}, gen_sas ignoring clamp_a, ...)
is reproduced.
"""
import torch

from .. import _lib, rng
from . import lim as _lim
from .dlpm import DLPM, LossType, ModelMeanType, ModelVarType, match_last_dims  # noqa: F401
from .lim import LIM_sampler, VPSDE


def compute_loss_terms(x, y, lploss):
    """Per-sample loss terms (GenerativeLevyProcess.py:19-31) via the fused reduction kernel."""
    if lploss not in (2.0, 1.0, -1):
        raise NotImplementedError("lploss must be 2., 1. or -1 (generic p-norm is not on the hot path)")
    dev = _lib.require_cuda(y.device)
    B = x.shape[0]
    D = x[0].numel()
    flags = _lib.STEP_EPS_BF16 if x.dtype == torch.bfloat16 else 0
    xx = x.contiguous() if flags else x.to(torch.float32).contiguous()
    out = torch.empty(B, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.call("dlpm_b200_loss_terms", _lib.ptr(out), _lib.ptr(xx), _lib.ptr(y.to(torch.float32).contiguous()), B, D,
                  float(lploss), flags, _lib.stream_ptr())
    return out


def _torch_loss_terms(x, y, lploss):
    """GenerativeLevyProcess.py:19-31 in torch ops (differentiable): used when the model output carries an autograd graph."""
    dims = list(range(1, x.dim()))
    if lploss == 2.0:
        return torch.sqrt(torch.nn.functional.mse_loss(x, y, reduction="none").mean(dim=dims))
    if lploss == 1.0:
        return torch.nn.functional.smooth_l1_loss(x, y, beta=1, reduction="none").mean(dim=dims)
    if lploss == -1:
        return torch.nn.functional.mse_loss(x, y, reduction="none").mean(dim=dims)
    raise NotImplementedError("lploss must be 2., 1. or -1 (generic p-norm is not on the hot path)")


def _wants_grad(model):
    """True when the caller is TRAINING this module: grad mode on, module in train() mode, trainable parameters
    (bem/TrainingManager.py:111-120 calls ``loss.backward()`` right after ``training_losses``)."""
    return (torch.is_grad_enabled() and isinstance(model, torch.nn.Module) and model.training
            and any(p.requires_grad for p in model.parameters()))


class _Net:
    """Adapter around ``models['default']``: picks the native engine for this package's score nets (or
    reference modules that can be ingested), otherwise calls the user's module on the device.

    ``train=True`` (training_losses under autograd): a torch module -- the reference's ``UNetModel`` / ``MLPModel``
    included -- is called as is so that the loss keeps its graph; this package's own mirrors are parameter containers
    without a torch forward and there are no backward kernels in this tier, so they raise."""

    def __init__(self, model, device, train=False):
        from .. import score_nets
        if train:
            if getattr(model, "native_kind", None) is not None:
                raise NotImplementedError(
                    "dlpm_b200's %s runs forward-only CUDA kernels (no backward in this tier, dropout ignored): call "
                    "training_losses under torch.no_grad() / model.eval() for the loss value, or train a torch module "
                    "(e.g. the reference's own model class) -- its weights are ingested for sampling" % type(model).__name__)
            self.model, self.kind = model, "module"
            return
        self.model = score_nets.as_native(model, device)
        self.kind = getattr(self.model, "native_kind", "module")

    def __call__(self, x, t_vec, **kw):
        return self.model(x, t_vec, **kw)


class GenerativeLevyProcess:
    def __init__(self, alpha, device, reverse_steps, model_mean_type=ModelMeanType.EPSILON,
                 model_var_type=ModelVarType.FIXED, time_spacing="linear", rescale_timesteps=False, isotropic=True,
                 LIM=False, scale="scale_preserving", input_scaling=False):
        self.alpha = alpha
        self.device = device
        self.reverse_steps = reverse_steps
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.time_spacing = time_spacing
        self.rescale_timesteps = rescale_timesteps
        self.isotropic = isotropic
        self.LIM = LIM
        self.input_scaling = input_scaling
        assert (self.model_mean_type == ModelMeanType.EPSILON) and (self.model_var_type == ModelVarType.FIXED), \
            "Only epsilon prediction and fixed variance are supported for the moment"
        if self.LIM:
            assert self.rescale_timesteps == True, \
                "LIM only supports epsilon prediction, fixed variance and rescaled timesteps"  # noqa: E712
            self.sde = VPSDE(alpha, "cosine")
            self.levy = None  # the reference instantiates torchlevy.LevyStable here but never calls it
        self.dlpm = DLPM(alpha, device, diffusion_steps=reverse_steps, time_spacing=time_spacing, isotropic=isotropic,
                         scale=scale)

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1.0 / self.reverse_steps)
        return t

    def _input_scale_table(self):
        """1 / (1 + barsigma_t) as a device (T,) table when the reference scales the model input
        (GenerativeLevyProcess.py:177-180, :651-654: ``input_scaling`` with the 'scale_exploding' schedule), else None."""
        if not (self.input_scaling and self.dlpm.scale == "scale_exploding"):
            return None
        return (1 / (1 + self.dlpm._sched_host[:, 3])).contiguous().to(self.device)

    def _scaled_input(self, x, table, t_vec=None, t=0):
        """x * table[t] (per-sample t_vec, or one batch-constant step) through dlpm_b200_scale_by_step."""
        x = x.contiguous()
        out = torch.empty_like(x)
        B = x.shape[0]
        _lib.call("dlpm_b200_scale_by_step", _lib.ptr(out), _lib.ptr(x), _lib.ptr(table),
                  None if t_vec is None else _lib.ptr(t_vec.to(torch.int64).contiguous()), int(t), None, self.reverse_steps, B,
                  x[0].numel(), _lib.stream_ptr())
        return out

    def get_timesteps(self, N, **kwargs):
        return self.dlpm.get_timesteps(N)

    def q_sample(self, x_start, t, eps=None):
        return self.dlpm.sample_x_t_from_xstart(x_start, t, eps)

    # ------------------------------------------------------------------------------------------ sampling
    def p_mean_variance(self, model, x, t, clip_denoised=False, denoised_fn=None, model_kwargs=None):
        """API-parity single step (GenerativeLevyProcess.py:154-219); the loops below use the fused kernels."""
        assert denoised_fn is None, "denoised_fn is not supported"
        B = x.shape[0]
        assert t.shape == (B,)
        net = _Net(model, self.device)
        table = self._input_scale_table()
        x_in = x if table is None else self._scaled_input(x.to(torch.float32), table, t_vec=t)
        eps = net(x_in, self._scale_timesteps(t), **(model_kwargs or {})).to(torch.float32).reshape(x.shape)
        if clip_denoised:
            eps = self.dlpm.predict_eps(x, t, self.dlpm.predict_xstart(x, t, eps).clamp(-1, 1))
        mean, var = self.dlpm.anterior_mean_variance_dlpm(x, t[0], eps)
        assert mean.shape == x.shape
        return {"eps": eps, "mean": mean, "variance": var}

    # ------------------------------------------------------------------------------------------ single-step API (:225-239, :332-373)
    def q_posterior_mean_variance(self, x_start, x_t, t):
        raise NotImplementedError("dead code in the reference (GenerativeLevyProcess.py:126-127 calls methods that do not exist)")

    def _single_step(self, model, x, t, clip_denoised, model_kwargs, deterministic, noise, state, z_offset):
        dev = _lib.require_cuda(self.device)
        B = x.shape[0]
        assert t.shape == (B,)
        d = self.dlpm
        T = self.reverse_steps
        tt = int(t[0])  # the reference indexes the posterior with t[0] (:210): the step is batch-constant
        assert 1 <= tt < T, "t out of range [1, T)"
        st = state or rng.default_state()
        net = _Net(model, dev)
        with torch.inference_mode(), torch.cuda.device(dev):
            xs = x.to(dev, torch.float32).contiguous()
            table = self._input_scale_table()
            x_in = xs if table is None else self._scaled_input(xs, table, t=tt)
            eps = net(x_in, self._scale_timesteps(t.to(dev)), **(model_kwargs or {}))
            fl = (_lib.STEP_CLIP_DENOISED if clip_denoised else 0) | (_lib.STEP_EPS_BF16 if eps.dtype == torch.bfloat16 else 0)
            eps = eps.contiguous() if eps.dtype == torch.bfloat16 else eps.to(torch.float32).contiguous()
            out = xs.clone()
            D = out[0].numel()
            if deterministic:
                _lib.call("dlpm_b200_dlim_step", _lib.ptr(out), _lib.ptr(eps), _lib.ptr(d.sched), tt, None, T, B, D, fl, None,
                          _lib.stream_ptr())
            else:
                assert d.Sigmas is not None, "sample_A / compute_Sigmas must run first (as in the reference, :308-309)"
                fl |= 0 if d.isotropic else _lib.STEP_SIGMA_FULL
                z = None if noise is None else noise.to(dev, torch.float32).contiguous()
                # one reserved block of T offsets per call: the kernel keys z by (offset + t), unique for every (call, t)
                zo = st.reserve(T) if z_offset is None else z_offset
                _lib.call("dlpm_b200_reverse_step", _lib.ptr(out), _lib.ptr(eps), _lib.ptr(d.Sigmas), _lib.ptr(d.sched), tt, None, T,
                          B, D, fl, _lib.ptr(z), st.seed, zo, st.sample_base, None, _lib.stream_ptr())
        return {"sample": out.view(x.shape)}

    def p_sample(self, model, x, t, clip_denoised=False, denoised_fn=None, model_kwargs=None, noise=None, state=None,
                 _z_offset=None):
        """:225-239: x_{t-1} = mean + 1[t != 1] sqrt(var) z -- one network evaluation + the fused K3 step.
        ``noise`` (extension) injects z for parity tests; otherwise it is drawn in-kernel."""
        assert denoised_fn is None, "denoised_fn is not supported"
        return self._single_step(model, x, t, clip_denoised, model_kwargs, False, noise, state, _z_offset)

    def ddim_sample(self, model, x, t, clip_denoised=False, denoised_fn=None, model_kwargs=None, eta=0.0):
        """:332-373 with eta = 0 (App. B.3): x_{t-1} = (x_t - bs_t eps) / g_t + bs_{t-1} eps."""
        assert denoised_fn is None, "denoised_fn is not supported"
        if eta != 0.0:
            raise NotImplementedError("dlim_eta != 0 is broken in the reference (dlpm.py:289-297); use eta=0.0")
        return self._single_step(model, x, t, clip_denoised, model_kwargs, True, None, None, None)

    def _progressive(self, model, shape, noise, clip_denoised, model_kwargs, deterministic, state):
        dev = _lib.require_cuda(self.device)
        assert isinstance(shape, (tuple, list))
        shape = [int(s) for s in shape]
        st = state or rng.default_state()
        T = self.reverse_steps
        self.dlpm.sample_A(shape, T, state=st)
        if noise is not None:
            img = noise.to(dev, torch.float32)
        else:
            img = self.dlpm.gen_eps.generate(size=shape, scale=float(self.dlpm._sched_host[-1, 3]), state=st)
        yield {"sample": img}
        z_offset = st.reserve(T)
        for i in range(T - 1, 0, -1):
            t = torch.tensor([i] * shape[0], device=dev)
            if deterministic:
                out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs, eta=0.0)
            else:
                out = self.p_sample(model, img, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs, state=st,
                                    _z_offset=z_offset)
            yield out
            img = out["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=False, denoised_fn=None, model_kwargs=None,
                                  state=None):
        """:291-330: generator over {'sample': x_t} for t = T-1 ... 0 (T entries).  Same Philox stream layout as
        ``p_sample_loop`` (A chain, x_T, then one block of T offsets for z), so both produce the same samples."""
        assert denoised_fn is None, "denoised_fn is not supported"
        return self._progressive(model, shape, noise, clip_denoised, model_kwargs, False, state)

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=False, denoised_fn=None, model_kwargs=None,
                                     eta=0.0, state=None):
        """:413-452 with eta = 0."""
        assert denoised_fn is None, "denoised_fn is not supported"
        if eta != 0.0:
            raise NotImplementedError("dlim_eta != 0 is broken in the reference (dlpm.py:289-297); use eta=0.0")
        return self._progressive(model, shape, noise, clip_denoised, model_kwargs, True, state)

    def _reverse_loop(self, model, shape, noise, clip_denoised, deterministic, get_sample_history, model_kwargs=None,
                      progress=False, injected_A=None, injected_z=None, state=None, postprocess=None):
        """p_sample_loop_progressive / ddim_sample_loop_progressive (:291-330, :413-452) on the fused kernels."""
        dev = _lib.require_cuda(self.device)
        assert isinstance(shape, (tuple, list))
        shape = [int(s) for s in shape]
        B = shape[0]
        D = 1
        for s in shape[1:]:
            D *= s
        T = self.reverse_steps
        d = self.dlpm
        if B == 0:  # same error as the reference, which indexes t[0] of the empty batch (GenerativeLevyProcess.py:210)
            raise IndexError("index 0 is out of bounds for dimension 0 with size 0")
        st = state or rng.default_state()
        net = _Net(model, dev)
        if hasattr(net.model, "eval"):
            net.model.eval()
        with torch.inference_mode(), torch.cuda.device(dev):
            # (a)+(b) A_{0:T-1} and the Sigma recursion (dlpm.py:226-239), one scan kernel
            if injected_A is not None:
                want = (T, B) if d.isotropic else (T, *shape)
                if tuple(injected_A.shape) == (T, *shape) and d.isotropic:
                    # the reference's full-shape layout (dlpm.py:227): isotropic A is constant per sample -> compact (T, B)
                    injected_A = injected_A.reshape(T, B, -1)[:, :, 0]
                assert tuple(injected_A.shape) == want, "injected_A must have shape %s (got %s)" % (want, tuple(injected_A.shape))
                d.A = injected_A.to(dev, torch.float32).contiguous()
                d._shape = shape
                d._sigma_src = None
                d.compute_Sigmas()
            else:
                d.sample_A(shape, T, state=st)
            # (c) x_{T-1} = barsigma_{T-1} * eps  (:313), scale folded into the noise kernel
            if noise is not None:
                x = noise.to(dev, torch.float32).contiguous().clone()
            else:
                x = d.gen_eps.generate(size=shape, scale=float(d._sched_host[-1, 3]), state=st)
            z = None if injected_z is None else injected_z.to(dev, torch.float32).contiguous()
            hist = None
            if get_sample_history:
                hist = torch.empty((T, *shape), device=dev, dtype=torch.float32)
                hist[0].copy_(x)
            flags = (_lib.STEP_CLIP_DENOISED if clip_denoised else 0) | (0 if d.isotropic else _lib.STEP_SIGMA_FULL)
            z_offset = st.reserve(T)
            mode = 1 if deterministic else 0
            in_scale = self._input_scale_table()
            post = None if postprocess is None else postprocess.bind(shape, dev)  # fused into the LAST step's store
            # (d) the hot loop
            if net.kind == "mlp" and not model_kwargs and d.isotropic and self.rescale_timesteps and in_scale is None:
                m = net.model
                _lib.call("dlpm_b200_mlp_sample_chain", _lib.ptr(x), _lib.ptr(m.packed_weights()), _lib.ptr(d.Sigmas),
                          _lib.ptr(d.sched), T, B, m.nfeatures, m.nunits, m.time_emb_size, m.nblocks_total, mode,
                          flags & _lib.STEP_CLIP_DENOISED, _lib.ptr(z), _lib.ptr(hist), st.seed, z_offset,
                          st.sample_base, _lib.stream_ptr())
                if postprocess is not None:  # the persistent 2-D chain kernel keeps x in registers: one tiny extra launch
                    postprocess.run_standalone(x)
            elif net.kind == "unet" and not model_kwargs and z is None and self.rescale_timesteps:
                net.model.sample_loop(x, d, T, mode, flags, hist, st.seed, z_offset, st.sample_base,
                                      progress=progress, input_scale=in_scale, post=post)
            else:
                bar = None
                if progress:
                    from tqdm import tqdm
                    bar = tqdm(total=T)
                for k, t in enumerate(range(T - 1, 0, -1)):
                    tv = torch.full((B,), t, device=dev, dtype=torch.int64)
                    x_in = x if in_scale is None else self._scaled_input(x, in_scale, t=t)
                    eps = net(x_in.view(shape), self._scale_timesteps(tv), **(model_kwargs or {}))
                    fl = flags | (_lib.STEP_EPS_BF16 if eps.dtype == torch.bfloat16 else 0)
                    eps = eps.contiguous() if eps.dtype == torch.bfloat16 else eps.to(torch.float32).contiguous()
                    h = _lib.ptr(hist[k + 1]) if hist is not None else None
                    if deterministic:
                        _lib.call("dlpm_b200_dlim_step_post", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(d.sched), t, None, T, B, D,
                                  fl, h, post, _lib.stream_ptr())
                    else:
                        _lib.call("dlpm_b200_reverse_step_post", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(d.Sigmas),
                                  _lib.ptr(d.sched), t, None, T, B, D, fl, _lib.ptr(z[k]) if z is not None else None,
                                  st.seed, z_offset, st.sample_base, h, post, _lib.stream_ptr())
                    if bar is not None:
                        bar.update(1)
                if bar is not None:
                    bar.close()
        x = x.view(shape)
        if get_sample_history:
            return x, hist
        return x

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=False, denoised_fn=None, model_kwargs=None,
                      progress=False, get_sample_history=False, injected_A=None, injected_z=None, state=None,
                      postprocess=None):
        """:241-289 (+ ``injected_A`` (T,B) / ``injected_z`` (T-1,*shape) for parity tests; ``postprocess``: see ``sample``)."""
        assert denoised_fn is None, "denoised_fn is not supported"
        return self._reverse_loop(model, shape, noise, clip_denoised, False, get_sample_history, model_kwargs, progress,
                                  injected_A, injected_z, state, postprocess)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=False, denoised_fn=None, model_kwargs=None,
                         progress=False, eta=0.0, get_sample_history=False, injected_A=None, state=None, postprocess=None):
        """:375-411.  eta must be 0."""
        assert denoised_fn is None, "denoised_fn is not supported"
        if eta != 0.0:
            raise NotImplementedError("dlim_eta != 0 is broken in the reference (dlpm.py:289-297); use dlim_eta=0.0")
        return self._reverse_loop(model, shape, noise, clip_denoised, True, get_sample_history, model_kwargs, progress,
                                  injected_A, None, state, postprocess)

    def lim_sample(self, model, shape, ddim=False, get_sample_history=False, clip_denoised=False, injected_x=None,
                   injected_noise=None, state=None, postprocess=None):
        """:454-506.  x_T ~ SaS is NOT scaled by barsigma (:464)."""
        dev = _lib.require_cuda(self.device)
        st = state or rng.default_state()
        net = _Net(model, dev)
        x = injected_x if injected_x is not None else self.dlpm.gen_eps.generate(size=shape, state=st)
        shape = list(x.shape)
        return LIM_sampler(ddim=ddim, x=x, y=None, model=net.model, sde=self.sde, levy=self.levy,
                           isotropic=self.isotropic, steps=self.reverse_steps, gen_a=self.dlpm.gen_a,
                           gen_eps=self.dlpm.gen_eps, device=dev, get_sample_history=get_sample_history,
                           injected_noise=injected_noise, net_call=lambda xx, tt: net(xx.view(shape), tt), state=st,
                           postprocess=postprocess)

    def sample(self, models, shape, reverse_steps, time_spacing=None, initial_data=None, clip_denoised=False,
               deterministic=False, dlim_eta=1.0, print_progression=False, get_sample_history=False, clamp_a=None,
               clamp_eps=None, postprocess=None):
        """Boundary entry point (:512-569); called by ``bem/GenerationManager.py:43-47``.

        ``postprocess`` (extension, default None = reference behaviour): a ``dlpm_b200.generation.FusedPost``; the
        caller's clamp / (x+1)/2 / uint8 quantisation (bem/GenerationManager.py:50-63) is then written by the LAST step
        kernel into ``postprocess.out`` while the returned tensor still holds the unprocessed x_0."""
        self.dlpm.gen_a.setParams(clamp_a=clamp_a)
        self.dlpm.gen_eps.setParams(clamp_eps=clamp_eps)
        model = models["default"]
        assert time_spacing is None, "Specific time spacing is not yet supported for diffusion reverse sampling"
        default_reverse_steps = self.reverse_steps
        rescaled = self.reverse_steps != reverse_steps
        if rescaled:
            assert self.rescale_timesteps, "Rescaling only works when rescale_timesteps is True"
            snapshot = self.dlpm.snapshot_schedule()
            # like the reference (dlpm.py:176-185, App. B.5) the rescaled schedule is always the cosine 'scale_preserving'
            # one, whatever self.dlpm.scale says (and input_scaling, if on, then uses the cosine barsigma: same quirk)
            self.dlpm.rescale_diffusion(reverse_steps, time_spacing=time_spacing)
            self.reverse_steps = reverse_steps
        try:
            if self.LIM:
                x = self.lim_sample(model, shape=shape, ddim=deterministic, get_sample_history=get_sample_history,
                                    clip_denoised=clip_denoised, postprocess=postprocess)
            elif deterministic:
                x = self.ddim_sample_loop(model, shape=initial_data.shape if initial_data is not None else shape,
                                          noise=initial_data, eta=dlim_eta, progress=print_progression,
                                          get_sample_history=get_sample_history, clip_denoised=clip_denoised,
                                          postprocess=postprocess)
            else:
                x = self.p_sample_loop(model, shape=shape, progress=print_progression,
                                       get_sample_history=get_sample_history, clip_denoised=clip_denoised,
                                       postprocess=postprocess)
        finally:
            if rescaled:  # the reference's restore guard never fires (App. B.4); we do restore -- the exact tables, whatever
                self.dlpm.restore_schedule(snapshot)  # their family ('scale_exploding' included)
                self.reverse_steps = default_reverse_steps
        return x

    # ------------------------------------------------------------------------------------------ training (forward + loss)
    def training_losses(self, models, x_start, model_kwargs=None, **kwargs):
        """:581-609.  The noise / x_t / eps_t elements always come from the fused kernels.  A torch module in train()
        mode under autograd (the reference's TrainingManager, bem/TrainingManager.py:111-120) is evaluated by torch and
        the loss is differentiable; this package's native nets are forward-only (no backward kernels in
        this tier) and raise NotImplementedError when asked to train."""
        model = models["default"]
        x_start = x_start.to(self.device)
        if model_kwargs is None:
            model_kwargs = {}
        if self.LIM:
            loss = self.training_losses_lim(model, x_start, **model_kwargs, **kwargs)
        else:
            loss = self.training_losses_dlpm(model, x_start, **model_kwargs, **kwargs)
        return {"loss": loss}

    def training_losses_dlpm(self, model, x_start, loss_type="EPSILON", lploss=2.0, loss_monte_carlo="mean",
                             monte_carlo_outer=1, monte_carlo_inner=1, model_kwargs=None, clamp_a=None, clamp_eps=None,
                             injected=None, state=None):
        """:612-677 with Proposition (9)'s one-r.v. elements.  ``injected`` = dict(t, A, z) for parity tests."""
        assert self.model_mean_type == ModelMeanType.EPSILON, "only epsilon model output is supported for the moment"
        assert loss_type == LossType.EPS_LOSS, "only epsilon loss is supported for the moment"
        dev = _lib.require_cuda(self.device)
        model_kwargs = model_kwargs or {}
        self.dlpm.gen_a.setParams(clamp_a=clamp_a)
        self.dlpm.gen_eps.setParams(clamp_eps=clamp_eps)
        st = state or rng.default_state()
        inj = injected or {}
        n0 = len(x_start)
        t = inj["t"].to(dev) if "t" in inj else torch.randint(1, self.reverse_steps, size=[n0]).to(dev)
        total = monte_carlo_outer * monte_carlo_inner
        x_ext = x_start.repeat(total, *([1] * len(x_start.shape[1:])))
        t_ext = t.repeat(total)
        A_ext = None
        if "A" in inj:
            A = inj["A"].to(dev, torch.float32).reshape(n0 * monte_carlo_outer, -1)[:, 0]
            A_ext = A.repeat(monte_carlo_inner)
        elif monte_carlo_inner > 1:
            from ..datasets.Distributions import gen_skewed_levy
            A = gen_skewed_levy(self.alpha, [n0 * monte_carlo_outer], device=dev, isotropic=True, clamp_a=clamp_a,
                                compact=True, state=st)
            A_ext = A.repeat(monte_carlo_inner)
        x_t, eps_t = self.dlpm.get_one_rv_loss_elements(t_ext, x_ext, A_ext, inj.get("z"), state=st)
        train = _wants_grad(model)
        net = _Net(model, dev, train=train)
        table = self._input_scale_table()
        x_in = x_t if table is None else self._scaled_input(x_t, table, t_vec=t_ext)
        model_eps = net(x_in, self._scale_timesteps(t_ext), **model_kwargs)
        if model_eps.requires_grad:  # a torch module under autograd: differentiable loss terms, as the reference computes them
            losses = _torch_loss_terms(model_eps.reshape(x_t.shape), eps_t, lploss)
        else:
            losses = compute_loss_terms(model_eps.reshape(x_t.shape), eps_t, lploss)
        assert not torch.isnan(losses).any(), "Nan in losses"
        if loss_monte_carlo == "mean":
            return losses.mean()
        elif loss_monte_carlo == "median":
            losses = losses.reshape(monte_carlo_outer, monte_carlo_inner, x_start.shape[0]).mean(dim=1)
            losses, _ = losses.median(dim=0)
            return losses.mean()
        raise NotImplementedError(loss_monte_carlo)

    def training_losses_lim(self, model, x_start, y=None, clamp_a=None, clamp_eps=None, injected=None, state=None):
        """:680-709 + LIM/functions/loss.py:13-39 (forward + loss): t ~ U(1e-5, T_max), e ~ SaS drawn in-kernel,
        x_t = x0 exp(l_t) + e sigma_t, target score = -e / alpha, loss = mean smooth-L1(model(x_t, t), score).
        ``injected`` = dict(u, e) (the uniform variates and the noise) for parity tests."""
        if self.sde.alpha == 2.0:
            raise NotImplementedError("LIM with alpha == 2 (plain Gaussian VPSDE branch) is out of scope")
        assert y is None, "class-conditional LIM training is not on the hot path"
        dev = _lib.require_cuda(self.device)
        self.dlpm.gen_a.setParams(clamp_a=clamp_a)
        self.dlpm.gen_eps.setParams(clamp_eps=clamp_eps)
        st = state or rng.default_state()
        inj = injected or {}
        x0 = x_start.to(dev, torch.float32).contiguous()
        n = x0.size(0)
        D = x0[0].numel()
        start_eps = 1e-5
        u = inj["u"].to(dev, torch.float32) if "u" in inj else torch.rand(n).to(dev)
        t = (u * (self.sde.T - start_eps) + start_eps).contiguous()
        e = inj["e"].to(dev, torch.float32).contiguous() if "e" in inj else None
        x_t = torch.empty_like(x0)
        score = torch.empty_like(x0)
        with torch.cuda.device(dev):
            _lib.call("dlpm_b200_lim_training_elements", _lib.ptr(x_t), _lib.ptr(score), _lib.ptr(x0), _lib.ptr(t), _lib.ptr(e), n, D,
                      float(self.sde.alpha), 1 if self.isotropic else 0, -1.0 if clamp_eps is None else float(clamp_eps), st.seed,
                      st.reserve(1), st.sample_base, _lib.stream_ptr())
        net = _Net(model, dev, train=_wants_grad(model))
        output = net(x_t, t)
        if output.requires_grad:
            losses = _torch_loss_terms(output.reshape(x_t.shape), score, 1.0)
        else:
            losses = compute_loss_terms(output.reshape(x_t.shape), score, 1.0)  # per-sample mean smooth-L1 (beta = 1)
        assert not torch.isnan(losses).any(), "Nan in losses"
        return losses.mean()  # all samples have D elements: mean of per-sample means == F.smooth_l1_loss(..., 'mean')
