"""Run the LIVE reference (``/root/reference`` or ``oracle/_ref``) with injected noise.  TEST INFRASTRUCTURE ONLY.

Noise is injected by monkeypatching only (no reference edits), exactly as ``tests/golden/make_golden.py`` does for the
committed fixtures: ``glp.dlpm.gen_a.generate`` / ``gen_eps.generate`` pop prepared tensors and the module attribute
``dlpm.methods.GenerativeLevyProcess.th`` is replaced by a proxy whose ``randn_like`` pops the injected z.  Works on any
device: on the GPU box ``device='cuda'`` gives the reference's own GPU path (PyTorch eager + cuDNN).
"""
import contextlib
import types

import numpy as np
import torch

from . import ref_import, stable


class TorchProxy(types.ModuleType):
    """Stands in for ``torch`` / ``th`` inside the reference module: ``randn_like`` pops injected tensors."""

    def __init__(self, randn_list):
        super().__init__("torch_proxy")
        self._randn = randn_list

    def __getattr__(self, name):
        return getattr(torch, name)

    def randn_like(self, x, **kw):
        return self._randn.pop(0).to(x.device, x.dtype)


@contextlib.contextmanager
def strict_fp32():
    """The north star's fp32 oracle: cuDNN / cuBLAS TF32 off (PyTorch enables TF32 convolutions by default and the
    reference never changes it)."""
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = c, m


def draw_inputs(alpha, shape, T, seed, clamp_a=None):
    """Injected tensors of a DLPM chain: A compact (T, B) from the oracle's scipy-equivalent CMS sampler, eps_init ~ SaS of
    the sample shape, z (T-1, *shape) ~ N(0, I).  Deterministic in ``seed``; small enough to regenerate anywhere."""
    B = shape[0]
    rs = np.random.RandomState(seed)
    A = torch.stack([torch.from_numpy(stable.gen_skewed_levy(alpha, (B,), isotropic=True, clamp_a=clamp_a, rng=rs).copy())
                     for _ in range(T)]).float()
    eps_init = torch.from_numpy(stable.gen_sas(alpha, tuple(shape), isotropic=True, rng=rs)).float()
    z = torch.randn((T - 1,) + tuple(shape), generator=torch.Generator().manual_seed(seed))
    return A, eps_init, z


def reference_dlpm_chain(model, shape, alpha, T, A_compact, eps_init, z, device="cpu", deterministic=False, clip_denoised=False,
                         keep=None):
    """Reference ``p_sample_loop`` / ``ddim_sample_loop`` (GenerativeLevyProcess.py:241-289, :375-411) with injected
    A (T, B), eps_init, z (T-1, *shape).  Returns (x_init, history (T, *shape) on the CPU) -- or, with ``keep`` = iterable
    of history indices, only those rows (dict index -> tensor): a T = 1000 history of a large batch need not be kept."""
    ns = ref_import.load()
    B = shape[0]
    glp = ns.glp.GenerativeLevyProcess(alpha, device, T, rescale_timesteps=True, isotropic=True)
    A_list = [a.view(B, *([1] * (len(shape) - 1))).expand(*shape).contiguous().to(device) for a in A_compact]
    eps_d = eps_init.to(device)
    glp.dlpm.gen_a.generate = lambda *a, **k: A_list.pop(0)
    glp.dlpm.gen_eps.generate = lambda *a, **k: eps_d
    ns.glp.th = TorchProxy(list(z))
    try:
        x_init = (glp.dlpm.barsigmas[-1] * eps_d).cpu()
        with torch.inference_mode():
            model.eval()
            gen = (glp.ddim_sample_loop_progressive(model, list(shape), eta=0.0, clip_denoised=clip_denoised) if deterministic
                   else glp.p_sample_loop_progressive(model, list(shape), clip_denoised=clip_denoised))
            rows = {} if keep is not None else []
            keep = None if keep is None else set(int(k) for k in keep)
            for k, out in enumerate(gen):
                if keep is None:
                    rows.append(out["sample"].cpu())
                elif k in keep:
                    rows[k] = out["sample"].cpu()
    finally:
        ns.glp.th = torch
    return x_init, (rows if isinstance(rows, dict) else torch.stack(rows))


def make_unet(cfg, in_ch, device, seed):
    """Reference UNetModel as dlpm_experiment.py:41-56 builds it, parameters re-randomised by the name-keyed recipe."""
    from dlpm_b200.init_utils import randomize_parameters_
    ns = ref_import.load()
    m = ns.unet.UNetModel(in_channels=in_ch, model_channels=cfg["model_channels"], out_channels=in_ch,
                          num_res_blocks=cfg["num_res_blocks"], attention_resolutions=cfg["attention_resolutions"], dropout=0.0,
                          channel_mult=cfg["channel_mult"], dims=2, num_classes=None, use_checkpoint=False,
                          num_heads=cfg["num_heads"], num_heads_upsample=-1, use_scale_shift_norm=True)
    randomize_parameters_(m, seed)
    return m.to(device).eval()


def mlp_params(nblocks=4, nunits=64, temb=32, device="cpu"):
    """The ``p`` dict of dlpm/configs/2d_data.yml:71-88 as ``MLPModel.__init__`` reads it (Model.py:24-42)."""
    return {"data": {"nfeatures": 2}, "method": "dlpm", "dlpm": {"isotropic": True}, "device": device,
            "model": dict(use_a_t=False, no_a=True, a_pos_emb=False, a_emb_size=32, time_emb_type="learnable",
                          time_emb_size=temb, nblocks=nblocks, nunits=nunits, skip_connection=True,
                          group_norm=True, dropout_rate=0.0, learn_variance=False)}


def make_mlp(device, seed=0, **kw):
    ns = ref_import.load()
    torch.manual_seed(seed)
    return ns.Model.MLPModel(mlp_params(device=str(device), **kw)).to(device).eval()
