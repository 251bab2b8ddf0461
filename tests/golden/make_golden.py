"""Generate golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference`` through ``oracle/ref_import.py`` (stub modules for absent,
non-hot-path dependencies), drives ``GenerativeLevyProcess`` / ``Generator`` /
``UNetModel`` / ``MLPModel`` with *injected* noise (monkeypatching only -- no reference
edits) and stores inputs + outputs as small ``.npz`` fixtures next to this file.  The
reference ships no tests or golden vectors; these files are the pin for
``oracle/`` (``tests/test_oracle_golden.py``) and for the CUDA path (``-m gpu`` tests).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

ns = ref_import.load()
torch.set_num_threads(8)


from dlpm_b200.init_utils import randomize_parameters_ as rerandomize_, parameter_checksum  # noqa: E402


def sd_np(module):
    return {"sd/" + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def mlp_params(nblocks=4, nunits=64, temb=32):
    return {"data": {"nfeatures": 2}, "method": "dlpm", "dlpm": {"isotropic": True}, "device": "cpu",
            "model": dict(use_a_t=False, no_a=True, a_pos_emb=False, a_emb_size=32, time_emb_type="learnable",
                          time_emb_size=temb, nblocks=nblocks, nunits=nunits, skip_connection=True,
                          group_norm=True, dropout_rate=0.0, learn_variance=False)}


def make_unet(cfg, in_ch):
    return ns.unet.UNetModel(in_channels=in_ch, model_channels=cfg["model_channels"], out_channels=in_ch,
                             num_res_blocks=cfg["num_res_blocks"],
                             attention_resolutions=cfg["attention_resolutions"], dropout=0.0,
                             channel_mult=cfg["channel_mult"], dims=2, num_classes=None, use_checkpoint=False,
                             num_heads=cfg["num_heads"], num_heads_upsample=-1, use_scale_shift_norm=True)


class TorchProxy(types.ModuleType):
    """Stands in for ``torch``/``th`` inside the reference module: pops injected tensors."""

    def __init__(self, randn_list=None, randint_list=None, rand_list=None):
        super().__init__("torch_proxy")
        self._randn = list(randn_list or [])
        self._randint = list(randint_list or [])
        self._rand = list(rand_list or [])

    def __getattr__(self, name):
        return getattr(torch, name)

    def randn_like(self, x, **kw):
        return self._randn.pop(0)

    def randint(self, *a, **kw):
        return self._randint.pop(0)

    def rand(self, *a, **kw):
        return self._rand.pop(0)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote", name, "%.1f KB" % (os.path.getsize(path) / 1024))


# ------------------------------------------------------------------------------------------
def golden_noise():
    out = {}
    for alpha in (1.5, 1.7, 1.9, 2.0):
        for iso in (True, False):
            np.random.seed(1234)
            torch.manual_seed(1234)
            a = ns.Distributions.gen_skewed_levy(alpha, (64, 3, 4), isotropic=iso, clamp_a=20.0 if iso else None)
            np.random.seed(77)
            torch.manual_seed(77)
            e = ns.Distributions.gen_sas(alpha, (64, 3, 4), isotropic=iso, clamp_eps=50.0)
            torch.manual_seed(77)
            g = torch.randn(size=(64, 3, 4))
            tag = "a%.1f_%s" % (alpha, "iso" if iso else "full")
            out["A_" + tag] = a.numpy()
            out["eps_" + tag] = e.numpy()
            out["G_" + tag] = g.numpy()
    save("noise", **out)


def golden_schedule():
    out = {}
    for alpha in (1.5, 1.7, 1.9):
        for T in (10, 100, 1000):
            for spacing in ("linear", "quadratic"):
                d = ns.dlpm.DLPM(alpha, "cpu", T, time_spacing=spacing)
                tag = "a%.1f_T%d_%s" % (alpha, T, spacing)
                out[tag] = torch.stack([d.gammas, d.bargammas, d.sigmas, d.barsigmas]).numpy()
    d = ns.dlpm.DLPM(1.7, "cpu", 100, scale="scale_exploding")
    out["a1.7_T100_exploding"] = torch.stack([d.gammas, d.bargammas, d.sigmas, d.barsigmas]).numpy()
    save("schedule", **out)


def run_dlpm(model, shape, alpha, T, seed, deterministic=False, clip_denoised=False, clamp_a=None):
    """Reference p_sample_loop / ddim_sample_loop with injected A, x_init, z."""
    B = shape[0]
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    from oracle import stable
    A_compact = torch.stack([torch.from_numpy(stable.gen_skewed_levy(alpha, (B,), isotropic=True, clamp_a=clamp_a, rng=rs).copy())
                             for _ in range(T)])  # (T, B)
    eps_init = torch.from_numpy(stable.gen_sas(alpha, shape, isotropic=True, rng=rs))
    z = torch.randn((T - 1,) + tuple(shape), generator=g)
    glp = ns.glp.GenerativeLevyProcess(alpha, "cpu", T, rescale_timesteps=True, isotropic=True)
    A_list = [a.view(B, *([1] * (len(shape) - 1))).expand(*shape).contiguous() for a in A_compact]
    glp.dlpm.gen_a.generate = lambda *a, **k: A_list.pop(0)
    glp.dlpm.gen_eps.generate = lambda *a, **k: eps_init
    proxy = TorchProxy(randn_list=list(z))
    ns.glp.th = proxy
    try:
        x_init = glp.dlpm.barsigmas[-1] * eps_init
        if deterministic:
            final, hist = glp.ddim_sample_loop(model, shape, eta=0.0, get_sample_history=True,
                                               clip_denoised=clip_denoised)
        else:
            final, hist = glp.p_sample_loop(model, shape, get_sample_history=True, clip_denoised=clip_denoised)
    finally:
        ns.glp.th = torch
    return dict(A=A_compact.numpy(), eps_init=eps_init.numpy(), x_init=x_init.numpy(), z=z.numpy(),
                final=final.numpy(), hist=hist.numpy(), Sigmas=glp.dlpm.Sigmas.reshape(T, B, -1)[:, :, 0].numpy())


def run_lim(model, shape, alpha, steps, seed, ode):
    rs = np.random.RandomState(seed)
    from oracle import stable
    noises = [torch.from_numpy(stable.gen_sas(alpha, shape, isotropic=True, rng=rs)) for _ in range(steps + 1)]
    x_init = noises[0]
    e_L = torch.stack(noises[1:])
    glp = ns.glp.GenerativeLevyProcess(alpha, "cpu", steps, rescale_timesteps=True, isotropic=True, LIM=True)
    pool = list(noises)
    glp.dlpm.gen_eps.generate = lambda *a, **k: pool.pop(0)
    model.eval()
    final, hist = glp.lim_sample(model, shape, ddim=ode, get_sample_history=True)
    return dict(x_init=x_init.numpy(), e_L=e_L.numpy(), final=final.numpy(), hist=hist.numpy())


def golden_mlp_chain():
    torch.manual_seed(0)
    model = rerandomize_(ns.Model.MLPModel(mlp_params()), 11).eval()
    out = sd_np(model)
    shape = (16, 1, 2)
    for tag, kw in (("dlpm", {}), ("dlpm_clip", dict(clip_denoised=True)), ("dlpm_clampa", dict(clamp_a=20.0)),
                    ("dlim", dict(deterministic=True))):
        r = run_dlpm(model, shape, 1.7, 50, seed=5, **kw)
        out.update({tag + "/" + k: v for k, v in r.items()})
    for tag, ode in (("lim_sde", False), ("lim_ode", True)):
        r = run_lim(model, shape, 1.7, 20, seed=6, ode=ode)
        out.update({tag + "/" + k: v for k, v in r.items()})
    # single forward at assorted t
    x = torch.randn(32, 1, 2, generator=torch.Generator().manual_seed(3)) * 2
    t = torch.rand(32, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        out["fwd/x"], out["fwd/t"], out["fwd/y"] = x.numpy(), t.numpy(), model(x, t).numpy()
    # training loss (Prop. 9), injected t / A / z
    B, T = 32, 100
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, 1, 2, generator=g)
    tt = torch.randint(1, T, size=[B], generator=g)
    from oracle import stable
    A = torch.from_numpy(stable.gen_skewed_levy(1.7, (B, 1, 2), isotropic=True, rng=np.random.RandomState(9)))
    zz = torch.randn(B, 1, 2, generator=g)
    glp = ns.glp.GenerativeLevyProcess(1.7, "cpu", T, rescale_timesteps=True, isotropic=True)
    glp.dlpm.gen_a.generate = lambda *a, **k: A
    ns.glp.torch = TorchProxy(randn_list=[zz], randint_list=[tt])
    try:
        with torch.no_grad():
            loss = glp.training_losses({"default": model}, x0, loss_type="EPS_LOSS", lploss=2.0)["loss"]
            x_t, eps_t = glp.dlpm.get_one_rv_loss_elements(tt, x0, A, zz)
    finally:
        ns.glp.torch = torch
    out.update({"train/x0": x0.numpy(), "train/t": tt.numpy(), "train/A": A[:, 0, 0].numpy(), "train/z": zz.numpy(),
                "train/x_t": x_t.numpy(), "train/eps_t": eps_t.numpy(), "train/loss": loss.numpy()})
    save("mlp_chain", **out)


UNET_CFGS = {
    # MNIST-like (mnist.yml:50-58): ch 32, attention at ds 2 and 4
    "mnist": dict(model_channels=32, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(2, 4),
                  num_heads=4, in_ch=1, res=32),
    # CIFAR-10-LT-like (cifar10_lt.yml:51-59) at half width so the fixture stays small: middle-block attention only
    "cifar_half": dict(model_channels=64, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,),
                       num_heads=4, in_ch=3, res=32),
}


def golden_unet():
    for name, cfg in UNET_CFGS.items():
        model = rerandomize_(make_unet(cfg, cfg["in_ch"]), 21).eval()
        out = {"weight_checksum": np.float64(parameter_checksum(model))}  # weights: seed recipe, not stored
        B = 2
        x = torch.randn(B, cfg["in_ch"], cfg["res"], cfg["res"], generator=torch.Generator().manual_seed(1))
        t = torch.tensor([0.731, 0.731])
        with torch.no_grad():
            out["fwd/x"], out["fwd/t"], out["fwd/y"] = x.numpy(), t.numpy(), model(x, t).numpy()
            t2 = torch.tensor([0.05, 0.9])
            out["fwd2/t"], out["fwd2/y"] = t2.numpy(), model(x, t2).numpy()
        r = run_dlpm(model, (B, cfg["in_ch"], cfg["res"], cfg["res"]), 1.7, 6, seed=8, clamp_a=20.0)
        out.update({"dlpm/" + k: v for k, v in r.items()})
        r = run_lim(model, (B, cfg["in_ch"], cfg["res"], cfg["res"]), 1.7, 4, seed=8, ode=False)
        out.update({"lim_sde/" + k: v for k, v in r.items()})
        save("unet_" + name, **out)


def golden_unet_full():
    """Full-width CIFAR-10-LT UNet (36.99 M params): weights regenerated from the seed recipe
    (``rerandomize_``), fixture holds only x / t / y and a weight checksum."""
    cfg = dict(model_channels=128, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,),
               num_heads=4, in_ch=3, res=32)
    model = rerandomize_(make_unet(cfg, 3), 21).eval()
    x = torch.randn(1, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    t = torch.tensor([0.5])
    with torch.no_grad():
        y = model(x, t)
    csum = parameter_checksum(model)
    save("unet_cifar_full", x=x.numpy(), t=t.numpy(), y=y.numpy(), weight_checksum=np.float64(csum),
         nparams=np.int64(sum(p.numel() for p in model.parameters())))


def golden_unet_cifar10():
    """The UNet of the reference's OTHER image config, cifar10.yml:51-59: full width with ``attn_resolutions: [4, 8, 16]`` =
    AttentionBlocks after every ResBlock of the 8x8 and 4x4 levels (ds 4 and 8; L = 64 and 16, four heads of 64) in both halves of
    the net, 11 AttentionBlocks in all.  Weights from the seed recipe; x / t / y for a batch-constant and a per-sample t."""
    cfg = dict(model_channels=128, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(4, 8, 16),
               num_heads=4, in_ch=3, res=32)
    model = rerandomize_(make_unet(cfg, 3), 21).eval()
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    t, t2 = torch.tensor([0.37, 0.37]), torch.tensor([0.05, 0.9])
    with torch.no_grad():
        y, y2 = model(x, t), model(x, t2)
    save("unet_cifar10_attn", x=x.numpy(), t=t.numpy(), y=y.numpy(), t2=t2.numpy(), y2=y2.numpy(),
         weight_checksum=np.float64(parameter_checksum(model)), nparams=np.int64(sum(p.numel() for p in model.parameters())))


def golden_next_rows():
    """scale_exploding schedule with input_scaling (sampling chain + training loss) and the LIM
    training loss (GenerativeLevyProcess.py:177-180, :651-654, :680-709; LIM/functions/loss.py)."""
    from oracle import stable
    torch.manual_seed(0)
    model = rerandomize_(ns.Model.MLPModel(mlp_params()), 11).eval()
    out = sd_np(model)
    alpha, T, B = 1.7, 30, 8
    shape = (B, 1, 2)
    # (1) sampling chain, exploding schedule, input scaling 1 / (1 + barsigma_t)
    g = torch.Generator().manual_seed(15)
    rs = np.random.RandomState(15)
    A_compact = torch.stack([torch.from_numpy(stable.gen_skewed_levy(alpha, (B,), isotropic=True, rng=rs).copy()) for _ in range(T)])
    eps_init = torch.from_numpy(stable.gen_sas(alpha, shape, isotropic=True, rng=rs))
    z = torch.randn((T - 1,) + shape, generator=g)
    glp = ns.glp.GenerativeLevyProcess(alpha, "cpu", T, rescale_timesteps=True, isotropic=True, scale="scale_exploding",
                                       input_scaling=True)
    A_list = [a.view(B, 1, 1).expand(*shape).contiguous() for a in A_compact]
    glp.dlpm.gen_a.generate = lambda *a, **k: A_list.pop(0)
    glp.dlpm.gen_eps.generate = lambda *a, **k: eps_init
    ns.glp.th = TorchProxy(randn_list=list(z))
    try:
        final, hist = glp.p_sample_loop(model, shape, get_sample_history=True)
    finally:
        ns.glp.th = torch
    out.update({"expl/A": A_compact.numpy(), "expl/eps_init": eps_init.numpy(), "expl/x_init": (glp.dlpm.barsigmas[-1] * eps_init).numpy(),
                "expl/z": z.numpy(), "expl/final": final.numpy(), "expl/hist": hist.numpy(),
                "expl/sched": torch.stack([glp.dlpm.gammas, glp.dlpm.bargammas, glp.dlpm.sigmas, glp.dlpm.barsigmas]).numpy()})
    # (2) DLPM training loss with input scaling
    Bt = 32
    g = torch.Generator().manual_seed(19)
    x0 = torch.randn(Bt, 1, 2, generator=g)
    tt = torch.randint(1, T, size=[Bt], generator=g)
    A = torch.from_numpy(stable.gen_skewed_levy(alpha, (Bt, 1, 2), isotropic=True, rng=np.random.RandomState(19)))
    zz = torch.randn(Bt, 1, 2, generator=g)
    glp.dlpm.gen_a.generate = lambda *a, **k: A
    ns.glp.torch = TorchProxy(randn_list=[zz], randint_list=[tt])
    try:
        with torch.no_grad():
            loss = glp.training_losses({"default": model}, x0, loss_type="EPS_LOSS", lploss=2.0)["loss"]
    finally:
        ns.glp.torch = torch
    out.update({"expl_train/x0": x0.numpy(), "expl_train/t": tt.numpy(), "expl_train/A": A[:, 0, 0].numpy(), "expl_train/z": zz.numpy(),
                "expl_train/loss": loss.numpy()})
    # (3) LIM training loss: injected e (SaS) and t = rand * (T - 1e-5) + 1e-5
    lim = ns.glp.GenerativeLevyProcess(alpha, "cpu", 50, rescale_timesteps=True, isotropic=True, LIM=True)
    e = torch.from_numpy(stable.gen_sas(alpha, (Bt, 1, 2), isotropic=True, rng=np.random.RandomState(23)))
    u = torch.rand(Bt, generator=torch.Generator().manual_seed(23))
    lim.dlpm.gen_eps.generate = lambda *a, **k: e
    ns.glp.torch = TorchProxy(rand_list=[u])
    try:
        with torch.no_grad():
            lim_loss = lim.training_losses({"default": model}, x0, clamp_eps=None)["loss"]
    finally:
        ns.glp.torch = torch
    t_c = u * (lim.sde.T - 1e-5) + 1e-5
    out.update({"lim_train/x0": x0.numpy(), "lim_train/e": e.numpy(), "lim_train/u": u.numpy(), "lim_train/t": t_c.numpy(),
                "lim_train/loss": lim_loss.numpy(), "lim_train/x_coeff": lim.sde.diffusion_coeff(t_c).numpy(),
                "lim_train/sigma": lim.sde.marginal_std(t_c).numpy()})
    save("next_rows", **out)


if __name__ == "__main__":
    if sys.argv[1:] == ["next_rows"]:
        golden_next_rows()
        sys.exit(0)
    if sys.argv[1:] == ["unet_cifar10"]:
        golden_unet_cifar10()
        sys.exit(0)
    golden_noise()
    golden_schedule()
    golden_mlp_chain()
    golden_unet()
    golden_unet_full()
    golden_unet_cifar10()
    golden_next_rows()
