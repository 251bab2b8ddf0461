"""Drive the UNMODIFIED reference (``/root/reference`` or its verbatim copy ``oracle/_ref``) on a bounded sample of the
benchmark workload.  TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference's stochastic sampler is ``GenerativeLevyProcess.sample`` -> ``p_sample_loop`` -> the generator
``p_sample_loop_progressive`` (``dlpm/methods/GenerativeLevyProcess.py:512-569, 241-330``).  A full CIFAR pass (999 UNet
evaluations) takes ~30 min on a CPU, so the bounded sample pulls the first ``n_steps + 1`` items from that very
generator -- i.e. the reference's own code executes, at the full T = 1000 configuration: ``sample_A`` (T scipy draws,
``dlpm.py:226-227``), ``compute_Sigmas`` (``:230-239``), x_T, then ``n_steps`` x [``p_sample`` = UNet forward + posterior
update + ``randn_like``] -- under the same ``model.eval()`` / ``th.inference_mode()`` / clamp settings ``sample()`` and
``p_sample_loop`` establish (``:526-527, :263-266``).  The measured set-up time is charged in full and the measured
per-step time is extrapolated linearly to the T - 1 steps of a pass.
"""
import os
import time

import torch

from . import ref_import

UNET_CIFAR = dict(model_channels=128, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,), num_heads=4)


def build_unet(cfg, in_channels, device, seed=0):
    """The reference's UNetModel exactly as ``dlpm/dlpm_experiment.py:41-56`` builds it, every parameter re-randomised
    with the name-keyed recipe of ``dlpm_b200/init_utils.py`` (the zero-initialised layers would make eps == 0)."""
    from dlpm_b200.init_utils import randomize_parameters_
    ns = ref_import.load()
    m = ns.unet.UNetModel(in_channels=in_channels, model_channels=cfg["model_channels"], out_channels=in_channels,
                          num_res_blocks=cfg["num_res_blocks"], attention_resolutions=cfg["attention_resolutions"],
                          dropout=0.0, channel_mult=cfg["channel_mult"], dims=2, num_classes=None, use_checkpoint=False,
                          num_heads=cfg["num_heads"], num_heads_upsample=-1, use_scale_shift_norm=True)
    randomize_parameters_(m, seed)
    return m.to(device).eval()


def build_method(alpha, device, T, **kw):
    ns = ref_import.load()
    return ns.glp.GenerativeLevyProcess(alpha, device, T, rescale_timesteps=True, isotropic=True, **kw)


def bounded_sample(glp, model, shape, n_steps, clamp_a=20, clamp_eps=200, warm_steps=1):
    """First ``warm_steps + n_steps`` reverse steps of the reference's own sampling generator.
    Returns dict(setup_s, step_s (mean over the timed steps), steps_s (each), extrapolated_pass_s, samples_per_s)."""
    dev = torch.device(glp.device)
    cuda = dev.type == "cuda"

    def sync():
        if cuda:
            torch.cuda.synchronize(dev)

    T = glp.reverse_steps
    # what sample() / p_sample_loop do before the loop (GenerativeLevyProcess.py:526-529, :263-266)
    glp.dlpm.gen_a.setParams(clamp_a=clamp_a)
    glp.dlpm.gen_eps.setParams(clamp_eps=clamp_eps)
    model.eval()
    each = []
    with torch.inference_mode():
        sync()
        t0 = time.perf_counter()
        gen = glp.p_sample_loop_progressive(model, list(shape))
        x = next(gen)["sample"]  # sample_A + compute_Sigmas + x_T
        sync()
        setup = time.perf_counter() - t0
        for k in range(warm_steps + n_steps):
            t1 = time.perf_counter()
            x = next(gen)["sample"]
            sync()
            if k >= warm_steps:
                each.append(time.perf_counter() - t1)
        gen.close()
    # the (T, B, C, H, W) tables are the reference's own memory quirk; free them between samples
    glp.dlpm.A = None
    glp.dlpm.Sigmas = None
    step = sum(each) / len(each)
    full = setup + step * (T - 1)
    return {"setup_s": setup, "step_s": step, "steps_s": each, "extrapolated_pass_s": full, "samples_per_s": shape[0] / full,
            "finite": bool(torch.isfinite(x).all())}


def describe(kind, batch, T, r, n_steps, threads=None):
    where = {"_ref": "the unmodified reference (verbatim copy under oracle/_ref)", "reference": "the unmodified reference (/root/reference)"}[kind]
    s = ("%s, GenerativeLevyProcess.p_sample_loop_progressive at the full T=%d configuration, batch %d: sample_A + compute_Sigmas + x_T in "
         "full (%.2f s) + %d of %d reverse steps (%.4f s/step), per-step time extrapolated linearly to a full pass"
         % (where, T, batch, r["setup_s"], n_steps, T - 1, r["step_s"]))
    if threads:
        s += "; torch threads %d, scipy draw single-threaded" % threads
    return s


def cpu_arm(batch, n_steps, T=1000, alpha=1.7, img=32, ch=3, seed=0):
    """One bounded sample on the host cores (all torch threads).  Returns (samples_per_s, seconds_spent, description, raw)."""
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    model = build_unet(UNET_CIFAR, ch, "cpu", seed=0)
    glp = build_method(alpha, "cpu", T)
    r = bounded_sample(glp, model, (batch, ch, img, img), n_steps)
    return r["samples_per_s"], time.perf_counter() - t0, describe(ref_import.kind(), batch, T, r, n_steps, torch.get_num_threads()), r


def gpu_eager_arm(device, batch, n_steps=20, T=1000, alpha=1.7, img=32, ch=3):
    """The same reference code with device='cuda': PyTorch eager + cuDNN (TF32 convs by PyTorch's default, which the
    reference never changes) -- the same-box GPU baseline of SURVEY.md sections 2.2 / 8d.  ``cudnn.benchmark`` as the
    reference sets it (``bem/Experiments.py:52``)."""
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        model = build_unet(UNET_CIFAR, ch, device, seed=0)
        glp = build_method(alpha, device, T)
        r = bounded_sample(glp, model, (batch, ch, img, img), n_steps, warm_steps=5)
    finally:
        torch.backends.cudnn.benchmark = prev
    del model, glp
    torch.cuda.empty_cache()
    return {"value": r["samples_per_s"], "unit": "samples/s", "batch": batch, "ms_per_reverse_step": 1e3 * r["step_s"],
            "setup_s": r["setup_s"], "timed_steps": n_steps, "kind": ref_import.kind(),
            "precision": "fp32 storage, cuDNN TF32 convolutions (PyTorch default), fp32 matmul", "finite": r["finite"],
            "sample": describe(ref_import.kind(), batch, T, r, n_steps) + "; device=cuda, eager PyTorch + cuDNN (cudnn.benchmark=True)"}
