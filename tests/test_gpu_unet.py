"""GPU: whole UNet forward (K5-K7 engine) and the graph-replayed image sampling loop vs the golden vectors from
the real reference (bf16 bar of the north star: rtol 2e-2)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sub

pytestmark = pytest.mark.gpu

CFGS = {
    "mnist": dict(model_channels=32, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(2, 4), num_heads=4, in_ch=1),
    "cifar_half": dict(model_channels=64, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,), num_heads=4, in_ch=3),
    "cifar_full": dict(model_channels=128, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,), num_heads=4, in_ch=3),
    # cifar10.yml:51-59: attention after every ResBlock of the 8x8 and 4x4 levels as well (11 AttentionBlocks)
    "cifar10_attn": dict(model_channels=128, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(4, 8, 16), num_heads=4, in_ch=3),
}


def make(name, seed=21):
    from dlpm_b200.init_utils import parameter_checksum, randomize_parameters_
    from dlpm_b200.score_nets import UNetModel
    c = CFGS[name]
    m = UNetModel(in_channels=c["in_ch"], model_channels=c["model_channels"], out_channels=c["in_ch"],
                  num_res_blocks=c["num_res_blocks"], attention_resolutions=c["attention_resolutions"],
                  channel_mult=c["channel_mult"], num_heads=c["num_heads"], use_scale_shift_norm=True)
    randomize_parameters_(m, seed)
    return m.cuda().eval(), parameter_checksum(m)


def rel_err(got, want):
    return float((got - want).abs().max() / want.abs().max())


@pytest.mark.parametrize("name", ["cifar_half", "mnist"])
def test_unet_forward_golden(name):
    g = load_golden("unet_" + name)
    m, csum = make(name)
    assert abs(csum - float(g["weight_checksum"])) < 1e-6 * max(1.0, abs(csum))
    x = torch.from_numpy(g["fwd/x"]).cuda()
    for tk, yk in (("fwd/t", "fwd/y"), ("fwd2/t", "fwd2/y")):  # batch-constant t, then per-sample t
        y = m(x, torch.from_numpy(g[tk]).cuda()).cpu()
        want = torch.from_numpy(g[yk])
        assert rel_err(y, want) < 2e-2, (name, tk, rel_err(y, want))
        np.testing.assert_allclose(y.numpy(), want.numpy(), rtol=2e-2, atol=2e-2 * float(want.abs().max()))


@pytest.mark.parametrize("opts", [{}, {"gne": 0, "dxs": 0, "idskip": 1024}, {"conv_gne": 0}, {"fuse": 1, "max_px": 64}, {"fuse": 1, "max_px": 1024}, {"gn_stats": 0},
                                  {"gn_fuse": 2}, {"gn_fuse": 1}],
                         ids=["default_groupnorm_in_the_epilogue_16x16_dx_stacked_out_conv", "separate_gn_apply_everywhere_plain_out_conv_identity_skips_in_the_k_loop",
                              "gne_targets_through_post_warps",
                              "producer_side_groupnorm_up_to_8x8", "producer_side_groupnorm_everywhere",
                              "gn_cluster_kernels", "normalise_on_load_everywhere", "normalise_on_load_final_conv"])
def test_unet_forward_full_width_golden(opts):
    """Benchmark-width UNet against the reference's output, with every GroupNorm strategy of the engine: statistics from
    the conv epilogues + one streaming pass (default), applied by the producing convolution's post warps (option), the
    stand-alone cluster kernel, and GroupNorm + SiLU applied inside the consuming conv's shared-memory pipeline."""
    from dlpm_b200 import _lib
    g = load_golden("unet_cifar_full")
    opts = dict(opts)
    fuse = bool(opts.pop("fuse", 0))
    max_px = opts.pop("max_px", 64)
    gne = bool(opts.pop("gne", 1))
    dxs = bool(opts.pop("dxs", 1))
    idskip = opts.pop("idskip", None)
    for k, v in opts.items():
        _lib.call("dlpm_b200_set_option", k.encode(), v)
    try:
        m, csum = make("cifar_full")
        m.fuse_groupnorm = fuse
        m.fuse_groupnorm_max_pixels = max_px
        m.fuse_groupnorm_epilogue = gne
        m.dx_stacked_out_conv = dxs
        if idskip is not None:
            m.identity_skip_as_conv_min_pixels = idskip
        assert abs(csum - float(g["weight_checksum"])) < 1e-6 * max(1.0, abs(csum))
        y = m(torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["t"]).cuda()).cpu()
    finally:
        _lib.call("dlpm_b200_set_option", b"gn_stats", 1)
        _lib.call("dlpm_b200_set_option", b"gn_fuse", 0)
        _lib.call("dlpm_b200_set_option", b"conv_gne", 1)
    want = torch.from_numpy(g["y"])
    assert rel_err(y, want) < 2e-2, rel_err(y, want)


@pytest.mark.parametrize("strategy", ["default", "separate_passes", "post_warps"])
def test_unet_forward_cifar10_yml_architecture_golden(strategy):
    """The reference's other image config (cifar10.yml: ``attn_resolutions: [4, 8, 16]`` = AttentionBlocks at 8x8, L = 64, and 4x4,
    L = 16, in both halves of the net) at full width against the reference's output: GroupNorms of the attention blocks carried by
    the producing convolutions' epilogues without SiLU (default), as separate passes, and through the post warps."""
    from dlpm_b200 import _lib
    g = load_golden("unet_cifar10_attn")
    m, csum = make("cifar10_attn")
    assert abs(csum - float(g["weight_checksum"])) < 1e-6 * max(1.0, abs(csum))
    if strategy == "separate_passes":
        m.fuse_groupnorm_epilogue, m.dx_stacked_out_conv = False, False
    if strategy == "post_warps":
        m.fuse_groupnorm, m.fuse_groupnorm_epilogue, m.fuse_groupnorm_max_pixels = True, False, 64
    x = torch.from_numpy(g["x"]).cuda()
    for tk, yk in (("t", "y"), ("t2", "y2")):  # batch-constant t, then per-sample t
        n0 = ctypes_stat(_lib, b"conv_gne_launches")
        y = m(x, torch.from_numpy(g[tk]).cuda()).cpu()
        want = torch.from_numpy(g[yk])
        assert rel_err(y, want) < 2e-2, (strategy, tk, rel_err(y, want))
        np.testing.assert_allclose(y.numpy(), want.numpy(), rtol=2e-2, atol=2e-2 * float(want.abs().max()))
        if strategy == "default":
            assert ctypes_stat(_lib, b"conv_gne_launches") > n0
    # a batch that changes the tiling (CTA pairs, several items per CTA): default engine against separate passes
    if strategy == "default":
        gen = torch.Generator().manual_seed(5)
        xb = torch.randn(96, 3, 32, 32, generator=gen).cuda()
        tb = torch.full((96,), 0.61).cuda()
        y_fused = m(xb, tb).cpu()
        m.fuse_groupnorm_epilogue, m.dx_stacked_out_conv = False, False
        y_plain = m(xb, tb).cpu()
        assert torch.isfinite(y_fused).all() and rel_err(y_fused, y_plain) < 2e-2, rel_err(y_fused, y_plain)


@pytest.mark.parametrize("B", [3, 17, 40, 100, 333])
def test_default_engine_matches_the_unfused_engine_across_batch_sizes(B):
    """The default engine (GroupNorm in the conv epilogues at 16x16 / 8x8 / 4x4, dx-stacked final conv) against the same weights with
    every fusion off (separate GroupNorm passes, 9-tap final conv), at batch sizes that change the tiling: odd batches (masked
    tails, idle halves of CTA pairs), fewer work items than SMs, several items per CTA; batch-constant and per-sample time steps."""
    from dlpm_b200 import _lib
    m, _ = make("cifar_full")
    g = torch.Generator().manual_seed(100 + B)
    x = torch.randn(B, 3, 32, 32, generator=g).cuda()
    for t in (torch.full((B,), 0.37), torch.rand(B, generator=g)):
        n0 = [ctypes_stat(_lib, b"conv_gne_launches")]
        y_fused = m(x, t.cuda()).cpu()
        n0.append(ctypes_stat(_lib, b"conv_gne_launches"))
        m.fuse_groupnorm_epilogue, m.dx_stacked_out_conv = False, False
        try:
            y_plain = m(x, t.cuda()).cpu()
        finally:
            m.fuse_groupnorm_epilogue, m.dx_stacked_out_conv = True, True
        assert n0[1] > n0[0], "the default engine did not launch a GroupNorm-in-the-epilogue kernel"
        assert torch.isfinite(y_fused).all()
        assert rel_err(y_fused, y_plain) < 2e-2, (B, rel_err(y_fused, y_plain))


def ctypes_stat(_lib, name):
    import ctypes
    v = ctypes.c_int64(0)
    _lib.call("dlpm_b200_get_stat", name, ctypes.byref(v))
    return v.value


def test_unet_layerwise_against_oracle():
    """Per-block parity through the engine's debug buffers (no scratch reuse) against the CPU oracle's hooks."""
    from oracle import nets
    name = "cifar_half"
    m, _ = make(name)
    c = CFGS[name]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(3))
    t = torch.tensor([0.4, 0.4])
    eng = m.engine(32, 32, 2, reuse_scratch=False)
    out = torch.empty(2, 3, 32, 32, device="cuda")
    eng.forward(x.cuda(), t[:1].cuda(), None, 0.0, out, 2)
    torch.cuda.synchronize()
    # oracle intermediates: re-run the functional forward, recording block outputs
    rec = {}
    cfg = dict(model_channels=c["model_channels"], channel_mult=c["channel_mult"], num_res_blocks=c["num_res_blocks"],
               attention_resolutions=c["attention_resolutions"], num_heads=c["num_heads"])
    import torch.nn.functional as F
    orig = nets._resblock

    def hook(sd_, p, xx, emb):
        y = orig(sd_, p, xx, emb)
        rec[p] = y
        return y
    nets._resblock = hook
    try:
        want = nets.unet_forward(sd, cfg, x, t)
    finally:
        nets._resblock = orig
    errs = {}
    for p, y in rec.items():
        if p in eng.names:
            got = eng.read_buffer(p, 2, (y.shape[2], y.shape[3], y.shape[1])).cpu()
            errs[p] = rel_err(got, y)
    errs["final"] = rel_err(out.cpu(), want)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/layerwise.txt", "w") as fh:
        for k, v in errs.items():
            fh.write("%-24s %.5f\n" % (k, v))
    assert len(rec) >= 20
    assert max(errs.values()) < 2.5e-2, errs


def test_image_chain_teacher_forced_and_free_running():
    from dlpm_b200 import GenerativeLevyProcess, _lib
    from oracle import process
    g = load_golden("unet_cifar_half")
    m, _ = make("cifar_half")
    r = sub(g, "dlpm")
    T, B = r["A"].shape
    hist = torch.from_numpy(r["hist"])
    glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
    glp.dlpm.A = torch.from_numpy(r["A"]).cuda()
    glp.dlpm._shape = list(hist.shape[1:])
    glp.dlpm._sigma_src = None
    glp.dlpm.compute_Sigmas()
    D = int(np.prod(hist.shape[2:]))
    for k, t in enumerate(range(T - 1, 0, -1)):  # teacher-forced: reference x_t in, x_{t-1} out
        x = hist[k].clone().cuda()
        eps = m(x, torch.full((B,), t / T, device="cuda"))
        z = torch.from_numpy(r["z"][k]).cuda()
        _lib.call("dlpm_b200_reverse_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(glp.dlpm.Sigmas), _lib.ptr(glp.dlpm.sched), t, None, T,
                  B, D, 0, _lib.ptr(z), 0, 0, 0, None, _lib.stream_ptr())
        assert rel_err(x.cpu(), hist[k + 1]) < 2e-2, (t, rel_err(x.cpu(), hist[k + 1]))
    final, h = glp.p_sample_loop(m, list(hist.shape[1:]), noise=torch.from_numpy(r["x_init"]), injected_A=torch.from_numpy(r["A"]),
                                 injected_z=torch.from_numpy(r["z"]), get_sample_history=True)
    assert rel_err(h.cpu(), hist) < 5e-2


def test_graph_replayed_sampling_matches_direct_launches():
    from dlpm_b200 import GenerativeLevyProcess, rng
    m, _ = make("cifar_half")
    outs = []
    for hist in (False, True):  # hist=True forces direct launches, False uses the captured CUDA graph
        glp = GenerativeLevyProcess(1.7, "cuda", 8, rescale_timesteps=True, isotropic=True)
        st = rng.PhiloxState(seed=5, offset=0)
        glp.dlpm.gen_a.setParams(clamp_a=20.0)
        glp.dlpm.gen_eps.setParams(clamp_eps=200.0)
        o = glp.p_sample_loop(m, [4, 3, 32, 32], get_sample_history=hist, state=st)
        outs.append(o[0] if hist else o)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])
    x = GenerativeLevyProcess(1.7, "cuda", 1000, rescale_timesteps=True, isotropic=True).sample(
        {"default": m}, [2, 3, 32, 32], reverse_steps=6, clamp_a=20, clamp_eps=200)
    assert x.shape == (2, 3, 32, 32)


def test_exploding_input_scaling_image_loop():
    """scale_exploding + input_scaling on the image net: the graph-replayed loop (scale kernel inside the
    captured step, step index from the device counter), direct launches and the generic per-step path (the network called
    as an opaque module on x * 1/(1+barsigma_t)) agree; and the scaling is really applied."""
    from dlpm_b200 import GenerativeLevyProcess, rng
    m, _ = make("cifar_half")

    class Opaque(torch.nn.Module):  # hides native_kind -> generic per-step path of _reverse_loop
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x, t):
            return self.inner(x, t)

    outs = {}
    for tag, model, hist, scaling in (("graph", m, False, True), ("direct", m, True, True), ("generic", Opaque(m), False, True),
                                      ("unscaled", m, False, False)):
        glp = GenerativeLevyProcess(1.7, "cuda", 8, rescale_timesteps=True, isotropic=True, scale="scale_exploding",
                                    input_scaling=scaling)
        glp.dlpm.gen_a.setParams(clamp_a=20.0)
        glp.dlpm.gen_eps.setParams(clamp_eps=200.0)
        o = glp.p_sample_loop(model, [4, 3, 32, 32], get_sample_history=hist, state=rng.PhiloxState(seed=5, offset=0))
        outs[tag] = o[0] if hist else o
    assert torch.isfinite(outs["graph"]).all()
    assert torch.equal(outs["graph"], outs["direct"])
    np.testing.assert_allclose(outs["graph"].cpu().numpy(), outs["generic"].cpu().numpy(), rtol=2e-2, atol=2e-2)
    assert not torch.allclose(outs["graph"], outs["unscaled"], rtol=1e-2, atol=1e-2)


def test_full_size_shard_invariance_and_determinism():
    """BASELINE.json C3 at the per-GPU size (512 x 3 x 32 x 32, full-width UNet), a few reverse steps through the graph-replayed
    loop: determinism (same key -> same bits), and shard invariance -- two 'ranks' of 256 samples with sample_base = rank * 256
    reproduce the 512-sample run (per-sample GroupNorm statistics, attention and Philox streams keyed by the GLOBAL sample
    index; the reduction order inside a sample does not depend on how the batch is tiled)."""
    from dlpm_b200 import GenerativeLevyProcess, rng
    m, _ = make("cifar_full")
    T = 5

    def run(B, base):
        glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
        glp.dlpm.gen_a.setParams(clamp_a=20.0)
        glp.dlpm.gen_eps.setParams(clamp_eps=200.0)
        return glp.p_sample_loop(m, [B, 3, 32, 32], state=rng.PhiloxState(seed=77, offset=0, sample_base=base))

    full = run(512, 0)
    assert full.shape == (512, 3, 32, 32) and torch.isfinite(full).all()
    assert torch.equal(run(512, 0), full)
    halves = torch.cat([run(256, 0), run(256, 256)])
    np.testing.assert_allclose(halves.cpu().numpy(), full.cpu().numpy(), rtol=2e-2, atol=2e-2)
    assert float((halves - full).abs().max()) <= 2e-2 * float(full.abs().max())


def test_lim_image_chain_golden_and_graph():
    """LIM SDE sampler on the image net: injected-noise chain vs the reference history, and the graph-replayed loop
    (times from a device table) vs direct launches."""
    from dlpm_b200 import GenerativeLevyProcess, rng
    g = load_golden("unet_cifar_half")
    m, _ = make("cifar_half")
    r = sub(g, "lim_sde")
    steps = r["e_L"].shape[0]
    glp = GenerativeLevyProcess(1.7, "cuda", steps, rescale_timesteps=True, isotropic=True, LIM=True)
    final, hist = glp.lim_sample(m, list(r["x_init"].shape), get_sample_history=True, injected_x=torch.from_numpy(r["x_init"]),
                                 injected_noise=torch.from_numpy(r["e_L"]))
    want = torch.from_numpy(r["hist"])
    assert rel_err(hist.cpu(), want) < 3e-2, rel_err(hist.cpu(), want)
    outs = []
    for with_hist in (False, True):  # False: captured graph; True: direct launches
        glp = GenerativeLevyProcess(1.7, "cuda", 6, rescale_timesteps=True, isotropic=True, LIM=True)
        o = glp.lim_sample(m, [4, 3, 32, 32], get_sample_history=with_hist, state=rng.PhiloxState(seed=11, offset=0))
        outs.append(o[0] if with_hist else o)
    assert torch.isfinite(outs[0]).all() and torch.equal(outs[0], outs[1])
    ode = GenerativeLevyProcess(1.7, "cuda", 6, rescale_timesteps=True, isotropic=True, LIM=True).sample(
        {"default": m}, [2, 3, 32, 32], reverse_steps=6, deterministic=True)
    assert ode.shape == (2, 3, 32, 32) and torch.isfinite(ode).all()


def test_training_loss_forward_with_unet():
    """Training-loss forward path (Prop. 9, GenerativeLevyProcess.py:612-677) through the UNet engine with per-sample
    timesteps, against the oracle on the same injected (t, A, z)."""
    from dlpm_b200 import GenerativeLevyProcess
    from oracle import nets, process, stable
    m, _ = make("cifar_half")
    c = CFGS["cifar_half"]
    cfg = dict(model_channels=c["model_channels"], channel_mult=c["channel_mult"], num_res_blocks=c["num_res_blocks"],
               attention_resolutions=c["attention_resolutions"], num_heads=c["num_heads"])
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    B, T = 6, 100
    g = torch.Generator().manual_seed(4)
    x0 = torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1)
    t = torch.randint(1, T, size=[B], generator=g)
    A = torch.from_numpy(stable.gen_skewed_levy(1.7, (B,), isotropic=True, clamp_a=20.0, rng=np.random.RandomState(4)).copy())
    z = torch.randn(B, 3, 32, 32, generator=g)
    glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
    got = glp.training_losses({"default": m}, x0, loss_type="EPS_LOSS", lploss=2.0, clamp_a=20.0,
                              injected=dict(t=t, A=A, z=z))["loss"].item()
    want = process.training_loss_dlpm(lambda xx, tt: nets.unet_forward(sd, cfg, xx, tt), x0, t,
                                      A.view(B, 1, 1, 1).expand(B, 3, 32, 32), z, 1.7, T).item()
    assert abs(got - want) < 2e-2 * abs(want), (got, want)
