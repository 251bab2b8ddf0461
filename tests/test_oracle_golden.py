"""CPU: pin the oracle (oracle/) against the golden vectors generated from the real reference
(tests/golden/make_golden.py) and, when /root/reference is present, against the live reference."""
import numpy as np
import pytest
import scipy.stats
import torch

from conftest import load_golden, sub
from oracle import nets, process, ref_import, stable

ALPHAS = (1.5, 1.7, 1.9)


def test_cms_restatement_matches_scipy_on_same_variates():
    """scipy _rvs_Z1 draws TH then W from random_state (scipy stats/_levy_stable/__init__.py:472-475)."""
    for alpha in ALPHAS:
        n = 20000
        rs = np.random.RandomState(11)
        ref = scipy.stats.levy_stable.rvs(alpha / 2, 1, loc=0, scale=stable.levy_scale(alpha), size=n, random_state=rs)
        rs = np.random.RandomState(11)
        TH = rs.uniform(-np.pi / 2, np.pi / 2, size=n)
        W = rs.standard_exponential(size=n)
        mine = stable.levy_scale(alpha) * stable.cms_totally_skewed(alpha / 2, TH, W)
        np.testing.assert_allclose(mine, ref, rtol=1e-9)
        # reduced Kanter form used by the CUDA kernel: identical pointwise with U = TH + pi/2
        np.testing.assert_allclose(stable.kanter_A(alpha, TH + np.pi / 2, W), ref, rtol=1e-7)


def test_laplace_transform_closed_form():
    """E exp(-A/2) = exp(-1) for every alpha (A = 2K, E exp(-sK) = exp(-s^(alpha/2)))."""
    for alpha in ALPHAS:
        A = stable.gen_skewed_levy(alpha, (400000,), isotropic=False, rng=np.random.RandomState(3))
        assert abs(np.exp(-A.astype(np.float64) / 2).mean() - np.exp(-1.0)) < 3e-3


def test_noise_golden():
    g = load_golden("noise")
    for alpha in (1.5, 1.7, 1.9, 2.0):
        for iso in (True, False):
            tag = "a%.1f_%s" % (alpha, "iso" if iso else "full")
            A = stable.gen_skewed_levy(alpha, (64, 3, 4), isotropic=iso, clamp_a=20.0 if iso else None,
                                       rng=np.random.RandomState(1234))
            np.testing.assert_allclose(A, g["A_" + tag], rtol=2e-6)
            e = stable.gen_sas(alpha, (64, 3, 4), isotropic=iso, clamp_eps=50.0, rng=np.random.RandomState(77),
                               G=g["G_" + tag])
            np.testing.assert_allclose(e, g["eps_" + tag], rtol=2e-6, atol=1e-7)


def test_schedule_golden_bit_exact():
    g = load_golden("schedule")
    for key in g.files:
        a, T, kind = key.split("_")
        alpha, T = float(a[1:]), int(T[1:])
        if kind == "exploding":
            sched = process.gen_noise_schedule(alpha, T, scale="scale_exploding")
        else:
            sched = process.gen_noise_schedule(alpha, T, time_spacing=kind)
        got = torch.stack(sched).numpy()
        assert np.array_equal(got, g[key], equal_nan=True), key


def _mlp(g):
    sd = {k: torch.from_numpy(v) for k, v in sub(g, "sd").items()}
    return lambda x, t: nets.mlp_forward(sd, 4, x, t)


def test_mlp_forward_golden():
    g = load_golden("mlp_chain")
    f = sub(g, "fwd")
    y = _mlp(g)(torch.from_numpy(f["x"]), torch.from_numpy(f["t"]))
    np.testing.assert_allclose(y.numpy(), f["y"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag,kw", [("dlpm", {}), ("dlpm_clip", dict(clip_denoised=True)), ("dlpm_clampa", {}),
                                    ("dlim", dict(deterministic=True))])
def test_dlpm_chain_golden(tag, kw):
    g = load_golden("mlp_chain")
    r = sub(g, tag)
    T, B = r["A"].shape
    shape = r["x_init"].shape
    A = torch.from_numpy(r["A"]).view(T, B, *([1] * (len(shape) - 1))).expand(T, *shape)
    net = _mlp(g)
    z = torch.from_numpy(r["z"])
    # free-running: 1-ulp differences are amplified by the chain, so the bar is the north-star rtol 1e-3 regime
    final, hist = process.dlpm_sample_loop(net, torch.from_numpy(r["x_init"]), A, z, 1.7, T, **kw)
    np.testing.assert_allclose(hist.numpy(), r["hist"], rtol=5e-3, atol=1e-3)
    sched = process.gen_noise_schedule(1.7, T)
    Sig = process.compute_Sigmas(torch.from_numpy(r["A"]), sched[0], sched[2])
    np.testing.assert_array_equal(Sig.numpy(), r["Sigmas"])
    # teacher-forced: feed the reference's x_t, compare x_{t-1} (tight)
    Sigf = process.compute_Sigmas(A, sched[0], sched[2])
    ref_hist = torch.from_numpy(r["hist"])
    for k, t in enumerate(range(T - 1, 0, -1)):
        x = ref_hist[k]
        eps = net(x, torch.tensor([t] * B).float() * (1.0 / T))
        if kw.get("deterministic"):
            xn = process.dlim_step(x, eps, t, sched, kw.get("clip_denoised", False))
        else:
            xn = process.dlpm_step(x, eps, z[k], t, Sigf, sched, kw.get("clip_denoised", False))
        np.testing.assert_allclose(xn.numpy(), r["hist"][k + 1], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("tag,ode", [("lim_sde", False), ("lim_ode", True)])
def test_lim_chain_golden(tag, ode):
    g = load_golden("mlp_chain")
    r = sub(g, tag)
    steps = r["e_L"].shape[0]
    final, hist = process.lim_sample_loop(_mlp(g), torch.from_numpy(r["x_init"]), torch.from_numpy(r["e_L"]), 1.7, steps,
                                          ode=ode)
    np.testing.assert_allclose(hist.numpy(), r["hist"], rtol=5e-3, atol=1e-3)


def test_training_loss_golden():
    g = load_golden("mlp_chain")
    r = sub(g, "train")
    x0, t = torch.from_numpy(r["x0"]), torch.from_numpy(r["t"])
    A = torch.from_numpy(r["A"]).view(-1, 1, 1).expand_as(x0)
    z = torch.from_numpy(r["z"])
    gs, bg, s, bs = process.gen_noise_schedule(1.7, 100)
    x_t, eps_t = process.one_rv_loss_elements(x0, t, A, z, bg, bs)
    np.testing.assert_allclose(x_t.numpy(), r["x_t"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(eps_t.numpy(), r["eps_t"], rtol=1e-5, atol=1e-6)
    loss = process.training_loss_dlpm(_mlp(g), x0, t, A, z, 1.7, 100)
    np.testing.assert_allclose(loss.numpy(), r["loss"], rtol=1e-5)


def test_exploding_schedule_input_scaling_golden():
    """scale_exploding schedule + input_scaling, sampling chain and training loss vs the reference."""
    g = load_golden("next_rows")
    r = sub(g, "expl")
    T, B = r["A"].shape
    shape = r["x_init"].shape
    sched = process.gen_noise_schedule(1.7, T, scale="scale_exploding")
    np.testing.assert_array_equal(torch.stack(sched).numpy(), r["sched"])
    A = torch.from_numpy(r["A"]).view(T, B, 1, 1).expand(T, *shape)
    final, hist = process.dlpm_sample_loop(_mlp(g), torch.from_numpy(r["x_init"]), A, torch.from_numpy(r["z"]), 1.7, T,
                                           scale="scale_exploding", input_scaling=True)
    np.testing.assert_allclose(hist.numpy(), r["hist"], rtol=5e-3, atol=1e-3)
    tr = sub(g, "expl_train")
    x0 = torch.from_numpy(tr["x0"])
    loss = process.training_loss_dlpm(_mlp(g), x0, torch.from_numpy(tr["t"]), torch.from_numpy(tr["A"]).view(-1, 1, 1).expand_as(x0),
                                      torch.from_numpy(tr["z"]), 1.7, T, scale="scale_exploding", input_scaling=True)
    np.testing.assert_allclose(loss.numpy(), tr["loss"], rtol=1e-5)


def test_lim_training_loss_golden():
    g = load_golden("next_rows")
    r = sub(g, "lim_train")
    x0, u, e = (torch.from_numpy(r[k]) for k in ("x0", "u", "e"))
    sde = process.VPSDE(1.7)
    t = u * (sde.T - 1e-5) + 1e-5
    np.testing.assert_array_equal(t.numpy(), r["t"])
    np.testing.assert_allclose(sde.diffusion_coeff(t).numpy(), r["x_coeff"], rtol=1e-6)
    np.testing.assert_allclose(sde.marginal_std(t).numpy(), r["sigma"], rtol=1e-6)
    loss = process.training_loss_lim(_mlp(g), x0, u, e, 1.7)
    np.testing.assert_allclose(loss.numpy(), r["loss"], rtol=1e-5)


UNET_CFGS = {
    "mnist": dict(model_channels=32, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(2, 4),
                  num_heads=4, in_ch=1),
    "cifar_half": dict(model_channels=64, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,),
                       num_heads=4, in_ch=3),
}


def unet_state_dict(cfg, seed=21):
    """Weights from the seed recipe (dlpm_b200.init_utils) on the parameter-container mirror."""
    from dlpm_b200.init_utils import parameter_checksum, randomize_parameters_
    from dlpm_b200.score_nets import UNetModel
    m = UNetModel(in_channels=cfg["in_ch"], model_channels=cfg["model_channels"], out_channels=cfg["in_ch"],
                  num_res_blocks=cfg["num_res_blocks"], attention_resolutions=cfg["attention_resolutions"],
                  channel_mult=cfg["channel_mult"], num_heads=cfg["num_heads"], use_scale_shift_norm=True)
    randomize_parameters_(m, seed)
    return {k: v.detach() for k, v in m.state_dict().items()}, parameter_checksum(m)


@pytest.mark.parametrize("name", ["mnist", "cifar_half"])
def test_unet_forward_and_chain_golden(name):
    g = load_golden("unet_" + name)
    cfg = UNET_CFGS[name]
    sd, csum = unet_state_dict(cfg)
    assert abs(csum - float(g["weight_checksum"])) < 1e-6 * max(1.0, abs(csum)), "seed recipe drifted"
    net = lambda x, t: nets.unet_forward(sd, cfg, x, t)
    x = torch.from_numpy(g["fwd/x"])
    np.testing.assert_allclose(net(x, torch.from_numpy(g["fwd/t"])).numpy(), g["fwd/y"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(net(x, torch.from_numpy(g["fwd2/t"])).numpy(), g["fwd2/y"], rtol=1e-4, atol=2e-5)
    r = sub(g, "dlpm")
    T, B = r["A"].shape
    shape = r["x_init"].shape
    A = torch.from_numpy(r["A"]).view(T, B, 1, 1, 1).expand(T, *shape)
    final, hist = process.dlpm_sample_loop(net, torch.from_numpy(r["x_init"]), A, torch.from_numpy(r["z"]), 1.7, T)
    np.testing.assert_allclose(hist.numpy(), r["hist"], rtol=1e-3, atol=1e-3)
    r = sub(g, "lim_sde")
    final, hist = process.lim_sample_loop(net, torch.from_numpy(r["x_init"]), torch.from_numpy(r["e_L"]), 1.7,
                                          r["e_L"].shape[0])
    np.testing.assert_allclose(hist.numpy(), r["hist"], rtol=1e-3, atol=1e-3)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_oracle_against_live_reference():
    """Fresh random case straight against the imported reference (build container only)."""
    ns = ref_import.load()
    for alpha in ALPHAS:
        d = ns.dlpm.DLPM(alpha, "cpu", 37)
        mine = process.gen_noise_schedule(alpha, 37)
        for a, b in zip(mine, (d.gammas, d.bargammas, d.sigmas, d.barsigmas)):
            assert torch.equal(a, b)
        np.random.seed(5)
        ref_A = ns.Distributions.gen_skewed_levy(alpha, (33, 5), isotropic=False).numpy()
        np.testing.assert_allclose(stable.gen_skewed_levy(alpha, (33, 5), isotropic=False, rng=np.random.RandomState(5)),
                                   ref_A, rtol=2e-6)
    sde_ref, sde = ns.sde.VPSDE(1.7, "cosine"), process.VPSDE(1.7)
    t = torch.linspace(0.9946, 1e-5, 11)
    assert torch.equal(sde_ref.marginal_std(t), sde.marginal_std(t))
    assert torch.equal(sde_ref.diffusion_coeff(t), sde.diffusion_coeff(t))


# ------------------------------------------------------------------------------------------------------------------
# Counter-based variates (oracle/philox.py): Philox known-answer vectors, and the fp32 evaluation order of the CUDA
# transform (rng.cuh::stable_A, restated in numpy float32) against the reference formula on identical lattice variates
# ------------------------------------------------------------------------------------------------------------------
def test_philox4x32_known_answers():
    """Random123 kat_vectors (Salmon et al. 2011): philox4x32-10 and philox4x32-7 (the kernels' default round count)."""
    from oracle import philox
    keys = [((0, 0), (0, 0, 0, 0)), ((0xffffffff, 0xffffffff), (0xffffffff,) * 4),
            ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344))]
    kat = {10: [(0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd),
                (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)],
           7: [(0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48), (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662),
               (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a)]}
    for rounds, exps in kat.items():
        for (key, ctr), exp in zip(keys, exps):
            out = philox.philox4x32(key, *[np.array([c]) for c in ctr], rounds=rounds)
            assert [int(o.reshape(-1)[0]) for o in out] == list(exp), rounds
    assert philox.ROUNDS == 7
    # counter layout: distinct (sample, position, offset, stream) never share a block
    w = philox.words(1234, philox.STREAM_A, 7, np.arange(4), 0)
    assert len({tuple(int(c[i]) for c in w) for i in range(4)}) == 4
    assert not np.array_equal(w[0], philox.words(1234, philox.STREAM_G, 7, np.arange(4), 0)[0])
    assert not np.array_equal(w[0], philox.words(1234, philox.STREAM_A, 8, np.arange(4), 0)[0])


def _device_stable_A_f32(alpha, xu, xw):
    """rng.cuh::stable_A operation by operation in numpy float32 (exact log2 / reciprocal in place of the MUFU approximations)."""
    f32 = np.float32
    C = [f32(c) for c in (3.14159263689534, -5.167709684792514, 2.5500697262138807, -0.5982421256741446, 0.07756038554456743)]

    def sinpi_half(v, scale):
        z = (v * v).astype(f32)
        p = np.full(v.shape, C[4] * scale, dtype=f32)
        for c in C[3::-1]:
            p = (p.astype(np.float64) * z + np.float64(c * scale)).astype(f32)  # fmaf: one rounding
        return (v * p).astype(f32)

    ap, r, om = f32(alpha / 2), f32((1 - alpha / 2) / (alpha / 2)), f32(1 - alpha / 2)
    top = (xu >> np.uint32(31)).astype(bool)
    xr = np.where(top, ~xu, xu).astype(np.uint32)
    v = (xr.astype(f32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(f32)
    u = np.where(top, (f32(1) - v).astype(f32), v)
    sinU = sinpi_half(v, f32(1))
    a1 = (ap * u).astype(f32)
    s1 = sinpi_half(np.minimum(a1, (f32(1) - a1).astype(f32)), f32(2))
    a2 = (om * u).astype(f32)
    if om > 0.5:
        a2 = np.minimum(a2, (f32(1) - a2).astype(f32))
    s2 = sinpi_half(a2, f32(1))
    d = np.minimum((xw.astype(f32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(f32), f32(0.99999994))
    series = (d * (f32(1) + d * (f32(0.5) + d * (f32(0.33333334) + d * f32(0.25))))).astype(f32)
    full = (f32(-0.6931471805599453) * np.log2((f32(1) - d).astype(f32)).astype(f32)).astype(f32)
    w = np.where(d < f32(0.03125), series, full).astype(f32)
    q = (f32(1) / (sinU * w).astype(f32)).astype(f32)
    l2 = (np.log2(((s1 * q).astype(f32) * w).astype(f32)).astype(f32) + r * np.log2((s2 * q).astype(f32)).astype(f32)).astype(f32)
    return np.exp2(l2).astype(f32)


@pytest.mark.parametrize("alpha", [0.8, 1.2, 1.5, 1.7, 1.9, 1.99])
def test_device_stable_formula_matches_reference_formula(alpha):
    """The kernel's merged-logarithm / polynomial-sine evaluation (fp32) against scipy's CMS branch restated in float64
    (oracle/stable.py::kanter_A through oracle/philox.py) on the same words, including both ends of both lattices."""
    from oracle import philox
    rs = np.random.RandomState(0)
    n = 400000
    xu = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    xw = rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    k = np.arange(1000, dtype=np.uint32)
    xu[:1000], xu[1000:2000] = k, np.uint32(2 ** 32 - 1) - k          # U -> 0 and U -> pi (the heavy tail)
    xw[2000:3000], xw[3000:4000] = k, np.uint32(2 ** 32 - 1) - k      # W -> 0 and W -> max
    xu[4000:5000], xw[4000:5000] = np.uint32(2 ** 32 - 1) - k, k      # both tails at once
    with np.errstate(over="ignore"):
        A = _device_stable_A_f32(alpha, xu, xw).astype(np.float64)
    ref = philox.stable_A_from_words(alpha, xu, xw)
    ok = ref < 1e38  # alpha < 1: the joint tail exceeds the fp32 range (inf, like the reference's float32 cast)
    assert ok.all() or alpha < 1.0
    assert np.all(np.isfinite(A[ok])) and A.min() > 0 and np.all(A[~ok] > 1e38)
    np.testing.assert_allclose(A[ok], ref[ok], rtol=2e-5 if alpha < 1.95 else 5e-5)
    # and the lattice restatement is the published formula: same as kanter_A on U = pi u, W
    top = (xu >> np.uint32(31)).astype(bool)
    v = philox._lattice(np.where(top, ~xu, xu))
    d = np.minimum(philox._lattice(xw), np.float64(np.float32(0.99999994)))
    lo = ~top  # on the lower half u == v exactly, so the two restatements must agree to rounding
    np.testing.assert_allclose(ref[lo], stable.kanter_A(alpha, np.pi * v[lo], -np.log1p(-d[lo])), rtol=1e-9)


# ------------------------------------------------------------------------------------------------------------------
# The caller of the boundary (bem/GenerationManager.generate) restated in oracle/caller.py, against the real one
# ------------------------------------------------------------------------------------------------------------------
class _RecordingMethod:
    """Stands in for GenerativeLevyProcess: records what the caller passes to sample() and returns fixed tensors."""

    def __init__(self, T=5):
        self.calls, self.T = [], T

    def sample(self, **kw):
        self.calls.append({k: (v if not isinstance(v, dict) else sorted(v)) for k, v in kw.items()})
        g = torch.Generator().manual_seed(3)
        shape = kw["shape"]
        x = torch.randn(*shape, generator=g) * 2.0
        if kw.get("get_sample_history"):
            hist = torch.randn(self.T, *shape, generator=g) * 2.0
            return x, hist
        return x


@pytest.mark.parametrize("is_image,shape", [(True, (4, 3, 8, 8)), (False, (4, 1, 2))])
@pytest.mark.parametrize("history", [False, True])
def test_restated_caller_matches_generation_manager(is_image, shape, history):
    from oracle import caller, ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_import.load()
    import importlib
    GM = importlib.import_module("bem.GenerationManager").GenerationManager
    loader = [(torch.zeros(*shape), torch.zeros(shape[0]))]
    mgr_kwargs = dict(reverse_steps=5, clamp_a=20, clamp_eps=200, deterministic=False)
    m_ref, m_new = _RecordingMethod(), _RecordingMethod()
    gm = GM(m_ref, loader, is_image, **mgr_kwargs)
    gm.generate({"default": None}, 7, get_sample_history=history, reverse_steps=3)
    samples, hist = caller.generation_manager_generate(m_new, {"default": None}, shape, 7, is_image, manager_kwargs=mgr_kwargs,
                                                       get_sample_history=history, reverse_steps=3)
    assert m_ref.calls == m_new.calls and m_new.calls[0]["shape"] == [7] + list(shape[1:]) and m_new.calls[0]["reverse_steps"] == 3
    assert torch.equal(gm.samples, samples)
    if history:
        assert torch.equal(gm.history, hist)
    else:
        assert len(gm.history) == 0 and len(hist) == 0
    if is_image:
        assert float(samples.min()) >= 0.0 and float(samples.max()) <= 1.0


def test_sextet_scheme_restatement_is_a_standard_normal_field():
    """oracle/philox.py::normal_sextet (six normals per Philox block, rng.cuh): every position of a row is written exactly once,
    the field is N(0,1) (KS, moments, tail bound of the 27-bit radius lattice) and differs from the quad scheme."""
    from scipy import stats
    from oracle import philox
    inner = 1152
    z = philox.normal_sextet(5, 3, np.arange(400), inner)
    assert z.shape == (400, inner) and np.isfinite(z).all()
    assert abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 1e-2 and abs(stats.kurtosis(z.ravel())) < 3e-2
    assert stats.kstest(z.ravel()[:200000], "norm").pvalue > 1e-3
    assert np.abs(z).max() <= np.sqrt(2 * 28 * np.log(2.0)) + 1e-9
    # neighbouring elements (a Box-Muller pair, and elements of different pairs) are uncorrelated
    assert abs(np.corrcoef(z[:, 0::2].ravel(), z[:, 1::2].ravel())[0, 1]) < 5e-3
    assert abs(np.corrcoef(z[:, :-4].ravel(), z[:, 4:].ravel())[0, 1]) < 5e-3
    # layout: quad 32 j + lane of granule g comes from generator 32 g + lane, normals 4 j .. 4 j + 3
    w = philox.words(5, philox.STREAM_Z, 3, np.array([7], dtype=np.uint64), np.array([2 * 33, 2 * 33 + 1], dtype=np.uint64))
    n12 = philox.normal6_from_words(w[0][0], w[1][0], w[2][0], w[3][0]) + philox.normal6_from_words(w[0][1], w[1][1], w[2][1], w[3][1])
    row = philox.normal_sextet(5, 3, np.array([7]), inner)[0].reshape(-1, 4)
    for j in range(3):  # generator 33 = granule 1, lane 1
        np.testing.assert_allclose(row[96 + 32 * j + 1], [float(v) for v in n12[4 * j:4 * j + 4]])
    assert not np.allclose(z[:4], philox.normal(5, 3, np.arange(4), inner))
