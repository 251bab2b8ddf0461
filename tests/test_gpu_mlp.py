"""GPU: K4 fused MLP forward and the persistent whole-chain sampler vs golden vectors from the reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sub

pytestmark = pytest.mark.gpu


def mlp_params():
    return {"data": {"nfeatures": 2}, "method": "dlpm", "dlpm": {"isotropic": True}, "device": "cuda",
            "model": dict(use_a_t=False, no_a=True, a_pos_emb=False, a_emb_size=32, time_emb_type="learnable",
                          time_emb_size=32, nblocks=4, nunits=64, skip_connection=True, group_norm=True, dropout_rate=0.0,
                          learn_variance=False)}


@pytest.fixture(scope="module")
def model():
    from dlpm_b200.score_nets import MLPModel
    g = load_golden("mlp_chain")
    m = MLPModel(mlp_params())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sub(g, "sd").items()}, strict=True)
    return m.cuda().eval()


def test_mlp_forward_golden(model):
    g = load_golden("mlp_chain")
    f = sub(g, "fwd")
    y = model(torch.from_numpy(f["x"]).cuda(), torch.from_numpy(f["t"]).cuda())
    np.testing.assert_allclose(y.cpu().numpy(), f["y"], rtol=1e-4, atol=1e-5)  # fp32 bar: rtol 1e-3
    # ragged batch (not a multiple of the 64-sample tile) and a batch spanning several CTAs
    x = torch.randn(1000, 1, 2).cuda()
    t = torch.rand(1000).cuda()
    big = model(x, t)
    small = model(x[:77], t[:77])
    assert torch.equal(big[:77], small)


@pytest.mark.parametrize("tag,kw", [("dlpm", {}), ("dlpm_clip", dict(clip_denoised=True)), ("dlpm_clampa", {})])
def test_chain_golden_injected_noise(model, tag, kw):
    """Free-running T=50 chain with the reference's A, x_T and z injected (north-star tolerance rtol 1e-3 regime;
    1-ulp differences are amplified along the chain, see tests/test_oracle_golden.py)."""
    from dlpm_b200 import GenerativeLevyProcess
    g = load_golden("mlp_chain")
    r = sub(g, tag)
    T, B = r["A"].shape
    glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
    final, hist = glp.p_sample_loop(model, list(r["x_init"].shape), noise=torch.from_numpy(r["x_init"]),
                                    injected_A=torch.from_numpy(r["A"]), injected_z=torch.from_numpy(r["z"]),
                                    get_sample_history=True, **kw)
    assert hist.shape == r["hist"].shape
    np.testing.assert_allclose(glp.dlpm.Sigmas.cpu().numpy(), r["Sigmas"], rtol=0, atol=0)
    np.testing.assert_allclose(hist.cpu().numpy(), r["hist"], rtol=5e-3, atol=2e-3)
    np.testing.assert_allclose(final.cpu().numpy(), r["final"], rtol=5e-3, atol=2e-3)


def test_dlim_chain_golden(model):
    from dlpm_b200 import GenerativeLevyProcess
    g = load_golden("mlp_chain")
    r = sub(g, "dlim")
    T, B = r["A"].shape
    glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
    final, hist = glp.ddim_sample_loop(model, list(r["x_init"].shape), noise=torch.from_numpy(r["x_init"]), eta=0.0,
                                       injected_A=torch.from_numpy(r["A"]), get_sample_history=True)
    np.testing.assert_allclose(hist.cpu().numpy(), r["hist"], rtol=5e-3, atol=2e-3)
    with pytest.raises(NotImplementedError):
        glp.ddim_sample_loop(model, [4, 1, 2], eta=1.0)


@pytest.mark.parametrize("tag,ode", [("lim_sde", False), ("lim_ode", True)])
def test_lim_chain_golden(model, tag, ode):
    from dlpm_b200 import GenerativeLevyProcess
    g = load_golden("mlp_chain")
    r = sub(g, tag)
    steps = r["e_L"].shape[0]
    glp = GenerativeLevyProcess(1.7, "cuda", steps, rescale_timesteps=True, isotropic=True, LIM=True)
    final, hist = glp.lim_sample(model, list(r["x_init"].shape), ddim=ode, get_sample_history=True,
                                 injected_x=torch.from_numpy(r["x_init"]), injected_noise=torch.from_numpy(r["e_L"]))
    np.testing.assert_allclose(hist.cpu().numpy(), r["hist"], rtol=5e-3, atol=2e-3)


def test_chain_kernel_equals_stepwise_kernels(model):
    """The persistent chain (one launch) and the per-step path (forward kernel + K3) draw the same Philox noise
    and must agree to rounding."""
    from dlpm_b200 import GenerativeLevyProcess, rng
    T, B = 30, 300
    outs = []
    for force_stepwise in (False, True):
        glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
        st = rng.PhiloxState(seed=99, offset=0)
        m = model
        if force_stepwise:
            class Wrap(torch.nn.Module):  # hides native_kind -> generic per-step path
                def __init__(self, inner):
                    super().__init__()
                    self.inner = inner

                def forward(self, x, t):
                    return self.inner(x, t)
            m = Wrap(model)
        x = glp.p_sample_loop(m, [B, 1, 2], state=st)
        outs.append(x)
    np.testing.assert_allclose(outs[0].cpu().numpy(), outs[1].cpu().numpy(), rtol=2e-3, atol=2e-3)


def test_sample_api_and_training_loss(model):
    from dlpm_b200 import GenerativeLevyProcess, manual_seed
    manual_seed(7)
    glp = GenerativeLevyProcess(1.7, "cuda", 100, rescale_timesteps=True, isotropic=True)
    x = glp.sample({"default": model}, [512, 1, 2], reverse_steps=100, clamp_a=None, clamp_eps=None)
    assert x.shape == (512, 1, 2) and x.is_cuda and torch.isfinite(x).all()
    x2, hist = glp.sample({"default": model}, [64, 1, 2], reverse_steps=20, get_sample_history=True)
    assert hist.shape == (20, 64, 1, 2) and glp.reverse_steps == 100 and glp.dlpm.gammas.shape[0] == 100
    xd = glp.sample({"default": model}, [64, 1, 2], reverse_steps=100, deterministic=True, dlim_eta=0.0)
    assert torch.isfinite(xd).all()
    # training loss with injected t / A / z vs the reference value
    g = load_golden("mlp_chain")
    r = sub(g, "train")
    loss = glp.training_losses({"default": model}, torch.from_numpy(r["x0"]), loss_type="EPS_LOSS", lploss=2.0,
                               injected=dict(t=torch.from_numpy(r["t"]), A=torch.from_numpy(r["A"]), z=torch.from_numpy(r["z"])))
    np.testing.assert_allclose(loss["loss"].item(), float(r["loss"]), rtol=1e-4)
    free = glp.training_losses({"default": model}, torch.randn(256, 1, 2), loss_type="EPS_LOSS")["loss"]
    assert torch.isfinite(free) and free.dim() == 0


def test_exploding_schedule_input_scaling_golden(model):
    """'scale_exploding' schedule with input_scaling = x / (1 + barsigma_t) fed to the network
    (GenerativeLevyProcess.py:177-180, :651-654): free-running chain and training loss vs the reference."""
    from dlpm_b200 import GenerativeLevyProcess
    g = load_golden("next_rows")
    r = sub(g, "expl")
    T, B = r["A"].shape
    glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True, scale="scale_exploding", input_scaling=True)
    np.testing.assert_array_equal(glp.dlpm._sched_host.t().numpy(), r["sched"])
    final, hist = glp.p_sample_loop(model, list(r["x_init"].shape), noise=torch.from_numpy(r["x_init"]),
                                    injected_A=torch.from_numpy(r["A"]), injected_z=torch.from_numpy(r["z"]), get_sample_history=True)
    np.testing.assert_allclose(hist.cpu().numpy(), r["hist"], rtol=5e-3, atol=2e-3)
    # without the scaling the chain must differ (the flag is really applied)
    glp0 = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True, scale="scale_exploding", input_scaling=False)
    _, hist0 = glp0.p_sample_loop(model, list(r["x_init"].shape), noise=torch.from_numpy(r["x_init"]),
                                  injected_A=torch.from_numpy(r["A"]), injected_z=torch.from_numpy(r["z"]), get_sample_history=True)
    assert not np.allclose(hist0.cpu().numpy(), r["hist"], rtol=5e-2, atol=2e-2)
    tr = sub(g, "expl_train")
    loss = glp.training_losses({"default": model}, torch.from_numpy(tr["x0"]), loss_type="EPS_LOSS", lploss=2.0,
                               injected=dict(t=torch.from_numpy(tr["t"]), A=torch.from_numpy(tr["A"]), z=torch.from_numpy(tr["z"])))
    np.testing.assert_allclose(loss["loss"].item(), float(tr["loss"]), rtol=1e-4)
    # p_mean_variance API parity with per-sample scaling
    x = torch.from_numpy(r["hist"][3]).cuda()
    t = torch.full((B,), T - 4, device="cuda", dtype=torch.int64)
    out = glp.p_mean_variance(model, x, t)
    assert out["mean"].shape == x.shape and torch.isfinite(out["mean"]).all()


def test_lim_training_loss_golden(model):
    """LIM training loss (GenerativeLevyProcess.py:680-709, LIM/functions/loss.py) with injected (u, e) vs the reference,
    in-kernel elements vs the oracle, and a free-running draw."""
    from dlpm_b200 import GenerativeLevyProcess, _lib
    from oracle import process
    g = load_golden("next_rows")
    r = sub(g, "lim_train")
    glp = GenerativeLevyProcess(1.7, "cuda", 50, rescale_timesteps=True, isotropic=True, LIM=True)
    x0, u, e = (torch.from_numpy(r[k]) for k in ("x0", "u", "e"))
    loss = glp.training_losses({"default": model}, x0, injected=dict(u=u, e=e))["loss"]
    np.testing.assert_allclose(loss.item(), float(r["loss"]), rtol=1e-4)
    # elements kernel vs oracle
    t = torch.from_numpy(r["t"])
    x_t_o, score_o = process.lim_training_elements(x0, t, e, 1.7)
    x_t, score = torch.empty_like(x0).cuda(), torch.empty_like(x0).cuda()
    x0d, td, ed = x0.cuda(), t.cuda(), e.cuda()  # keep the device tensors alive across the call
    _lib.call("dlpm_b200_lim_training_elements", _lib.ptr(x_t), _lib.ptr(score), _lib.ptr(x0d), _lib.ptr(td), _lib.ptr(ed),
              x0.shape[0], 2, 1.7, 1, -1.0, 0, 0, 0, _lib.stream_ptr())
    np.testing.assert_allclose(x_t.cpu().numpy(), x_t_o.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(score.cpu().numpy(), score_o.numpy(), rtol=1e-6)
    free = glp.training_losses({"default": model}, torch.randn(256, 1, 2), clamp_eps=20.0)["loss"]
    assert torch.isfinite(free) and free.dim() == 0


def test_single_step_and_progressive_api(model):
    """API parity of the per-step entry points (GenerativeLevyProcess.py:225-239 p_sample, :332-373 ddim_sample, :291-330 /
    :413-452 the progressive generators): teacher-forced against the reference history with injected z, and the generators
    against the fused loops on the same Philox key."""
    from dlpm_b200 import GenerativeLevyProcess, rng
    g = load_golden("mlp_chain")
    r = sub(g, "dlpm")
    T, B = r["A"].shape
    glp = GenerativeLevyProcess(1.7, "cuda", T, rescale_timesteps=True, isotropic=True)
    glp.dlpm.A = torch.from_numpy(r["A"]).cuda()
    glp.dlpm._shape = list(r["x_init"].shape)
    glp.dlpm._sigma_src = None
    glp.dlpm.compute_Sigmas()
    hist = torch.from_numpy(r["hist"])
    for k, t in enumerate(range(T - 1, 0, -1)):
        tv = torch.full((B,), t, device="cuda", dtype=torch.int64)
        out = glp.p_sample(model, hist[k].cuda(), tv, noise=torch.from_numpy(r["z"][k]))
        np.testing.assert_allclose(out["sample"].cpu().numpy(), r["hist"][k + 1], rtol=2e-4, atol=2e-4)
    rd = sub(g, "dlim")
    histd = torch.from_numpy(rd["hist"])
    for k, t in enumerate(range(T - 1, 0, -1)):
        tv = torch.full((B,), t, device="cuda", dtype=torch.int64)
        out = glp.ddim_sample(model, histd[k].cuda(), tv, eta=0.0)
        np.testing.assert_allclose(out["sample"].cpu().numpy(), rd["hist"][k + 1], rtol=2e-4, atol=2e-4)
    with pytest.raises(NotImplementedError):
        glp.ddim_sample(model, histd[0].cuda(), torch.full((B,), 3, device="cuda"), eta=0.5)
    # generators: T entries, same samples as the fused loop on the same key
    T2, B2 = 12, 64
    glp2 = GenerativeLevyProcess(1.7, "cuda", T2, rescale_timesteps=True, isotropic=True)
    steps = list(glp2.p_sample_loop_progressive(model, [B2, 1, 2], state=rng.PhiloxState(seed=31, offset=0)))
    assert len(steps) == T2 and all(s["sample"].shape == (B2, 1, 2) for s in steps)
    fused, fh = glp2.p_sample_loop(model, [B2, 1, 2], get_sample_history=True, state=rng.PhiloxState(seed=31, offset=0))
    np.testing.assert_allclose(torch.stack([s["sample"] for s in steps]).cpu().numpy(), fh.cpu().numpy(), rtol=2e-3, atol=2e-3)
    dsteps = list(glp2.ddim_sample_loop_progressive(model, [B2, 1, 2], state=rng.PhiloxState(seed=31, offset=0)))
    dfused = glp2.ddim_sample_loop(model, [B2, 1, 2], eta=0.0, state=rng.PhiloxState(seed=31, offset=0))
    np.testing.assert_allclose(dsteps[-1]["sample"].cpu().numpy(), dfused.cpu().numpy(), rtol=2e-3, atol=2e-3)
    # DLPM helper parity (dlpm.py:199-202, :243-270, :377-382) against the oracle expressions
    d = glp.dlpm
    x = hist[3].cuda()
    tv = torch.full((B,), T - 4, device="cuda", dtype=torch.int64)
    eps = torch.randn_like(x)
    mean, var = d.anterior_mean_variance_dlpm(x, T - 4, eps)
    Gamma = d.compute_Gamma_t(T - 4, d.Sigmas_full()[T - 5], d.Sigmas_full()[T - 4])
    np.testing.assert_allclose(d.compute_m_tilde_t_1(x, tv, Gamma, eps).cpu().numpy(), mean.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(d.predict_eps_from_m_tilde(x, tv, mean).cpu().numpy(), eps.cpu().numpy(), rtol=2e-3, atol=2e-3)
    a = torch.rand(B, 1, 2, device="cuda") + 0.5
    np.testing.assert_allclose(d.compute_one_rv_Sigma_prime_t(tv, a).cpu().numpy(), (a * d.barsigmas[T - 4] ** 2).cpu().numpy(), rtol=1e-6)
    xs = d.sample_x_t_from_xstart_given_Sigma(x, tv, a, z_t=eps)
    np.testing.assert_allclose(xs.cpu().numpy(), (d.bargammas[T - 4] * x + a.sqrt() * eps).cpu().numpy(), rtol=1e-6, atol=1e-6)
    assert len(d.update_constants(x.shape)) == 4


def test_empty_and_single_sample_batches(model):
    """Batch edge cases through ``sample()``.  An EMPTY batch behaves as in the reference (checked against it on CPU): the DLPM /
    DLIM loops index ``t[0]`` of the empty batch and raise IndexError (GenerativeLevyProcess.py:210), LIM's loop runs over the
    empty tensors and returns them (history of steps + 1 entries, sampler.py:218-258).  A single sample runs every loop."""
    from dlpm_b200 import GenerativeLevyProcess
    glp = GenerativeLevyProcess(1.7, "cuda", 20, rescale_timesteps=True, isotropic=True)
    lim = GenerativeLevyProcess(1.7, "cuda", 20, rescale_timesteps=True, isotropic=True, LIM=True)
    for det in (False, True):
        with pytest.raises(IndexError):
            glp.sample({"default": model}, [0, 1, 2], reverse_steps=20, deterministic=det, dlim_eta=0.0)
    x, h = lim.sample({"default": model}, [0, 1, 2], reverse_steps=5, get_sample_history=True)
    assert tuple(x.shape) == (0, 1, 2) and tuple(h.shape) == (6, 0, 1, 2)
    assert tuple(lim.sample({"default": model}, [0, 1, 2], reverse_steps=5).shape) == (0, 1, 2)
    for proc, steps, n_hist in ((glp, 20, 20), (lim, 5, 6)):
        for det in (False, True):
            x, h = proc.sample({"default": model}, [1, 1, 2], reverse_steps=steps, deterministic=det, dlim_eta=0.0, get_sample_history=True)
            assert tuple(x.shape) == (1, 1, 2) and tuple(h.shape) == (n_hist, 1, 1, 2)
            assert torch.isfinite(x).all() and torch.equal(h[-1], x)
