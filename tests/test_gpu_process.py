"""GPU: K2 (Sigma scan), K3 (fused DLPM / DLIM / LIM steps) and the training-forward kernels against the
golden vectors generated from the real reference (teacher-forced: bit-level; free-running: rtol 1e-3)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, sub

pytestmark = pytest.mark.gpu

from oracle import nets, process  # noqa: E402


@pytest.fixture(scope="module")
def L():
    import dlpm_b200
    from dlpm_b200 import _lib
    assert torch.cuda.is_available()
    return _lib


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def sched_dev(alpha, T):
    s = torch.stack(process.gen_noise_schedule(alpha, T), dim=1).contiguous()
    return s, s.cuda()


def test_sigma_scan_matches_reference_bitwise(L):
    g = load_golden("mlp_chain")
    r = sub(g, "dlpm")
    T, B = r["A"].shape
    _, sd = sched_dev(1.7, T)
    A = cu(r["A"])
    Sig = torch.empty_like(A)
    L.call("dlpm_b200_sigma_scan", L.ptr(Sig), L.ptr(A), None, L.ptr(sd), T, B, 1, 0, 1.7, -1.0, 0, 0, 0, L.stream_ptr())
    assert np.array_equal(Sig.cpu().numpy(), r["Sigmas"])  # same fp32 op order as dlpm.py:230-239


@pytest.mark.parametrize("tag,clip,det", [("dlpm", False, False), ("dlpm_clip", True, False), ("dlpm_clampa", False, False),
                                          ("dlim", False, True)])
def test_reverse_step_teacher_forced(L, tag, clip, det):
    """Feed the reference's x_t and eps(x_t) (oracle net on CPU), compare x_{t-1} with the reference history."""
    g = load_golden("mlp_chain")
    r = sub(g, tag)
    sdict = {k: torch.from_numpy(v) for k, v in sub(g, "sd").items()}
    T, B = r["A"].shape
    D = int(np.prod(r["x_init"].shape[1:]))
    _, sd = sched_dev(1.7, T)
    Sig = cu(r["Sigmas"])
    hist = torch.from_numpy(r["hist"])
    flags = L.STEP_CLIP_DENOISED if clip else 0
    worst = 0.0
    for k, t in enumerate(range(T - 1, 0, -1)):
        x = hist[k].clone()
        eps = nets.mlp_forward(sdict, 4, x, torch.tensor([t] * B).float() * (1.0 / T))
        xd, ed = x.cuda().contiguous(), eps.cuda().contiguous()
        if det:
            L.call("dlpm_b200_dlim_step", L.ptr(xd), L.ptr(ed), L.ptr(sd), t, None, T, B, D, flags, None, L.stream_ptr())
        else:
            zd = cu(r["z"][k])
            L.call("dlpm_b200_reverse_step", L.ptr(xd), L.ptr(ed), L.ptr(Sig), L.ptr(sd), t, None, T, B, D, flags, L.ptr(zd), 0, 0, 0,
                   None, L.stream_ptr())
        got, want = xd.cpu().numpy(), r["hist"][k + 1]
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
        worst = max(worst, float(np.abs(got - want).max()))
    assert worst < 1e-4


def test_reverse_step_variants_agree(L):
    """vector / scalar paths, bf16 eps, device-side step counter and history output."""
    torch.manual_seed(0)
    T, B, D = 20, 37, 48
    _, sd = sched_dev(1.7, T)
    A = torch.rand(T, B).cuda() * 3 + 0.1
    Sig = torch.empty_like(A)
    L.call("dlpm_b200_sigma_scan", L.ptr(Sig), L.ptr(A), None, L.ptr(sd), T, B, 1, 0, 1.7, -1.0, 0, 0, 0, L.stream_ptr())
    x0 = torch.randn(B, D).cuda()
    eps = torch.randn(B, D).cuda()
    z = torch.randn(B, D).cuda()
    t = 7

    def run(x, e, flags=0, zz=z, t_dev=None, hist=None, Dd=D, Bb=B):
        x = x.clone()
        L.call("dlpm_b200_reverse_step", L.ptr(x), L.ptr(e), L.ptr(Sig), L.ptr(sd), t, L.ptr(t_dev), T, Bb, Dd, flags, L.ptr(zz), 5, 11,
               3, L.ptr(hist), L.stream_ptr())
        return x
    ref = run(x0, eps)
    # oracle
    sched = process.gen_noise_schedule(1.7, T)
    want = process.dlpm_step(x0.cpu(), eps.cpu(), z.cpu(), t, Sig.cpu()[:, :, None], sched)
    np.testing.assert_allclose(ref.cpu().numpy(), want.numpy(), rtol=1e-6, atol=1e-6)
    # device counter + history
    td = torch.tensor([t], dtype=torch.int32).cuda()
    h = torch.empty_like(x0)
    out = run(x0, eps, t_dev=td, hist=h)
    assert torch.equal(out, ref) and torch.equal(h, ref)
    # scalar path (D not multiple of 4): same numbers on the overlapping layout
    x1, e1, z1 = x0[:, :47].contiguous(), eps[:, :47].contiguous(), z[:, :47].contiguous()
    out1 = run(x1, e1, zz=z1, Dd=47)
    assert torch.equal(out1, ref[:, :47])
    # bf16 eps
    eb = eps.to(torch.bfloat16)
    outb = run(x0, eb, flags=L.STEP_EPS_BF16)
    refb = run(x0, eb.float())
    assert torch.equal(outb, refb)
    # in-kernel noise: deterministic, N(0,1)-driven: (x' - mean)/sd is standard normal
    a = run(x0, eps, zz=None)
    b = run(x0, eps, zz=None)
    assert torch.equal(a, b)
    mean = run(x0, eps, zz=torch.zeros_like(z))
    one = run(x0, eps, zz=torch.ones_like(z))
    zhat = ((a - mean) / (one - mean)).cpu().numpy().ravel()
    assert abs(zhat.mean()) < 0.1 and abs(zhat.std() - 1) < 0.1
    # vector and scalar paths draw the same in-kernel noise
    a47 = run(x1, e1, zz=None, Dd=47)  # scalar path keeps the IEEE divide, the vector path uses reciprocal + FMA (<= 2 ulp)
    np.testing.assert_allclose(a47.cpu().numpy(), a[:, :47].cpu().numpy(), rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("tag,ode", [("lim_sde", False), ("lim_ode", True)])
def test_lim_step_teacher_forced(L, tag, ode):
    from dlpm_b200.methods.lim import VPSDE, lim_step_table
    g = load_golden("mlp_chain")
    r = sub(g, tag)
    sdict = {k: torch.from_numpy(v) for k, v in sub(g, "sd").items()}
    steps = r["e_L"].shape[0]
    ts, coef = lim_step_table(VPSDE(1.7), steps, ode)
    cd = coef.cuda()
    hist = torch.from_numpy(r["hist"])
    B = hist.shape[1]
    D = int(np.prod(hist.shape[2:]))
    for i in range(steps):
        x = hist[i].clone()
        out = nets.mlp_forward(sdict, 4, x, torch.ones(B) * ts[i])
        xd, od, ed = x.cuda().contiguous(), out.cuda().contiguous(), cu(r["e_L"][i])
        L.call("dlpm_b200_lim_step", L.ptr(xd), L.ptr(od), L.ptr(cd), i, None, B, D, 0, 1 if ode else 0, 1, 1.7, -1.0, L.ptr(ed), 0, 0,
               0, None, L.stream_ptr())
        np.testing.assert_allclose(xd.cpu().numpy(), r["hist"][i + 1], rtol=2e-5, atol=2e-5)


def test_training_elements_and_loss(L):
    g = load_golden("mlp_chain")
    r = sub(g, "train")
    _, sd = sched_dev(1.7, 100)
    B, D = r["x0"].shape[0], int(np.prod(r["x0"].shape[1:]))
    x_t = torch.empty(B, D).cuda()
    eps_t = torch.empty(B, D).cuda()
    x0d, td, Ad, zd = cu(r["x0"]), cu(r["t"]), cu(r["A"]), cu(r["z"])
    L.call("dlpm_b200_training_elements", L.ptr(x_t), L.ptr(eps_t), L.ptr(x0d), L.ptr(td), L.ptr(Ad), L.ptr(zd), L.ptr(sd), 100, B, D,
           1.7, -1.0, 0, 0, 0, L.stream_ptr())
    np.testing.assert_allclose(x_t.cpu().numpy().reshape(r["x_t"].shape), r["x_t"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(eps_t.cpu().numpy().reshape(r["eps_t"].shape), r["eps_t"], rtol=1e-5, atol=1e-6)
    # loss terms vs oracle for the three supported exponents
    pred = torch.randn(B, 96).cuda()
    tgt = torch.randn(B, 96).cuda()
    for lp in (2.0, 1.0, -1):
        out = torch.empty(B).cuda()
        L.call("dlpm_b200_loss_terms", L.ptr(out), L.ptr(pred), L.ptr(tgt), B, 96, float(lp), 0, L.stream_ptr())
        want = process.compute_loss_terms(pred.cpu(), tgt.cpu(), lp)
        np.testing.assert_allclose(out.cpu().numpy(), want.numpy(), rtol=1e-5)


def test_postprocess(L):
    x = torch.randn(1000).cuda() * 2
    y = torch.empty_like(x)
    L.call("dlpm_b200_postprocess", L.ptr(y), L.ptr(x), x.numel(), 1.0, 1, L.stream_ptr())
    assert torch.equal(y, (x.clamp(-1, 1) + 1) / 2)


def test_argument_errors(L):
    x = torch.zeros(4, 4).cuda()
    with pytest.raises(L.DlpmB200Error, match="t out of range"):
        L.call("dlpm_b200_reverse_step", L.ptr(x), L.ptr(x), L.ptr(x), L.ptr(x), 0, None, 4, 4, 4, 0, None, 0, 0, 0, None, L.stream_ptr())
    with pytest.raises(L.DlpmB200Error):
        L.ptr(torch.zeros(3))  # CPU tensor: no fallback


def test_non_isotropic_sigma_chain_and_step(L):
    """isotropic=False: per-element A / Sigma tables (the reference's full-tensor layout, dlpm.py:226-239) and the fused
    step with DLPM_STEP_SIGMA_FULL, against the oracle on the A table the kernel drew."""
    from dlpm_b200.methods.dlpm import DLPM
    from dlpm_b200 import rng
    T, shape = 12, (5, 3, 4, 4)
    d = DLPM(1.7, "cuda", T, isotropic=False)
    d.gen_a.setParams(clamp_a=20.0)
    d.sample_A(shape, T, state=rng.PhiloxState(seed=3, offset=0))
    assert d.A.shape == (T, *shape) and d.Sigmas.shape == (T, *shape)
    A = d.A.cpu()
    assert float(A.max()) <= 20.0 and float(A.min()) >= 0.0 and A[0].flatten().unique().numel() > 100  # per-element draws
    sched = process.gen_noise_schedule(1.7, T)
    want = process.compute_Sigmas(A, sched[0], sched[2])
    assert torch.equal(d.Sigmas.cpu(), want)
    x, eps, z = (torch.randn(shape) for _ in range(3))
    t = 5
    xd, ed, zd = x.cuda().contiguous(), eps.cuda().contiguous(), z.cuda().contiguous()
    B, D = shape[0], 48
    L.call("dlpm_b200_reverse_step", L.ptr(xd), L.ptr(ed), L.ptr(d.Sigmas), L.ptr(d.sched), t, None, T, B, D, L.STEP_SIGMA_FULL,
           L.ptr(zd), 0, 0, 0, None, L.stream_ptr())
    ref = process.dlpm_step(x, eps, z, t, want, sched)
    np.testing.assert_allclose(xd.cpu().numpy(), ref.numpy(), rtol=1e-6, atol=1e-6)
    # replacing A by hand re-runs the scan (compute_Sigmas) on the injected table
    d.A = (A * 0.5).cuda()
    d.compute_Sigmas()
    assert torch.equal(d.Sigmas.cpu(), process.compute_Sigmas(A * 0.5, sched[0], sched[2]))


# ------------------------------------------------------------------------------------------------------------------
# In-kernel noise paths (the production mode): the variates each kernel draws are the documented Philox words
# (oracle/philox.py), so the kernel's result must equal the ORACLE step fed with the oracle's own evaluation of those variates.
# ------------------------------------------------------------------------------------------------------------------
def test_in_kernel_noise_is_the_documented_stream(L):
    from oracle import philox
    alpha, T, B, D = 1.7, 20, 37, 48
    seed, off, base = 0xABCDEF0123, 11, 3
    sched = process.gen_noise_schedule(alpha, T)
    _, sd = sched_dev(alpha, T)
    samples = base + np.arange(B)

    # K2: A_t drawn in-kernel at call offset off + t, clamped; Sigma follows dlpm.py:230-239 on those draws bit for bit
    Sig, A_out = torch.empty(T, B).cuda(), torch.empty(T, B).cuda()
    L.call("dlpm_b200_sigma_scan", L.ptr(Sig), None, L.ptr(A_out), L.ptr(sd), T, B, 1, 0, alpha, 20.0, seed, off, base, L.stream_ptr())
    A_ref = np.stack([philox.sample_A(alpha, seed, off + t, samples, clamp_a=20.0) for t in range(T)])
    np.testing.assert_allclose(A_out.cpu().numpy(), A_ref, rtol=3e-5)
    want_S = process.compute_Sigmas(A_out.cpu()[:, :, None], sched[0], sched[2])[:, :, 0]
    assert np.array_equal(Sig.cpu().numpy(), want_S.numpy())

    # K3 (vector fast path and scalar path): z = N(0,1) of stream Z at call offset off + t
    torch.manual_seed(1)
    x0, eps = torch.randn(B, D), torch.randn(B, D)
    for t in (7, 1):
        for Dd in (D, D - 1):
            x = x0[:, :Dd].contiguous().cuda()
            e = eps[:, :Dd].contiguous().cuda()
            L.call("dlpm_b200_reverse_step", L.ptr(x), L.ptr(e), L.ptr(Sig), L.ptr(sd), t, None, T, B, Dd, 0, None, seed, off, base, None,
                   L.stream_ptr())
            z = torch.from_numpy(philox.normal(seed, off + t, samples, D, stream=philox.STREAM_Z)[:, :Dd]).float()
            want = process.dlpm_step(x0[:, :Dd], eps[:, :Dd], z, t, Sig.cpu()[:, :, None], sched)
            np.testing.assert_allclose(x.cpu().numpy(), want.numpy(), rtol=2e-5, atol=2e-5)

    # LIM SDE step: eps_L = sqrt(A_b) G with A from stream EPS_A and G from stream G at call offset off + step
    coef = torch.tensor([[1.3, 0.97, -0.02, 0.11]] * 4).cuda()
    step = 2
    x = x0.clone().cuda()
    mo = eps.clone().cuda()
    L.call("dlpm_b200_lim_step", L.ptr(x), L.ptr(mo), L.ptr(coef), step, None, B, D, 0, 0, 1, alpha, 200.0, None, seed, off, base, None,
           L.stream_ptr())
    eL = philox.sas_isotropic(alpha, seed, off + step, samples, D, clamp_eps=200.0)
    sc, a, c_score, c_noise = 1.3, 0.97, -0.02, 0.11
    want = a * x0.numpy().astype(np.float64) + c_score * (sc * eps.numpy().astype(np.float64)) + c_noise * eL
    np.testing.assert_allclose(x.cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.sqrt(np.abs(eL).max()))

    # training elements (dlpm.py:384-401): A from stream A, z from stream Z, both at the call offset
    tt = torch.randint(1, T, (B,), dtype=torch.int64)
    xs = torch.rand(B, D) * 2 - 1
    x_t, e_t = torch.empty(B, D).cuda(), torch.empty(B, D).cuda()
    xs_d, tt_d = xs.cuda(), tt.cuda()  # keep both alive: a freed temporary's block is handed to the next allocation
    L.call("dlpm_b200_training_elements", L.ptr(x_t), L.ptr(e_t), L.ptr(xs_d), L.ptr(tt_d), None, None, L.ptr(sd), T, B, D, alpha,
           20.0, seed, off, base, L.stream_ptr())
    A_tr = philox.sample_A(alpha, seed, off, samples, clamp_a=20.0)
    z_tr = philox.normal(seed, off, samples, D, stream=philox.STREAM_Z)
    bg, bs = sched[1].numpy().astype(np.float64)[tt.numpy()], sched[3].numpy().astype(np.float64)[tt.numpy()]
    want_eps = np.sqrt(A_tr)[:, None] * z_tr
    want_xt = bg[:, None] * xs.numpy() + bs[:, None] * want_eps
    np.testing.assert_allclose(x_t.cpu().numpy(), want_xt, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(e_t.cpu().numpy(), want_eps, rtol=1e-4, atol=2e-4)
