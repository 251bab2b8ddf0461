"""GPU: the UNet building blocks (K5 tcgen05 conv, K6 GroupNorm, K7 attention, conv_in, upsample, time embedding)
against plain PyTorch fp32 references of the same op evaluated on the bf16-rounded operands (bf16 bar: rtol 2e-2)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from dlpm_b200 import _lib
    _lib.load()
    assert torch.cuda.is_available()
    return _lib


def bf(x):
    return x.to(torch.bfloat16)


def nhwc(x):  # NCHW fp32 -> NHWC bf16 (device)
    return bf(x.permute(0, 2, 3, 1).contiguous()).cuda()


def from_nhwc(y):
    return y.float().permute(0, 3, 1, 2).cpu()


def run_conv(L, x, w, b, stride=1, skips=(), skip_w=None, skip_b=None, residual=None, f32_out=False):
    """x NCHW fp32 (bf16-representable), w torch conv weight.  Returns NCHW fp32 (cpu)."""
    B, C_in, H, W = x.shape
    C_out, _, k, _ = w.shape
    wk = w.permute(0, 2, 3, 1).reshape(C_out, -1)
    bias = b.clone()
    sk = []
    if skips:
        wk = torch.cat([wk, skip_w.reshape(C_out, -1)], dim=1)
        bias = bias + skip_b
        sk = [(nhwc(s), s.shape[1]) for s in skips]
    if f32_out:
        wk = torch.cat([wk, torch.zeros(16 - C_out, wk.shape[1])])
        bias = torch.cat([bias, torch.zeros(16 - C_out)])
    wd, bd, xd = bf(wk).contiguous().cuda(), bias.float().cuda(), nhwc(x)
    Ho, Wo = H // stride, W // stride
    out = (torch.zeros(B, C_out, Ho, Wo, device="cuda") if f32_out
           else torch.zeros(B, Ho, Wo, C_out, device="cuda", dtype=torch.bfloat16))
    rd = nhwc(residual) if residual is not None else None
    s = sk + [(None, 0)] * (2 - len(sk))
    L.call("dlpm_b200_conv2d", L.ptr(xd), L.ptr(wd), L.ptr(bd), L.ptr(s[0][0]), s[0][1], L.ptr(s[1][0]), s[1][1], L.ptr(rd),
           L.ptr(out), 1 if f32_out else 0, B, H, W, C_in, C_out, k, stride, L.stream_ptr())
    torch.cuda.synchronize()
    return out.cpu() if f32_out else from_nhwc(out)


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return bf(torch.randn(*shape, generator=g) * scale).float()


CONV_CASES = [
    # B, H, C_in, C_out, k, stride
    (3, 32, 128, 128, 3, 1),   # 7x in the CIFAR UNet
    (2, 16, 256, 256, 3, 1),
    (5, 8, 256, 256, 3, 1),    # two images per tile, odd batch -> masked tail
    (11, 4, 256, 256, 3, 1),   # eight images per tile
    (2, 16, 512, 256, 3, 1),
    (2, 32, 128, 128, 3, 2),   # Downsample (unet.py:96)
    (3, 16, 256, 256, 3, 2),
    (9, 4, 256, 768, 1, 1),    # attention qkv
    (2, 32, 64, 64, 3, 1),     # tile N = 64
    (2, 32, 32, 32, 3, 1),     # MNIST width: 64-byte swizzle, K block 32
    (2, 16, 96, 64, 3, 1),
    (200, 32, 128, 128, 3, 1),  # > 148 tiles per... persistent loop with several tiles per CTA
]


@pytest.fixture(params=[1, 2], ids=["cta1", "cta_pair"])
def cta_group(request, L):
    """Run the conv tests with single-CTA MMAs and with CTA pairs (tcgen05 cta_group::2)."""
    L.call("dlpm_b200_set_option", b"conv_cta_group", request.param)
    yield request.param
    L.call("dlpm_b200_set_option", b"conv_cta_group", 0)


@pytest.mark.parametrize("B,H,C_in,C_out,k,stride", CONV_CASES)
def test_conv_matches_torch(L, cta_group, B, H, C_in, C_out, k, stride):
    x = rnd(B, C_in, H, H)
    w = rnd(C_out, C_in, k, k, scale=1.0 / math.sqrt(C_in * k * k))
    b = torch.randn(C_out) * 0.1
    got = run_conv(L, x, w, b, stride=stride)
    want = F.conv2d(x, w, b, stride=stride, padding=k // 2)
    err = (got - want).abs().max().item()
    assert got.shape == want.shape
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-2, atol=2e-2, err_msg="max err %g" % err)
    assert (got - want).abs().mean().item() < 4e-3


def test_conv_fused_skip_residual_and_final(L, cta_group):
    # second conv of a ResBlock with channel change: 3x3 on h plus the 1x1 skip conv over cat([x1, x2]) (unet.py:161-195,489)
    B, H, C_out = 2, 32, 128
    h, x1, x2 = rnd(B, 128, H, H, seed=1), rnd(B, 256, H, H, seed=2), rnd(B, 128, H, H, seed=3)
    w = rnd(C_out, 128, 3, 3, scale=1 / math.sqrt(1152))
    sw = rnd(C_out, 384, 1, 1, scale=1 / math.sqrt(384))
    b, sb = torch.randn(C_out) * 0.1, torch.randn(C_out) * 0.1
    got = run_conv(L, h, w, b, skips=(x1, x2), skip_w=sw, skip_b=sb)
    want = F.conv2d(h, w, b, padding=1) + F.conv2d(torch.cat([x1, x2], 1), sw, sb)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-2, atol=3e-2)
    # identity skip: residual added in the epilogue
    res = rnd(B, 128, H, H, seed=4)
    got = run_conv(L, h, w, b, residual=res)
    np.testing.assert_allclose(got.numpy(), (F.conv2d(h, w, b, padding=1) + res).numpy(), rtol=2e-2, atol=3e-2)
    # final conv: 128 -> 3, fp32 NCHW output (unet.py:435)
    w3 = rnd(3, 128, 3, 3, scale=1 / math.sqrt(1152))
    b3 = torch.randn(3) * 0.1
    got = run_conv(L, h, w3, b3, f32_out=True)
    np.testing.assert_allclose(got.numpy(), F.conv2d(h, w3, b3, padding=1).numpy(), rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("C0,C1,HW,ss,silu", [(128, 0, 1024, False, True), (256, 128, 1024, False, True), (256, 256, 64, True, True),
                                              (256, 0, 16, False, False), (32, 0, 1024, True, True), (64, 32, 256, False, True)])
def test_groupnorm_silu(L, C0, C1, HW, ss, silu):
    B, C = 5, C0 + C1
    H = int(math.isqrt(HW))
    x0 = rnd(B, C0, H, H, seed=5) * 2 + 0.3
    x1 = rnd(B, C1, H, H, seed=6) if C1 else None
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    table = torch.randn(B, 3 * C + 7) * 0.3
    off = 5
    out = torch.zeros(B, H, H, C, device="cuda", dtype=torch.bfloat16)
    x0d, x1d, gd, bd = nhwc(x0), (nhwc(x1) if C1 else None), gamma.cuda(), beta.cuda()
    for rows in ((1, B) if ss else (1,)):
        tb = table[:rows].contiguous().cuda()
        L.call("dlpm_b200_groupnorm_silu", L.ptr(out), L.ptr(x0d), C0, L.ptr(x1d), C1, B, HW,
               L.ptr(gd), L.ptr(bd), L.ptr(tb) if ss else None, rows, table.shape[1], off, 1 if silu else 0,
               L.stream_ptr())
        torch.cuda.synchronize()
        x = torch.cat([x0d.float().cpu().permute(0, 3, 1, 2), x1d.float().cpu().permute(0, 3, 1, 2)], 1) if C1 \
            else x0d.float().cpu().permute(0, 3, 1, 2)
        want = F.group_norm(x, min(32, C), gamma, beta, eps=1e-5)
        if ss:
            t = table[:rows]
            scale, shift = t[:, off:off + C, None, None], t[:, off + C:off + 2 * C, None, None]
            want = want * (1 + scale) + shift
        if silu:
            want = want * torch.sigmoid(want)
        np.testing.assert_allclose(from_nhwc(out).numpy(), want.numpy(), rtol=1.5e-2, atol=1.5e-2)


@pytest.mark.parametrize("poly", [-1, 0, 1])  # softmax exponentials: auto / all MUFU.EX2 / a quarter on the FMA pipes
@pytest.mark.parametrize("L_,C,heads", [(16, 256, 4), (64, 256, 4), (256, 64, 4), (64, 64, 4), (16, 64, 4), (1024, 128, 4)])
def test_attention(L, L_, C, heads, poly):
    B = 3
    qkv = rnd(B, 3 * C, L_, seed=7) * (3.0 if L_ == 1024 else 1.0)  # wide score range: exercises the running-maximum rescale
    qd = bf(qkv.permute(0, 2, 1).contiguous()).cuda()  # [B, L, 3C]
    qkv = qd.float().cpu().permute(0, 2, 1).contiguous()  # the reference sees the same bf16-rounded operands
    out = torch.zeros(B, L_, C, device="cuda", dtype=torch.bfloat16)
    L.call("dlpm_b200_set_option", b"attention_poly", poly)
    try:
        L.call("dlpm_b200_attention", L.ptr(out), L.ptr(qd), B, L_, C, heads, L.stream_ptr())
    finally:
        L.call("dlpm_b200_set_option", b"attention_poly", -1)
    q = qkv.reshape(B * heads, -1, L_)
    ch = q.shape[1] // 3
    qq, kk, vv = torch.split(q, ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", qq * s, kk * s), dim=-1)
    want = torch.einsum("bts,bcs->bct", wgt, vv).reshape(B, -1, L_)
    np.testing.assert_allclose(out.float().permute(0, 2, 1).cpu().numpy(), want.numpy(), rtol=1.5e-2, atol=1.5e-2)


def test_conv_in_upsample_time_embedding(L):
    B, H = 3, 32
    x = torch.randn(B, 3, H, H)
    w = torch.randn(128, 3, 3, 3) / math.sqrt(27)
    b = torch.randn(128) * 0.1
    out = torch.zeros(B, H, H, 128, device="cuda", dtype=torch.bfloat16)
    xd, wd, bd = x.cuda(), w.reshape(128, -1).t().contiguous().cuda(), b.cuda()
    L.call("dlpm_b200_conv_in", L.ptr(out), L.ptr(xd), L.ptr(wd), L.ptr(bd), B, 3, 128, H, H, L.stream_ptr())
    np.testing.assert_allclose(from_nhwc(out).numpy(), F.conv2d(x, w, b, padding=1).numpy(), rtol=1e-2, atol=1e-2)
    # the same conv leaving GroupNorm partial statistics (per 4-row band, per channel quad)
    import ctypes
    parts = ctypes.c_int(0)
    L.call("dlpm_b200_conv_in_stats", None, None, None, None, B, 3, 128, H, H, None, ctypes.byref(parts), L.stream_ptr())
    assert parts.value == H // 4
    st = torch.full((B, parts.value, 32, 2), float("nan"), device="cuda")
    out2 = torch.zeros_like(out)
    L.call("dlpm_b200_conv_in_stats", L.ptr(out2), L.ptr(xd), L.ptr(wd), L.ptr(bd), B, 3, 128, H, H, L.ptr(st), None, L.stream_ptr())
    assert torch.equal(out, out2)
    ref = F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1).reshape(B, H * H, 32, 4)
    np.testing.assert_allclose(st.sum(1)[..., 0].cpu().numpy(), ref.sum((1, 3)).numpy(), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(st.sum(1)[..., 1].cpu().numpy(), (ref * ref).sum((1, 3)).numpy(), rtol=1e-3, atol=1e-2)
    # upsample
    y = rnd(B, 64, 8, 8, seed=8)
    up = torch.zeros(B, 16, 16, 64, device="cuda", dtype=torch.bfloat16)
    yd = nhwc(y)
    L.call("dlpm_b200_upsample2x", L.ptr(up), L.ptr(yd), B, 8, 8, 64, L.stream_ptr())
    assert torch.equal(from_nhwc(up), F.interpolate(y, scale_factor=2, mode="nearest"))
    # time embedding + emb_layers
    from oracle import nets
    mc, sst = 128, 1000
    w0, b0 = torch.randn(4 * mc, mc) / math.sqrt(mc), torch.randn(4 * mc) * 0.1
    w2, b2 = torch.randn(4 * mc, 4 * mc) / math.sqrt(4 * mc), torch.randn(4 * mc) * 0.1
    wa, ba = torch.randn(sst, 4 * mc) / math.sqrt(4 * mc), torch.randn(sst) * 0.1
    t = torch.tensor([0.731, 0.05, 0.9])
    ss = torch.zeros(3, sst, device="cuda")
    semb = torch.zeros(3, 4 * mc, device="cuda")
    cu = lambda v: v.contiguous().cuda()
    td_, w0T, b0d, w2T, b2d, waT, bad = t.cuda(), cu(w0.t()), cu(b0), cu(w2.t()), cu(b2), cu(wa.t()), cu(ba)
    L.call("dlpm_b200_time_embedding", L.ptr(ss), L.ptr(semb), L.ptr(td_), None, 0.0, 3, mc, sst, L.ptr(w0T), L.ptr(b0d),
           L.ptr(w2T), L.ptr(b2d), L.ptr(waT), L.ptr(bad), L.stream_ptr())
    emb = nets.timestep_embedding(t, mc)
    emb = F.linear(F.silu(F.linear(emb, w0, b0)), w2, b2)
    want = F.linear(F.silu(emb), wa, ba)
    np.testing.assert_allclose(ss.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    # device-side step counter: t = *t_dev * inv_T
    td = torch.tensor([731], dtype=torch.int32).cuda()
    ss1 = torch.zeros(1, sst, device="cuda")
    L.call("dlpm_b200_time_embedding", L.ptr(ss1), L.ptr(semb), None, L.ptr(td), 0.001, 1, mc, sst, L.ptr(w0T), L.ptr(b0d),
           L.ptr(w2T), L.ptr(b2d), L.ptr(waT), L.ptr(bad), L.stream_ptr())
    np.testing.assert_allclose(ss1.cpu().numpy()[0], want.numpy()[0], rtol=1e-4, atol=1e-4)


STATS_CASES = [
    # B, H, C_in, C_out, stride, C1 (second GroupNorm source: a plain tensor through its own conv), ss
    (3, 32, 128, 128, 1, 0, True),     # tall, two sub-tiles per CTA
    (5, 16, 256, 256, 1, 0, False),    # tall N = 256 pairs / single CTAs
    (2, 32, 128, 128, 2, 0, True),     # Downsample conv emits statistics too
    (3, 32, 128, 256, 1, 128, True),   # concat 256 + 128 -> groups of 12 channels straddle the two sources
    (2, 16, 256, 256, 1, 256, False),
    (5, 8, 256, 256, 1, 0, True),      # two images per tile (odd batch: masked tail): one statistics row per epilogue warp
    (3, 8, 256, 256, 1, 256, False),
]


@pytest.mark.parametrize("B,H,C_in,C_out,stride,C1,ss", STATS_CASES)
def test_conv_epilogue_statistics_feed_groupnorm(L, cta_group, B, H, C_in, C_out, stride, C1, ss):
    """conv2d_stats leaves per-quad (sum, sum of squares) partial rows; groupnorm_from_stats / groupnorm_fold must
    reproduce GroupNorm32 (+ scale-shift, SiLU) of the stored tensor without a statistics pass."""
    import ctypes

    def conv_with_stats(x, C_o, seed):
        Bc, Ci, Hc, _ = x.shape
        w = rnd(C_o, Ci, 3, 3, scale=1.0 / math.sqrt(Ci * 9), seed=seed)
        b = torch.randn(C_o) * 0.1
        wd, bd, xd = bf(w.permute(0, 2, 3, 1).reshape(C_o, -1)).contiguous().cuda(), b.cuda(), nhwc(x)
        Ho = Hc // stride
        out = torch.zeros(Bc, Ho, Ho, C_o, device="cuda", dtype=torch.bfloat16)
        parts = ctypes.c_int(0)
        args = (L.ptr(xd), L.ptr(wd), L.ptr(bd), None, 0, None, 0, None, L.ptr(out), 0, Bc, Hc, Hc, Ci, C_o, 3, stride)
        L.call("dlpm_b200_conv2d_stats", *args, None, ctypes.byref(parts), L.stream_ptr())
        assert parts.value > 0
        st = torch.full((Bc, parts.value, C_o // 4, 2), float("nan"), device="cuda")
        L.call("dlpm_b200_conv2d_stats", *args, L.ptr(st), ctypes.byref(parts), L.stream_ptr())
        torch.cuda.synchronize()
        return out, st, parts.value

    x = rnd(B, C_in, H, H, seed=11)
    y0, st0, p0 = conv_with_stats(x, C_out, 1)
    Ho = H // stride
    # the partial rows add up to the per-quad sums of the stored tensor (fp32 values before the bf16 rounding)
    yq = y0.float().reshape(B, Ho * Ho, C_out // 4, 4)
    tot = st0.sum(1).cpu()
    assert torch.isfinite(tot).all()
    np.testing.assert_allclose(tot[..., 0].numpy(), yq.sum((1, 3)).cpu().numpy(), rtol=2e-2, atol=0.02 * Ho * Ho ** 0.5)
    np.testing.assert_allclose(tot[..., 1].numpy(), (yq * yq).sum((1, 3)).cpu().numpy(), rtol=2e-2, atol=1.0)
    y1 = st1 = None
    p1 = 0
    if C1:
        y1, st1, p1 = conv_with_stats(rnd(B, 128 if H > 8 else 256, H, H, seed=12) * 1.7, C1, 2)
    C = C_out + C1
    gamma, beta = (1 + 0.1 * torch.randn(C)).cuda(), (0.1 * torch.randn(C)).cuda()
    table = (torch.randn(B, 2 * C + 3) * 0.3).cuda()
    want = torch.zeros(B, Ho, Ho, C, device="cuda", dtype=torch.bfloat16)
    got = torch.zeros_like(want)
    tail = (L.ptr(gamma), L.ptr(beta), L.ptr(table) if ss else None, B, table.shape[1], 2)
    L.call("dlpm_b200_groupnorm_silu", L.ptr(want), L.ptr(y0), C_out, L.ptr(y1), C1, B, Ho * Ho, *tail, 1, L.stream_ptr())
    L.call("dlpm_b200_groupnorm_from_stats", L.ptr(got), L.ptr(y0), C_out, L.ptr(st0), p0, L.ptr(y1), C1, L.ptr(st1), p1, B,
           Ho * Ho, *tail, 1, L.stream_ptr())
    ab = torch.zeros(B, C, 2, device="cuda")
    L.call("dlpm_b200_groupnorm_fold", L.ptr(ab), C_out, L.ptr(st0), p0, C1, L.ptr(st1), p1, B, Ho * Ho, *tail, 0, L.stream_ptr())
    torch.cuda.synchronize()
    # statistics of the unrounded fp32 conv results vs statistics of the bf16 tensor: well inside the bf16 bar
    np.testing.assert_allclose(got.float().cpu().numpy(), want.float().cpu().numpy(), rtol=1.5e-2, atol=1.5e-2)
    yc = torch.cat([y0, y1], -1).float() if C1 else y0.float()
    z = yc * ab[:, None, None, :, 0] + ab[:, None, None, :, 1]
    np.testing.assert_allclose((z * torch.sigmoid(z)).cpu().numpy(), want.float().cpu().numpy(), rtol=1.5e-2, atol=1.5e-2)


GN_CONV_CASES = [
    # B, H, C0, C1 (second raw source), C_out, residual, f32_out
    (3, 32, 128, 0, 128, True, False),
    (2, 32, 256, 128, 128, False, False),   # concat 256 + 128: groups straddle the sources, two activation maps
    (5, 16, 256, 0, 256, True, False),      # N = 256 (CTA pairs in auto mode; single CTAs fall back to ... see below)
    (2, 16, 256, 256, 256, False, False),
    (3, 32, 128, 0, 3, False, True),        # final GroupNorm + SiLU + conv to 3 fp32 channels
    (150, 32, 128, 0, 128, False, False),   # several work items per CTA: barrier phases wrap
]


@pytest.mark.parametrize("B,H,C0,C1,C_out,residual,f32_out", GN_CONV_CASES)
def test_conv_normalise_on_load(L, B, H, C0, C1, C_out, residual, f32_out):
    """conv2d_gn == conv3x3(SiLU(GroupNorm32([x0 | x1]) * (1 + scale) + shift)): the normalised tensor only ever exists in
    shared memory.  Statistics come from the epilogues of the convs that produced x0 / x1."""
    import ctypes

    def produce(Ci, Co, seed, scale):
        x = rnd(B, Ci, H, H, seed=seed)
        w = rnd(Co, Ci, 3, 3, scale=scale / math.sqrt(Ci * 9), seed=seed + 1)
        b = torch.randn(Co) * 0.3
        wd, bd, xd = bf(w.permute(0, 2, 3, 1).reshape(Co, -1)).contiguous().cuda(), b.cuda(), nhwc(x)
        out = torch.zeros(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
        parts = ctypes.c_int(0)
        args = (L.ptr(xd), L.ptr(wd), L.ptr(bd), None, 0, None, 0, None, L.ptr(out), 0, B, H, H, Ci, Co, 3, 1)
        L.call("dlpm_b200_conv2d_stats", *args, None, ctypes.byref(parts), L.stream_ptr())
        st = torch.zeros(B, parts.value, Co // 4, 2, device="cuda")
        L.call("dlpm_b200_conv2d_stats", *args, L.ptr(st), ctypes.byref(parts), L.stream_ptr())
        return out, st, parts.value

    y0, st0, p0 = produce(128, C0, 21, 2.0)
    y1, st1, p1 = produce(128, C1, 23, 1.0) if C1 else (None, None, 0)
    C = C0 + C1
    gamma, beta = (1 + 0.1 * torch.randn(C)).cuda(), (0.1 * torch.randn(C)).cuda()
    table = (torch.randn(B, 2 * C + 3) * 0.3).cuda()
    ab = torch.zeros(B, C, 2, device="cuda")
    L.call("dlpm_b200_groupnorm_fold", L.ptr(ab), C0, L.ptr(st0), p0, C1, L.ptr(st1), p1, B, H * H, L.ptr(gamma), L.ptr(beta),
           L.ptr(table), B, table.shape[1], 2, 1, L.stream_ptr())
    w = rnd(C_out, C, 3, 3, scale=1.0 / math.sqrt(C * 9), seed=31)
    b = torch.randn(C_out) * 0.1
    wk = w.permute(0, 2, 3, 1).reshape(C_out, -1)
    bias = b.clone()
    if f32_out:
        wk = torch.cat([wk, torch.zeros(16 - C_out, wk.shape[1])])
        bias = torch.cat([bias, torch.zeros(16 - C_out)])
    wd, bd = bf(wk).contiguous().cuda(), bias.cuda()
    res = rnd(B, C_out, H, H, seed=33) if residual else None
    rd = nhwc(res) if residual else None
    out = (torch.zeros(B, C_out, H, H, device="cuda") if f32_out else torch.zeros(B, H, H, C_out, device="cuda", dtype=torch.bfloat16))
    L.call("dlpm_b200_conv2d_gn", L.ptr(y0), L.ptr(y1), C1, L.ptr(ab), L.ptr(wd), L.ptr(bd), None, 0, None, 0, L.ptr(rd), L.ptr(out),
           1 if f32_out else 0, B, H, H, C0, C_out, None, None, L.stream_ptr())
    torch.cuda.synchronize()
    got = out.cpu() if f32_out else from_nhwc(out)
    # reference: fp32 GroupNorm of the bf16 tensors, SiLU, rounded to bf16 like the operand the tensor core sees
    x = torch.cat([y0, y1], -1).float().cpu() if C1 else y0.float().cpu()
    x = x.permute(0, 3, 1, 2)
    t = table.cpu()
    gn = F.group_norm(x, 32, gamma.cpu(), beta.cpu(), eps=1e-5) * (1 + t[:, 2:2 + C, None, None]) + t[:, 2 + C:2 + 2 * C, None, None]
    act = bf(gn * torch.sigmoid(gn)).float()
    want = F.conv2d(act, bf(w).float(), b, padding=1)
    if residual:
        want = want + res
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-2, atol=3e-2)
    assert (got - want).abs().mean().item() < 5e-3


# ------------------------------------------------------------------------------------------------------------------
# K5 + K6 fused on the PRODUCER side: the convolution's post warps apply the consumer's GroupNorm (+ scale-shift, SiLU)
# to every finished sample ("GroupNorm in the producer's tail", conv_tc.cu POST kernels)
# ------------------------------------------------------------------------------------------------------------------
POST_CASES = [
    # B, H, C_in, C_out, k, stride, consumer channels (dst_C), c_off, residual, per-sample scale-shift
    (3, 32, 128, 128, 3, 1, 128, 0, False, False),    # 32x32: a sample = 2 (pairs) / 4 (single CTAs) work items
    (37, 32, 128, 128, 3, 1, 128, 0, True, True),     # odd batch, CTA pairs with an idle half, identity residual
    (150, 32, 128, 128, 3, 1, 256, 128, False, False),  # > 148 samples: several units per CTA (slot recycling); second half of a concat
    (5, 16, 256, 256, 3, 1, 256, 0, False, True),
    (40, 16, 256, 256, 3, 1, 512, 0, True, False),    # first half of a 512-channel concat (16-channel groups)
    (2, 32, 128, 128, 3, 2, 256, 128, False, False),  # Downsample output (hs entry) -> second half of a concat
    (5, 8, 256, 256, 3, 1, 256, 0, False, True),      # two images per tile, odd batch
    (300, 8, 256, 256, 3, 1, 512, 256, False, False),
    (11, 4, 256, 256, 3, 1, 256, 0, True, True),      # eight images per tile: statistics per half warp
    (600, 4, 256, 256, 3, 1, 512, 0, False, False),
    (9, 4, 256, 256, 1, 1, 256, 0, True, False),      # attention proj_out (1x1, residual)
    (3, 32, 32, 128, 3, 1, 128, 0, False, False),     # K blocks of 32 channels (the split input conv)
]


# GroupNorm in the EPILOGUE (conv_tc.cu GNE kernels): maps of 256 pixels with CTA pairs -- the pair's accumulator stage is one
# whole sample, the statistics are exchanged through distributed shared memory and the normalised rows come straight from TMEM
GNE_CASES = [
    (40, 16, 256, 256, 3, 1, 256, 0, False, False),   # conv1-type: batch-constant scale / shift (tables filled once)
    (33, 16, 128, 256, 3, 1, 256, 0, False, True),    # odd batch, per-sample scale / shift (tables refilled per item)
    (70, 16, 512, 256, 3, 1, 512, 0, True, False),    # K = 4608, identity residual, first half of a 512-channel concat (16-channel groups)
    (300, 16, 256, 256, 3, 1, 512, 256, False, False),  # several items per CTA pair (parity double buffer), second half of a concat
    (40, 32, 128, 128, 3, 2, 128, 0, False, False),   # Downsample output at 16x16: N tile of 128 channels, 4-channel groups
    (36, 16, 256, 256, 1, 1, 256, 0, True, True),     # 1x1
    # 4x4 maps (mode 2): eight samples per tile, a sample = half a warp, one TMEM pass, statistics by half-warp shuffles
    (11, 4, 256, 256, 3, 1, 256, 0, True, False),     # odd batch (masked tail), identity residual, single CTAs
    (600, 4, 256, 256, 3, 1, 512, 0, False, False),   # CTA pairs, several items per CTA, first half of a concat (16-channel groups)
    (300, 4, 512, 256, 3, 1, 512, 256, False, False), # K = 4608, second half of a concat
    (9, 4, 256, 256, 1, 1, 256, 0, True, False),      # attention proj_out (1x1, residual)
    (64, 8, 256, 256, 3, 2, 256, 0, False, False),    # Downsample output at 4x4
    # 8x8 maps (mode 2): two samples per tile, a sample = two warps, which swap their totals through shared memory
    (5, 8, 256, 256, 3, 1, 256, 0, False, False),     # odd batch: the last tile's second sample is masked
    (300, 8, 256, 256, 3, 1, 512, 256, False, False), # CTA pairs, several items per CTA, second half of a concat
    (130, 8, 512, 256, 3, 1, 256, 0, True, False),    # K = 4608, identity residual
    (64, 16, 256, 256, 3, 2, 256, 0, False, False),   # Downsample output at 8x8
]


def _stat(L, name):
    import ctypes
    v = ctypes.c_int64(0)
    L.call("dlpm_b200_get_stat", name, ctypes.byref(v))
    return v.value


_CASE_ID = lambda c: "B%d_%dx%d_%d-%d_k%d_s%d_dst%d@%d_res%d_ss%d" % (c[0], c[1], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9])


@pytest.mark.parametrize("case", GNE_CASES, ids=_CASE_ID)
def test_conv_with_groupnorm_in_the_epilogue(L, case):
    n0 = _stat(L, b"conv_gne_launches")
    test_conv_with_producer_side_groupnorm(L, case)
    assert _stat(L, b"conv_gne_launches") > n0, "the GNE kernel did not run for this shape"
    L.call("dlpm_b200_set_option", b"conv_gne", 0)  # the same case through the post warps: both flavours stay covered
    try:
        n1 = _stat(L, b"conv_gne_launches")
        test_conv_with_producer_side_groupnorm(L, case)
        assert _stat(L, b"conv_gne_launches") == n1
    finally:
        L.call("dlpm_b200_set_option", b"conv_gne", 1)


@pytest.mark.parametrize("case", POST_CASES, ids=_CASE_ID)
def test_conv_with_producer_side_groupnorm(L, case):
    import ctypes
    B, H, C_in, C_out, k, stride, dst_C, c_off, use_res, per_sample = case
    x = rnd(B, C_in, H, H, seed=1)
    w = rnd(C_out, C_in, k, k, scale=1.0 / math.sqrt(C_in * k * k), seed=2)
    b = rnd(C_out, scale=0.1, seed=3)
    Ho = H // stride
    res = rnd(B, C_out, Ho, Ho, seed=4) if use_res else None
    g = torch.Generator().manual_seed(7)
    gamma, beta = 1 + 0.2 * torch.randn(dst_C, generator=g), 0.2 * torch.randn(dst_C, generator=g)
    ss_rows = B if per_sample else 1
    ss = 0.3 * torch.randn(ss_rows, 2 * dst_C + 5, generator=g)
    ss_off = 3
    silu = 1
    cpg = dst_C // 32
    xd, wd, bd = nhwc(x), bf(w.permute(0, 2, 3, 1).reshape(C_out, -1)).contiguous().cuda(), b.cuda()
    rd = nhwc(res) if use_res else None
    out = torch.zeros(B, Ho, Ho, C_out, device="cuda", dtype=torch.bfloat16)
    dst = torch.full((B, Ho, Ho, dst_C), 7.0, device="cuda", dtype=torch.bfloat16)
    parts = ctypes.c_int(0)
    args = lambda st: (L.ptr(xd), L.ptr(wd), L.ptr(bd), None, 0, None, 0, L.ptr(rd), L.ptr(out), B, H, H, C_in, C_out, k, stride, st,
                       ctypes.byref(parts), L.ptr(dst), dst_C, c_off, cpg, L.ptr(gamma.cuda()), L.ptr(beta.cuda()), L.ptr(ss.cuda()),
                       ss_rows, ss.shape[1], ss_off, silu, L.stream_ptr())
    gd, bed, ssd = gamma.cuda(), beta.cuda(), ss.cuda().contiguous()
    call = lambda st: L.call("dlpm_b200_conv2d_post", L.ptr(xd), L.ptr(wd), L.ptr(bd), None, 0, None, 0, L.ptr(rd), L.ptr(out), B, H, H,
                             C_in, C_out, k, stride, st, ctypes.byref(parts), L.ptr(dst), dst_C, c_off, cpg, L.ptr(gd), L.ptr(bed),
                             L.ptr(ssd), ss_rows, ssd.shape[1], ss_off, silu, L.stream_ptr())
    call(None)
    assert parts.value > 0
    stats = torch.zeros(B * parts.value * (C_out // 4) * 2, device="cuda")
    for rep in range(2):  # second run on a dirty destination: every element must be rewritten
        dst.fill_(7.0 + rep)
        call(L.ptr(stats))
        torch.cuda.synchronize()
    y = F.conv2d(x, w, b, stride=stride, padding=k // 2)
    if use_res:
        y = y + res
    raw = from_nhwc(out)
    assert float((raw - y).abs().max() / y.abs().max()) < 2e-2
    # the consumer's GroupNorm restricted to this conv's channels (statistics of the fp32 values, applied to the bf16 rows)
    G = C_out // cpg
    yg = y.reshape(B, G, -1)
    mean, var = yg.mean(-1), yg.var(-1, unbiased=False)
    yn = ((raw.reshape(B, G, -1) - mean[..., None]) / torch.sqrt(var[..., None] + 1e-5)).reshape(y.shape)
    sl = slice(c_off, c_off + C_out)
    yn = yn * gamma[sl][None, :, None, None] + beta[sl][None, :, None, None]
    scale = ss[:, ss_off:ss_off + dst_C][:, sl]
    shift = ss[:, ss_off + dst_C:ss_off + 2 * dst_C][:, sl]
    yn = F.silu(yn * (1 + scale[:, :, None, None]) + shift[:, :, None, None])
    got = from_nhwc(dst)
    np.testing.assert_allclose(got[:, sl].numpy(), yn.numpy(), rtol=2e-2, atol=2e-2 * float(yn.abs().max()))
    # channels outside [c_off, c_off + C_out) belong to the other producer of the concatenation: untouched
    other = torch.ones(dst_C, dtype=torch.bool)
    other[sl] = False
    if other.any():
        assert torch.equal(got[:, other], torch.full_like(got[:, other], 8.0))
