"""Test helper: a plain-PyTorch (CPU fp32) interpreter of the UNet op list emitted by
``dlpm_b200.score_nets.UNetModel.build_program``.  It validates the architecture walk, the weight packing and the
offsets independently of the CUDA kernels (the C++ engine executes the very same list)."""
import math

import torch
import torch.nn.functional as F

OP_CONV_IN, OP_GN, OP_CONV, OP_UP, OP_ATTN, OP_SPLIT = 0, 1, 2, 3, 4, 5


def interpret(prog, x, t, mc, emulate_bf16=False):
    """x: (B, C, H, W) fp32 NCHW; t: (B,) floats.  Returns (eps NCHW fp32, {buffer id: NHWC tensor})."""
    hd = prog["header"]
    wb = prog["wb"].float().cpu()
    wf = prog["wf"].float().cpu()
    B = x.shape[0]
    ss_total = hd[7]
    E = 4 * mc
    q = (lambda v: v.to(torch.bfloat16).float()) if emulate_bf16 else (lambda v: v)
    # time embedding (k_time_embed / k_emb_layers)
    half = mc // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    e0 = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    w0T = wf[hd[8]:hd[8] + mc * E].reshape(mc, E)
    b0 = wf[hd[9]:hd[9] + E]
    w2T = wf[hd[10]:hd[10] + E * E].reshape(E, E)
    b2 = wf[hd[11]:hd[11] + E]
    wallT = wf[hd[12]:hd[12] + E * ss_total].reshape(E, ss_total)
    ball = wf[hd[13]:hd[13] + ss_total]
    semb = F.silu(F.silu(e0 @ w0T + b0) @ w2T + b2)
    ss = semb @ wallT + ball  # [B, ss_total]
    bufs = {}
    out = None
    for f in prog["ops"]:
        code = f[0]
        if code == OP_CONV_IN:
            _, o, cin, cout, H, W, woff, boff = f[:8]
            w = wf[woff:woff + cout * cin * 9].reshape(cin * 9, cout).t().reshape(cout, cin, 3, 3)
            b = wf[boff:boff + cout]
            bufs[o] = q(F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1))
        elif code == OP_GN:
            _, i0, i1, o, C0, C1, HW, goff, boff, ssoff, silu = f[:11]
            C = C0 + C1
            v = bufs[i0] if i1 < 0 else torch.cat([bufs[i0], bufs[i1]], dim=-1)
            v = v.permute(0, 3, 1, 2)
            y = F.group_norm(v, min(32, C), wf[goff:goff + C], wf[boff:boff + C], eps=1e-5)
            if ssoff >= 0:
                y = y * (1 + ss[:, ssoff:ssoff + C, None, None]) + ss[:, ssoff + C:ssoff + 2 * C, None, None]
            if silu:
                y = F.silu(y)
            bufs[o] = q(y.permute(0, 2, 3, 1))
        elif code == OP_CONV:
            _, i, o, s0, C0, s1, C1, res, H, W, cin, cout, k, stride, woff, boff = f[:16]
            tr, tc_, dy0b, dx0b, oscale, oyb, oxb, n_par = f[16:24]
            dxs = n_par == 3  # dx-stacked thin conv: weights [3 (dx) x 16][3 (dy) x C_in] -> back to the plain [16][(dy, dx, c)] layout
            n_par = 4 if n_par == 4 else 1
            rows = 16 if o < 0 else cout
            ntap = tr * tc_
            K = ntap * cin + C0 + C1
            if dxs:
                w48 = wb[woff:woff + 48 * 3 * cin].reshape(3, 16, 3, cin)          # [dx][co][dy][c]
                wall = w48.permute(1, 2, 0, 3).reshape(1, 16, 9 * cin)            # [co][(dy, dx, c)]
            else:
                wall = wb[woff:woff + n_par * rows * K].reshape(n_par, rows, K)
            bias = wf[boff:boff + cout]
            xin = bufs[i].permute(0, 3, 1, 2)
            Ho, Wo = H // stride, W // stride
            xp = F.pad(xin, (2, 2, 2, 2))
            for par in range(n_par):
                wk = wall[par][:cout]
                dy0, dx0 = (dy0b + (par >> 1), dx0b + (par & 1)) if n_par == 4 else (dy0b, dx0b)
                oy, ox = ((par >> 1), (par & 1)) if n_par == 4 else (oyb, oxb)
                y = torch.zeros(B, cout, Ho, Wo)
                for tI in range(ntap):
                    dy, dx = dy0 + tI // tc_, dx0 + tI % tc_
                    sl = xp[:, :, 2 + dy:2 + dy + H:stride, 2 + dx:2 + dx + W:stride]
                    y = y + torch.einsum("bchw,oc->bohw", sl, wk[:, tI * cin:(tI + 1) * cin])
                col = ntap * cin
                for sid, Cs in ((s0, C0), (s1, C1)):
                    if sid >= 0:
                        y = y + F.conv2d(bufs[sid].permute(0, 3, 1, 2), wk[:, col:col + Cs].reshape(cout, Cs, 1, 1))
                        col += Cs
                y = y + bias[None, :, None, None]
                if res >= 0:
                    y = y + bufs[res].permute(0, 3, 1, 2)
                if o < 0:
                    out = y
                elif oscale == 1:
                    bufs[o] = q(y.permute(0, 2, 3, 1))
                    # fused GroupNorm targets (fields 24..): the consumer's GroupNorm restricted to this conv's channels --
                    # its groups never straddle the halves of a concatenation -- written at channel offset c_off of dst
                    for kk in range(2):
                        dst, dC, c_off, cpg, goff, boff2, ssoff, silu = f[24 + 8 * kk: 32 + 8 * kk]
                        if dst < 0:
                            continue
                        assert cout % cpg == 0 and c_off % cpg == 0
                        # statistics from the fp32 values the epilogue holds (as the kernel does), applied to the bf16-rounded tensor
                        yn = F.group_norm(y, cout // cpg, None, None, eps=1e-5)
                        if emulate_bf16:
                            mean = y.reshape(B, cout // cpg, -1).mean(-1)
                            var = y.reshape(B, cout // cpg, -1).var(-1, unbiased=False)
                            yb = bufs[o].permute(0, 3, 1, 2)
                            yn = ((yb.reshape(B, cout // cpg, -1) - mean[..., None]) / torch.sqrt(var[..., None] + 1e-5)).reshape(y.shape)
                        yn = yn * wf[goff + c_off:goff + c_off + cout][None, :, None, None] + wf[boff2 + c_off:boff2 + c_off + cout][None, :, None, None]
                        if ssoff >= 0:
                            yn = yn * (1 + ss[:, ssoff + c_off:ssoff + c_off + cout, None, None]) + ss[:, ssoff + dC + c_off:ssoff + dC + c_off + cout, None, None]
                        if silu & 1:  # bit 1: "sole reader of the raw output" (GNE kernels skip the raw store)
                            yn = F.silu(yn)
                        if dst not in bufs or bufs[dst].shape[-1] != dC or bufs[dst].shape[1] != Ho:
                            bufs[dst] = torch.zeros(B, Ho, Wo, dC)
                        bufs[dst][..., c_off:c_off + cout] = q(yn.permute(0, 2, 3, 1))
                else:
                    if o not in bufs or bufs[o].shape[1] != Ho * oscale or bufs[o].shape[3] != cout:
                        bufs[o] = torch.zeros(B, Ho * oscale, Wo * oscale, cout)
                    bufs[o][:, oy::oscale, ox::oscale, :] = q(y.permute(0, 2, 3, 1))
        elif code == OP_SPLIT:
            _, o, cin, H, W = f[:5]
            hi = x.to(torch.bfloat16).float()
            lo = (x - hi).to(torch.bfloat16).float()
            v = torch.zeros(B, 32, H, W)
            v[:, 0:cin], v[:, cin:2 * cin], v[:, 2 * cin:3 * cin] = hi, lo, hi
            bufs[o] = v.permute(0, 2, 3, 1)
        elif code == OP_UP:
            _, i, o, H, W, C = f[:6]
            bufs[o] = F.interpolate(bufs[i].permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
        elif code == OP_ATTN:
            _, i, o, L, C, heads = f[:6]
            qkv = bufs[i].reshape(B, L, 3 * C).permute(0, 2, 1)  # [B, 3C, L]
            qkv = qkv.reshape(B * heads, -1, L)
            ch = qkv.shape[1] // 3
            qq, kk, vv = torch.split(qkv, ch, dim=1)
            s = 1 / math.sqrt(math.sqrt(ch))
            wgt = torch.softmax(torch.einsum("bct,bcs->bts", qq * s, kk * s), dim=-1)
            a = torch.einsum("bts,bcs->bct", wgt, vv).reshape(B, C, L)
            hh = int(math.isqrt(L))
            bufs[o] = q(a.permute(0, 2, 1).reshape(B, hh, hh, C))
        else:
            raise ValueError(code)
    return out, bufs
