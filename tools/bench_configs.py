"""Secondary measurements for the other BASELINE.json configs (C1 2-D MLP, C2 MNIST UNet, C4 LIM vs DLPM at 25/100/1000
steps, C5 noise sweep).  Prints one JSON object per line; results are committed under profiles/.  Not the headline bench
(that is bench.py / C3)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlpm_b200  # noqa: E402
from dlpm_b200 import GenerativeLevyProcess, _lib  # noqa: E402
from dlpm_b200.init_utils import randomize_parameters_  # noqa: E402
from dlpm_b200.score_nets import MLPModel, UNetModel  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dlpm_b200.manual_seed(1234)
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def c1():
    p = {"data": {"nfeatures": 2}, "method": "dlpm", "dlpm": {"isotropic": True}, "device": "cuda",
         "model": dict(use_a_t=False, no_a=True, a_pos_emb=False, a_emb_size=32, time_emb_type="learnable", time_emb_size=32,
                       nblocks=4, nunits=64, skip_connection=True, group_norm=True, dropout_rate=0.0, learn_variance=False)}
    torch.manual_seed(0)
    m = MLPModel(p).to(dev).eval()
    glp = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True)
    for n in (10000, 100000):
        ms = timed(lambda: glp.sample({"default": m}, [n, 1, 2], reverse_steps=1000), reps=3)
        print(json.dumps({"config": "C1 2-D MLP, DLPM alpha=1.7, T=1000", "samples": n, "ms": ms, "samples_per_s": n / ms * 1e3,
                          "reference_cpu_samples_per_s_survey": 841}), flush=True)


def unet(ch, in_ch, attn):
    m = UNetModel(in_ch, ch, in_ch, 2, attn, channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
    randomize_parameters_(m, 0)
    return m.to(dev).eval()


def c2():
    m = unet(32, 1, (2, 4))
    glp = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True)
    B = 2000
    ms = timed(lambda: glp.sample({"default": m}, [B, 1, 32, 32], reverse_steps=1000, clamp_a=20, clamp_eps=200), reps=1)
    print(json.dumps({"config": "C2 MNIST 32x32x1 UNet(ch32, attn at 16x16 and 8x8), DLPM alpha=1.7, T=1000, batch 2000", "ms": ms,
                      "samples_per_s": B / ms * 1e3}), flush=True)


def c3b():
    """The reference's other image config, cifar10.yml: the C3 UNet with AttentionBlocks at 8x8 and 4x4 as well (attn_resolutions [4, 8, 16])."""
    m = unet(128, 3, (4, 8, 16))
    glp = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True)
    B = 512
    ms = timed(lambda: glp.sample({"default": m}, [B, 3, 32, 32], reverse_steps=1000, clamp_a=20, clamp_eps=200), reps=1)
    print(json.dumps({"config": "cifar10.yml UNet (ch128, attention at 8x8 / 4x4 / middle: 11 AttentionBlocks), DLPM alpha=1.7, T=1000, batch 512",
                      "ms": ms, "samples_per_s": B / ms * 1e3, "ms_per_network_eval": ms / 999.0}), flush=True)


def each(fn, reps, warm=1):
    """Per-call device times (ms) of `reps` calls after `warm` warm-up calls."""
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return out


def c4():
    """LIM vs DLPM at 25 / 100 / 1000 steps (BASELINE.json configs[3]).  From the second call on the captured loop is an in-place
    update of the engine's cached executable graph (dlpm_b200_graph_sample), so short loops are not dominated by set-up:
    every call's time is listed next to steps x the per-step time of the 1000-step run."""
    m = unet(128, 3, (16,))
    B = 512
    rows = []
    for steps in (1000, 100, 25):
        d = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True)
        li = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True, LIM=True)
        reps = 2 if steps == 1000 else 6
        t_d = each(lambda: d.sample({"default": m}, [B, 3, 32, 32], reverse_steps=steps, clamp_a=20, clamp_eps=200), reps)
        t_l = each(lambda: li.sample({"default": m}, [B, 3, 32, 32], reverse_steps=steps, clamp_eps=200), reps)
        t_o = each(lambda: li.sample({"default": m}, [B, 3, 32, 32], reverse_steps=steps, deterministic=True), reps)
        med = lambda v: sorted(v)[len(v) // 2]
        rows.append({"config": "C4 CIFAR shape, batch 512, UNet ch128", "steps": steps, "reps": reps,
                     "dlpm_ms": med(t_d), "lim_sde_ms": med(t_l), "lim_ode_ms": med(t_o),
                     "dlpm_ms_each": [round(v, 1) for v in t_d], "lim_sde_ms_each": [round(v, 1) for v in t_l],
                     "lim_ode_ms_each": [round(v, 1) for v in t_o],
                     "dlpm_samples_per_s": B / med(t_d) * 1e3, "lim_sde_samples_per_s": B / med(t_l) * 1e3,
                     "lim_ode_samples_per_s": B / med(t_o) * 1e3})
    per_step = rows[0]["dlpm_ms"] / 999.0
    for r in rows:
        r["dlpm_ms_per_network_eval"] = r["dlpm_ms"] / max(r["steps"] - 1, 1)
        r["lim_ms_per_network_eval"] = r["lim_sde_ms"] / r["steps"]
        r["dlpm_vs_steps_x_per_step"] = r["dlpm_ms"] / (per_step * max(r["steps"] - 1, 1))
        print(json.dumps(r), flush=True)


def c5():
    big = torch.empty(64 * 1024 * 1024, device=dev)
    for alpha in (1.5, 1.7, 1.9, 2.0):
        for logn in (20, 24, 28, 30):
            n = 1 << logn
            inner = 3072
            outer = n // inner
            buf = torch.empty(outer * inner, device=dev)
            res = {"config": "C5 noise sweep", "alpha": alpha, "draws": outer * inner}
            for name, call in (
                    ("A_per_element", lambda: _lib.call("dlpm_b200_stable_A", _lib.ptr(buf), outer, inner, 2, alpha, -1.0, 1, 2, 0, _lib.stream_ptr())),
                    ("sas_per_element", lambda: _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 0, alpha, -1.0, 1.0, 1, 2, 0, _lib.stream_ptr())),
                    ("sas_isotropic", lambda: _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 1, alpha, -1.0, 1.0, 1, 2, 0, _lib.stream_ptr()))):
                t = 0.0
                reps = 5
                for _ in range(reps):
                    big.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); call(); b.record()
                    torch.cuda.synchronize()
                    t += a.elapsed_time(b)
                ms = t / reps
                res[name] = {"ms": ms, "GB/s": 4 * outer * inner / ms / 1e6, "frac_of_measured_hbm_peak": 4 * outer * inner / ms / 1e6 / PEAK}
            print(json.dumps(res), flush=True)
            del buf


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
    for w in which:
        {"c1": c1, "c2": c2, "c3b": c3b, "c4": c4, "c5": c5}[w]()
