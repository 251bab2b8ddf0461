"""Micro-benchmark of the HBM-bound streaming kernels (K1 noise fills, K3 reverse / LIM steps), each timed alone with
CUDA events and an L2 flush (256 MB write) in between.  One JSON object per line.

    python tools/bench_stream.py            # all
    python tools/bench_stream.py k3         # only the step kernels
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlpm_b200  # noqa: E402
from dlpm_b200 import GenerativeLevyProcess, _lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dlpm_b200.manual_seed(1234)
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
big = torch.empty(64 * 1024 * 1024, device=dev)
VARIANTS = [int(v) for v in os.environ.get("K3_VARIANTS", "0").split(",")]


def time_kernel(fn, reps=10):
    fn()
    best, tot = 1e9, 0.0
    for _ in range(reps):
        big.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best, tot = min(best, ms), tot + ms
    return tot / reps, best


def report(name, nbytes, fn, **kw):
    ms, best = time_kernel(fn)
    print(json.dumps({"kernel": name, "bytes": nbytes, "ms": round(ms, 5), "ms_best": round(best, 5), "GB/s": round(nbytes / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(nbytes / ms / 1e6 / PEAK, 4), **kw}), flush=True)


def k1():
    inner = 3072
    for logn in (24, 28):
        outer = (1 << logn) // inner
        n = outer * inner
        buf = torch.empty(n, device=dev)
        sp = _lib.stream_ptr()
        report("normal_fill", 4 * n, lambda: _lib.call("dlpm_b200_normal", _lib.ptr(buf), outer, inner, 1, 2, 0, sp), draws=n)
        for alpha in (1.7, 2.0):
            report("sas_isotropic", 4 * n, lambda: _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 1, alpha, 200.0, 1.0, 1, 2, 0, sp), draws=n, alpha=alpha)
            report("A_isotropic", 4 * n, lambda: _lib.call("dlpm_b200_stable_A", _lib.ptr(buf), outer, inner, 1, alpha, 20.0, 1, 2, 0, sp), draws=n, alpha=alpha)
        report("sas_per_element", 4 * n, lambda: _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 0, 1.7, 200.0, 1.0, 1, 2, 0, sp), draws=n, alpha=1.7,
               note="XU-bound: 9 MUFU per draw (3 sin, 5 lg2, 1 ex2) + 2 per normal")
        report("A_per_element", 4 * n, lambda: _lib.call("dlpm_b200_stable_A", _lib.ptr(buf), outer, inner, 2, 1.7, 20.0, 1, 2, 0, sp), draws=n, alpha=1.7,
               note="XU-bound: 9 MUFU per draw")
        del buf


def k3():
    T, D = 1000, 3072
    for B in (512, 4096):
        shape = [B, 3, 32, 32]
        glp = GenerativeLevyProcess(1.7, dev, T, rescale_timesteps=True, isotropic=True)
        d = glp.dlpm
        d.sample_A(shape, T)
        x = torch.randn(shape, device=dev)
        eps = torch.randn(shape, device=dev)
        eps16 = eps.to(torch.bfloat16)
        sp = _lib.stream_ptr()
        n = B * D
        # same-traffic library baseline: torch's elementwise add (2 reads + 1 write of the same tensors, no RNG)
        report("torch_add_same_traffic", 12 * n, lambda: torch.add(x, eps, out=x), B=B)
        x.normal_()
        for variant in VARIANTS:
            _lib.call("dlpm_b200_set_option", b"k3_variant", variant)
            report("reverse_step_fp32eps", 12 * n, lambda: _lib.call("dlpm_b200_reverse_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(d.Sigmas), _lib.ptr(d.sched), 500,
                                                                      None, T, B, D, 0, None, 1, 2, 0, None, sp), B=B, k3_variant=variant)
            x.normal_()
        _lib.call("dlpm_b200_set_option", b"k3_variant", 0)
        x.normal_()
        report("reverse_step_bf16eps", 10 * n, lambda: _lib.call("dlpm_b200_reverse_step", _lib.ptr(x), _lib.ptr(eps16), _lib.ptr(d.Sigmas), _lib.ptr(d.sched), 500,
                                                                  None, T, B, D, _lib.STEP_EPS_BF16, None, 1, 2, 0, None, sp), B=B)
        x.normal_()
        report("dlim_step", 12 * n, lambda: _lib.call("dlpm_b200_dlim_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(d.sched), 500, None, T, B, D, 0, None, sp), B=B)
        coef = torch.tensor([[1.0, 0.999, -0.001, 0.01]] * 8, device=dev, dtype=torch.float32)
        x.normal_()
        report("lim_sde_step", 12 * n, lambda: _lib.call("dlpm_b200_lim_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(coef), 3, None, B, D, 0, 0, 1, 1.7, 200.0, None,
                                                          1, 2, 0, None, sp), B=B)


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["k1", "k3"]):
        {"k1": k1, "k3": k3}[w]()
