"""Micro-benchmark of "GroupNorm in the epilogue" (conv_tc.cu GNE kernels) on the 16x16 layers of the benchmark UNet:
the convolution with its consumer's GroupNorm applied from TMEM against the convolution + a separate k_gn_apply pass.

    python tools/bench_conv_gne.py [--ncu]     # --ncu: one GNE launch per shape only (run under ncu)
"""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlpm_b200 import _lib as L  # noqa: E402

L.load()
ncu = "--ncu" in sys.argv
for a in sys.argv[1:]:
    if a.startswith("xf_dbg="):
        L.call("dlpm_b200_set_option", b"xf_dbg", int(a.split("=")[1]))
B, H = 512, 16
for (Ci, Co, k, res) in [(128, 256, 3, False), (256, 256, 3, False), (256, 256, 3, True), (512, 256, 3, False)]:
    x = torch.randn(B, H, H, Ci, device="cuda").bfloat16()
    w = (torch.randn(Co, k * k * Ci, device="cuda") / math.sqrt(k * k * Ci)).bfloat16()
    b = torch.randn(Co, device="cuda")
    r = torch.randn(B, H, H, Co, device="cuda").bfloat16() if res else None
    out = torch.zeros(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
    dst = torch.zeros(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
    gamma, beta = torch.randn(Co, device="cuda"), torch.randn(Co, device="cuda")
    ss = torch.randn(1, 2 * Co, device="cuda")
    parts = ctypes.c_int(0)
    base = (L.ptr(x), L.ptr(w), L.ptr(b), None, 0, None, 0, L.ptr(r), L.ptr(out))
    L.call("dlpm_b200_conv2d_stats", *base, 0, B, H, H, Ci, Co, k, 1, None, ctypes.byref(parts), L.stream_ptr())
    st = torch.zeros(B, parts.value, Co // 4, 2, device="cuda")

    def fused():
        L.call("dlpm_b200_conv2d_post", *base, B, H, H, Ci, Co, k, 1, L.ptr(st), None, L.ptr(dst), Co, 0, Co // 32, L.ptr(gamma), L.ptr(beta),
               L.ptr(ss), 1, ss.shape[1], 0, 1, L.stream_ptr())

    def separate():
        L.call("dlpm_b200_conv2d_stats", *base, 0, B, H, H, Ci, Co, k, 1, L.ptr(st), None, L.stream_ptr())
        L.call("dlpm_b200_groupnorm_from_stats", L.ptr(dst), L.ptr(out), Co, L.ptr(st), parts.value, None, 0, None, 0, B, H * H, L.ptr(gamma),
               L.ptr(beta), L.ptr(ss), 1, ss.shape[1], 0, 1, L.stream_ptr())

    if ncu:
        fused()
        torch.cuda.synchronize()
        continue
    times = []
    for fn in (separate, fused):
        best = 1e9
        for _ in range(5):
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        times.append(best)
    print("B=%d %dx%d %d->%d k%d res=%d  conv + k_gn_apply %.4f ms   GNE %.4f ms" % (B, H, H, Ci, Co, k, int(res), times[0], times[1]))
