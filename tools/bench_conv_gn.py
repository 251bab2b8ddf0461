"""Micro-benchmark: conv2d vs conv2d_gn (normalise-on-load) for the two dominant layer shapes."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlpm_b200 import _lib as L  # noqa: E402

L.load()
for o in sys.argv[1:]:
    k, v = o.split("=")
    L.call("dlpm_b200_set_option", k.encode(), int(v))


def bench(fn):
    best = 1e9
    for _ in range(4):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    return best


for (B, H, Ci, Co) in [(512, 16, 256, 256), (512, 32, 128, 128), (512, 32, 256, 128)]:
    x = torch.randn(B, H, H, Ci, device="cuda").bfloat16()
    w = (torch.randn(Co, 9 * Ci, device="cuda") / math.sqrt(9 * Ci)).bfloat16()
    b = torch.randn(Co, device="cuda")
    ab = torch.randn(B, Ci, 2, device="cuda") * 0.3
    out = torch.zeros(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
    plain = bench(lambda: L.call("dlpm_b200_conv2d", L.ptr(x), L.ptr(w), L.ptr(b), None, 0, None, 0, None, L.ptr(out), 0, B, H, H, Ci, Co, 3, 1,
                                 L.stream_ptr()))
    fused = bench(lambda: L.call("dlpm_b200_conv2d_gn", L.ptr(x), None, 0, L.ptr(ab), L.ptr(w), L.ptr(b), None, 0, None, 0, None, L.ptr(out), 0,
                                 B, H, H, Ci, Co, None, None, L.stream_ptr()))
    fl = 2.0 * B * H * H * Co * 9 * Ci
    print("%s B=%d H=%d %d->%d  plain %.4f ms (%.0f TF)  fused %.4f ms (%.0f TF)" % (sys.argv[1:], B, H, Ci, Co, plain, fl / plain / 1e9, fused,
                                                                                fl / fused / 1e9))
