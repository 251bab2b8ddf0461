"""Micro-benchmark: cost of the GroupNorm-statistics epilogue of k_conv_tc (with / without, small and large batch)."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlpm_b200 import _lib as L  # noqa: E402

L.load()
for (B, H, Ci, Co) in [(2, 16, 256, 256), (512, 16, 256, 256), (2, 32, 128, 128), (512, 32, 128, 128), (1024, 32, 128, 128)]:
    x = torch.randn(B, H, H, Ci, device="cuda").bfloat16()
    w = (torch.randn(Co, 9 * Ci, device="cuda") / math.sqrt(9 * Ci)).bfloat16()
    b = torch.randn(Co, device="cuda")
    out = torch.zeros(B, H, H, Co, device="cuda", dtype=torch.bfloat16)
    parts = ctypes.c_int(0)
    args = (L.ptr(x), L.ptr(w), L.ptr(b), None, 0, None, 0, None, L.ptr(out), 0, B, H, H, Ci, Co, 3, 1)
    L.call("dlpm_b200_conv2d_stats", *args, None, ctypes.byref(parts), L.stream_ptr())
    st = torch.zeros(B, parts.value, Co // 4, 2, device="cuda")
    res = []
    for stats in (None, st):
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            L.call("dlpm_b200_conv2d_stats", *args, L.ptr(stats) if stats is not None else None, None, L.stream_ptr())
            e0.record()
            for _ in range(10):
                L.call("dlpm_b200_conv2d_stats", *args, L.ptr(stats) if stats is not None else None, None, L.stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        res.append(best)
    print("B=%d H=%d %d->%d  plain %.4f ms  stats %.4f ms  (+%.4f)" % (B, H, Ci, Co, res[0], res[1], res[1] - res[0]))
