"""Per-op CUDA-event profile of one UNet forward (benchmark configuration), with library options for A/B runs.

    python tools/profile_ops.py                        # table of every op
    python tools/profile_ops.py --opt conv_tall256=0   # same with an option of dlpm_b200_set_option changed
    python tools/profile_ops.py --summary              # only the per-kind totals
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlpm_b200 import _lib  # noqa: E402
from dlpm_b200.init_utils import randomize_parameters_  # noqa: E402
from dlpm_b200.score_nets import UNetModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--opt", action="append", default=[])
ap.add_argument("--summary", action="store_true")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--config", default="cifar", choices=["cifar", "mnist", "cifar10"])
ap.add_argument("--fuse", action="store_true")  # GroupNorms that cannot run in a conv epilogue: the producers' post warps (maps <= 8x8) instead of separate launches
ap.add_argument("--no-fuse", action="store_true")  # (the default; kept for old command lines)
ap.add_argument("--no-gne", action="store_true")  # ... and no GroupNorm in the epilogue on the 16x16 maps either
args = ap.parse_args()
for o in args.opt:
    k, v = o.split("=")
    _lib.call("dlpm_b200_set_option", k.encode(), int(v))
if args.config == "mnist":  # mnist.yml:50-58: ch 32, attention at ds 2 and 4 (16x16 and 8x8)
    m = UNetModel(1, 32, 1, 2, (2, 4), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
elif args.config == "cifar10":  # cifar10.yml:51-59: the benchmark UNet with attention at 8x8 and 4x4 as well
    m = UNetModel(3, 128, 3, 2, (4, 8, 16), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
else:
    m = UNetModel(3, 128, 3, 2, (16,), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
randomize_parameters_(m, 0)
m = m.cuda().eval()
m.fuse_groupnorm = bool(args.fuse) and not args.no_fuse  # default: the product's default engine
m.fuse_groupnorm_epilogue = not args.no_gne
B = args.batch
eng = m.engine(32, 32, B)
x = torch.randn(B, 1 if args.config == "mnist" else 3, 32, 32, device="cuda")
t = torch.full((1,), 0.5, device="cuda")
out = torch.empty_like(x)
eng.profile(x, t, out, B)
acc = None
for _ in range(args.reps):
    prof = eng.profile(x, t, out, B)
    acc = [p[1] for p in prof] if acc is None else [min(a, p[1]) for a, p in zip(acc, prof)]
names = {-1: "time_embedding", 0: "conv_in", 1: "groupnorm_silu", 2: "conv_tc", 3: "upsample2x", 4: "attention", 5: "conv_in_split"}
tot, flops = {}, {}
for i, ((code, _, fl), ms) in enumerate(zip(prof, acc)):
    tot[names[code]] = tot.get(names[code], 0.0) + ms
    flops[names[code]] = flops.get(names[code], 0.0) + fl
    if not args.summary:
        op = eng.prog["ops"][i - 1] if i > 0 else []
        print("%3d %-15s %8.4f ms %8.1f TFLOP/s  %s" % (i, names[code], ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, list(op)[:14]))
print("opts", args.opt, "| total %.3f ms |" % sum(acc), " ".join("%s %.3f" % kv for kv in sorted(tot.items())),
      "| conv %.1f TFLOP/s" % (flops["conv_tc"] / tot["conv_tc"] / 1e9))
