"""Run a few UNet forwards of the benchmark configuration (for ncu captures; see profiles/README.md)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlpm_b200.init_utils import randomize_parameters_  # noqa: E402
from dlpm_b200.score_nets import UNetModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--forwards", type=int, default=2)
args = ap.parse_args()
m = UNetModel(3, 128, 3, 2, (16,), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
randomize_parameters_(m, 0)
m = m.cuda().eval()
x = torch.randn(args.batch, 3, 32, 32, device="cuda")
t = torch.full((args.batch,), 0.5, device="cuda")
for _ in range(args.forwards):
    y = m(x, t)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
