"""A/B of library options on the graph-replayed sampling loop (CIFAR shape, batch 512), alternating in ONE process so
that box-to-box and thermal differences cancel.   python tools/bench_options.py pdl=1 [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlpm_b200  # noqa: E402
from dlpm_b200 import GenerativeLevyProcess, _lib  # noqa: E402
from dlpm_b200.init_utils import randomize_parameters_  # noqa: E402
from dlpm_b200.score_nets import UNetModel  # noqa: E402

opt = sys.argv[1] if len(sys.argv) > 1 else "pdl=1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 150
name, val = opt.split("=")
dev = torch.device("cuda", 0)
dlpm_b200.manual_seed(1)
glp = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True)
res = {0: [], 1: []}
for rnd in range(4):
    for on in ((0, 1) if rnd % 2 == 0 else (1, 0)):
        if name not in ("fuse", "gne", "gne8", "dxs", "idskip"):  # "fuse=<max pixels>": GroupNorm applied by the producing convolution's post warps (model attribute)
            _lib.call("dlpm_b200_set_option", name.encode(), int(val) if on else (0 if name != "gn_stats" else 1))
        m = UNetModel(3, 128, 3, 2, (16,), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
        randomize_parameters_(m, 0)
        m = m.to(dev).eval()
        if name == "fuse":
            m.fuse_groupnorm, m.fuse_groupnorm_max_pixels = bool(on), int(val)
        if name == "dxs":  # "dxs=1": the final conv with its horizontal taps stacked along N (the default) against the 9-tap form
            m.dx_stacked_out_conv = bool(on)
        if name == "idskip":  # "idskip=1024": identity skips of maps with >= 1024 pixels as unit-weight 1x1 skip convs in the K loop
            m.identity_skip_as_conv_min_pixels = int(val) if on else (1 << 30)
        if name == "gne8":  # "gne8=1": GroupNorm in the epilogue on the 8x8 maps too (the default)
            m.fuse_groupnorm_epilogue_8x8 = bool(on)
        if name == "gne":  # "gne=1": GroupNorm in the epilogue on the 16x16 maps (the default) against separate passes
            m.fuse_groupnorm_epilogue = bool(on)
        fn = lambda: glp.sample({"default": m}, [512, 3, 32, 32], reverse_steps=steps, clamp_a=20, clamp_eps=200)
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        res[on].append(a.elapsed_time(b) / (steps - 1))
        del m
print(opt, "off: ms/step", ["%.3f" % v for v in res[0]], " on:", ["%.3f" % v for v in res[1]])
