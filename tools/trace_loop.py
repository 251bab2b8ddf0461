"""Kernel timeline of a few graph-replayed sampling steps through torch.profiler (CUPTI activity records; no nsys in the
image): per-kernel durations inside the loop and the idle gaps between consecutive kernels.

    python tools/trace_loop.py [reverse_steps]
"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlpm_b200  # noqa: E402
from dlpm_b200 import GenerativeLevyProcess  # noqa: E402
from dlpm_b200.init_utils import randomize_parameters_  # noqa: E402
from dlpm_b200.score_nets import UNetModel  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 12
for a in sys.argv[1:]:  # library options for A/B runs: --opt=name=value
    if a.startswith("--opt="):
        from dlpm_b200 import _lib
        k, v = a[6:].split("=")
        _lib.call("dlpm_b200_set_option", k.encode(), int(v))
dev = torch.device("cuda", 0)
dlpm_b200.manual_seed(1)
m = UNetModel(3, 128, 3, 2, (16,), channel_mult=(1, 2, 2, 2), num_heads=4, use_scale_shift_norm=True)
randomize_parameters_(m, 0)
m = m.to(dev).eval()
m.fuse_groupnorm = "--fuse" in sys.argv  # producer-side GroupNorm (off by default)
glp = GenerativeLevyProcess(1.7, dev, 1000, rescale_timesteps=True, isotropic=True)
fn = lambda: glp.sample({"default": m}, [512, 3, 32, 32], reverse_steps=steps, clamp_a=20, clamp_eps=200)
fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ev = sorted(ev, key=lambda e: e.time_range.start)
ks = [(e.name, e.time_range.start, e.time_range.end) for e in ev if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
# keep the steady part: the last (steps - 2) * per-step kernels
names = [k[0] for k in ks]
last = max(i for i, n in enumerate(names) if "k_advance" in n)
first = [i for i, n in enumerate(names) if "k_advance" in n][2]  # skip the warm-up step and the first replay
seg = ks[first + 1:last + 1]
n_steps = sum(1 for k in seg if "k_advance" in k[0])
busy = sum(e - s for _, s, e in seg)
span = seg[-1][2] - seg[0][1]
gaps = [seg[i + 1][1] - seg[i][2] for i in range(len(seg) - 1)]
agg = collections.OrderedDict()
for n, s, e in seg:
    key = n.split("(")[0].replace("void ", "").replace("dlpm::", "")[:40]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e - s
print("steps %d  kernels/step %.1f  span/step %.1f us  busy/step %.1f us  idle/step %.1f us (%.1f%%)  mean gap %.2f us" % (
    n_steps, len(seg) / n_steps, span / n_steps, busy / n_steps, (span - busy) / n_steps, 100 * (span - busy) / span,
    sum(gaps) / len(gaps)))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s %6.1f launches/step %9.1f us/step" % (k, c / n_steps, t / n_steps))

if "--sequence" in sys.argv:  # kernel sequence of one steady step with in-loop durations (us), averaged over the steps
    per = len(seg) // n_steps
    print("--- per-kernel in-loop duration (mean over %d steps) ---" % n_steps)
    for i in range(per):
        d = [seg[s * per + i][2] - seg[s * per + i][1] for s in range(n_steps)]
        print("%3d %-40s %8.1f" % (i, seg[i][0].split("(")[0].replace("void ", "").replace("dlpm::", "")[:40], sum(d) / len(d)))
