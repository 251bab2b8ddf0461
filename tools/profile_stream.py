"""Launches each streaming kernel (K1 fills, K3 steps) a few times at the bench sizes: the target of the ncu captures in
profiles/ (`ncu --set full -k regex:"k_fill6|k_sas_vec|k_stable_A_vec|k_reverse_step_fast|k_lim_step_vec" python tools/profile_stream.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dlpm_b200  # noqa: E402
from dlpm_b200 import GenerativeLevyProcess, _lib  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dlpm_b200.manual_seed(1234)
sp = _lib.stream_ptr()
inner, outer = 3072, (1 << 28) // 3072
buf = torch.empty(outer * inner, device=dev)
T, B, D = 1000, 4096, 3072
glp = GenerativeLevyProcess(1.7, dev, T, rescale_timesteps=True, isotropic=True)
glp.dlpm.sample_A([B, 3, 32, 32], T)
x = torch.randn(B, D, device=dev)
eps = torch.randn(B, D, device=dev)
coef = torch.tensor([[1.0, 0.999, -0.001, 0.01]] * 8, device=dev, dtype=torch.float32)
for _ in range(2):
    _lib.call("dlpm_b200_normal", _lib.ptr(buf), outer, inner, 1, 2, 0, sp)
    _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 1, 1.7, 200.0, 1.0, 1, 2, 0, sp)
    _lib.call("dlpm_b200_stable_A", _lib.ptr(buf), outer, inner, 1, 1.7, 20.0, 1, 2, 0, sp)
    _lib.call("dlpm_b200_stable_A", _lib.ptr(buf), outer, inner, 2, 1.7, 20.0, 1, 2, 0, sp)
    _lib.call("dlpm_b200_reverse_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(glp.dlpm.Sigmas), _lib.ptr(glp.dlpm.sched), 500, None, T, B, D, 0,
              None, 1, 2, 0, None, sp)
    _lib.call("dlpm_b200_lim_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(coef), 3, None, B, D, 0, 0, 1, 1.7, 200.0, None, 1, 2, 0, None, sp)
torch.cuda.synchronize()
