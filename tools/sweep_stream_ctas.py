import os, sys, json, torch
sys.path.insert(0, "/root/repo")
import dlpm_b200
from dlpm_b200 import GenerativeLevyProcess, _lib
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
big = torch.empty(64 * 1024 * 1024, device=dev)
def tk(fn, reps=8):
    fn(); tot = 0.0; best = 1e9
    for _ in range(reps):
        big.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b); tot += ms; best = min(best, ms)
    return tot / reps, best
sp = _lib.stream_ptr()
inner, outer = 3072, (1 << 28) // 3072
n = inner * outer
buf = torch.empty(n, device=dev)
T, B, D = 1000, 4096, 3072
glp = GenerativeLevyProcess(1.7, dev, T, rescale_timesteps=True, isotropic=True)
glp.dlpm.sample_A([B, 3, 32, 32], T)
x = torch.randn(B, D, device=dev); eps = torch.randn(B, D, device=dev)
for ctas in [int(v) for v in sys.argv[1].split(",")]:
    _lib.call("dlpm_b200_set_option", b"stream_ctas", ctas)
    m1 = tk(lambda: _lib.call("dlpm_b200_normal", _lib.ptr(buf), outer, inner, 1, 2, 0, sp))
    m2 = tk(lambda: _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 1, 1.7, 200.0, 1.0, 1, 2, 0, sp))
    m3 = tk(lambda: _lib.call("dlpm_b200_reverse_step", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(glp.dlpm.Sigmas), _lib.ptr(glp.dlpm.sched), 500, None, T, B, D, 0, None, 1, 2, 0, None, sp))
    x.normal_()
    print(json.dumps({"stream_ctas": ctas, "normal_ms": m1, "normal_GBs": 4 * n / m1[0] / 1e6, "sas_ms": m2, "sas_GBs": 4 * n / m2[0] / 1e6, "k3_ms": m3, "k3_GBs": 12 * B * D / m3[0] / 1e6}), flush=True)
