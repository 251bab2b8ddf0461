"""ctypes side of the UNet engine (``include/dlpm_b200_unet.h``) and the graph-replayed sampling loop."""
import ctypes

import torch

from . import _lib

c_i64, c_int, c_f32, c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p

UNET_SIGNATURES = {
    "dlpm_b200_conv2d": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int,
                         c_int, c_int, c_vp],
    "dlpm_b200_conv2d_stats": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_vp, ctypes.POINTER(c_int), c_vp],
    "dlpm_b200_conv2d_post": [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp,
                              ctypes.POINTER(c_int), c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_int, c_vp],
    "dlpm_b200_groupnorm_from_stats": [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp,
                                       c_int, c_i64, c_i64, c_int, c_vp],
    "dlpm_b200_groupnorm_fold": [c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_i64, c_i64,
                                 c_int, c_vp],
    "dlpm_b200_conv2d_gn": [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_i64, c_int, c_int,
                            c_int, c_int, c_vp, ctypes.POINTER(c_int), c_vp],
    "dlpm_b200_set_option": [ctypes.c_char_p, c_int],
    "dlpm_b200_get_stat": [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int64)],
    "dlpm_b200_groupnorm_silu": [c_vp, c_vp, c_int, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_int, c_i64, c_i64, c_int,
                                 c_vp],
    "dlpm_b200_attention": [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp],
    "dlpm_b200_conv_in": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp],
    "dlpm_b200_conv_in_stats": [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, ctypes.POINTER(c_int), c_vp],
    "dlpm_b200_split_input": [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp],
    "dlpm_b200_upsample2x": [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp],
    "dlpm_b200_time_embedding": [c_vp, c_vp, c_vp, c_vp, c_f32, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "dlpm_b200_unet_create": [ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64), c_vp,
                              c_i64, c_vp, c_i64, c_i64],
    "dlpm_b200_unet_forward": [c_vp, c_vp, c_vp, c_int, c_vp, c_f32, c_vp, c_i64, c_vp],
    "dlpm_b200_graph_sample": [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_int, c_f32, c_f32, ctypes.c_uint64,
                               ctypes.c_uint64, c_i64, c_vp, c_vp],
    "dlpm_b200_graph_sample_stats": [c_vp, ctypes.POINTER(c_int), ctypes.POINTER(c_int)],
    "dlpm_b200_unet_copy_buffer": [c_vp, c_int, c_vp, c_i64, c_vp],
    "dlpm_b200_unet_profile": [c_vp, c_vp, c_vp, c_int, c_vp, c_i64, ctypes.POINTER(c_f32), ctypes.POINTER(ctypes.c_double), c_vp],
    "dlpm_b200_unet_num_launches": [c_vp],
    "dlpm_b200_unet_destroy": [c_vp],
}


def declare(lib):
    for name, argtypes in UNET_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.dlpm_b200_unet_workspace_bytes.argtypes = [c_vp]
    lib.dlpm_b200_unet_workspace_bytes.restype = c_i64


class Engine:
    """Owns one ``dlpm_b200_unet_create`` handle."""

    def __init__(self, prog, max_batch, version):
        lib = _lib.load()
        self.prog = prog
        self.max_batch = int(max_batch)
        self.version = version
        self.names = prog["names"]
        self.bufs = prog["bufs"]
        header = (c_i64 * 16)(*prog["header"])
        flat_ops = [v for op in prog["ops"] for v in op]
        ops = (c_i64 * len(flat_ops))(*flat_ops)
        bufs = (c_i64 * len(prog["bufs"]))(*prog["bufs"])
        self.handle = c_vp()
        wb, wf = prog["wb"], prog["wf"]
        _lib.call("dlpm_b200_unet_create", ctypes.byref(self.handle), header, ops, bufs, _lib.ptr(wb), wb.numel(), _lib.ptr(wf),
                  wf.numel(), self.max_batch)
        torch.cuda.synchronize()
        self.workspace_bytes = lib.dlpm_b200_unet_workspace_bytes(self.handle)

    def forward(self, x, t, t_dev, inv_T, out, B):
        _lib.call("dlpm_b200_unet_forward", self.handle, _lib.ptr(x), _lib.ptr(t), 0 if t is None else t.numel(), _lib.ptr(t_dev),
                  float(inv_T), _lib.ptr(out), B, _lib.stream_ptr())

    def profile(self, x, t, out, B):
        """One forward with per-op CUDA-event timing.  Returns [(opcode, ms, flops)], entry 0 = time embedding."""
        n = len(self.prog["ops"]) + 1
        ms = (c_f32 * n)()
        fl = (ctypes.c_double * n)()
        _lib.call("dlpm_b200_unet_profile", self.handle, _lib.ptr(x), _lib.ptr(t), t.numel(), _lib.ptr(out), B, ms, fl,
                  _lib.stream_ptr())
        codes = [-1] + [op[0] for op in self.prog["ops"]]
        return [(codes[i], float(ms[i]), float(fl[i])) for i in range(n)]

    def num_launches(self):
        return _lib.load().dlpm_b200_unet_num_launches(self.handle)

    def read_buffer(self, name_or_id, B, shape_hwc):
        """Debug: NHWC bf16 activation of the last forward as a float32 NCHW tensor."""
        bid = self.names[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        H, W, C = shape_hwc
        dst = torch.empty((B, H, W, C), device="cuda", dtype=torch.bfloat16)
        assert H * W * C == self.bufs[bid], (H * W * C, self.bufs[bid])
        _lib.call("dlpm_b200_unet_copy_buffer", self.handle, bid, _lib.ptr(dst), B, _lib.stream_ptr())
        return dst.float().permute(0, 3, 1, 2).contiguous()

    def close(self):
        if self.handle:
            torch.cuda.synchronize()
            _lib.load().dlpm_b200_unet_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


LOOP_DLPM, LOOP_DLIM, LOOP_LIM_SDE, LOOP_LIM_ODE = 0, 1, 2, 3


def graph_stats(eng):
    """(instantiations, in-place updates) of the engine's cached executable graph (dlpm_b200_graph_sample_stats)."""
    a, b = c_int(), c_int()
    _lib.call("dlpm_b200_graph_sample_stats", eng.handle, ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def run_lim_loop(model, x, coef_d, t_table, steps, ode, isotropic, alpha, clamp_eps, hist, seed, offset, sample_base,
                 use_graph=True, post=None):
    """LIM reverse loop (sampler.py:218-258) for the image net: per step UNet forward at the continuous time
    t_table[step] (device table, device step counter) + fused LIM update.  Without history the whole loop is ONE call of
    ``dlpm_b200_graph_sample`` (capture + cached executable graph inside the library); ``post`` = ctypes pointer from
    ``_lib.make_post`` fuses the GenerationManager post-processing into the last step."""
    dev = x.device
    B, C, H, W = x.shape
    D = C * H * W
    eng = model.engine(H, W, B)
    ce = -1.0 if clamp_eps is None else float(clamp_eps)
    if hist is None and use_graph:
        _lib.call("dlpm_b200_graph_sample", eng.handle, LOOP_LIM_ODE if ode else LOOP_LIM_SDE, _lib.ptr(x), None, _lib.ptr(coef_d),
                  _lib.ptr(t_table), int(steps), B, 0, 1 if isotropic else 0, float(alpha), ce, seed, offset, sample_base, post,
                  _lib.stream_ptr())
        return
    out = torch.empty((B, model.out_channels, H, W), device=dev, dtype=torch.float32)
    step_dev = torch.zeros((1,), device=dev, dtype=torch.int32)
    for k in range(steps):
        eng.forward(x, t_table, step_dev, 0.0, out, B)
        _lib.call("dlpm_b200_lim_step_post", _lib.ptr(x), _lib.ptr(out), _lib.ptr(coef_d), 0, _lib.ptr(step_dev), B, D, 0, 1 if ode else 0,
                  1 if isotropic else 0, float(alpha), ce, None, seed, offset, sample_base,
                  _lib.ptr(hist[k + 1]) if hist is not None else None, post, steps - 1, _lib.stream_ptr())
        _lib.call("dlpm_b200_advance_counter", _lib.ptr(step_dev), 1, _lib.stream_ptr())


def run_sample_loop(model, x, dlpm, T, mode, flags, hist, seed, z_offset, sample_base, progress=False,
                    use_graph=True, input_scale=None, post=None):
    """x: (B, C, H, W) fp32, updated in place to x_0.  One step = UNet forward (t from the device counter) + fused
    update + counter decrement.  Without history the whole loop is ONE call of ``dlpm_b200_graph_sample``: the library
    captures the step on the current stream, updates its cached executable graph in place and replays it T-1 times
    (nothing is re-instantiated from the second call on; ``graph_stats``).  With history every step needs its own
    destination, so the launches are issued directly."""
    dev = x.device
    B, C, H, W = x.shape
    D = C * H * W
    eng = model.engine(H, W, B)
    if hist is None and use_graph:
        _lib.call("dlpm_b200_graph_sample", eng.handle, LOOP_DLIM if mode == 1 else LOOP_DLPM, _lib.ptr(x),
                  _lib.ptr(dlpm.Sigmas) if mode != 1 else None, _lib.ptr(dlpm.sched), _lib.ptr(input_scale), int(T), B, int(flags), 1,
                  0.0, -1.0, seed, z_offset, sample_base, post, _lib.stream_ptr())
        return
    eps = torch.empty((B, model.out_channels, H, W), device=dev, dtype=torch.float32)
    t_dev = torch.full((1,), T - 1, device=dev, dtype=torch.int32)
    inv_T = float(torch.tensor(1.0 / T, dtype=torch.float32))
    x_in = x if input_scale is None else torch.empty_like(x)  # scale_exploding + input_scaling: net sees x / (1 + barsigma_t)
    for k in range(T - 1):
        h_ptr = _lib.ptr(hist[k + 1]) if hist is not None else None
        if input_scale is not None:
            _lib.call("dlpm_b200_scale_by_step", _lib.ptr(x_in), _lib.ptr(x), _lib.ptr(input_scale), None, 0, _lib.ptr(t_dev), T, B, D,
                      _lib.stream_ptr())
        eng.forward(x_in, None, t_dev, inv_T, eps, B)
        if mode == 1:
            _lib.call("dlpm_b200_dlim_step_post", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(dlpm.sched), 0, _lib.ptr(t_dev), T, B, D, flags,
                      h_ptr, post, _lib.stream_ptr())
        else:
            _lib.call("dlpm_b200_reverse_step_post", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(dlpm.Sigmas), _lib.ptr(dlpm.sched), 0,
                      _lib.ptr(t_dev), T, B, D, flags, None, seed, z_offset, sample_base, h_ptr, post, _lib.stream_ptr())
        _lib.call("dlpm_b200_advance_counter", _lib.ptr(t_dev), -1, _lib.stream_ptr())
