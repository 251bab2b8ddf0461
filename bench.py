#!/usr/bin/env python
"""Benchmark of the DLPM sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is ONE full pass of the hot path over one batch of synthetic input: `GenerativeLevyProcess.sample()`
= alpha-stable noise + Sigma chain + 999 score-network evaluations + 999 fused posterior updates, for the
CIFAR-10-LT configuration (BASELINE.json configs[2]: 32x32x3, UNet ch128 mult (1,2,2,2), alpha=1.7, T=1000,
clamp_a=20, clamp_eps=200; 4096 samples over 8 GPUs = 512 samples per GPU, weak scaling).  Weights are random
(every parameter re-randomised, see dlpm_b200/init_utils.py), data is synthetic (there is no input data: the
sampler starts from in-kernel noise).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALPHA, T_STEPS, IMG, CH = 1.7, 1000, 32, 3
UNET = dict(model_channels=128, channel_mult=(1, 2, 2, 2), num_res_blocks=2, attention_resolutions=(16,), num_heads=4)
GFLOP_PER_SAMPLE_STEP = 11.44  # SURVEY.md section 6 (2*MAC of conv/linear/bmm, torch flop counter on the reference UNet)


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Rows logged so far belong to the warm-up: only later ones are reported."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.rows = self.rows[getattr(self, "first", 0):]
        try:  # diagnostic trace of the timed region (scratch, not tracked)
            if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
                with open(os.path.join(ROOT, "gpurun_out", "clock_rows.csv"), "w") as fh:
                    fh.write(self.Q + "\n" + "\n".join(",".join(r) for r in self.rows) + "\n")
        except OSError:
            pass
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref = verbatim copy installed by oracle/install_ref.py; /root/reference in
# the build container) on the box's host cores; the oracle port only if no copy of the reference is present
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(b_cpu, n_sub, seed=0):
    """One bounded sample of the workload on the host cores.  Returns (samples_per_sec, seconds_spent, description, kind)."""
    from oracle import ref_import
    if ref_import.available():
        from oracle import ref_driver
        v, spent, desc, _ = ref_driver.cpu_arm(b_cpu, n_sub, T=T_STEPS, alpha=ALPHA, img=IMG, ch=CH, seed=seed)
        # contract vocabulary: "reference" = the reference's own code "port" = the oracle restatement
        return v, spent, desc, "reference"
    v, spent, desc = cpu_port_step(b_cpu, n_sub, seed)
    return v, spent, desc, "port"


def reference_source():
    """Where the reference arm's code comes from: '/root/reference', 'oracle/_ref' (verbatim installed copy) or 'oracle port'."""
    from oracle import ref_import
    return {"reference": "/root/reference", "_ref": "oracle/_ref"}.get(ref_import.kind(), "oracle port")


def cpu_port_step(b_cpu, n_sub, seed=0):
    """Fallback when neither /root/reference nor oracle/_ref exists: the oracle port of the reference CPU path --
    the full A/Sigma set-up for T=1000 at batch b_cpu, then n_sub reverse steps, extrapolated linearly."""
    import numpy as np
    import torch
    from dlpm_b200.init_utils import randomize_parameters_
    from dlpm_b200.score_nets import UNetModel
    from oracle import nets, process, stable
    torch.set_num_threads(os.cpu_count() or 1)
    m = UNetModel(CH, UNET["model_channels"], CH, UNET["num_res_blocks"], UNET["attention_resolutions"],
                  channel_mult=UNET["channel_mult"], num_heads=UNET["num_heads"], use_scale_shift_norm=True)
    randomize_parameters_(m, 0)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = dict(UNET)
    rs = np.random.RandomState(seed)
    shape = (b_cpu, CH, IMG, IMG)
    t0 = time.perf_counter()
    sched = process.gen_noise_schedule(ALPHA, T_STEPS)
    A = torch.stack([torch.from_numpy(stable.gen_skewed_levy(ALPHA, shape, isotropic=True, clamp_a=20.0, rng=rs)) for _ in range(T_STEPS)])
    Sig = process.compute_Sigmas(A, sched[0], sched[2])
    x = sched[3][-1] * torch.from_numpy(stable.gen_sas(ALPHA, shape, isotropic=True, clamp_eps=200.0, rng=rs))
    t_setup = time.perf_counter() - t0
    with torch.inference_mode():
        def one(t):
            eps = nets.unet_forward(sd, cfg, x, torch.full((b_cpu,), t / T_STEPS))
            return process.dlpm_step(x, eps, torch.randn(shape), t, Sig, sched)
        x = one(T_STEPS - 1)  # warm-up (thread pools, allocator)
        t1 = time.perf_counter()
        for k in range(n_sub):
            x = one(T_STEPS - 2 - k)
        t_step = (time.perf_counter() - t1) / n_sub
    full = t_setup + t_step * (T_STEPS - 1)
    desc = ("oracle port of the reference CPU path: batch %d, Sigma set-up for T=%d in full (%.2f s) + %d of 999 reverse steps "
            "(%.3f s/step), extrapolated linearly to a full pass" % (b_cpu, T_STEPS, t_setup, n_sub, t_step))
    return b_cpu / full, time.perf_counter() - t0, desc


def run_reference(args, rank):
    if rank != 0:
        return
    vals, spent = [], 0.0
    desc, kind = "", "port"
    for i in range(args.warmup + args.steps):
        v, s, desc, kind = cpu_reference_step(args.cpu_batch, args.cpu_substeps, seed=i)
        spent += s
        if i >= args.warmup:
            vals.append(v)
        if spent > 240 and vals:  # keep the whole arm within a few minutes
            break
    value = sum(vals) / len(vals)
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": "DLPM samples/sec (1000 reverse steps, CIFAR-10 shape)", "value": value, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1000.0 * args.cpu_batch / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.cpu_batch, 1),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "source": reference_source(), "sample": desc},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def workload_config(args, batch_per_gpu, n):
    return {"workload": "CIFAR-10-LT 32x32x3 UNet(ch128, mult 1-2-2-2, 2 res blocks, middle attention) DLPM alpha=1.7, "
                        "T=1000 (999 network evals), clamp_a=20, clamp_eps=200; BASELINE.json configs[2] = 4096 samples over 8 GPUs",
            "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * n, "reverse_steps": args.reverse_steps,
            "parallelism": "batch-sharded x%d (independent Philox streams keyed by global sample index, one NCCL all_gather)" % n,
            "l2_note": "per-step working set (activations %.1f GB at batch %d) exceeds the 126 MB L2; no extra flush needed"
                       % (5.1e-3 * batch_per_gpu, batch_per_gpu)}


def hbm_kernels(torch, _lib, glp, dev, T, hbm_peak, sets=4, reps=5):
    """K3 (fused reverse step, 12 B per element) at the full C3 batch 4096 x 3 x 32 x 32 and K1 (isotropic SaS fill,
    4 B per draw) on 2^28 draws: mean launch duration over graph-replayed launches cycling over `sets` different buffer
    sets (K3: 100 MB each, 400 MB together > the 126 MB L2; K1: 1 GB each), CUDA events on the replaying stream, best of
    `reps` replays.  A library kernel with exactly K3's traffic (torch.add) is timed the same way beside it."""
    out = {}
    Bk = 4096
    kshape = [Bk, CH, IMG, IMG]
    n_el = Bk * CH * IMG * IMG
    d = glp.dlpm
    d.sample_A(kshape, T)
    xs = [torch.randn(kshape, device=dev) for _ in range(sets)]
    es = [torch.randn(kshape, device=dev) for _ in range(sets)]

    def timed(fn_of_set, n_sets, launches=None):
        """Mean launch duration: `launches` launches cycling over the buffer sets, captured in ONE CUDA graph (the way the
        sampling loop issues them: no host launch gap between consecutive kernels) and timed with CUDA events around the
        replay on the replaying stream; best of `reps` replays after a warm-up replay."""
        launches = launches or 4 * n_sets
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for k in range(n_sets):
                fn_of_set(k)  # warm-up outside capture
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(launches):
                    fn_of_set(i % n_sets)
            best = float("inf")
            for r in range(reps + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(side)
                g.replay()
                b.record(side)
                side.synchronize()
                if r > 0:
                    best = min(best, a.elapsed_time(b) / launches)
        torch.cuda.current_stream().wait_stream(side)
        del g
        return best
    ms_k3 = timed(lambda k: _lib.call("dlpm_b200_reverse_step", _lib.ptr(xs[k]), _lib.ptr(es[k]), _lib.ptr(d.Sigmas), _lib.ptr(d.sched),
                                      min(500, T - 1), None, T, Bk, CH * IMG * IMG, 0, None, 1, 2, 0, None, _lib.stream_ptr()), sets)
    ms_add = timed(lambda k: torch.add(xs[k], es[k], out=xs[k]), sets)
    out["reverse_step"] = {"bytes_per_launch": 12 * n_el, "ms": ms_k3, "GB/s": 12 * n_el / ms_k3 / 1e6, "frac": 12 * n_el / ms_k3 / 1e6 / hbm_peak,
                           "same_traffic_torch_add_GB/s": 12 * n_el / ms_add / 1e6,
                           "how": "mean of %d graph-replayed launches cycling over %d buffer sets (%.0f MB together > L2), best of %d replays, before the sampling loop"
                                  % (4 * sets, sets, sets * 8 * n_el / 1e6, reps)}
    del xs, es
    torch.cuda.empty_cache()
    n_noise = 1 << 28
    nn = (n_noise // 3072) * 3072
    bufs = [torch.empty(n_noise, device=dev) for _ in range(2)]
    for name, iso, per_el in (("sas_noise_isotropic", 1, False), ("sas_noise_per_element", 0, False), ("A_per_element", 0, True)):
        if per_el:
            ms = timed(lambda k: _lib.call("dlpm_b200_stable_A", _lib.ptr(bufs[k]), n_noise // 3072, 3072, 2, ALPHA, -1.0, 1, 2, 0,
                                           _lib.stream_ptr()), 2)
        else:
            ms = timed(lambda k: _lib.call("dlpm_b200_sas", _lib.ptr(bufs[k]), None, n_noise // 3072, 3072, iso, ALPHA, 200.0, 1.0, 1, 2, 0,
                                           _lib.stream_ptr()), 2)
        out[name] = {"bytes_per_launch": 4 * nn, "ms": ms, "GB/s": 4 * nn / ms / 1e6, "frac": 4 * nn / ms / 1e6 / hbm_peak}
    out["sas_noise_isotropic"]["generator"] = "Philox4x32-%d" % _lib.load().dlpm_b200_philox_rounds()
    out["sas_noise_isotropic"]["scheme"] = ("sextet: six normals per Philox block (27-bit radius lattice, 15-bit angle), rows of 3072 = 8 x 384 "
                                            "elements; csrc/rng.cuh, oracle/philox.py::normal_sextet")
    out["sas_noise_isotropic"]["limiter"] = ("HBM write (the quad scheme -- four normals per block -- was dispatch-bound at 0.66: 2 quarter-rate "
                                             "IMAD.WIDE per Philox round, profiles/r02_noise.md)")
    out["sas_noise_per_element"]["limiter"] = out["A_per_element"]["limiter"] = (
        "instruction dispatch: one CMS / Kanter transform (3 polynomial sines, 2 lg2, rcp, ex2, Exp(1) tail series) and 64 random bits per draw")
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import dlpm_b200
    from dlpm_b200 import GenerativeLevyProcess, _lib, rng
    from dlpm_b200.init_utils import randomize_parameters_
    from dlpm_b200.score_nets import UNetModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()  # fail loudly if the CUDA library is missing
    B, T = args.batch_per_gpu, args.reverse_steps
    model = UNetModel(CH, UNET["model_channels"], CH, UNET["num_res_blocks"], UNET["attention_resolutions"],
                      channel_mult=UNET["channel_mult"], num_heads=UNET["num_heads"], use_scale_shift_norm=True)
    randomize_parameters_(model, 0)
    model = model.to(dev).eval()
    glp = GenerativeLevyProcess(ALPHA, dev, T, rescale_timesteps=True, isotropic=True)
    models = {"default": model}
    shape = [B, CH, IMG, IMG]
    state = rng.default_state()
    dlpm_b200.manual_seed(1234)
    dlpm_b200.set_sample_base(rank * B)  # global sample index of this rank's first sample
    gathered = torch.empty((world * B, CH, IMG, IMG), device=dev) if world > 1 else None
    host_sched = glp.dlpm._sched_host.clone().pin_memory()
    # the caller of the boundary (bem/GenerationManager.py:29-63) with its post-processing fused into the last step kernel
    # and the device -> host copy of the result into pinned memory issued inside generate()
    manager = dlpm_b200.GenerationManager(glp, shape, is_image=True, reverse_steps=T, clamp_a=20, clamp_eps=200)

    def step(e2e=False):
        if e2e:
            # host -> device: the per-call inputs of sample() are the schedule table and the RNG key / offset
            glp.dlpm.sched.copy_(host_sched, non_blocking=True)
            manager.generate(models, B)  # sample() + fused clamp, (x+1)/2 + pinned async D2H + stream sync
            if world > 1:
                dist.all_gather_into_tensor(gathered, manager.device_samples)
            return manager.samples
        x = glp.sample(models, shape, reverse_steps=T, clamp_a=20, clamp_eps=200)
        if world > 1:
            dist.all_gather_into_tensor(gathered, x)  # the path's only collective
        return x

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    each = {}  # per-step device times of this rank (diagnostic: shows clock / power-cap drift across the run)

    def timed(n, e2e, tag=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [e0]
        w0 = time.perf_counter()
        e0.record()
        for i in range(n):
            step(e2e)
            m = e1 if i == n - 1 else torch.cuda.Event(enable_timing=True)
            m.record()
            marks.append(m)
        if n == 0:
            e1.record()
        barrier()
        wall = time.perf_counter() - w0
        if tag:
            each[tag] = [round(marks[i].elapsed_time(marks[i + 1]), 1) for i in range(len(marks) - 1)]
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1000.0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    # ---- HBM-bound kernels: fused reverse step (12 B/element) and SaS noise (4 B/element).  Timed BEFORE the long loop
    # (the loop leaves the GPU power-capped at ~1590 MHz; these kernels are quoted against the BURST copy peak of
    # MEASURED_PEAKS.json, which was taken on a cool GPU), as the mean of back-to-back launches over buffer sets that together
    # exceed the 126 MB L2 (each launch misses L2; one launch's ~7 us launch latency is not billed to a 25 us kernel)
    pk, pk_src = peaks()
    hbm_peak = float(pk["hbm_gbs"])
    hbm = hbm_kernels(torch, _lib, glp, dev, T, hbm_peak)
    torch.cuda.empty_cache()

    # nvidia-smi needs ~1 s to initialise NVML and briefly contends with kernel launches while it does: start it before
    # the last warm-up step so that only its steady 200 ms polling runs during the timed region; rows logged before the
    # timed region starts are dropped.
    clocks = ClockSampler(local_rank)
    if args.warmup > 1:
        timed(args.warmup - 1, e2e=False, tag="warmup")
    if rank == 0:
        clocks.start()
    if args.warmup > 0:
        timed(1, e2e=False, tag="warmup_last")
    clocks.mark()
    ms_dev, _ = timed(args.steps, e2e=False, tag="timed")
    clk = clocks.stop() if rank == 0 else None
    _, ms_e2e_wall = timed(args.e2e_steps, e2e=True, tag="e2e") if args.e2e_steps > 0 else (0.0, float("nan"))
    ms_per_step = ms_dev / args.steps
    value = world * B / (ms_per_step / 1000.0)
    e2e_value = world * B / (ms_e2e_wall / args.e2e_steps / 1000.0) if args.e2e_steps > 0 else None

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), measured live with CUDA events per op
    eng = model.engine(IMG, IMG, B)
    xs = torch.randn(shape, device=dev)
    tt = torch.full((1,), 0.5, device=dev)
    out = torch.empty_like(xs)
    eng.profile(xs, tt, out, B)
    prof = eng.profile(xs, tt, out, B)
    names = {-1: "time_embedding", 0: "conv_in", 1: "groupnorm_silu", 2: "conv_tc", 3: "upsample2x", 4: "attention", 5: "conv_in_split"}
    breakdown = {}
    for code, ms, fl in prof:
        breakdown[names.get(code, "op%d" % code)] = breakdown.get(names.get(code, "op%d" % code), 0.0) + ms
    if rank == 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "ops_profile.txt"), "w") as fh:
            for i, (code, ms, fl) in enumerate(prof):
                op = eng.prog["ops"][i - 1] if i > 0 else []
                fh.write("%3d %-15s %8.4f ms %8.1f TFLOP/s  %s\n" % (i, names.get(code, "op%d" % code), ms, fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0, op))
    conv = [(ms, fl) for code, ms, fl in prof if code == 2]
    conv_ms, conv_fl = sum(m for m, _ in conv), sum(f for _, f in conv)
    fwd_ms = sum(ms for _, ms, _ in prof)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12
    peak = float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"]))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json"))).get("dram_bytes_per_launch_avg")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "k_conv_tc (tcgen05 implicit-GEMM conv, %d launches per UNet forward)" % len(conv),
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % pk_src,
                "flop_per_launch_avg": conv_fl / len(conv), "ms_per_launch_avg": conv_ms / len(conv),
                "conv_share_of_forward": conv_ms / fwd_ms, "forward_ms_serialised": fwd_ms, "forward_ms_by_op": breakdown,
                "end_to_end_tflops": GFLOP_PER_SAMPLE_STEP * 1e9 * B * (T - 1) / (ms_per_step * 1e-3) / 1e12}

    launches_per_pass = (T - 1) * (eng.num_launches() + 2) + 3
    # ---- same-box GPU baseline: the UNMODIFIED reference with device='cuda' (PyTorch eager + cuDNN, TF32 convs), in the
    # chunk size its own eval config uses (64, dlpm/configs/cifar10_lt.yml:35) and at 256 (its (T,B,C,H,W) tables: 6 GB)
    gpu_eager = None
    if rank == 0 and world == 1 and args.gpu_eager:
        try:
            from oracle import ref_import
            if ref_import.available():
                from oracle import ref_driver
                del xs, out
                torch.cuda.empty_cache()
                runs = [ref_driver.gpu_eager_arm(dev, b, n_steps=20, T=T, alpha=ALPHA, img=IMG, ch=CH) for b in (64, 256)]
                best = max(runs, key=lambda r: r["value"])
                gpu_eager = dict(best, runs=[{"batch": r["batch"], "value": r["value"], "ms_per_reverse_step": r["ms_per_reverse_step"],
                                              "setup_s": r["setup_s"]} for r in runs])
            else:
                gpu_eager = {"unavailable": "no copy of the reference on this box (oracle/_ref missing)"}
        except Exception as e:  # the baseline must never take the bench line down
            gpu_eager = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        cpu_val, _, cpu_desc, cpu_kind = cpu_reference_step(args.cpu_batch, args.cpu_substeps)
        line = {"metric": "DLPM samples/sec (1000 reverse steps, CIFAR-10 shape)", "value": value, "unit": "samples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "ms_each_step": each, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B, world),
                "precision": {"ours": "bf16 activations and conv weights, fp32 accumulation / GroupNorm statistics / x_t state / noise",
                              "reference": "fp32 storage (cuDNN TF32 convolutions when it runs on a GPU)",
                              "parity": "tests/test_gpu_reference_live.py: T=1000 chains + distribution of 1000-step samples vs the live reference"},
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(host_sched.numel() * 4 * world),
                        "d2h_bytes_per_step": int(B * CH * IMG * IMG * 4 * world), "steps": args.e2e_steps,
                        "api": "dlpm_b200.GenerationManager.generate() = GenerativeLevyProcess.sample() with the clamp / (x+1)/2 of "
                               "bem/GenerationManager.py:50-63 fused into the last step kernel + async D2H into pinned memory"},
                "gpu_launches": int(launches_per_pass * (args.steps + args.e2e_steps)), "clocks": clk, "roofline": roofline,
                "hbm_kernels": hbm, "gpu_eager_baseline": gpu_eager,
                "cpu_baseline": {"value": cpu_val, "unit": "samples/s", "cores": os.cpu_count(), "kind": cpu_kind, "source": reference_source(), "sample": cpu_desc},
                "workspace_gb": eng.workspace_bytes / 1e9}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# --workload noise: BASELINE.json configs[4] (alpha-stable noise sweep, 1M - 1B draws, 1/2/4/8 GPUs vs the HBM roofline)
# ------------------------------------------------------------------------------------------------------------------
def run_noise(args, rank, local_rank, world):
    """Independent shards: rank r fills its own buffer with the variates of the GLOBAL samples [r * n_outer, (r+1) * n_outer)
    (sample_base), no collective on the data path; every number is the MAX over ranks of the device-timed mean launch."""
    import torch
    import torch.distributed as dist
    import dlpm_b200
    from dlpm_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    pk, pk_src = peaks()
    hbm_peak = float(pk["hbm_gbs"])
    inner = 3072
    seed = 1234

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fill(buf, mode, alpha, outer, base):
        if mode == "A_per_element":
            _lib.call("dlpm_b200_stable_A", _lib.ptr(buf), outer, inner, 2, alpha, -1.0, seed, 2, base, _lib.stream_ptr())
        else:
            _lib.call("dlpm_b200_sas", _lib.ptr(buf), None, outer, inner, 1 if mode == "sas_isotropic" else 0, alpha, -1.0, 1.0, seed, 2,
                      base, _lib.stream_ptr())

    def timed(mode, alpha, logn, launches):
        outer = (1 << logn) // inner
        n = outer * inner
        # two buffers when they fit comfortably (a 2^30 fill is 4 GB: far beyond L2 by itself)
        bufs = [torch.empty(n, device=dev) for _ in range(2 if logn <= 28 else 1)]
        flush = torch.empty(64 << 20, device=dev) if logn < 26 else None  # small fills: flush L2 between launches instead
        base = rank * outer
        for b in bufs:
            fill(b, mode, alpha, outer, base)  # warm-up
        barrier()
        tot = 0.0
        if flush is not None:
            for k in range(launches):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fill(bufs[k % len(bufs)], mode, alpha, outer, base); b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for k in range(launches):
                fill(bufs[k % len(bufs)], mode, alpha, outer, base)
            b.record()
            torch.cuda.synchronize()
            tot = a.elapsed_time(b)
        ms = torch.tensor([tot / launches], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ok = bool(torch.isfinite(bufs[0][: 1 << 16]).all())
        del bufs, flush
        return float(ms[0]), n, ok

    for _ in range(max(args.warmup, 3)):
        timed("sas_isotropic", ALPHA, 28, 2)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(1.0)
    clocks.mark()
    ms_head, n_head, _ = timed("sas_isotropic", ALPHA, 30, max(args.steps, 5))
    clk = clocks.stop() if rank == 0 else None
    sweep = []
    for alpha in (1.5, 1.7, 1.9, 2.0):
        for logn in (20, 24, 28, 30):
            row = {"alpha": alpha, "draws_per_gpu": ((1 << logn) // inner) * inner}
            for mode in ("A_per_element", "sas_per_element", "sas_isotropic"):
                ms, n, ok = timed(mode, alpha, logn, 3 if logn == 30 else 5)
                row[mode] = {"ms": ms, "GB/s": world * 4 * n / ms / 1e6, "frac_per_gpu": 4 * n / ms / 1e6 / hbm_peak, "finite": ok}
            sweep.append(row)
    # end to end through the public API with a host destination (the reference's gen_sas returns a device tensor; its caller
    # moves samples to the host): fill + pinned D2H, 2^26 draws
    n_e = ((1 << 26) // inner) * inner
    host = torch.empty(n_e, dtype=torch.float32).pin_memory()
    dlpm_b200.manual_seed(seed)

    def e2e_once():
        x = dlpm_b200.gen_sas(ALPHA, (n_e // inner, inner), device=dev, isotropic=True)
        host.copy_(x.view(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_once()
    barrier()
    w0 = time.perf_counter()
    for _ in range(3):
        e2e_once()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - w0) / 3 * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        value = world * 4 * n_head / ms_head / 1e6
        cpu = None
        try:
            from oracle import ref_import
            if ref_import.available():
                import numpy as np
                ns = ref_import.load()
                np.random.seed(0)
                t0 = time.perf_counter()
                n_cpu = 1 << 20
                ns.Distributions.gen_sas(ALPHA, (n_cpu // inner, inner), device="cpu", isotropic=False)
                dt = time.perf_counter() - t0
                cpu = {"value": 4 * n_cpu / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": "reference", "source": reference_source(),
                       "sample": "bem/datasets/Distributions.py gen_sas(alpha=1.7, isotropic=False) of the unmodified reference, 2^20 draws "
                                 "(scipy levy_stable.rvs, single-threaded) in %.2f s" % dt}
        except Exception as e:
            cpu = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        emit({"metric": "alpha-stable noise GB/s (isotropic SaS fill, alpha=1.7, 2^30 draws per GPU)", "value": value, "unit": "GB/s",
              "n_gpus": world, "steps": max(args.steps, 5), "warmup": max(args.warmup, 3), "ms_per_step": ms_head, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": "BASELINE.json configs[4]: alpha-stable noise sweep alpha in {1.5,1.7,1.9,2.0}, 2^20..2^30 draws per GPU, "
                                     "modes A per element / SaS per element / SaS isotropic (inner 3072); independent shards per GPU",
                         "generator": "Philox4x32-%d" % _lib.load().dlpm_b200_philox_rounds(),
                         "l2_note": "fills >= 2^26 draws exceed L2 by themselves; smaller fills flush L2 between launches"},
              "roofline": {"bound": "hbm", "kernel": "k_fill6 (isotropic SaS fill, sextet scheme: six normals per Philox block)", "achieved": 4 * n_head / ms_head / 1e6, "peak": hbm_peak,
                           "unit": "GB/s", "frac": 4 * n_head / ms_head / 1e6 / hbm_peak, "traffic": None,
                           "peak_source": "MEASURED_PEAKS.json hbm_gbs (%s)" % pk_src},
              "e2e": {"value": world * 4 * n_e / float(e2e_ms[0]) / 1e6, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n_e * world,
                      "api": "dlpm_b200.gen_sas(...) + pinned D2H of the tensor (PCIe-bound)"},
              "gpu_launches": int(world * (max(args.steps, 5) + 2)), "clocks": clk, "cpu_baseline": cpu, "sweep": sweep})
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sampling", choices=["sampling", "noise"])  # noise = BASELINE.json configs[4]
    ap.add_argument("--batch-per-gpu", type=int, default=512)
    ap.add_argument("--reverse-steps", type=int, default=T_STEPS)
    ap.add_argument("--e2e-steps", type=int, default=2)  # two passes: one pass alone carries the +-3 % power-cap jitter
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--cpu-substeps", type=int, default=3)
    ap.add_argument("--gpu-eager", type=int, default=1)  # 0 skips the reference-on-GPU (PyTorch eager) baseline
    args = ap.parse_args()
    # stdout carries exactly one JSON line.  NCCL prints its version banner to the process's fd 1 at communicator creation
    # (NCCL_DEBUG_FILE does not redirect it), so fd 1 is pointed at stderr for the whole run and the JSON line is written
    # to a duplicate of the original stdout.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # not launched through torchrun: re-exec under it (one process per GPU)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_JSON_OUT))
    if args.workload == "noise":
        run_noise(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
