"""Turn ncu CSV logs into the small markdown summaries committed under profiles/.

    python tools/summarize_ncu.py launches  gpurun_out/launches.csv            # gpu__time_duration per launch -> share per kernel
    python tools/summarize_ncu.py metrics   gpurun_out/r01_unet_kernels_metrics.csv
"""
import collections
import csv
import re
import sys


def read_rows(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    r = list(csv.reader(lines))
    return r[0], r[1:]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("dlpm::", "")


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def launches(path):
    hdr, rows = read_rows(path)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in rows:
        a = agg.setdefault(short(row[ki]), [0, 0.0])
        a[0] += 1
        a[1] += to_us(row[vi], row[ui])
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k[:90], n, t, 100 * t / tot))
    print("| **total** | %d | %.1f | 100%% |" % (sum(a[0] for a in agg.values()), tot))


def steps(path):
    """Launch list restricted to the COMPLETE reverse steps found between consecutive k_advance launches (per-step averages)."""
    hdr, rows = read_rows(path)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(short(r[ki]), to_us(r[vi], r[ui])) for r in rows]
    cuts = [i for i, (k, _) in enumerate(seq) if k.startswith("k_advance")]
    segs = [seq[a + 1:b + 1] for a, b in zip(cuts[:-1], cuts[1:])]
    # a pure step is the SHORTEST segment (the first step of a pass also carries the pass set-up: Sigma scan, x_T fill, ...)
    n_max = min(len(sg) for sg in segs if any(k.startswith("k_conv_tc") for k, _ in sg))
    segs = [sg for sg in segs if len(sg) == n_max]
    agg = collections.OrderedDict()
    for sg in segs:
        for k, us in sg:
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += us
    n = len(segs)
    tot = sum(a[1] for a in agg.values()) / n
    print("%d complete steps of %d launches\n" % (n, n_max))
    print("| kernel | launches / step | us / step | share of the step |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %.1f | %.1f | %.1f %% |" % (k[:90], c / n, t / n, 100 * t / n / tot))
    print("| **total** | %d | %.1f | 100 %% |" % (n_max, tot))
    conv = sum(t for k, (c, t) in agg.items() if k.startswith("k_conv_tc")) / n
    print("\nconvolutions: %.1f %% of the step" % (100 * conv / tot))


def metrics(path):
    hdr, rows = read_rows(path)
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for row in rows:
        key = (row[idx["ID"]], short(row[idx["Kernel Name"]]), row[idx["Grid Size"]])
        per.setdefault(key, {})[row[idx["Metric Name"]]] = (row[idx["Metric Value"]], row[idx["Metric Unit"]])
    agg = collections.OrderedDict()
    print("| # | kernel | grid | us | DRAM rd MB | DRAM wr MB | L2 MB | TMA MB | tensor pipe % |\n|---|---|---|---:|---:|---:|---:|---:|---:|")
    for (i, name, grid), m in per.items():
        us = to_us(*m["gpu__time_duration.sum"])
        rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
        l2 = to_bytes(*m["lts__t_bytes.sum"])
        tma = to_bytes(*m.get("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", ("0", "byte")))
        tp = float(m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0].replace(",", ""))
        print("| %s | `%s` | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |" % (i, name[:40], grid, us, rd / 1e6, wr / 1e6, l2 / 1e6, tma / 1e6, tp))
        a = agg.setdefault(name.split("<")[0], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += rd + wr; a[3] += tp * us
    print("\n| kernel family | launches | total us | DRAM bytes/launch (avg) | time-weighted tensor pipe % |\n|---|---:|---:|---:|---:|")
    for k, (n, us, b, tp) in agg.items():
        print("| `%s` | %d | %.1f | %.1f MB | %.1f |" % (k, n, us, b / n / 1e6, tp / us))


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics, "steps": steps}[sys.argv[1]](sys.argv[2])
