"""Build libdlpm_b200.so in-tree with nvcc for sm_100a (B200).

    python -m dlpm_b200.build [--force] [--verbose]

Plain ``nvcc -shared``: no torch headers, no pybind -- the library exposes only the C ABI declared
in ``include/dlpm_b200.h`` / ``include/dlpm_b200_unet.h``.  Objects are cached under ``build/`` by
source mtime; nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
LIB_PATH = os.path.join(PKG, "libdlpm_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
] + os.environ.get("DLPM_B200_NVCC_EXTRA", "").split()  # e.g. -DDLPM_CONV_COALESCED_STORES=1 for A/B builds


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libdlpm_b200.so cannot be built")


def sources():
    out = []
    for base, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith(".cu"):
                out.append(os.path.join(base, f))
    return sorted(out)


def _headers_mtime():
    m = 0.0
    for base in (CSRC, os.path.join(ROOT, "include")):
        for b, _, files in os.walk(base):
            for f in files:
                if f.endswith((".cuh", ".h", ".hpp")):
                    m = max(m, os.path.getmtime(os.path.join(b, f)))
    return m


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hm = _headers_mtime()
    srcs = sources()
    jobs = []
    objs = []
    for src in srcs:
        obj = os.path.join(OBJ_DIR, os.path.relpath(src, CSRC).replace(os.sep, "_")[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((src, cmd))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return src, r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(run, jobs):
                if verbose and log:
                    print("== %s\n%s" % (os.path.relpath(src, ROOT), log))
    if jobs or force or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
