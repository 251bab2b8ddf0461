"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm (the UNMODIFIED reference on the host
cores: /root/reference here, its verbatim copy oracle/_ref on the GPU box; the oracle port only when neither exists) prints
ONE JSON line with the required keys, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_arm_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "2", "--cpu-substeps", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("DLPM samples/sec") and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("CIFAR-10-LT") and d["config"]["reverse_steps"] == 1000
    from oracle import ref_import
    want_kind = "reference" if ref_import.available() else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    if ref_import.available():
        assert d["cpu_baseline"]["source"] in ("/root/reference", "oracle/_ref")
        assert "unmodified reference" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_runs_from_the_installed_copy():
    """oracle/_ref (what travels to the GPU box) alone is enough: hide /root/reference through DLPM_REFERENCE_ROOT."""
    sys.path.insert(0, ROOT)
    from oracle import install_ref
    if install_ref.install() is None:
        pytest.skip("no reference source and no installed copy")
    env = dict(os.environ, DLPM_REFERENCE_ROOT=install_ref.DEST, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-batch", "2", "--cpu-substeps", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["source"] == "oracle/_ref"


def test_installed_copy_is_verbatim():
    """Every file under oracle/_ref is byte-identical to its source in /root/reference (when both exist)."""
    import hashlib
    sys.path.insert(0, ROOT)
    from oracle import install_ref
    if not (os.path.isdir(install_ref.SOURCE) and install_ref.install()):
        pytest.skip("reference source not present")
    man = json.load(open(os.path.join(install_ref.DEST, "MANIFEST.json")))
    assert len(man["files"]) > 50
    for rel, sha in man["files"].items():
        for root in (install_ref.SOURCE, install_ref.DEST):
            assert hashlib.sha256(open(os.path.join(root, rel), "rb").read()).hexdigest() == sha, rel


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a machine without a GPU")
def test_product_arm_needs_cuda():
    r = run_bench("--steps", "1", "--warmup", "0", "--e2e-steps", "0", timeout=300)
    assert r.returncode != 0, "the product arm must fail loudly without CUDA, not fall back to the CPU"
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
