"""Install the UNMODIFIED reference into ``oracle/_ref/`` so that it travels to the GPU box.

    python -m oracle.install_ref [--force]

TEST / BENCH INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference`` exists only in the build
container; ``oracle/_ref/`` is git-ignored (nothing of the reference enters the history) but NOT gpurun-ignored, so --
exactly like the built ``libdlpm_b200.so`` -- it ships with the snapshot.  ``__graft_entry__.build()`` calls this.

The reference is a research repository without packaging metadata (no ``setup.py`` / ``pyproject.toml``), so the
contract's ``pip install --target`` has nothing to build: what an install of a pure-Python project does -- place its
importable packages on a path -- is done here directly: every ``*.py`` (and the YAML configs) of the two packages
``dlpm/`` and ``bem/`` is copied verbatim, byte for byte, with a manifest of SHA-256 sums (``MANIFEST.json``) that
``tests/test_oracle_golden.py`` re-checks against ``/root/reference`` whenever both exist.  The seven third-party
modules the hot path never touches stay absent and are stubbed at import time by
``oracle/ref_import.py``.  Uses: (i) ``bench.py --impl reference`` (the reference's own CPU path on the box's host
cores), (ii) ``bench.py``'s ``gpu_eager_baseline`` (the same code with ``device='cuda'``: PyTorch eager + cuDNN, the
same-box bar of SURVEY.md section 2.2), (iii) GPU parity tests against the live reference.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("DLPM_REFERENCE_SOURCE", "/root/reference")
PACKAGES = ("dlpm", "bem")
KEEP_EXT = (".py", ".yml", ".yaml")
SKIP_DIRS = {"__pycache__", ".git"}


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def installed():
    return os.path.isfile(os.path.join(DEST, "MANIFEST.json"))


def install(force=False, quiet=True):
    """Copy the reference's packages into oracle/_ref.  Returns the destination, or None when the source is absent
    (the GPU box: the prebuilt copy, if any, is used as is)."""
    if not os.path.isdir(os.path.join(SOURCE, "dlpm", "methods")):
        return DEST if installed() else None
    if installed() and not force:
        try:
            man = json.load(open(os.path.join(DEST, "MANIFEST.json")))
            if all(os.path.isfile(os.path.join(SOURCE, rel)) and _sha(os.path.join(SOURCE, rel)) == sha and
                   os.path.isfile(os.path.join(DEST, rel)) for rel, sha in man["files"].items()):
                return DEST
        except Exception:
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    files = {}
    for pkg in PACKAGES:
        for base, dirs, names in os.walk(os.path.join(SOURCE, pkg)):
            dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
            for n in sorted(names):
                if not n.endswith(KEEP_EXT):
                    continue
                src = os.path.join(base, n)
                rel = os.path.relpath(src, SOURCE)
                dst = os.path.join(DEST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                files[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SOURCE, "packages": PACKAGES, "files": files,
                   "note": "verbatim copy of the reference's python packages (git-ignored); see oracle/install_ref.py"}, fh,
                  indent=1, sort_keys=True)
    if not quiet:
        print("installed %d reference files into %s" % (len(files), DEST))
    return DEST


if __name__ == "__main__":
    out = install(force="--force" in sys.argv, quiet=False)
    print(out if out else "reference source not present and no prebuilt copy")
