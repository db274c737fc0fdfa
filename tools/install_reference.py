#!/usr/bin/env python
"""Install the UNMODIFIED reference (bwittmann/transoar) into the git-ignored ``baseline/_ref/`` so that it travels to the GPU box.

    python tools/install_reference.py            # build container only (needs /root/reference)

What it does, in order, and why:

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`` -- the contract's
   recipe.  It succeeds but installs *metadata only*: the reference's ``setup.py`` uses ``find_packages()`` and ``transoar/`` has no
   ``__init__.py`` (it is used as a namespace package from the repository root, README "pip install -e ."), so the wheel is empty.
2. Therefore the package tree (``transoar/**/*.py``) and ``config/*.yaml`` are copied verbatim next to that metadata.  Nothing is
   edited; ``baseline/_ref/MANIFEST.json`` records the sha256 of every file so tests can show the tree is the reference's.
3. ``timm`` (pinned 0.4.12 by the reference, absent from this image and its wheelhouse) is used for two names only
   (``DropPath``, ``trunc_normal_``: encoder_blocks.py:10, focused_decoder.py:9).  A shim package with exactly those two names is
   written to ``baseline/_ref/timm/`` -- our code, not the reference's; the reference's files are not touched.

``baseline/_ref`` is in .gitignore (never committed) and not in .gpurunignore (ships with the gpurun snapshot).  Only
``oracle/reference_model.py`` (test infrastructure) puts it on ``sys.path``."""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("TRANSOAR_REFERENCE", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")

TIMM_LAYERS = '''"""Shim for the two timm 0.4.12 names the reference imports (written by tools/install_reference.py; not reference code)."""
import torch
from torch import nn

trunc_normal_ = nn.init.trunc_normal_


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath semantics: scale by 1/keep, identity in eval mode)."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if not self.drop_prob or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x / keep * mask
'''


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    src = os.path.join(REFERENCE, "transoar")
    if not os.path.isdir(src):
        sys.exit(f"{REFERENCE} is not mounted: the reference can only be installed in the build container")
    if os.path.isdir(TARGET):
        shutil.rmtree(TARGET)
    os.makedirs(TARGET)
    # 1. the contract's pip recipe, from a writable copy (the mount is read-only)
    pip_outcome = "not run"
    with tempfile.TemporaryDirectory() as tmp:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, copy, ignore=shutil.ignore_patterns(".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, copy]
        r = subprocess.run(cmd, capture_output=True, text=True)
        pip_outcome = "ok (metadata only: find_packages() finds no package, transoar/ has no __init__.py)" if r.returncode == 0 \
            else f"failed rc={r.returncode}: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}"
    # 2. the package tree and the yaml configs, verbatim
    manifest = {}
    for sub, pattern in (("transoar", ".py"), ("config", ".yaml")):
        for dirpath, _, files in os.walk(os.path.join(REFERENCE, sub)):
            for f in files:
                if not f.endswith(pattern):
                    continue
                s = os.path.join(dirpath, f)
                rel = os.path.relpath(s, REFERENCE)
                d = os.path.join(TARGET, rel)
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                manifest[rel] = sha256(d)
    # 3. timm shim
    layers = os.path.join(TARGET, "timm", "models")
    os.makedirs(layers)
    for p in (os.path.join(TARGET, "timm", "__init__.py"), os.path.join(layers, "__init__.py")):
        open(p, "w").close()
    with open(os.path.join(layers, "layers.py"), "w") as f:
        f.write(TIMM_LAYERS)
    with open(os.path.join(TARGET, "MANIFEST.json"), "w") as f:
        json.dump({"reference": REFERENCE, "pip": pip_outcome, "files": manifest, "timm": "shim (DropPath, trunc_normal_)"}, f, indent=1)
    print(f"installed {len(manifest)} reference files into {TARGET}; pip: {pip_outcome}")


if __name__ == "__main__":
    main()
