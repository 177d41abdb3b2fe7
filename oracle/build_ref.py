"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference's Python package (src/) so that the UNMODIFIED
reference sampler can be timed on the GPU box's host cores (`bench.py --impl reference`, kind "reference").

TEST / BENCH INFRASTRUCTURE ONLY.  /root/reference exists in the authoring container, not on the GPU box; built
artefacts under oracle/_ref/ travel with the gpurun snapshot (git-ignored, not gpurun-ignored), exactly like the built
librgfm.so.  Nothing is compiled (the reference is 100 % Python) and no reference source enters the repo's history.
`__graft_entry__.build()` calls build(); the product path never imports anything from here.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REFERENCE = "/root/reference"
DEST = os.path.join(ROOT, "oracle", "_ref")


def build(verbose: bool = True) -> bool:
    """Copy /root/reference/src -> oracle/_ref/src when the reference tree is present.  Returns True when oracle/_ref is usable."""
    src = os.path.join(REFERENCE, "src")
    if os.path.isdir(src):
        dst = os.path.join(DEST, "src")
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        if verbose:
            n = sum(len(f) for _, _, f in os.walk(dst))
            print(f"oracle/_ref: copied {n} files of the unmodified reference package from {src}")
    return os.path.isfile(os.path.join(DEST, "src", "sample_mnist_svhn.py"))


if __name__ == "__main__":
    raise SystemExit(0 if build() else 1)
