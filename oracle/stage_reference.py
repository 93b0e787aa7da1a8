"""Stage the UNMODIFIED reference package into oracle/_ref/ so the reference's own MeshGraphNet can run on the GPU box's
host cores (bench.py --impl reference, cpu_baseline.kind = "reference").

    python oracle/stage_reference.py            # needs /root/reference (build container only)

TEST / MEASUREMENT INFRASTRUCTURE.  oracle/_ref/ is git-ignored (never part of the repo's history) but travels with the
gpurun snapshot, like the built .so.  What is staged: a verbatim copy of /root/reference/physicsnemo (pure Python, 3 MB,
no build step) plus a MANIFEST with the sha256 of every file copied.  The absent third-party imports of its import chain
(dgl, treelib, s3fs, timm, ...) are served by the stand-ins under oracle/ref_shim/ -- of those only `dgl` carries
arithmetic (the row gather and the destination segment sum, ~10 lines); Linear / LayerNorm / ReLU / autograd are the
reference's own code on real PyTorch.  The product (modulus_b200/) never imports anything from here.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/physicsnemo"
DST = os.path.join(HERE, "_ref")


def stage(force: bool = False) -> str:
    if not os.path.isdir(SRC):
        raise RuntimeError(f"{SRC} not found: the reference can only be staged in the build container")
    man_path = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(man_path) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, os.path.join(DST, "physicsnemo"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(os.path.join(DST, "physicsnemo")):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(man_path, "w"), indent=0, sort_keys=True)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
