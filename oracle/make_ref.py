"""Populate oracle/_ref/ with the UNMODIFIED reference modules of the hot path.  TEST / BENCH INFRASTRUCTURE ONLY.

The reference (cudagu/aero-gnn) is plain Python with no build step, so "building the reference" here means
placing byte-identical copies of its path modules where they can travel to the GPU box: /root/reference does not
exist there, oracle/_ref/ (git-ignored, NOT gpurun-ignored) does.  Nothing is edited; MANIFEST.json records the
sha256 of every file so a reader can check the copies against the upstream checkout.  Only
`bench.py --impl reference` (the reference arm) and the cpu_baseline leg execute what lands here -- never the
product path.

    python oracle/make_ref.py            # run where /root/reference exists (this container; __graft_entry__.build())

Third-party packages the modules import and this image lacks (`torch_scatter`, `torch_geometric`) are stood in by
oracle/standins.py at import time (published semantics restated).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("AERO_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["models/mlp.py", "models/mgnLayer.py", "models/mgn.py", "models/bsms_mgn.py", "models/poolmgn.py",
         "models/fouriermgn.py", "config.yaml",
         # the callers, for the "train.py runs unchanged on the drop-in models" demonstration (scripts/run_reference_train.py)
         "train.py", "utils.py", "dataset.py", "inference.py"]


def make_ref(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/ is populated (now or before), False when there is no reference checkout."""
    if not os.path.isdir(REF_SRC):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": REF_SRC, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files copied unmodified from {REF_SRC}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DST, "models", "mgnLayer.py"))


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
