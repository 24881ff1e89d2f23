"""Recipe for oracle/_ref: a runnable copy of the UNMODIFIED reference separator.  TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            # build container only (needs /root/reference); __graft_entry__.build() calls it

The reference is pure Python: there is nothing to compile, but /root/reference does not exist on the GPU box, so the
`--impl reference` arm of bench.py and the CPU baseline could only run the oracle PORT there.  This recipe copies the
reference's own model files byte for byte into oracle/_ref/ (git-ignored - reference sources never enter the history -
but not gpurun-ignored, so the directory travels to the GPU box like a built .so) and records their sha256 in
oracle/_ref/MANIFEST.json.  oracle/ref_runner.py imports them with the third-party stand-ins of oracle/shims
(asteroid_filterbanks / espnet2 are not installed in this image).  Nothing in the product package reads oracle/_ref.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = [
    "src/__init__.py",
    "src/models/tfgridnet_realtime_clean_dis_embd3/net.py",
    "src/models/tfgridnet_realtime_clean_dis_embd3/tfgridnet_causal.py",
    "src/models/tfgridnet_realtime_clean_optim/net.py",
    "src/models/tfgridnet_realtime_clean_optim/tfgridnet_causal.py",
]


def build_ref(verbose=False):
    """Returns the path of oracle/_ref, or None when the reference is not mounted (the GPU box: prebuilt copy is used)."""
    if not os.path.isdir(REF):
        return OUT if os.path.exists(os.path.join(OUT, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    if verbose:
        print("oracle/_ref: %d reference files copied" % len(manifest))
    return OUT


if __name__ == "__main__":
    print(build_ref(verbose=True))
    sys.exit(0)
