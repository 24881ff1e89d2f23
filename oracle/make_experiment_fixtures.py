"""Writes tests/golden/experiments.json: the `pl_module_args` (+ grad_clip, batch sizes) of every shipped experiment JSON
of the reference (syn_experiments/*.json, real_experiments/*.json).  Build container only (needs /root/reference); the
fixture lets the GPU box construct `train_dist.TrainModule` exactly as src/train_pt.py:85 constructs PLModule.

    python oracle/make_experiment_fixtures.py
"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "syn_experiments", "*.json")) + glob.glob(os.path.join(REF, "real_experiments", "*.json"))):
        cfg = json.load(open(path))
        rel = os.path.relpath(path, REF)
        out[rel] = {k: cfg[k] for k in ("pl_module", "pl_module_args", "grad_clip", "batch_size", "eval_batch_size", "epochs") if k in cfg}
    dst = os.path.join(ROOT, "tests", "golden", "experiments.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print(dst, list(out))


if __name__ == "__main__":
    main()
