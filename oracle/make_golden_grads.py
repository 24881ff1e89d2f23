"""Generate tests/golden/grad_*.npz: parameter GRADIENTS of the UNMODIFIED reference model.  Build-container only.

    python oracle/make_golden_grads.py      # needs /root/reference (read-only) - absent on the GPU box

Same recipe as make_golden.py (reference ``Net`` + oracle/shims stand-ins + deterministic weights, strict load), but the
module is in train() mode and the call is differentiated the way PLModule._step does it
(src/hl_modules/distance_based_hl_module.py:303-330): ``out = model(inputs)['output']``, a scalar loss, ``backward()``.
The loss is L = sum(out * R) with a seeded R, so that every output sample matters and the fixture is loss-independent.
It also asserts that autograd through the oracle restatement reproduces the reference's gradients before writing.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "shims"), REF]

from oracle.tfgridnet_oracle import OracleConfig, net_forward                               # noqa: E402
from oracle.weights import make_state_dict, state_dict_digest, synthetic_mixture, radius_one_hot  # noqa: E402
from oracle.cases import GRAD_CASES                                                          # noqa: E402


def loss_weights(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def main():
    from src.models.tfgridnet_realtime_clean_dis_embd3.net import Net as NetDE3
    from src.models.tfgridnet_realtime_clean_optim.net import Net as NetOPT

    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, case in GRAD_CASES.items():
        variant, kw = case["variant"], case["kwargs"]
        cfg = OracleConfig.from_kwargs(variant, **kw)
        sd = make_state_dict(cfg, case.get("seed", 0))
        ref = (NetDE3 if variant == "dis_embed" else NetOPT)(**kw).train()
        ref.load_state_dict(sd, strict=True)
        mix = synthetic_mixture(case["batch"], kw["num_ch"], case["n_samples"], seed=1234)
        dis = radius_one_hot(mix.shape[0])
        out = ref({"mixture": mix, "dis_embed": dis})["output"]
        R = loss_weights(out.shape, case["loss_seed"])
        (out * R).sum().backward()
        grads = {k: p.grad for k, p in ref.named_parameters()}
        assert all(g is not None for g in grads.values())

        leaf = {k: (v.clone().requires_grad_(True) if k in grads else v) for k, v in sd.items()}
        o = net_forward(leaf, cfg, {"mixture": mix, "dis_embed": dis})["output"]
        (o * R).sum().backward()
        worst = max(float((leaf[k].grad - g).abs().max() / g.abs().max()) for k, g in grads.items())
        assert worst <= 2e-5, (name, worst)

        rec = {"mixture": mix.numpy(), "dis_embed": dis.numpy(), "output": out.detach().numpy(),
               "meta": np.array(json.dumps({"variant": variant, "kwargs": kw, "seed": case.get("seed", 0),
                                            "loss_seed": case["loss_seed"], "weights_digest": state_dict_digest(sd),
                                            "oracle_vs_reference_relerr": worst}))}
        for k, g in grads.items():
            rec["grad::" + k] = g.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(f"{name:16s} out {tuple(out.shape)} params {len(grads)} oracle-vs-ref {worst:.2e}")


if __name__ == "__main__":
    main()
