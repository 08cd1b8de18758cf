"""Golden fixture for the FiLM variant of the generator (``use_film=True``), from the REAL reference.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden_film.py
Same recipe as make_golden.py (whose helpers it reuses): seeded inputs and weights, the reference's own
AFGSANet / L1ReconstructionLoss, the CPU oracle checked against them, small fixtures written next to this file.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402


def main():
    from oracle import afgsa_oracle as O
    ref_model, ref_losses, ref_pre = MG.import_reference()
    torch.set_num_threads(8)
    noisy_hwc, gt_hwc, aux_hwc = MG.synth_inputs(2, 16, MG.SEED + 1)
    x, gt, aux = O.preprocess_batch(noisy_hwc, gt_hwc, aux_hwc)
    torch.manual_seed(MG.SEED)
    G = ref_model.AFGSANet(3, 7, 256, num_sa=2, num_gcp=0, padding_mode="replicate", use_film=True)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    out = G(x.clone(), aux.clone())
    loss = ref_losses.L1ReconstructionLoss()(out, gt)
    loss.backward()
    o_out, o_loss, o_grads = O.g_only_train_step(x, aux, gt, sd, "replicate", num_sa=2)
    pins = {"film_out_maxdiff": float((o_out - out.detach()).abs().max()),
            "film_loss_diff": float(abs(o_loss - loss.detach())),
            "film_grad_rel_maxdiff": max(float((o_grads[k] - p.grad).abs().max() / (p.grad.abs().max() + 1e-30))
                                         for k, p in G.named_parameters() if p.grad is not None),
            "film_params_without_grad": [k for k, p in G.named_parameters() if p.grad is None]}
    assert pins["film_out_maxdiff"] < 1e-5 and pins["film_grad_rel_maxdiff"] < 1e-4, pins
    grads = {}
    for k, p in G.named_parameters():
        g = (p.grad if p.grad is not None else torch.zeros_like(p)).flatten()
        idx = MG.grad_probe_indices(g.numel())
        grads[k] = {"sum": float(g.double().sum()), "abssum": float(g.double().abs().sum()),
                    "absmax": float(g.abs().max()), "probe_idx": idx, "probe": [float(g[i]) for i in idx]}
    meta = {"seed": MG.SEED, "torch": torch.__version__, "pins": pins,
            "param_order": [k for k, _ in G.named_parameters()],
            "param_shapes": {k: list(v.shape) for k, v in G.named_parameters()},
            "param_checksums": {k: [float(v.double().sum()), float(v.double().abs().sum()), float(v.flatten()[0])]
                                for k, v in G.named_parameters()}}
    with open(os.path.join(HERE, "net_film_grads.json"), "w") as f:
        json.dump(grads, f)
    with open(os.path.join(HERE, "net_film_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "net_film.npz"), x=x.numpy(), aux=aux.numpy(), gt=gt.numpy(),
                        out=out.detach().numpy(), loss=np.float64(loss.item()))
    print(json.dumps(pins, indent=1))


if __name__ == "__main__":
    main()
