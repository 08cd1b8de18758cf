"""Golden fixtures of the REAL reference generator at the BASELINE.json patch shapes.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden_shapes.py

dev (8 x 32 x 32), stag (8 x 64 x 64) and prod (2 x 128 x 128) batches through the reference's own
``AFGSANet`` + ``L1ReconstructionLoss`` (pht/models/afgsa/model.py:585-733, pht/models/losses.py:175-184) with the
reference's random init under ``torch.manual_seed(990819)``: output, loss and per-parameter gradient statistics
(L2 norm, abs-sum, abs-max, 8 probes).  The inputs are regenerated from the seed at test time (``shape_inputs``; a
checksum in the fixture detects generator drift), so only the outputs are committed.  The CPU oracle is checked against
the reference at the same shapes and the deviations are recorded under ``pins`` -- the GPU tests use the oracle for the
full per-tensor gradient comparison and this fixture for the direct pin to the reference.
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

from make_golden import SEED, grad_probe_indices, import_reference, synth_inputs  # noqa: E402

SHAPES = {"dev": (8, 32), "stag": (8, 64), "prod": (2, 128)}


def shape_inputs(name: str):
    """(x, gt, aux) NCHW fp32, preprocessed as base_trainer.py:373-383 does -- shared by this script and the tests."""
    from oracle import afgsa_oracle as O
    b, p = SHAPES[name]
    noisy_hwc, gt_hwc, aux_hwc = synth_inputs(b, p, SEED + 100 + p)
    return O.preprocess_batch(noisy_hwc, gt_hwc, aux_hwc)


def checksum(*tensors) -> float:
    return float(sum(t.double().abs().sum() for t in tensors))


def main():
    from oracle import afgsa_oracle as O
    ref_model, ref_losses, _ = import_reference()
    torch.set_num_threads(os.cpu_count() or 8)
    l1 = ref_losses.L1ReconstructionLoss()
    meta = {"seed": SEED, "torch": torch.__version__, "shapes": SHAPES, "pins": {}, "inputs_checksum": {}, "grads": {}}
    outs = {}
    for name in SHAPES:
        x, gt, aux = shape_inputs(name)
        meta["inputs_checksum"][name] = checksum(x, gt, aux)
        torch.manual_seed(SEED)
        G = ref_model.AFGSANet(3, 7, 256, num_gcp=0, padding_mode="replicate")
        sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
        out = G(x.clone(), aux.clone())
        loss = l1(out, gt)
        loss.backward()
        o_out, o_loss, o_grads = O.g_only_train_step(x, aux, gt, sd, "replicate")
        meta["pins"][f"{name}_oracle_out_rel"] = float((o_out - out.detach()).abs().max() / out.detach().abs().max())
        meta["pins"][f"{name}_oracle_loss_rel"] = float(abs(o_loss - loss.detach()) / loss.detach())
        meta["pins"][f"{name}_oracle_grad_rel_l2_worst"] = max(
            float((o_grads[k] - p.grad).norm() / (p.grad.norm() + 1e-30)) for k, p in G.named_parameters())
        assert meta["pins"][f"{name}_oracle_out_rel"] < 1e-5 and meta["pins"][f"{name}_oracle_grad_rel_l2_worst"] < 5e-3
        grads = {}
        for k, p in G.named_parameters():
            g = p.grad.flatten()
            idx = grad_probe_indices(g.numel())
            grads[k] = {"l2": float(g.double().norm()), "abssum": float(g.double().abs().sum()),
                        "absmax": float(g.abs().max()), "probe_idx": idx, "probe": [float(g[i]) for i in idx]}
        meta["grads"][name] = grads
        meta[f"{name}_loss"] = float(loss.item())
        outs[f"{name}_out"] = out.detach().numpy()
        print(name, {k: v for k, v in meta["pins"].items() if k.startswith(name)}, "loss", meta[f"{name}_loss"], flush=True)
    np.savez_compressed(os.path.join(HERE, "net_shapes.npz"), **outs)
    with open(os.path.join(HERE, "net_shapes_meta.json"), "w") as f:
        json.dump(meta, f)


if __name__ == "__main__":
    main()
