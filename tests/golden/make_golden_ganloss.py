"""Golden fixture for GANLoss, from the REAL reference (pht/models/losses.py:103-172).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_ganloss.py
"""
from __future__ import annotations

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402


def scores(loss_type: str) -> torch.Tensor:
    """Seeded critic scores [8, 1]; probabilities in (0, 1) for the BCE variant."""
    g = torch.Generator().manual_seed(MG.SEED)
    s = torch.randn(8, 1, generator=g) * 1.5
    return torch.sigmoid(s) if loss_type == "nsgan" else s


def main():
    MG.import_reference()
    from pht.models.losses import GANLoss as RefGANLoss
    out = {}
    for t in ("nsgan", "wgan", "lsgan", "hinge"):
        ref = RefGANLoss(t)
        x = scores(t).requires_grad_(True)
        for real in (True, False):
            for disc in ((True, False) if t == "hinge" else (None,)):
                loss = ref(x, real, disc) if t == "hinge" else ref(x, real)
                (grad,) = torch.autograd.grad(loss, x)
                out[f"{t}/{int(real)}/{disc}"] = {"loss": float(loss), "grad": [float(v) for v in grad.reshape(-1)]}
    with open(os.path.join(HERE, "ganloss.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote ganloss.json:", len(out), "cases")


if __name__ == "__main__":
    main()
