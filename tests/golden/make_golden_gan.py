"""Golden fixture of ONE FULL GAN ITERATION of the REAL reference (critic step + generator step).

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden_gan.py

Restates the loop body of ``BaseTrainer.train`` (pht/models/base_trainer.py:388-457) around the reference's OWN modules:
``AFGSANet`` (model.py:585-733), ``DiscriminatorVGG`` (model.py:264-344), ``GANLoss("wgan")`` (losses.py:103-172),
``GradientPenaltyLoss`` (losses.py:12-57), ``L1ReconstructionLoss`` (losses.py:175-184), Adam(lr 1e-4, betas 0.9/0.999)
for both nets (base_trainer.py:177-204), loss weights l1 1.0 / gan 0.005 / gp 10 (config/base.py:66-68), batch 4 of 32 x 32
patches, padding_mode "replicate".  G is initialised under ``torch.manual_seed(SEED)``, D under ``SEED + 1``; the gradient
penalty's interpolation coefficients (``torch.rand`` from the global generator in the reference) are drawn under
``SEED + 2`` and STORED, because a CUDA generator yields different numbers: the GPU test injects them.

Committed: tests/golden/gan_step.json -- d_loss and its three terms, g_loss and its two terms, per-parameter statistics
(L2, abs-max, 8 probes) of the critic's gradients after ``d_loss.backward()`` and of the generator's after
``g_loss.backward()``, the critic's BatchNorm running statistics after the iteration, parameter checksums after both Adam
steps, and a checksum of the inputs (regenerated from the seed at test time).
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

from make_golden import SEED, grad_probe_indices, import_reference, synth_inputs  # noqa: E402

B, P = 4, 32
L1_W, GAN_W, GP_W, LR = 1.0, 0.005, 10.0, 1e-4


def gan_inputs():
    """(x, gt, aux) NCHW fp32, preprocessed as base_trainer.py:373-383 does -- shared by this script and the test."""
    from oracle import afgsa_oracle as O
    noisy_hwc, gt_hwc, aux_hwc = synth_inputs(B, P, SEED + 700)
    return O.preprocess_batch(noisy_hwc, gt_hwc, aux_hwc)


def gp_alphas():
    g = torch.Generator().manual_seed(SEED + 2)
    return torch.rand((B, 1, 1, 1), dtype=torch.float32, generator=g)


def stats(named):
    out = {}
    for k, t in named:
        g = t.detach().flatten()
        idx = grad_probe_indices(g.numel())
        out[k] = {"l2": float(g.double().norm()), "absmax": float(g.abs().max()), "probe_idx": idx,
                  "probe": [float(g[i]) for i in idx]}
    return out


def main():
    ref_model, ref_losses, _ = import_reference()
    torch.set_num_threads(os.cpu_count() or 8)
    x, gt, aux = gan_inputs()
    torch.manual_seed(SEED)
    G = ref_model.AFGSANet(3, 7, 256, num_gcp=0, padding_mode="replicate")
    torch.manual_seed(SEED + 1)
    D = ref_model.DiscriminatorVGG(3, 64, P)
    dev = torch.device("cpu")
    l1, gan, gp = ref_losses.L1ReconstructionLoss(), ref_losses.GANLoss("wgan"), ref_losses.GradientPenaltyLoss(dev)
    opt_g = torch.optim.Adam(G.parameters(), lr=LR, betas=(0.9, 0.999))
    opt_d = torch.optim.Adam(D.parameters(), lr=LR, betas=(0.9, 0.999))
    alphas = gp_alphas()
    real_rand = torch.rand
    torch.rand = lambda *a, **k: alphas.clone()          # the ONE torch.rand call of the iteration (losses.py:35-39)
    try:
        meta = {"seed": SEED, "torch": torch.__version__, "B": B, "P": P, "weights": {"l1": L1_W, "gan": GAN_W, "gp": GP_W},
                "lr": LR, "inputs_checksum": float(sum(t.double().abs().sum() for t in (x, gt, aux))),
                "d_init_checksum": float(sum(p.detach().double().abs().sum() for p in D.parameters()))}
        # ---- base_trainer.py:388-412
        output = G(x, aux)
        opt_d.zero_grad()
        pred_fake, pred_real = D(output.detach()), D(gt)
        loss_real, loss_fake = gan(pred_real, True), gan(pred_fake, False)
        loss_gp = gp(D, gt, output.detach())
        d_loss = (loss_fake + loss_real) / 2 + GP_W * loss_gp
        d_loss.backward()
        meta.update(d_loss=float(d_loss), loss_d_real=float(loss_real), loss_d_fake=float(loss_fake), loss_gp=float(loss_gp))
        meta["d_grads"] = stats((k, p.grad) for k, p in D.named_parameters())
        opt_d.step()
        # ---- base_trainer.py:414-457
        opt_g.zero_grad()
        pred_g_fake = D(output)
        loss_g_fake, loss_l1 = gan(pred_g_fake, True), l1(output, gt)
        g_loss = GAN_W * loss_g_fake + L1_W * loss_l1
        g_loss.backward()
        meta.update(g_loss=float(g_loss), loss_g_fake=float(loss_g_fake), loss_l1=float(loss_l1))
        meta["g_grads"] = stats((k, p.grad) for k, p in G.named_parameters())
        opt_g.step()
    finally:
        torch.rand = real_rand
    meta["d_buffers"] = {k: {"l2": float(b.double().norm()), "first": float(b.flatten()[0])} for k, b in D.named_buffers()}
    meta["d_params_after"] = float(sum(p.detach().double().abs().sum() for p in D.parameters()))
    meta["g_params_after"] = float(sum(p.detach().double().abs().sum() for p in G.parameters()))
    meta["out_l2"] = float(output.detach().double().norm())
    with open(os.path.join(HERE, "gan_step.json"), "w") as f:
        json.dump(meta, f)
    print({k: v for k, v in meta.items() if isinstance(v, float)})


if __name__ == "__main__":
    main()
