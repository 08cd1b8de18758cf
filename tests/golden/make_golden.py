"""Generate the golden fixtures in tests/golden/ from the REAL reference.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
It imports the reference's own modules (with stubs for the third-party deps that
are not installed and not on the hot path), runs them on seeded inputs, checks
the CPU oracle (oracle/) against them, and writes small fixtures so the tests
on the GPU box never need the reference checkout.
"""
from __future__ import annotations

import json
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SEED = 990819  # reference default seed, config/default.yaml:11


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Missing:
    def __init__(self, *a, **k):
        raise NotImplementedError("third-party dependency not installed")


def import_reference():
    _stub("hilbertcurve").hilbertcurve = _stub("hilbertcurve.hilbertcurve", HilbertCurve=_Missing)
    _stub("kornia").losses = _stub("kornia.losses", MS_SSIMLoss=_Missing)
    _stub("pyexr")
    _stub("matplotlib", use=lambda *a, **k: None).pyplot = _stub("matplotlib.pyplot")
    sys.path.insert(0, REF)
    from pht.models.afgsa import model as ref_model
    from pht.models import losses as ref_losses
    from pht.models.afgsa import preprocessing as ref_pre
    return ref_model, ref_losses, ref_pre


def synth_inputs(b, p, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.exp(torch.randn(b, p, p, 3, generator=g) * 0.5)
    noisy = gt * (torch.rand(b, p, p, 3, generator=g) + torch.rand(b, p, p, 3, generator=g))
    normal = torch.rand(b, p, p, 3, generator=g) * 2 - 1
    normal[torch.rand(b, p, p, 3, generator=g) < 0.01] = float("nan")
    depth = torch.rand(b, p, p, 1, generator=g)
    albedo = torch.rand(b, p, p, 3, generator=g)
    aux = torch.cat([normal, depth, albedo], -1)
    return noisy, gt, aux


def grad_probe_indices(numel, k=8):
    # fixed, seed-free probe positions
    return [(i * 2654435761 + 12345) % numel for i in range(k)]


def main():
    sys.path.insert(0, ROOT)
    from oracle import afgsa_oracle as O
    from oracle import sampler_oracle as S

    ref_model, ref_losses, ref_pre = import_reference()
    torch.set_num_threads(8)
    meta = {"seed": SEED, "torch": torch.__version__, "pins": {}}

    # ---- 1. parameter init pin ------------------------------------------------
    torch.manual_seed(SEED)
    G = ref_model.AFGSANet(3, 7, 256, num_sa=5, block_size=8, halo_size=3, num_heads=4, num_gcp=0,
                           padding_mode="replicate", curve_order=ref_model.CurveOrder.RASTER, use_film=False)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    meta["param_order"] = [k for k, _ in G.named_parameters()]
    meta["param_shapes"] = {k: list(v.shape) for k, v in G.named_parameters()}
    meta["buffers"] = {k: v.tolist() for k, v in G.named_buffers()}
    meta["param_checksums"] = {
        k: [float(v.double().sum()), float(v.double().abs().sum()), float(v.flatten()[0])]
        for k, v in G.named_parameters()
    }
    meta["num_params"] = int(sum(p.numel() for p in G.parameters()))

    # ---- 2. preprocessing (numpy reference) -----------------------------------
    noisy_hwc, gt_hwc, aux_hwc = synth_inputs(2, 16, SEED + 1)
    ref_aux = aux_hwc.clone()
    ref_aux[:, :, :, :3] = torch.FloatTensor(ref_pre.preprocess_normal(ref_aux[:, :, :, :3]))
    ref_aux = ref_aux.permute(0, 3, 1, 2).contiguous()
    ref_noisy = ref_pre.preprocess_specular(noisy_hwc).permute(0, 3, 1, 2).contiguous()
    ref_gt = ref_pre.preprocess_specular(gt_hwc).permute(0, 3, 1, 2).contiguous()
    o_noisy, o_gt, o_aux = O.preprocess_batch(noisy_hwc, gt_hwc, aux_hwc)
    meta["pins"]["preprocess_maxdiff"] = float(max((o_noisy - ref_noisy).abs().max(), (o_gt - ref_gt).abs().max(),
                                                     (o_aux - ref_aux).abs().max()))
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), noisy_hwc=noisy_hwc.numpy(), gt_hwc=gt_hwc.numpy(),
                        aux_hwc=aux_hwc.numpy(), noisy=ref_noisy.numpy(), gt=ref_gt.numpy(), aux=ref_aux.numpy())

    # ---- 3. full generator fwd + L1 + bwd, both padding modes -----------------
    x, gt, aux = ref_noisy, ref_gt, ref_aux
    l1 = ref_losses.L1ReconstructionLoss()
    for mode in ("replicate", "reflect"):
        torch.manual_seed(SEED)
        Gm = ref_model.AFGSANet(3, 7, 256, num_gcp=0, padding_mode=mode)
        Gm.zero_grad()
        out = Gm(x.clone(), aux.clone())
        loss = l1(out, gt)
        loss.backward()
        o_out, o_loss, o_grads = O.g_only_train_step(x, aux, gt, sd, mode)
        d_out = float((o_out - out.detach()).abs().max())
        d_grad = max(float((o_grads[k] - p.grad).abs().max() / (p.grad.abs().max() + 1e-30))
                     for k, p in Gm.named_parameters())
        meta["pins"][f"net_{mode}_out_maxdiff"] = d_out
        meta["pins"][f"net_{mode}_loss_diff"] = float(abs(o_loss - loss.detach()))
        meta["pins"][f"net_{mode}_grad_rel_maxdiff"] = d_grad
        grads = {}
        for k, p in Gm.named_parameters():
            g = p.grad.flatten()
            idx = grad_probe_indices(g.numel())
            grads[k] = {"sum": float(g.double().sum()), "abssum": float(g.double().abs().sum()),
                        "absmax": float(g.abs().max()), "probe_idx": idx, "probe": [float(g[i]) for i in idx]}
        with open(os.path.join(HERE, f"net_{mode}_grads.json"), "w") as f:
            json.dump(grads, f)
        np.savez_compressed(os.path.join(HERE, f"net_{mode}.npz"), x=x.numpy(), aux=aux.numpy(), gt=gt.numpy(),
                            out=out.detach().numpy(), loss=np.float64(loss.item()))

    # ---- 4. AFGSA module alone (small channels; exercises zero-pad + rel-pos) --
    torch.manual_seed(SEED + 2)
    A = ref_model.AFGSA(32, block_size=8, halo_size=3, num_heads=4)
    an = torch.randn(2, 32, 24, 16)
    aa = torch.randn(2, 32, 24, 16)
    a_out = A(an, aa).detach()
    asd = {"p." + k: v.detach().clone() for k, v in A.state_dict().items()}
    o_a = O.afgsa(an, aa, asd, "p.", 8, 3, 4)
    meta["pins"]["afgsa_module_maxdiff"] = float((o_a - a_out).abs().max())
    np.savez_compressed(os.path.join(HERE, "afgsa_module.npz"), noisy=an.numpy(), aux=aa.numpy(), out=a_out.numpy(),
                        **{k.replace(".", "__"): v.numpy() for k, v in asd.items() if v.dtype.is_floating_point})

    # z-order curve == raster: record that the reference agrees
    torch.manual_seed(SEED + 2)
    Az = ref_model.AFGSA(32, block_size=8, halo_size=3, num_heads=4, curve_order=ref_model.CurveOrder.ZORDER)
    meta["pins"]["zorder_vs_raster_maxdiff"] = float((Az(an, aa).detach() - a_out).abs().max())

    # ---- 5. sampler ------------------------------------------------------------
    samp = {}
    for (hw, p, n) in (((512, 512), 32, 100), ((512, 512), 64, 200), ((1024, 1024), 128, 400), ((300, 420), 32, 60)):
        ref_pts = ref_pre.sample_patches_dart_throwing(hw, p, n, random.Random(SEED))
        o_pts = S.dart_throwing(hw, p, n, S.MT19937(SEED))
        assert np.array_equal(ref_pts, o_pts), "oracle sampler differs from reference"
        samp[f"h{hw[0]}_w{hw[1]}_p{p}_n{n}"] = ref_pts.astype(np.int32)
    meta["pins"]["sampler_bit_exact"] = True
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **samp)
    r = random.Random(SEED)
    meta["mt_kat"] = {"getrandbits32": [r.getrandbits(32) for _ in range(3)],
                      "randint_0_479": [r.randint(0, 479) for _ in range(4)], "random": r.random()}

    # ---- 6. two Adam steps (G-only, L1) on the reference --------------------------
    torch.manual_seed(SEED)
    Gt = ref_model.AFGSANet(3, 7, 256, num_gcp=0, padding_mode="replicate")
    opt = torch.optim.Adam(Gt.parameters(), lr=1e-4, betas=(0.9, 0.999), eps=1e-8)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        ls = l1(Gt(x.clone(), aux.clone()), gt)
        ls.backward()
        opt.step()
        losses.append(float(ls.item()))
    meta["adam_l1_losses_3steps"] = losses
    w = dict(Gt.named_parameters())["decoder.2.0.weight"].detach().flatten()
    meta["adam_decoder2_weight_probe"] = [float(w[i]) for i in grad_probe_indices(w.numel())]

    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta["pins"], indent=1))


if __name__ == "__main__":
    main()
