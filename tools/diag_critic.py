"""Per-tensor gradient difference of the critic step: stock torch modules vs the fused BatchNorm + LeakyReLU kernels,
with and without the convolution bias folded into the BatchNorm (fp32 cuDNN, TF32 off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200.models.afgsa.discriminator import DiscriminatorVGG  # noqa: E402
from pixel_heal_thyself_b200.models.losses import GANLoss, GradientPenaltyLoss  # noqa: E402

DEV = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(5)
D = DiscriminatorVGG(3, 64, 64).to(DEV)
real, fake = torch.rand(4, 3, 64, 64, device=DEV), torch.rand(4, 3, 64, 64, device=DEV)
gan, gp = GANLoss("wgan").to(DEV), GradientPenaltyLoss(torch.device(DEV))
sd = {k: v.clone() for k, v in D.state_dict().items()}
res = {}
for name, fused, fold in (("stock", False, False), ("stock2", False, False), ("fused", True, False), ("fused+fold", True, True)):
    D.load_state_dict(sd)
    D.fused_bn_act, D.fold_conv_bias = fused, fold
    D.zero_grad(set_to_none=True)
    torch.manual_seed(77)
    loss = (gan(D(fake), False) + gan(D(real), True)) / 2 + 10.0 * gp(D, real, fake)
    loss.backward()
    res[name] = (float(loss), {n: p.grad.clone() for n, p in D.named_parameters()})
# forward only: intermediate activations of D(fake) in the three modes
acts = {}
for name, fused, fold in (("stock", False, False), ("fused", True, False), ("fused+fold", True, True)):
    D.load_state_dict(sd)
    D.fused_bn_act, D.fold_conv_bias = fused, fold
    outs = []
    hooks = [blk.register_forward_hook(lambda m, i, o, outs=outs: outs.append(o.detach().float().clone())) for blk in D.features]
    with torch.no_grad():
        x = fake
        if fused:
            y = D(fake)
        else:
            y = D(fake)
    for h in hooks:
        h.remove()
    acts[name] = (outs, y.detach().clone(), {n: b.clone() for n, b in D.named_buffers()})
print("hooked blocks per mode:", {k: len(v[0]) for k, v in acts.items()})
for k in ("fused", "fused+fold"):
    print(k, "logit diff", float((acts[k][1] - acts["stock"][1]).abs().max()), "of", float(acts["stock"][1].abs().max()))
    for n, b in acts["stock"][2].items():
        if "running" in n:
            d = float((acts[k][2][n].float() - b.float()).abs().max()) / (float(b.float().abs().max()) + 1e-30)
            print(f"   {n:32s} {d:9.2e}")
gmax = max(float(g.abs().max()) for g in res["stock"][1].values())
print("loss", {k: v[0] for k, v in res.items()})
for n, g in res["stock"][1].items():
    row = []
    for k in ("stock2", "fused", "fused+fold"):
        d = res[k][1][n] - g
        row.append((float(d.abs().max()) / max(float(g.abs().max()), 1e-4 * gmax),
                    float(d.norm()) / max(float(g.norm()), 1e-4 * gmax * g.numel() ** 0.5)))
    print(f"{n:28s} |g|max {float(g.abs().max()):9.3e}  " + "  ".join(f"{k} max {e:8.2e} l2 {l:8.2e}" for k, (e, l) in zip(("stock2", "fused", "fused+fold"), row)))
