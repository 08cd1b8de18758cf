"""Critic step (3 x D forward, gradient penalty, backward) on one B200: CUDA time by kernel name (torch.profiler)."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200.models.afgsa.discriminator import DiscriminatorVGG  # noqa: E402
from pixel_heal_thyself_b200.models.losses import GANLoss, GradientPenaltyLoss  # noqa: E402

torch.backends.cudnn.deterministic = True
torch.backends.cudnn.benchmark = False
dev = torch.device("cuda")
D = DiscriminatorVGG(3, 64, 128).to(dev)
gan, gp = GANLoss("wgan").to(dev), GradientPenaltyLoss(dev)
real, fake = torch.rand(8, 3, 128, 128, device=dev), torch.rand(8, 3, 128, 128, device=dev)


def step():
    D.zero_grad()
    loss = (gan(D(fake), False) + gan(D(real), True)) / 2 + 10.0 * gp(D, real, fake)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.device_time > 0:
        tot[e.name[:90]][0] += 1
        tot[e.name[:90]][1] += e.device_time
total = sum(v[1] for v in tot.values())
print(f"critic step: {total / 1e3:.2f} ms of kernels, {sum(v[0] for v in tot.values())} launches")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t / 1e3:8.3f} ms {100 * t / total:5.1f}% n={n:4d}  {k}")
