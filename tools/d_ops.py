"""Critic step: CUDA time by aten operator (torch.profiler, CPU+CUDA activities) -- which autograd ops the small
elementwise kernels of the step come from."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200.models.afgsa.discriminator import DiscriminatorVGG  # noqa: E402
from pixel_heal_thyself_b200.models.losses import GANLoss, GradientPenaltyLoss  # noqa: E402

torch.backends.cudnn.deterministic = True
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.self_device_time_total) for e in prof.key_averages() if e.key.startswith("aten::") or e.key[0] == "_"]
for k, n, t in sorted(rows, key=lambda r: -r[2])[:40]:
    print(f"{t / 1e3:8.3f} ms  n={n:4d}  {k}")
print("---- by input shape, elementwise-ish ops only")
rows = [(e.key, str(e.input_shapes)[:110], e.count, e.self_device_time_total) for e in prof.key_averages(group_by_input_shape=True)
        if e.key in ("aten::copy_", "aten::add_", "aten::add", "aten::mul", "aten::fill_", "aten::sum", "aten::div", "aten::sub", "aten::neg",
                     "aten::constant_pad_nd", "aten::leaky_relu", "aten::leaky_relu_backward", "aten::zero_", "aten::mul_", "aten::clone")]
for k, sh, n, t in sorted(rows, key=lambda r: -r[3])[:40]:
    print(f"{t / 1e3:8.3f} ms  n={n:4d}  {k:28s} {sh}")
