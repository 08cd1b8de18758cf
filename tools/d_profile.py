"""Which critic ops are slow: aten-level CUDA time by input shape for one critic step (D fwd x3, GP, backward)."""
import os
import sys

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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="device_time_total", row_limit=14, max_name_column_width=40,
                                                         max_shapes_column_width=90))
