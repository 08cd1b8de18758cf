"""Kernel time breakdown of one full GAN training iteration (G on the hand-written path + the PyTorch critic step)."""
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200.config import load_config  # noqa: E402
from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer  # noqa: E402

cfg = load_config("prod", ["trainer.batch_size=8", "data.synthetic.num_images=1", "model.afgsa.compute_dtype=bf16"])
tr = AFGSATrainer(cfg)
tr.setup(g_only=False)
ds = tr.setup_data()
batch = ds.batch_device(torch.arange(8, device=tr.device))
for _ in range(3):
    tr.train_step(*batch)
torch.cuda.synchronize()
print("allow_tf32 cudnn/matmul:", torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
      "cudnn.benchmark:", torch.backends.cudnn.benchmark, "deterministic:", torch.backends.cudnn.deterministic)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(*batch)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = defaultdict(lambda: [0, 0.0])
for e in evs:
    k = e.name[:70]
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
total = sum(v[1] for v in tot.values())
span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
print(f"{len(evs)} GPU activities, busy {total / 1e3:.2f} ms, span {span / 1e3:.2f} ms")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{k:72s} {n:5d} {t / 1e3:9.3f} ms {100 * t / total:5.1f}%")
