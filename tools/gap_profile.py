"""Kernel timeline of one training step from torch.profiler (CUPTI): per-kernel durations and the idle gaps between
consecutive kernels on the stream, grouped by the kernel that follows the gap."""
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
tr.setup(g_only=True)
ds = tr.setup_data()
batch = ds.batch_device(torch.arange(8, device=tr.device))
for _ in range(3):
    tr.train_step(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_step(*batch)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()
       and "memset" not in e.name.lower()]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // 3
other = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and ("memcpy" in e.name.lower()
         or "memset" in e.name.lower())]
print("non-kernel GPU activities:", [(e.name[:30], round((e.time_range.end - e.time_range.start), 1)) for e in other][:20])
print("step boundaries: last kernel end -> next step's first kernel start (us):",
      [round(evs[(k + 1) * n].time_range.start - evs[(k + 1) * n - 1].time_range.end, 1) for k in range(2)],
      "| first/last kernels:", evs[n].name[:40], "/", evs[2 * n - 1].name[:40])
evs = evs[n:2 * n]   # the middle step
busy = sum(e.time_range.end - e.time_range.start for e in evs)
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"{len(evs)} kernels, busy {busy / 1e3:.3f} ms, span {span / 1e3:.3f} ms, idle {(span - busy) / 1e3:.3f} ms")
gaps = defaultdict(list)
for a, b in zip(evs[:-1], evs[1:]):
    gaps[(a.name[:40], b.name[:40])].append(b.time_range.start - a.time_range.end)
rows = sorted(gaps.items(), key=lambda kv: -sum(kv[1]))
print(f"{'prev kernel':42s} {'next kernel':42s} {'n':>4s} {'avg gap us':>10s} {'total us':>9s}")
for (a, b), g in rows[:30]:
    print(f"{a:42s} {b:42s} {len(g):4d} {sum(g) / len(g):10.2f} {sum(g):9.1f}")
