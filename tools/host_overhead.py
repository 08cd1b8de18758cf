"""Host-side enqueue time of one training step vs its GPU time (is the step launch-bound?)."""
import os
import sys
import time

import torch

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
for i in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    tr.train_step(*batch)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: host enqueue {1e3 * (t1 - t0):.2f} ms, until GPU done {1e3 * (t2 - t0):.2f} ms, GPU events {e0.elapsed_time(e1):.2f} ms")
# back-to-back without per-step sync
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(10):
    tr.train_step(*batch)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"10 steps back to back: host enqueue {1e2 * (t1 - t0):.2f} ms/step, GPU {e0.elapsed_time(e1) / 10:.2f} ms/step")
if "--cprofile" in sys.argv:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(5):
        tr.train_step(*batch)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
