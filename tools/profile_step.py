"""One training step of the hot path inside a cudaProfilerStart/Stop range, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches.csv python tools/profile_step.py [--workload prod] [--dtype bf16]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200.config import load_config  # noqa: E402
from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="prod")
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
cfg = load_config(args.workload, [f"trainer.batch_size={args.batch}", "data.synthetic.num_images=1",
                                  f"model.afgsa.compute_dtype={args.dtype}"])
tr = AFGSATrainer(cfg)
tr.setup(g_only=True)
ds = tr.setup_data()
batch = ds.batch_device(torch.arange(args.batch, device=tr.device))
for _ in range(args.warmup):
    tr.train_step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
g, _ = tr.train_step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(g))
