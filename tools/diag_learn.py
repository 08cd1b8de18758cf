"""Prints the G-only loss trajectory of the ci preset (what test_trainer_gan_step_runs_and_learns asserts on)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200 import _lib  # noqa: E402
from pixel_heal_thyself_b200.config import load_config  # noqa: E402
from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer  # noqa: E402

cfg = load_config("ci", ["data.synthetic.num_images=1", "data.synthetic.height=128", "data.synthetic.width=128",
                         "data.patches.num_patches=16"])
for simple in (0, 0):
    _lib.lib.pht_set_force_simple(simple)
    tr = AFGSATrainer(cfg)
    tr.setup(g_only=True)
    ds = tr.setup_data()
    noisy, gt, aux = ds.batch_device(torch.arange(2, device=tr.device))
    losses = [float(tr.train_step(noisy, gt, aux)[0]) for _ in range(8)]
    print("force_simple", simple, ["%.6f" % l for l in losses])
