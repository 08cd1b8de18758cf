"""Per-tensor bf16 / fp32 gradient errors against the CPU oracle for small smoke-sized batches (which tensors are noisy)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import afgsa_oracle as O  # noqa: E402
from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet  # noqa: E402
from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss  # noqa: E402
from make_golden_shapes import shape_inputs  # noqa: E402

dev = torch.device("cuda:0")


def run(tag, x, aux, gt, num_sa):
    for dt in ("fp32", "bf16"):
        torch.manual_seed(990819)
        net = AFGSANet(3, 7, 256, num_sa=num_sa, num_gcp=0, padding_mode="replicate", compute_dtype=dt).to(dev)
        sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        out = net(x.to(dev), aux.to(dev))
        L1ReconstructionLoss()(out, gt.to(dev)).backward()
        o_out, o_loss, o_g = O.g_only_train_step(x, aux, gt, sd, "replicate", num_sa=num_sa)
        errs = sorted(((float((p.grad.cpu() - o_g[n]).norm() / (o_g[n].norm() + 1e-30)), n) for n, p in net.named_parameters()), reverse=True)
        allg = torch.cat([p.grad.cpu().flatten() for _, p in net.named_parameters()])
        allo = torch.cat([o_g[n].flatten() for n, _ in net.named_parameters()])
        print(f"[{tag} {dt}] whole-gradient rel L2 {float((allg - allo).norm() / allo.norm()):.3e}; worst tensors:",
              ", ".join(f"{n} {e:.3e}" for e, n in errs[:5]), flush=True)


torch.manual_seed(990819)
run("randn 1x16x16 sa2", torch.randn(1, 3, 16, 16) * 0.5, torch.rand(1, 7, 16, 16), torch.randn(1, 3, 16, 16) * 0.5, 2)
torch.manual_seed(990819)
run("randn 2x32x32 sa2", torch.randn(2, 3, 32, 32) * 0.5, torch.rand(2, 7, 32, 32), torch.randn(2, 3, 32, 32) * 0.5, 2)
x, gt, aux = shape_inputs("dev")
run("dev[:2] sa5", x[:2], aux[:2], gt[:2], 5)
run("dev[:2] sa2", x[:2], aux[:2], gt[:2], 2)
run("dev[:8] sa5", x, aux, gt, 5)
