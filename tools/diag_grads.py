"""Per-parameter gradient deviation of the fp32 CUDA path against the CPU oracle for several shapes/depths.
Usage (GPU box): python tools/diag_grads.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import afgsa_oracle as O  # noqa: E402
from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet  # noqa: E402
from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss  # noqa: E402


def run(num_sa, B, H, W, mode="replicate", seed=3, dtype="fp32", in64=True):
    torch.manual_seed(990819)
    net = AFGSANet(3, 7, 256, num_sa=num_sa, num_gcp=0, padding_mode=mode, compute_dtype=dtype).cuda()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    torch.manual_seed(seed)
    x, aux, gt = torch.randn(B, 3, H, W) * 0.5, torch.rand(B, 7, H, W), torch.randn(B, 3, H, W) * 0.5
    if in64:
        sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
        o_out, o_loss, o_grads = O.g_only_train_step(x.double(), aux.double(), gt.double(), sd64, mode, num_sa=num_sa)
    else:
        o_out, o_loss, o_grads = O.g_only_train_step(x, aux, gt, sd, mode, num_sa=num_sa)
    out = net(x.cuda(), aux.cuda())
    loss = L1ReconstructionLoss()(out, gt.cuda())
    loss.backward()
    print(f"--- num_sa={num_sa} B={B} {H}x{W} {mode} {dtype}: out rel "
          f"{float((out.detach().cpu().double() - o_out.double()).abs().max() / o_out.abs().max()):.2e} "
          f"loss {float(loss):.7f} vs {float(o_loss):.7f}")
    rows = []
    for n, p in net.named_parameters():
        ref = o_grads[n].double()
        g = p.grad.cpu().double()
        rows.append((float((g - ref).abs().max() / (ref.abs().max() + 1e-300)), float((g - ref).norm() / (ref.norm() + 1e-300)), n))
    rows.sort(reverse=True)
    for e, l2, n in rows[:12]:
        print(f"   max {e:.2e}  l2 {l2:.2e}  {n}")
    print(f"   median max-rel {sorted(r[0] for r in rows)[len(rows) // 2]:.2e}")


if __name__ == "__main__":
    run(2, 1, 16, 16)
    run(2, 2, 24, 40)
    run(1, 2, 24, 40)
    run(0, 2, 24, 40)
    run(5, 1, 16, 16, "reflect")
    run(5, 1, 32, 32)
