"""Compares the decoder-side backward intermediates of the fp32 CUDA path with fp64 autograd of the oracle.
Usage (GPU box): python tools/diag_decoder.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import afgsa_oracle as O  # noqa: E402
from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet  # noqa: E402
from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300)), float((a - b).norm() / (b.norm() + 1e-300))


def run(num_sa, B, H, W, seed=3):
    torch.manual_seed(990819)
    net = AFGSANet(3, 7, 256, num_sa=num_sa, num_gcp=0, padding_mode="replicate", compute_dtype="fp32").cuda()
    sd = {k: (v.detach().cpu().double() if v.dtype.is_floating_point else v.cpu()) for k, v in net.state_dict().items()}
    torch.manual_seed(seed)
    x, aux, gt = torch.randn(B, 3, H, W) * 0.5, torch.rand(B, 7, H, W), torch.randn(B, 3, H, W) * 0.5
    xd, ad, gd = x.double(), aux.double(), gt.double()
    with torch.no_grad():
        f = O.encoder_noisy(xd, sd, "replicate")
        a = O.encoder_aux(ad, sd, "replicate")
        for i in range(num_sa):
            f = O.transformer_block(f, a, sd, f"transformer_blocks.{i}.", "replicate")
    f = f.clone().requires_grad_(True)
    d1 = O.conv_block(f, sd["decoder.0.0.weight"], sd["decoder.0.0.bias"], "relu", "replicate")
    d1.retain_grad()
    d2 = O.conv_block(d1, sd["decoder.1.0.weight"], sd["decoder.1.0.bias"], "relu", "replicate")
    d2.retain_grad()
    out = O.conv_block(d2, sd["decoder.2.0.weight"], sd["decoder.2.0.bias"], None, "zeros") + xd
    out.retain_grad()
    loss = O.l1_loss(out, gd)
    loss.backward()
    sink = {}
    net.engine.debug_sink = sink
    o = net(x.cuda(), aux.cuda())
    L1ReconstructionLoss()(o, gt.cuda()).backward()
    torch.cuda.synchronize()
    nhwc = lambda t: t.permute(0, 2, 3, 1)
    print(f"--- num_sa={num_sa} B={B} {H}x{W}")
    print("   out         ", rel(o.detach(), out.detach()))
    print("   d_out       ", rel(sink["d_out"], out.grad))
    nflip = int((torch.sign(sink["d_out"].cpu().double()) != torch.sign(out.grad)).sum())
    print("   d_out sign flips", nflip, "of", out.numel(), " min|out-gt|", float((out.detach() - gd).abs().min()))
    print("   D2          ", rel(sink["D2"], nhwc(d2.detach())))
    print("   D1 (interior)", rel(sink["D1"][:, 1:-1, 1:-1, :], nhwc(d1.detach())))
    print("   dD2pre      ", rel(sink["dD2pre"], nhwc(d2.grad * (d2.detach() > 0))))
    print("   dD1pre      ", rel(sink["dD1pre"], nhwc(d1.grad * (d1.detach() > 0))))
    if num_sa > 0:
        print("   dXlast      ", rel(sink["dXlast"], nhwc(f.grad)))
    m1 = (sink["D1"][:, 1:-1, 1:-1, :].cpu() > 0) != (nhwc(d1.detach()) > 0)
    m2 = (sink["D2"].cpu() > 0) != (nhwc(d2.detach()) > 0)
    print("   relu mask flips D1:", int(m1.sum()), "D2:", int(m2.sum()), "of", m1.numel())


if __name__ == "__main__":
    run(1, 2, 24, 40)
    run(2, 2, 24, 40)
    run(2, 1, 16, 16)
