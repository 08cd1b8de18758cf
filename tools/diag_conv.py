"""1x1 / 3x3 conv_gemm at the prod shape, standalone: CUDA-event timing and the per-tile pipeline trace of CTA 0
(pht_set_option("conv_trace", 1) + pht_conv_gemm_trace).   python tools/diag_conv.py [--k 256] [--n 256] [--ks 1]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=256)
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--ks", type=int, default=1)
ap.add_argument("--bias", type=int, default=0)
args = ap.parse_args()
dev = "cuda"
B, H, W = 8, 128, 128
torch.manual_seed(0)
srcs = [torch.randn(B, H, W, 256, device=dev).bfloat16() for _ in range(args.k // 256)]
w = (torch.randn(args.ks * args.ks, args.n, args.k, device=dev) * 0.05).bfloat16()
out = torch.empty(B, H, W, args.n, device=dev, dtype=torch.bfloat16)
bias = torch.randn(args.n, device=dev) if args.bias else None
slope = torch.zeros(args.n, device=dev) if args.bias else None
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
fn = lambda: ops.conv_gemm(srcs, w, args.n, ksize=args.ks, bias=bias, slope=slope, out1=out)
fn()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"conv_gemm K={args.k} N={args.n} ks={args.ks}: median {sorted(ts)[5]:.1f} us  min {min(ts):.1f} us")
_lib.lib.pht_set_option(b"conv_trace", 1)
flush.zero_()
fn()
buf = (C.c_int64 * (32 * 8 + 4 * 160))()
n = _lib.lib.pht_conv_gemm_trace(buf, 32 * 8 + 4 * 160)
_lib.lib.pht_set_option(b"conv_trace", 0)
ctas = torch.tensor(list(buf)[32 * 8:n]).view(-1, 4)[:148]
t = torch.tensor(list(buf)[:32 * 8]).view(-1, 8)
g0 = int(ctas[:, 0].min())
st, en = (ctas[:, 0] - g0).float() / 1e3, (ctas[:, 1] - g0).float() / 1e3
q = lambda v: " ".join(f"{float(x):7.1f}" for x in torch.quantile(v, torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0])))
print(f"per-CTA start us (min p10 p50 p90 max): {q(st)}")
print(f"per-CTA end   us (min p10 p50 p90 max): {q(en)}")
print(f"per-CTA busy  us (min p10 p50 p90 max): {q(en - st)}")
order = torch.argsort(en, descending=True)[:8]
print("slowest CTAs (cta, sm, end us): " + " ".join(f"({int(i)},{int(ctas[i, 2])},{float(en[i]):.1f})" for i in order))
t0 = int(t[0, 0])
print("per tile: 0=first stage issued 1=last stage issued 2=first MMA 3=MMAs issued 4=epi0 sees acc 5=epi0 done 6=epi1 sees acc 7=epi1 done")
for i in range(8):
    print(f"tile {i}: " + " ".join(f"{int(x) - t0:8d}" for x in t[i]))
