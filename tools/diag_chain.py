"""Back-to-back conv_gemm launches under CUDA-graph replay: effective time per launch (events around the whole chain)
against the per-CTA busy time the kernel's own trace reports -- the difference is the launch / prologue / drain time
that the step pays per kernel.   python tools/diag_chain.py [--ks 1] [--n 20]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ks", type=int, default=1)
ap.add_argument("--n", type=int, default=20)
ap.add_argument("--bufs", type=int, default=2, help="ring of activation buffers the chain walks (2 = ping-pong)")
args = ap.parse_args()
dev = "cuda"
B, H, W = 8, 128, 128
torch.manual_seed(0)
bufs = [torch.randn(B, H, W, 256, device=dev).bfloat16() for _ in range(args.bufs)]
w = (torch.randn(args.ks * args.ks, 256, 256, device=dev) * 0.02).bfloat16()
bias = torch.zeros(256, device=dev)
slope = torch.zeros(256, device=dev)


def chain():
    for i in range(args.n):
        ops.conv_gemm([bufs[i % args.bufs]], w, 256, ksize=args.ks, bias=bias, slope=slope, out1=bufs[(i + 1) % args.bufs])


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        chain()
    for mode in ("graph", "eager", "graph", "eager"):
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay() if mode == "graph" else chain()
            e1.record(s)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / args.n)
        print(f"ks={args.ks} bufs={args.bufs} {mode}: {sorted(ts)[3]:.2f} us per launch (chain of {args.n})")
    _lib.lib.pht_set_option(b"conv_trace", 1)
    chain()
    torch.cuda.synchronize()
    buf = (C.c_int64 * (32 * 8 + 4 * 160))()
    n = _lib.lib.pht_conv_gemm_trace(buf, 32 * 8 + 4 * 160)
    _lib.lib.pht_set_option(b"conv_trace", 0)
    ctas = torch.tensor(list(buf)[32 * 8:n]).view(-1, 4)[:148]
    g0 = int(ctas[:, 0].min())
    st, en = (ctas[:, 0] - g0).float() / 1e3, (ctas[:, 1] - g0).float() / 1e3
    q = lambda v: " ".join(f"{float(x):7.1f}" for x in torch.quantile(v, torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0])))
    print(f"last launch of an eager chain, per-CTA start us (min p10 p50 p90 max): {q(st)}")
    print(f"                               per-CTA end   us (min p10 p50 p90 max): {q(en)}")
    print(f"                               entry -> start us (min p10 p50 p90 max): {q((ctas[:, 0] - ctas[:, 3]).float() / 1e3)}")
