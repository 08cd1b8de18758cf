"""Attention kernels at the prod shape (B=8, 128x128, C=256), standalone: CUDA-event timings and, for the backward
kernel, the pipeline trace of CTA 0 (pht_set_option("attn_trace", 1) + pht_attn_bwd_trace).
   python tools/diag_attn.py [--trace] [--iters 20]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trace", action="store_true")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--P", type=int, default=128)
args = ap.parse_args()
dev = "cuda"
B, H, W, Cc = args.B, args.P, args.P, 256
torch.manual_seed(0)
mk = lambda s=1.0: (torch.randn(B, H, W, Cc, device=dev) * s).bfloat16()
q, k, v, do = mk(0.125), mk(), mk(), mk()
rel_h, rel_w = torch.randn(14, 32, device=dev), torch.randn(14, 32, device=dev)
out = torch.empty_like(q)
lse = torch.empty(B, H, W, 4, device=dev)
dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
EV = 12   # AB_TRACE_EVENTS
drh, drw = torch.empty_like(rel_h), torch.empty_like(rel_w)
ws = torch.empty(ops.attn_bwd_workspace_bytes(q) // 4 + 64, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


fwd = lambda: ops.attn_fwd(q, k, v, rel_h, rel_w, out, lse=lse, resid=do)
bwd = lambda: ops.attn_bwd(q, k, v, rel_h, rel_w, lse, do, dq, dk, dv, drh, drw, ws)
print("attn_fwd  median %.1f us  min %.1f us" % timed(fwd, args.iters))
print("attn_bwd (zero + kernel + rel reduce), direct vector reductions  median %.1f us  min %.1f us" % timed(bwd, args.iters))
_lib.lib.pht_set_option(b"attn_bwd_direct", 0)
ws = torch.empty(ops.attn_bwd_workspace_bytes(q) // 4 + 64, device=dev)
print("attn_bwd (kernel + fold + rel reduce), scratch + fixed-order fold median %.1f us  min %.1f us" % timed(bwd, args.iters))
_lib.lib.pht_set_option(b"attn_bwd_direct", 1)
if args.trace:
    _lib.lib.pht_set_option(b"attn_trace", 2)
    fwd()
    buf = (C.c_int64 * (48 * EV))()
    n = _lib.lib.pht_attn_bwd_trace(buf, 48 * EV)
    _lib.lib.pht_set_option(b"attn_trace", 0)
    t = torch.tensor(list(buf)[:n]).view(-1, EV)
    t0 = int(t[0, 0])
    print("FORWARD events: 0=S issue(it) 1=PV issue(it) 2=s_full seen 3=max exchanged 4=epilogue(it-1) done 5=p_full arrive "
          "6=pv_done(it) seen [in epilogue, during it+1]")
    for i in range(4, 14):
        print(f"it {i:2d}: " + " ".join(f"{int(x) - t0:8d}" for x in t[i, :7]))
    d = t[4:24]
    seg = lambda a, b: float((d[:, b] - d[:, a]).float().mean())
    print("period %.0f | s_full seen->max %.0f | epilogue(it-1) %.0f | pass2 %.0f | p_full->PV issue %.0f | PV issue->pv_done seen %.0f"
          % (float((t[24, 2] - t[4, 2]) / 20), seg(2, 3), seg(3, 4), seg(4, 5), seg(5, 1), seg(1, 6)))
    _lib.lib.pht_set_option(b"attn_trace", 1)
    bwd()
    buf = (C.c_int64 * (48 * EV))()
    n = _lib.lib.pht_attn_bwd_trace(buf, 48 * EV)
    _lib.lib.pht_set_option(b"attn_trace", 0)
    t = torch.tensor(list(buf)[:n]).view(-1, EV)
    t0 = int(t[0, 0])
    names = ["M1 issue", "M2 issue", "sdp_full seen", "ds arrive", "dq_full seen", "out_full seen", "R done", "M1 done (MMA thread)"]
    print("events (clocks since M1 issue of iteration 0): " + ", ".join(f"{i}={n}" for i, n in enumerate(names)))
    for i in range(min(24, t.shape[0])):
        print(f"it {i:2d}: " + " ".join(f"{int(x) - t0:8d}" for x in t[i, :8]))
    d = t[8:40]
    print("mean per-iteration period (clk): %.0f" % float((t[40, 0] - t[8, 0]) / 32))
    seg = lambda a, b: float((d[:, b] - d[:, a]).float().mean())
    print("M1 issue->sdp_full seen %.0f | softmax/dS %.0f | ds arrive->M2 issue %.0f | M2 issue->dq_full %.0f | "
          "dq_full->out_full %.0f | out_full->R done %.0f" % (seg(0, 2), seg(2, 3), seg(3, 1), seg(1, 4), seg(4, 5), seg(5, 6)))
    print("M1 issue -> M1 done %.0f | M2 issue -> dQ done (MMA thread) %.0f -> dV/dK done %.0f | ds arrive -> dK(it-1) read-out done %.0f"
          % (seg(0, 7), seg(1, 8), seg(8, 9), seg(3, 10)))
    print("dK read-out: ds arrive -> first ld done %.0f -> first chunk stored %.0f -> all done %.0f" % (seg(3, 11), seg(11, 6), seg(6, 10)))
    nxt = (t[9:41, 0] - d[:, 6]).float().mean()
    print("R done -> next M1 issue %.0f (negative = M1 already issued)" % float(nxt))
