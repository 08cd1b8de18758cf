"""Achieved HBM bandwidth of the bandwidth-bound kernels at FRAME scale (at train shapes they move only a
few MB per call and are launch-latency bound, so their GB/s is measured on 2048x2048-sized problems), against the
measured copy peak in MEASURED_PEAKS.json.  Algorithmic bytes per element are the ones stated in DESIGN.md.
    python tools/bench_bandwidth_kernels.py > gpurun_out/bandwidth.txt"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixel_heal_thyself_b200 import ops  # noqa: E402
from pixel_heal_thyself_b200._lib import PAD_MODES  # noqa: E402

dev = torch.device("cuda:0")
peak = 6541.1
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]


def timeit(fn, iters=20):
    iters = int(os.environ.get('PHT_BW_ITERS', iters))   # (2 under ncu: every launch is replayed per metric pass)
    # L2 flush by READING a buffer larger than the 126 MB L2 (clean lines: a write flush would leave dirty lines whose
    # write-back is then charged to the timed kernel)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    for _ in range(1 if 'PHT_BW_ITERS' in os.environ else 3):
        fn()
    ts = []
    for _ in range(iters):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:44s} {nbytes / 1e6:9.1f} MB  {ms * 1e3:8.1f} us  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of measured copy peak ({peak:.0f})")


torch.manual_seed(0)
S = 2048
npx = S * S
# L1 loss fused fwd+bwd: 36 B/px (3 ch fp32: read out + gt, write grad)
a = torch.randn(1, 3, S, S, device=dev)
b = torch.randn(1, 3, S, S, device=dev)
g = torch.empty_like(a)
loss = torch.empty(1, device=dev)
report("l1_loss fwd+bwd (3ch fp32)", npx * 36, timeit(lambda: ops.l1_loss(a, b, loss, g, 1.0)))
a8, b8 = torch.randn(8, 3, S, S, device=dev), torch.randn(8, 3, S, S, device=dev)
g8 = torch.empty_like(a8)
report("l1_loss fwd+bwd (8 frames, 1.2 GB)", 8 * npx * 36, timeit(lambda: ops.l1_loss(a8, b8, loss, g8, 1.0), 10))
del a8, b8, g8
# preprocess: 13 channels fp32 NHWC -> NCHW: 104 B/px
n_ = torch.rand(1, S, S, 3, device=dev)
t_ = torch.rand(1, S, S, 3, device=dev)
x_ = torch.rand(1, S, S, 7, device=dev)
no, go, ao = torch.empty(1, 3, S, S, device=dev), torch.empty(1, 3, S, S, device=dev), torch.empty(1, 7, S, S, device=dev)
report("preprocess (log1p, normal remap, NHWC->NCHW)", npx * 104, timeit(lambda: ops.preprocess(n_, t_, x_, no, go, ao)))
# crop + preprocess: 256 patches of 128^2 = 4.19 Mpx
P, n = 128, 256
cen = (torch.randint(64, S - 64, (n, 2), device=dev, dtype=torch.int32)).contiguous()
idx = torch.zeros(n, dtype=torch.int32, device=dev)
no2, go2, ao2 = torch.empty(n, 3, P, P, device=dev), torch.empty(n, 3, P, P, device=dev), torch.empty(n, 7, P, P, device=dev)
report("crop_preprocess (256 patches 128x128)", n * P * P * 104, timeit(lambda: ops.crop_preprocess(n_, t_, x_, cen, P, no2, go2, ao2, idx)))
# Adam: 28 B/param over 64 M params
N = 64 << 20
pp, gg, mm, vv = (torch.randn(N, device=dev) for _ in range(4))
vv.abs_()
report("adam (flat fp32 arena, 64M params)", N * 28, timeit(lambda: ops.adam(pp, gg, mm, vv, lr=1e-4, step=3)))
# border fill of a padded bf16 activation [8,130,130,256]: touches only the frame (read+write 2 * 516 px * 512 B per image)
buf = torch.randn(8, 130, 130, 256, device=dev).bfloat16()
report("border_fill (8x130x130x256 bf16, frame only)", 8 * 516 * 512 * 2, timeit(lambda: ops.border_fill(buf, PAD_MODES["replicate"])))
# pad_fold: read padded grad + mask, write 1 output: 3 * 512 B/px
gp = torch.randn(8, 130, 130, 256, device=dev).bfloat16()
mk = torch.randn(8, 128, 128, 256, device=dev).bfloat16()
o2 = torch.empty(8, 128, 128, 256, device=dev).bfloat16()
zs = torch.zeros(256, device=dev)
report("pad_fold + relu mask (8x128x128x256 bf16)", 8 * 128 * 128 * 512 * 3, timeit(lambda: ops.pad_fold(gp, PAD_MODES["replicate"], mask=mk, mslope=zs, out2=o2)))
# im2col5: write Kpad bf16 per px (reads are tiny): 8 x 128 x 128 x 192
x7 = torch.rand(8, 7, 128, 128, device=dev)
col = torch.empty(8, 128, 128, 192, device=dev).bfloat16()
report("im2col5 (7ch -> 192 bf16, 8x128x128)", 8 * 128 * 128 * (192 * 2 + 28), timeit(lambda: ops.im2col5(x7, col, PAD_MODES["replicate"])))
# decoder tail finish / tail im2col
y = torch.randn(8, 128, 128, 64, device=dev)
bias = torch.randn(3, device=dev)
x3 = torch.randn(8, 3, 128, 128, device=dev)
o3 = torch.empty_like(x3)
report("tail_finish (8x128x128)", 8 * 128 * 128 * (12 + 12 + 12), timeit(lambda: ops.tail_finish(y, bias, x3, o3)))
ta = torch.empty(8, 128, 128, 64, device=dev).bfloat16()
db = torch.empty(3, device=dev)
report("tail_im2col_bwd (8x128x128)", 8 * 128 * 128 * (12 + 128), timeit(lambda: ops.tail_im2col_bwd(x3, ta, db)))
# attention backward fold: reads 2 x window-major scratch (3.06x) + writes dk, dv
