"""On-GPU diagnostics for the tcgen05 paths: compares every tensor-core kernel with the CUDA-core kernel of the
same C-ABI op on the same inputs (pht_set_force_simple), prints where they differ, and times the hot shapes.
Usage (on the GPU box): python tools/diag_tc.py [--perf]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_heal_thyself_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


def both(fn):
    _lib.lib.pht_set_force_simple(1)
    a = fn()
    _lib.lib.pht_set_force_simple(0)
    before = _lib.counters()
    b = fn()
    after = _lib.counters()
    used_tc = any(after[k] > before[k] for k in ("gemm_tc", "wgrad_tc", "attn_tc"))
    torch.cuda.synchronize()
    return a, b, used_tc


def report(name, a, b, used_tc):
    outs_a = a if isinstance(a, (list, tuple)) else [a]
    outs_b = b if isinstance(b, (list, tuple)) else [b]
    worst = 0.0
    for i, (x, y) in enumerate(zip(outs_a, outs_b)):
        x, y = x.float(), y.float()
        d = (x - y).abs()
        rel = float(d.max() / (x.abs().max() + 1e-30))
        worst = max(worst, rel)
        flag = "OK " if rel < 2e-2 else "BAD"
        print(f"  [{flag}] {name} out{i}: rel={rel:.3e} max|ref|={float(x.abs().max()):.3e} tc={used_tc} "
              f"nan={bool(torch.isnan(y).any())}")
        if rel >= 2e-2:
            bad = (d > 2e-2 * x.abs().max()).nonzero()
            print(f"        {bad.shape[0]} bad of {d.numel()}; first: {bad[:6].tolist()} last: {bad[-3:].tolist()}")
            if d.dim() == 4:
                print("        bad fraction by y:", [round(float(v), 2) for v in (d > 2e-2 * x.abs().max()).float().mean((0, 2, 3))[:20]])
                print("        bad fraction by x:", [round(float(v), 2) for v in (d > 2e-2 * x.abs().max()).float().mean((0, 1, 3))[:20]])
                cfrac = (d > 2e-2 * x.abs().max()).float().mean((0, 1, 2))
                print("        bad fraction by c/16:", [round(float(v), 2) for v in cfrac.view(-1, 16).mean(1)[:32]])
    return worst


def conv_case(B, H, W, Cs, N, ks, epi=False, pad_domain=False, seed=0):
    torch.manual_seed(seed)
    srcs = [torch.randn(B, H, W, c, device=DEV).bfloat16() for c in Cs]
    K = sum(Cs)
    w = (torch.randn(ks * ks, N, K, device=DEV) / (K * ks * ks) ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    slope = torch.full((N,), 0.2, device=DEV)
    Ho, Wo = (H + 2, W + 2) if pad_domain else (H, W)
    resid = torch.randn(B, Ho, Wo, N, device=DEV).bfloat16()
    mask = torch.randn(B, Ho, Wo, N, device=DEV).bfloat16()
    offs = [(-1, -1)] * len(Cs) if pad_domain else None

    def run():
        o1 = torch.zeros(B, Ho, Wo, N, device=DEV, dtype=torch.bfloat16)
        o2 = torch.zeros_like(o1)
        if epi:
            ops.conv_gemm(srcs, w, N, ksize=ks, bias=bias, slope=slope, resid=resid, resid_mode="post", mask=mask,
                          mslope=slope, out1=o1, out2=o2, src_offsets=offs)
        else:
            ops.conv_gemm(srcs, w, N, ksize=ks, out1=o1, src_offsets=offs)
        return [o1, o2]

    return run


def wgrad_case(B, H, W, Cs, N, ks, seed=0):
    torch.manual_seed(seed)
    pad = ks // 2
    srcs = [torch.randn(B, H + 2 * pad, W + 2 * pad, c, device=DEV).bfloat16() for c in Cs]
    dy = torch.randn(B, H, W, N, device=DEV).bfloat16()
    K = sum(Cs)
    ws = torch.empty(64 * 1024 * 1024 // 4, device=DEV)

    def run():
        dw = torch.zeros(ks * ks, N, K, device=DEV)
        db = torch.zeros(N, device=DEV)
        ops.wgrad(dy, srcs, dw, ksize=ks, dbias=db, workspace=ws, src_offsets=[(pad, pad)] * len(Cs))
        return [dw, db]

    return run


def attn_case(B, H, W, seed=0, bwd=False):
    torch.manual_seed(seed)
    C = 256
    qk = torch.randn(B, H, W, 2 * C, device=DEV)
    qk[..., :C] *= 0.125
    qk = qk.bfloat16()
    v = torch.randn(B, H, W, C, device=DEV).bfloat16()
    resid = torch.randn(B, H, W, C, device=DEV).bfloat16()
    rel_h, rel_w = torch.randn(14, 32, device=DEV), torch.randn(14, 32, device=DEV)
    do = torch.randn(B, H, W, C, device=DEV).bfloat16()

    def run():
        out = torch.zeros(B, H + 2, W + 2, C, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, H, W, 4, device=DEV)
        ops.attn_fwd(qk[..., :C], qk[..., C:], v, rel_h, rel_w, out[:, 1:-1, 1:-1, :], resid=resid, lse=lse)
        res = [out, lse]
        if bwd:
            dq = torch.zeros(B, H, W, 2 * C, device=DEV, dtype=torch.bfloat16)
            dk = torch.zeros(B, H, W, C, device=DEV, dtype=torch.bfloat16)
            dv = torch.zeros_like(dk)
            drh, drw = torch.zeros(14, 32, device=DEV), torch.zeros(14, 32, device=DEV)
            ws = torch.empty(max(ops.attn_bwd_workspace_bytes(v), 16) // 4, device=DEV)
            ops.attn_bwd(qk[..., :C], qk[..., C:], v, rel_h, rel_w, lse, do, dq[..., :C], dk, dv, drh, drw, ws)
            res += [dq, dk, dv, drh, drw]
        return res

    return run


def main():
    print(torch.cuda.get_device_name(0))
    cases = [
        ("1x1 C64 N64 8x16", conv_case(1, 8, 16, [64], 64, 1)),
        ("1x1 C64 N128 8x16", conv_case(1, 8, 16, [64], 128, 1)),
        ("1x1 C64 N256 8x16", conv_case(1, 8, 16, [64], 256, 1)),
        ("1x1 C256 N256 16x32 B2", conv_case(2, 16, 32, [256], 256, 1)),
        ("1x1 C128+256 N512 24x40 epi", conv_case(2, 24, 40, [128, 256], 512, 1, epi=True)),
        ("3x3 C64 N64 8x16", conv_case(1, 8, 16, [64], 64, 3)),
        ("3x3 C256 N256 32x32 B2 epi", conv_case(2, 32, 32, [256], 256, 3, epi=True)),
        ("3x3 C256 N256 24x40 padded-domain", conv_case(1, 24, 40, [256], 256, 3, pad_domain=True)),
        ("1x1 C192 N768 16x16", conv_case(1, 16, 16, [192], 768, 1)),
    ]
    cases += [
        ("wgrad 1x1 C64 N128 16x16", wgrad_case(1, 16, 16, [64], 128, 1)),
        ("wgrad 1x1 C256 N256 32x32 B2", wgrad_case(2, 32, 32, [256], 256, 1)),
        ("wgrad 3x3 C256 N256 24x40 B2", wgrad_case(2, 24, 40, [256], 256, 3)),
        ("wgrad 1x1 C256+256 N256 24x40", wgrad_case(1, 24, 40, [256, 256], 256, 1)),
        ("wgrad 1x1 C256 N512 16x48", wgrad_case(1, 16, 48, [256], 512, 1)),
        ("wgrad 1x1 C192 N768 16x16", wgrad_case(2, 16, 16, [192], 768, 1)),
        ("wgrad 1x1 C768 N256 8x8", wgrad_case(1, 8, 8, [768], 256, 1)),
    ]
    cases += [
        ("attn fwd 8x8", attn_case(1, 8, 8)),
        ("attn fwd 16x24 B2", attn_case(2, 16, 24)),
        ("attn fwd+bwd 32x32", attn_case(1, 32, 32, bwd=True)),
    ]
    if "--only-attn" in sys.argv:
        cases = [c for c in cases if c[0].startswith("attn")]
    worst = 0.0
    for name, fn in cases:
        try:
            a, b, tc = both(fn)
            worst = max(worst, report(name, a, b, tc))
        except Exception as e:  # noqa: BLE001
            print(f"  [EXC] {name}: {e}")
            worst = float("inf")
    print("WORST", worst)
    if "--perf" in sys.argv and worst < 2e-2:
        for (B, H, W, Cs, N, ks) in ((8, 128, 128, [256], 256, 3), (8, 128, 128, [256, 256], 256, 1),
                                     (8, 128, 128, [256], 512, 1), (8, 128, 128, [768], 256, 1), (8, 32, 32, [256], 256, 3)):
            fn = conv_case(B, H, W, Cs, N, ks)
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = 2.0 * B * H * W * sum(Cs) * N * ks * ks
            print(f"  perf ks={ks} C={Cs} N={N} {B}x{H}x{W}: {ms:.3f} ms/launch (incl. 2 memsets) = {fl / ms / 1e9:.1f} TFLOP/s")
        for bwd in (False, True):
            fn = attn_case(8, 128, 128, bwd=bwd)
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print(f"  perf attn {'fwd+bwd' if bwd else 'fwd'} 8x128x128: {e0.elapsed_time(e1) / 5:.3f} ms (incl. output memsets)")
        for (B, H, W, Cs, N, ks) in ((8, 128, 128, [256], 256, 3), (8, 128, 128, [256, 256], 256, 1),
                                     (8, 128, 128, [256], 512, 1), (8, 128, 128, [768], 256, 1)):
            fn = wgrad_case(B, H, W, Cs, N, ks)
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fl = 2.0 * B * H * W * sum(Cs) * N * ks * ks
            print(f"  perf wgrad ks={ks} C={Cs} N={N} {B}x{H}x{W}: {ms:.3f} ms (incl. colsum+memsets+reduce) = {fl / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
