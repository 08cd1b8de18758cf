"""Per-op parity of the CUDA kernels (through the C ABI) against the CPU oracle / a torch fp32 reference.

fp32 ops: tolerance 1e-5 relative to the reference's max magnitude (north_star's fp32 bar).
bf16 ops: inputs are bf16-rounded first, the reference is computed in fp32 from those rounded inputs,
and the result may differ by bf16 output rounding (2^-8 relative) plus accumulation-order noise.
Integer work (sampler) is bit-exact.
"""
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_json, load_npz
from oracle import afgsa_oracle as O
from oracle import sampler_oracle as S

pytestmark = pytest.mark.gpu

DEV = "cuda"
F32_TOL = 1e-5
BF16_TOL = 1.2e-2


def _ops():
    from pixel_heal_thyself_b200 import ops
    return ops


@pytest.fixture
def cuda_core_bf16_allowed():
    """Some per-op cases are deliberately NOT tensor-core shaped (64-wide weight-gradient, head_dim 8): they check the
    CUDA-core kernels' bf16 instantiations, which the library only runs when the "bf16_fallback" option is on."""
    from pixel_heal_thyself_b200 import _lib
    assert _lib.lib.pht_set_option(b"bf16_fallback", 1) == 0
    yield
    _lib.lib.pht_set_option(b"bf16_fallback", 0)



def rel_err(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-30))


def tol(dtype):
    return F32_TOL if dtype == torch.float32 else BF16_TOL


def nhwc(t):  # NCHW -> NHWC contiguous
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def pack(w, dtype, **kw):
    ops = _ops()
    O_, I_, ks, _ = w.shape
    T = ks * ks
    if kw.get("transpose"):
        packed = torch.zeros(T, I_, O_, dtype=dtype, device=DEV)
        ops.pack_weight(w, packed, ksize=ks, Ntot=I_, Ktot=O_, **kw)
    else:
        packed = torch.zeros(T, O_, I_, dtype=dtype, device=DEV)
        ops.pack_weight(w, packed, ksize=ks, Ntot=O_, Ktot=I_, **kw)
    return packed


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ks,cin,cout,B,H,W", [(1, 64, 64, 2, 8, 16), (3, 64, 128, 2, 16, 24), (3, 256, 256, 1, 16, 16),
                                                (1, 768, 256, 1, 8, 8)])
def test_conv_gemm_zero_pad_matches_conv2d(dtype, ks, cin, cout, B, H, W):
    ops = _ops()
    torch.manual_seed(1)
    x = torch.randn(B, cin, H, W, device=DEV).to(dtype)
    w = (torch.randn(cout, cin, ks, ks, device=DEV) / (cin * ks * ks) ** 0.5)
    bias = torch.randn(cout, device=DEV)
    wp = pack(w, dtype)
    ref = F.relu(F.conv2d(x.float(), wp.float().view(ks, ks, cout, cin).permute(2, 3, 0, 1), bias, padding=ks // 2))
    out = torch.empty(B, H, W, cout, dtype=dtype, device=DEV)
    ops.conv_gemm([nhwc(x)], wp, cout, ksize=ks, bias=bias, slope=torch.zeros(cout, device=DEV), out1=out)
    assert rel_err(nchw(out), ref) < tol(dtype)


@pytest.mark.parametrize("option", [b"cta_pairs", b"half_ring"])
@pytest.mark.parametrize("ks,cin,cout,B,H,W,fused", [(3, 256, 256, 2, 32, 32, False), (1, 512, 256, 3, 24, 16, True),
                                                      (3, 256, 512, 1, 24, 16, True), (1, 256, 256, 2, 128, 128, False),
                                                      (3, 256, 256, 1, 40, 48, True)])
def test_conv_gemm_pipeline_variants_bit_identical(ks, cin, cout, B, H, W, fused, option):
    """Option "cta_pairs" = 1 runs the 256-wide GEMMs as CTA pairs (tcgen05 cta_group::2, M = 256); "half_ring" = 1 runs
    the 3x3 ones on the 8 x 24 KB half-stage operand ring (64-byte swizzle rows).  Same K order, same epilogue: the
    outputs must be bit-identical to the default kernel, an odd number of pixel tiles (one CTA of the last pair has no
    tile) and two output tiles (N = 512) included."""
    ops = _ops()
    from pixel_heal_thyself_b200 import _lib
    torch.manual_seed(5)
    dtype = torch.bfloat16
    x = torch.randn(B, H, W, cin, device=DEV).to(dtype)
    wp = pack(torch.randn(cout, cin, ks, ks, device=DEV) / (cin * ks * ks) ** 0.5, dtype)
    bias = torch.randn(cout, device=DEV)
    slope = torch.full((cout,), 0.2, device=DEV)
    resid = torch.randn(B, H, W, cout, device=DEV).to(dtype) if fused else None

    def run():
        o1 = torch.zeros(B, H, W, cout, dtype=dtype, device=DEV)
        o2 = torch.zeros(B, H, W, cout, dtype=dtype, device=DEV) if fused else None
        kw = dict(resid=resid, resid_mode="post", out2=o2) if fused else {}
        before = _lib.counters()["gemm_tc"]
        ops.conv_gemm([x], wp, cout, ksize=ks, bias=bias, slope=slope, out1=o1, **kw)
        assert _lib.counters()["gemm_tc"] == before + 1
        torch.cuda.synchronize()
        return o1, o2

    a = run()
    _lib.lib.pht_set_option(option, 1)
    try:
        b = run()
    finally:
        _lib.lib.pht_set_option(option, 0)
    assert torch.equal(a[0], b[0])
    if fused:
        assert torch.equal(a[1], b[1])
    assert float(a[0].float().abs().max()) > 0


@pytest.mark.parametrize("mode", ["replicate", "reflect"])
@pytest.mark.parametrize("ks,B,H,W,two_out", [(3, 2, 32, 48, False), (3, 1, 24, 16, True), (1, 3, 8, 16, True), (3, 8, 128, 128, True)])
def test_conv_gemm_ring_epilogue_equals_border_fill(mode, ks, B, H, W, two_out):
    """PHT_EPI_RING1 / RING2: the GEMM's epilogue also writes the 1-pixel padding frame of a padded output buffer.  The
    whole padded buffer must be bit-identical to (same GEMM into the interior) + pht_border_fill, for both outputs and both
    padding modes; the flags are refused off the tensor-core path."""
    ops = _ops()
    torch.manual_seed(9)
    dtype, C = torch.bfloat16, 256
    x = torch.randn(B, H + 2, W + 2, C, device=DEV).to(dtype) if ks == 3 else torch.randn(B, H, W, C, device=DEV).to(dtype)
    wp = pack(torch.randn(C, C, ks, ks, device=DEV) / (C * ks * ks) ** 0.5, dtype)
    bias, slope = torch.randn(C, device=DEV), torch.zeros(C, device=DEV)
    resid = torch.randn(B, H, W, C, device=DEV).to(dtype) if two_out else None
    pmode = {"replicate": 0, "reflect": 1}[mode]

    def run(fused):
        p1 = torch.full((B, H + 2, W + 2, C), 7.0, dtype=dtype, device=DEV)
        p2 = torch.full((B, H + 2, W + 2, C), 7.0, dtype=dtype, device=DEV) if two_out else None
        kw = dict(resid=resid, resid_mode="post", out2=p2[:, 1:-1, 1:-1]) if two_out else {}
        ops.conv_gemm([x], wp, C, ksize=ks, src_offsets=[(1, 1)] if ks == 3 else None, out_domain=(B, H, W), bias=bias,
                      slope=slope, out1=p1[:, 1:-1, 1:-1], ring1=mode if fused else None,
                      ring2=mode if (fused and two_out) else None, **kw)
        if not fused:
            ops.border_fill(p1, pmode)
            if two_out:
                ops.border_fill(p2, pmode)
        torch.cuda.synchronize()
        return p1, p2

    a, b = run(False), run(True)
    assert torch.equal(a[0], b[0])
    if two_out:
        assert torch.equal(a[1], b[1])
    assert float(a[0][:, 0].float().abs().max()) > 0          # (the frame was written)
    with pytest.raises(RuntimeError):                          # fp32 (CUDA-core path): refused, not ignored
        xf = torch.randn(1, 8, 8, 64, device=DEV)
        pf = torch.zeros(1, 10, 10, 64, device=DEV)
        ops.conv_gemm([xf], pack(torch.randn(64, 64, 1, 1, device=DEV), torch.float32), 64, out1=pf[:, 1:-1, 1:-1], ring1=mode)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["replicate", "reflect"])
def test_padded_conv_forward_and_backward(dtype, mode, cuda_core_bf16_allowed):
    """3x3 conv with replicate/reflect padding: forward through border_fill + conv_gemm, data-grad through the
    padded-domain conv_gemm + pad_fold, weight-grad through wgrad.  Reference: autograd of F.pad + F.conv2d."""
    ops = _ops()
    from pixel_heal_thyself_b200._lib import PAD_MODES
    torch.manual_seed(2)
    B, C, N, H, W = 2, 64, 64, 16, 24
    x = torch.randn(B, C, H, W, device=DEV).to(dtype)
    w = torch.randn(N, C, 3, 3, device=DEV) / (9 * C) ** 0.5
    bias = torch.randn(N, device=DEV)
    dy = torch.randn(B, N, H, W, device=DEV).to(dtype)
    wq = pack(w, dtype).float().view(3, 3, N, C).permute(2, 3, 0, 1).contiguous()  # weights as the kernel sees them
    xr = x.float().requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    yr = F.relu(F.conv2d(F.pad(xr, (1, 1, 1, 1), mode=mode), wr, br))
    # fused: d(pre-activation) = dy * (y > 0)
    yr.backward(dy.float())
    m = PAD_MODES[mode]
    xp = torch.zeros(B, H + 2, W + 2, C, dtype=dtype, device=DEV)
    xp[:, 1:-1, 1:-1, :] = nhwc(x)
    ops.border_fill(xp, m)
    assert torch.equal(nchw(xp).float(), F.pad(x.float(), (1, 1, 1, 1), mode=mode))
    y = torch.empty(B, H, W, N, dtype=dtype, device=DEV)
    ops.conv_gemm([xp], pack(w, dtype), N, ksize=3, src_offsets=[(1, 1)], bias=bias, slope=torch.zeros(N, device=DEV), out1=y)
    assert rel_err(nchw(y), yr) < tol(dtype)
    # backward: dpre = dy * (y>0) (reference mask), dgrad over padded domain, fold
    dpre = (dy.float() * (yr > 0)).to(dtype)
    gp = torch.empty(B, H + 2, W + 2, C, dtype=dtype, device=DEV)
    ops.conv_gemm([nhwc(dpre)], pack(w, dtype, transpose=1), C, ksize=3, out_domain=(B, H + 2, W + 2),
                  src_offsets=[(-1, -1)], out1=gp)
    dx = torch.empty(B, H, W, C, dtype=dtype, device=DEV)
    ops.pad_fold(gp, m, out1=dx)
    assert rel_err(nchw(dx), xr.grad) < (tol(dtype) if dtype == torch.float32 else 2.5e-2)
    dw = torch.empty(9, N, C, dtype=torch.float32, device=DEV)
    db = torch.empty(N, dtype=torch.float32, device=DEV)
    ws = torch.empty(1 << 22, dtype=torch.float32, device=DEV)
    ops.wgrad(nhwc(dpre), [xp], dw, ksize=3, dbias=db, workspace=ws, src_offsets=[(1, 1)])
    dw_oihw = torch.empty(N, C, 3, 3, device=DEV)
    ops.unpack_wgrad(dw_oihw, dw, ksize=3, Ntot=N, Ktot=C)
    wgrad_ref = torch.autograd.grad(F.conv2d(F.pad(x.float(), (1, 1, 1, 1), mode=mode), wr), wr, dpre.float())[0]
    assert rel_err(dw_oihw, wgrad_ref) < (1e-4 if dtype == torch.float32 else 1e-3)
    assert rel_err(db, dpre.float().sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("B,C,N,H,W", [(2, 64, 64, 16, 24), (1, 256, 256, 32, 32), (2, 128, 64, 8, 40), (1, 64, 64, 24, 136)])
def test_fused_replicate_padfold_dgrad(B, C, N, H, W):
    """PHT_EPI_PADFOLD: data-gradient of a replicate-padded 3x3 conv with the padding backward, the residual add and
    the ReLU mask fused into the GEMM epilogue == autograd of F.pad(replicate) + conv2d, and == the unfused
    conv_gemm + pht_pad_fold pair.  (W = 24 / 40 exercise the half-filled last 16-pixel tile column, W = 136 three 2x64 strip tiles on the last two rows.)"""
    ops = _ops()
    from pixel_heal_thyself_b200._lib import PAD_MODES
    torch.manual_seed(5)
    dt = torch.bfloat16
    w = torch.randn(N, C, 3, 3, device=DEV) / (9 * C) ** 0.5
    dy = torch.randn(B, N, H, W, device=DEV).to(dt)
    resid = torch.randn(B, C, H, W, device=DEV).to(dt)
    mask = torch.randn(B, C, H, W, device=DEV).to(dt)
    wq = pack(w, dt).float().view(3, 3, N, C).permute(2, 3, 0, 1).contiguous()
    xr = torch.zeros(B, C, H, W, device=DEV, requires_grad=True)
    F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="replicate"), wq).backward(dy.float())
    ref1 = xr.grad + resid.float()
    ref2 = ref1 * (mask.float() > 0)
    wT = pack(w, dt, transpose=1)
    o1p = torch.zeros(B, H + 2, W + 2, C, dtype=dt, device=DEV)   # outputs are padded frames, results in the interior
    o2p = torch.zeros(B, H + 2, W + 2, C, dtype=dt, device=DEV)
    o1, o2 = o1p[:, 1:-1, 1:-1, :], o2p[:, 1:-1, 1:-1, :]
    ops.conv_gemm([nhwc(dy)], wT, C, ksize=3, out_domain=(B, H + 2, W + 2), src_offsets=[(-1, -1)], padfold=True,
                  resid=nhwc(resid), resid_mode="pre", mask=nhwc(mask), mslope=torch.zeros(C, device=DEV), out1=o1p, out2=o2p)
    assert rel_err(nchw(o1), ref1) < 2.5e-2
    assert rel_err(nchw(o2), ref2) < 2.5e-2
    # against the unfused pair (same bf16 rounding points up to the border values)
    gp = torch.empty(B, H + 2, W + 2, C, dtype=dt, device=DEV)
    ops.conv_gemm([nhwc(dy)], wT, C, ksize=3, out_domain=(B, H + 2, W + 2), src_offsets=[(-1, -1)], out1=gp)
    u1 = torch.empty_like(o1)
    u2 = torch.empty_like(o2)
    ops.pad_fold(gp, PAD_MODES["replicate"], resid=nhwc(resid), mask=nhwc(mask), mslope=torch.zeros(C, device=DEV),
                 out1=u1, out2=u2)
    assert rel_err(o1.contiguous(), u1) < 1e-2 and rel_err(o2.contiguous(), u2) < 1e-2
    # single output, mask only (the decoder / feed-forward uses)
    o3p = torch.zeros(B, H + 2, W + 2, C, dtype=dt, device=DEV)
    ops.conv_gemm([nhwc(dy)], wT, C, ksize=3, out_domain=(B, H + 2, W + 2), src_offsets=[(-1, -1)], padfold=True,
                  mask=nhwc(mask), mslope=torch.zeros(C, device=DEV), out2=o3p)
    assert rel_err(nchw(o3p[:, 1:-1, 1:-1, :]), xr.grad * (mask.float() > 0)) < 2.5e-2


def test_padfold_flag_is_refused_off_the_tensor_core_path():
    ops = _ops()
    x = torch.randn(1, 10, 10, 64, device=DEV)
    w = torch.randn(9, 64, 64, device=DEV)
    out = torch.empty(1, 8, 8, 64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.conv_gemm([x], w, 64, ksize=3, out_domain=(1, 10, 10), padfold=True, out1=out)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_gemm_virtual_concat_and_epilogues(dtype):
    ops = _ops()
    torch.manual_seed(3)
    B, H, W, C1, C2, N = 2, 8, 8, 64, 128, 64
    a = torch.randn(B, H, W, C1, device=DEV).to(dtype)
    b = torch.randn(B, H, W, C2, device=DEV).to(dtype)
    w = torch.randn(N, C1 + C2, 1, 1, device=DEV) / (C1 + C2) ** 0.5
    bias = torch.randn(N, device=DEV)
    slope = torch.full((N,), 0.2, device=DEV)
    resid = torch.randn(B, H, W, N, device=DEV).to(dtype)
    mask = torch.randn(B, H, W, N, device=DEV).to(dtype)
    wp = pack(w, dtype)
    acc = torch.cat([a, b], -1).float() @ wp.float()[0].t() + bias
    # forward-style: out1 = act(acc), out2 = out1 + resid
    o1 = torch.empty(B, H, W, N, dtype=dtype, device=DEV)
    o2 = torch.empty_like(o1)
    ops.conv_gemm([a, b], wp, N, bias=bias, slope=slope, resid=resid, resid_mode="post", out1=o1, out2=o2)
    ref1 = F.leaky_relu(acc, 0.2)
    assert rel_err(o1, ref1) < tol(dtype)
    assert rel_err(o2, ref1.to(dtype).float() + resid.float()) < tol(dtype)
    # backward-style: v = acc + resid ; out1 = v ; out2 = v * dact(mask) (leaky 0.2)
    ops.conv_gemm([a, b], wp, N, bias=bias, resid=resid, resid_mode="pre", mask=mask, mslope=slope, out1=o1, out2=o2)
    v = acc + resid.float()
    assert rel_err(o1, v) < tol(dtype)
    assert rel_err(o2, v.to(dtype).float() * torch.where(mask.float() > 0, 1.0, 0.2)) < tol(dtype)
    # strided views: write into the interior of a padded buffer and read a channel slice
    buf = torch.zeros(B, H + 2, W + 2, N, dtype=dtype, device=DEV)
    wide = torch.cat([a, b], -1).contiguous()
    ops.conv_gemm([wide[..., :C1], wide[..., C1:]], wp, N, bias=bias, out1=buf[:, 1:-1, 1:-1, :])
    assert rel_err(buf[:, 1:-1, 1:-1, :], acc) < tol(dtype)
    assert float(buf[:, 0].abs().max()) == 0.0 and float(buf[:, :, 0].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_wgrad_1x1_two_sources(dtype):
    ops = _ops()
    torch.manual_seed(4)
    B, H, W, C1, C2, N = 2, 16, 16, 64, 64, 128
    a = torch.randn(B, H, W, C1, device=DEV).to(dtype)
    b = torch.randn(B, H, W, C2, device=DEV).to(dtype)
    dy = torch.randn(B, H, W, N, device=DEV).to(dtype)
    dw = torch.empty(1, N, C1 + C2, device=DEV)
    db = torch.empty(N, device=DEV)
    ops.wgrad(dy, [a, b], dw, dbias=db, workspace=torch.empty(1 << 22, device=DEV))
    ref = dy.float().reshape(-1, N).t() @ torch.cat([a, b], -1).float().reshape(-1, C1 + C2)
    assert rel_err(dw[0], ref) < (1e-4 if dtype == torch.float32 else 1e-3)
    assert rel_err(db, dy.float().sum((0, 1, 2))) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["replicate", "reflect"])
def test_encoder_im2col_gemm_matches_three_convs(dtype, mode):
    """conv1 / conv3 / conv5 (model.py:606-622) as one GEMM over the 5x5 im2col."""
    ops = _ops()
    from pixel_heal_thyself_b200._lib import PAD_MODES
    torch.manual_seed(5)
    B, Cin, H, W = 2, 7, 16, 24
    x = torch.randn(B, Cin, H, W, device=DEV)
    ws = [torch.randn(256, Cin, k, k, device=DEV) / (Cin * k * k) ** 0.5 for k in (1, 3, 5)]
    kpad = 192
    wp = torch.zeros(1, 768, kpad, dtype=dtype, device=DEV)
    for j, (w, k) in enumerate(zip(ws, (1, 3, 5))):
        ops.pack_weight(w, wp, ksize=k, Ntot=768, Ktot=kpad, n_off=256 * j, grid=5)
    col = torch.empty(B, H, W, kpad, dtype=dtype, device=DEV)
    ops.im2col5(x, col, PAD_MODES[mode])
    out = torch.empty(B, H, W, 768, dtype=dtype, device=DEV)
    ops.conv_gemm([col], wp, 768, out1=out)
    xq = x.to(dtype).float()
    refs = []
    for w, k in zip(ws, (1, 3, 5)):
        wq = w.to(dtype).float()
        xp_ = F.pad(xq, (k // 2,) * 4, mode=mode) if k > 1 else xq
        refs.append(F.conv2d(xp_, wq))
    assert rel_err(nchw(out), torch.cat(refs, 1)) < tol(dtype)
    # the embedded kernels unpack back to the original OIHW gradients layout
    back = torch.empty_like(ws[1])
    ops.unpack_wgrad(back, wp.float(), ksize=3, Ntot=768, Ktot=kpad, n_off=256, grid=5)
    assert rel_err(back, ws[1].to(dtype)) < 1e-6


def _attn_inputs(B, C, H, W, dtype, seed):
    torch.manual_seed(seed)
    heads, d = 4, C // 4
    q = (torch.randn(B, C, H, W) * d ** -0.5).to(dtype)
    k = torch.randn(B, C, H, W).to(dtype)
    v = torch.randn(B, C, H, W).to(dtype)
    rel_h = torch.randn(1, 14, 1, d // 2)
    rel_w = torch.randn(1, 1, 14, d // 2)
    return q, k, v, rel_h, rel_w


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,H,W", [(1, 256, 16, 24), (2, 32, 8, 8), (1, 256, 8, 8), (2, 256, 96, 104)])
def test_attention_forward_backward_matches_oracle(dtype, B, C, H, W, cuda_core_bf16_allowed):
    ops = _ops()
    q, k, v, rel_h, rel_w = _attn_inputs(B, C, H, W, dtype, 6)
    do = torch.randn(B, C, H, W).to(dtype)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    rh, rw = rel_h.clone().requires_grad_(True), rel_w.clone().requires_grad_(True)
    ref = O.attention_core(qr, kr, vr, rh, rw, 8, 3, 4)
    ref.backward(do.float())
    g = lambda t: nhwc(t).to(DEV)
    qd, kd, vd = g(q), g(k), g(v)
    out = torch.empty(B, H, W, C, dtype=dtype, device=DEV)
    lse = torch.empty(B, H, W, 4, device=DEV)
    rhd, rwd = rel_h.to(DEV), rel_w.to(DEV)
    ops.attn_fwd(qd, kd, vd, rhd, rwd, out, lse=lse)
    assert rel_err(nchw(out).cpu(), ref.detach()) < (1e-5 if dtype == torch.float32 else 1e-2)
    if dtype == torch.float32:  # fused residual (x + attention, model.py:579)
        resid = torch.randn(B, H, W, C, device=DEV)
        out2 = torch.empty_like(out)
        ops.attn_fwd(qd, kd, vd, rhd, rwd, out2, resid=resid)
        assert rel_err(out2 - resid, out) < 1e-5
    dq = torch.empty_like(qd)
    dk = torch.zeros(B, H, W, C, dtype=dtype, device=DEV)
    dv = torch.zeros(B, H, W, C, dtype=dtype, device=DEV)
    drh, drw = torch.empty_like(rhd), torch.empty_like(rwd)
    ws = torch.empty(max(ops.attn_bwd_workspace_bytes(qd), 16) // 4, device=DEV)
    ops.attn_bwd(qd, kd, vd, rhd, rwd, lse, g(do), dq, dk, dv, drh, drw, ws)
    t = 2e-5 if dtype == torch.float32 else 3e-2
    assert rel_err(nchw(dq).cpu(), qr.grad) < t
    assert rel_err(nchw(dk).cpu(), kr.grad) < t
    assert rel_err(nchw(dv).cpu(), vr.grad) < t
    assert rel_err(drh.cpu(), rh.grad) < t
    assert rel_err(drw.cpu(), rw.grad) < t


@pytest.mark.parametrize("mode", ["replicate", "reflect"])
@pytest.mark.parametrize("B,H,W", [(2, 16, 24), (1, 8, 8), (8, 128, 128)])
def test_attention_forward_ring_equals_border_fill(mode, B, H, W):
    """pht_attn_fwd with ``ring``: the kernel also writes the 1-pixel frame of the padded buffer its output lives in;
    the padded buffer must be bit-identical to (attention into the interior) + pht_border_fill.  Refused in fp32 mode."""
    ops = _ops()
    torch.manual_seed(12)
    C, dt = 256, torch.bfloat16
    q, k, v, x = (torch.randn(B, H, W, C, device=DEV).to(dt) for _ in range(4))
    rh, rw = torch.randn(1, 14, 1, 32, device=DEV), torch.randn(1, 1, 14, 32, device=DEV)

    def run(fused):
        p = torch.full((B, H + 2, W + 2, C), 3.0, dtype=dt, device=DEV)
        ops.attn_fwd(q, k, v, rh, rw, p[:, 1:-1, 1:-1], resid=x, ring=mode if fused else None)
        if not fused:
            ops.border_fill(p, {"replicate": 0, "reflect": 1}[mode])
        torch.cuda.synchronize()
        return p

    a, b = run(False), run(True)
    assert torch.equal(a, b)
    assert float((a[:, 0].float() - 3.0).abs().max()) > 0
    with pytest.raises(RuntimeError):
        qf = torch.randn(1, 8, 8, C, device=DEV)
        pf = torch.zeros(1, 10, 10, C, device=DEV)
        ops.attn_fwd(qf, qf, qf, rh, rw, pf[:, 1:-1, 1:-1], ring=mode)


def test_attention_rejects_unaligned_maps():
    ops = _ops()
    q = torch.zeros(1, 12, 16, 256, device=DEV)
    with pytest.raises(RuntimeError, match="divisible by the block size"):
        ops.attn_fwd(q, q, q, torch.zeros(14, 32, device=DEV), torch.zeros(14, 32, device=DEV), torch.empty_like(q))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decoder_tail_forward_backward(dtype):
    ops = _ops()
    torch.manual_seed(7)
    B, C, H, W = 2, 256, 16, 24
    h = torch.relu(torch.randn(B, C, H, W, device=DEV)).to(dtype)
    w = torch.randn(3, C, 3, 3, device=DEV) / (9 * C) ** 0.5
    bias = torch.randn(3, device=DEV)
    x = torch.randn(B, 3, H, W, device=DEV)
    dout = torch.randn(B, 3, H, W, device=DEV)
    hr = h.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    ref = F.conv2d(hr, wr, br, padding=1) + x
    ref.backward(dout)
    wk = w.permute(0, 2, 3, 1).reshape(3, 9, C).contiguous()
    out = torch.empty_like(x)
    hn = nhwc(h)
    ops.dec_tail_fwd(hn, wk, bias, x, out)
    assert rel_err(out, ref) < 1e-5
    dh = torch.empty_like(hn)
    ops.dec_tail_bwd_data(dout, wk, hn, dh)
    assert rel_err(nchw(dh), hr.grad * (h.float() > 0)) < tol(dtype)
    dw = torch.empty(3, 9, C, device=DEV)
    db = torch.empty(3, device=DEV)
    ws = torch.empty(ops.dec_tail_ws_bytes(B, H, W, C) // 4, device=DEV)
    ops.dec_tail_bwd_weight(dout, hn, dw, db, ws)
    assert rel_err(dw.view(3, 3, 3, C).permute(0, 3, 1, 2), wr.grad) < 1e-4
    assert rel_err(db, br.grad) < 1e-4


@pytest.mark.parametrize("n", [1, 7, 3 * 128 * 128 * 8 + 3, 5_000_000])
def test_l1_loss_and_grad(n):
    ops = _ops()
    torch.manual_seed(8)
    a, b = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    b[: n // 7] = a[: n // 7]  # exact ties -> sign 0
    loss = torch.empty(1, device=DEV)
    grad = torch.empty_like(a)
    ops.l1_loss(a, b, loss, grad)
    ref = float(O.l1_loss(a.double().cpu(), b.double().cpu()))
    assert abs(float(loss) - ref) / ref < 1e-6
    assert torch.equal(grad, torch.sign(a - b) / n)


def test_preprocess_matches_reference_golden():
    ops = _ops()
    g = load_npz("preprocess.npz")
    n, t, a = (torch.from_numpy(g[k]).to(DEV) for k in ("noisy_hwc", "gt_hwc", "aux_hwc"))
    B, P = n.shape[0], n.shape[1]
    no, to_, ao = (torch.empty(B, c, P, P, device=DEV) for c in (3, 3, 7))
    ops.preprocess(n, t, a, no, to_, ao)
    assert np.abs(no.cpu().numpy() - g["noisy"]).max() < 1e-6   # logf vs np.log: <= 1 ulp
    assert np.abs(to_.cpu().numpy() - g["gt"]).max() < 1e-6
    assert np.array_equal(ao.cpu().numpy(), g["aux"])           # clamp / nan_to_num are exact


def test_crop_preprocess_matches_oracle_crop():
    ops = _ops()
    from pixel_heal_thyself_b200.data import synthetic_frames
    fr = synthetic_frames(2, 96, 128, 11, DEV)
    P = 32
    centres = torch.tensor([[16, 16], [100, 50], [64, 80], [111, 79]], dtype=torch.int32, device=DEV)
    img = torch.tensor([0, 1, 1, 0], dtype=torch.int32, device=DEV)
    no, to_, ao = (torch.empty(4, c, P, P, device=DEV) for c in (3, 3, 7))
    ops.crop_preprocess(fr["noisy"], fr["gt"], fr["aux"], centres, P, no, to_, ao, img)
    rn, rg, ra = O.preprocess_batch(*(torch.from_numpy(np.stack([S.crop_patches(
        fr[k][int(img[i])].cpu().numpy(), centres[i:i + 1].cpu().numpy(), P)[0] for i in range(4)]))
        for k in ("noisy", "gt", "aux")))
    assert (no.cpu() - rn).abs().max() < 1e-6 and (to_.cpu() - rg).abs().max() < 1e-6
    assert torch.equal(ao.cpu(), ra)


def test_adam_matches_torch_adam():
    ops = _ops()
    torch.manual_seed(9)
    n = 100_003  # exercises the vector tail
    p = torch.randn(n, device=DEV)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.9, 0.999), eps=1e-8)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, device=DEV) * 10 ** random.Random(step).uniform(-4, 0)
        ref.grad = g.clone()
        opt.step()
        ops.adam(p, g, m, v, lr=1e-4, step=step)
    assert rel_err(p, ref.detach()) < 1e-6
    # grad_scale == averaging over ranks
    p2, m2, v2 = torch.ones(8, device=DEV), torch.zeros(8, device=DEV), torch.zeros(8, device=DEV)
    p3, m3, v3 = torch.ones(8, device=DEV), torch.zeros(8, device=DEV), torch.zeros(8, device=DEV)
    g = torch.arange(8, device=DEV, dtype=torch.float32) + 1
    ops.adam(p2, g * 4, m2, v2, lr=1e-3, step=1, grad_scale=0.25)
    ops.adam(p3, g, m3, v3, lr=1e-3, step=1)
    assert torch.allclose(p2, p3)


def test_pack_transpose_roundtrip():
    ops = _ops()
    torch.manual_seed(10)
    w = torch.randn(32, 48, 3, 3, device=DEV)
    fwd = torch.zeros(9, 32, 48, device=DEV)
    ops.pack_weight(w, fwd, ksize=3, Ntot=32, Ktot=48)
    assert torch.equal(fwd, w.permute(2, 3, 0, 1).reshape(9, 32, 48))
    tr = torch.zeros(9, 48, 32, device=DEV)
    ops.pack_weight(w, tr, ksize=3, Ntot=48, Ktot=32, transpose=1)
    assert torch.equal(tr, w.flip(2, 3).permute(2, 3, 1, 0).reshape(9, 48, 32))
    sl = torch.zeros(1, 16, 32, device=DEV)
    w1 = torch.randn(32, 48, 1, 1, device=DEV)
    ops.pack_weight(w1, sl, ksize=1, Ntot=16, Ktot=32, transpose=1, i_begin=32, i_count=16, scale=0.5)
    assert torch.equal(sl[0], 0.5 * w1[:, 32:, 0, 0].t())
    back = torch.empty_like(w)
    ops.unpack_wgrad(back, fwd, ksize=3, Ntot=32, Ktot=48, scale=2.0)
    assert torch.equal(back, 2 * w)


def test_sampler_bit_exact_with_reference_golden():
    ops = _ops()
    g = load_npz("sampler.npz")
    for key, ref in g.items():
        parts = key.split("_")
        h, w, p, n = int(parts[0][1:]), int(parts[1][1:]), int(parts[2][1:]), int(parts[3][1:])
        seeds = torch.tensor([990819], dtype=torch.int64, device=DEV)
        out = ops.sample_patches(seeds, (h, w), p, n)
        assert np.array_equal(out[0].cpu().numpy(), ref), key


def test_sampler_many_images_matches_oracle():
    ops = _ops()
    seeds = torch.tensor([0, 1, 2, 77, 2 ** 33 + 9, 990820], dtype=torch.int64, device=DEV)
    out = ops.sample_patches(seeds, (200, 264), 32, 40).cpu().numpy()
    for i, s in enumerate(seeds.tolist()):
        ref = S.dart_throwing((200, 264), 32, 40, S.MT19937(s))
        assert np.array_equal(out[i], ref), s
    # Poisson-disk property at full size: all accepted points are distinct
    big = ops.sample_patches(torch.tensor([5], dtype=torch.int64, device=DEV), (1024, 1024), 128, 400)[0].cpu().numpy()
    assert len({tuple(p) for p in big}) == 400 and big.min() >= 0 and big.max() <= 1024 - 128 - 1


def test_importance_map_and_sampling_match_reference_golden():
    """pht_importance_map ~ reference map (float, <= 2e-6 abs on a [0,1] map) and pht_importance_sample == the
    reference's kept patch centres bit for bit (dart throwing + prune on one MT19937 stream), from the GPU map."""
    ops = _ops()
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from importance_inputs import CASES, SEEDS, frames
    g = load_npz("importance.npz")
    for name, (h, w, p, n) in CASES.items():
        noisy, normal, aux = frames(h, w)
        # raw frames may hold NaN normals / negative radiance: the kernel applies preprocess_data's cleaning
        raw_aux = aux.copy()
        raw_noisy = noisy.copy()
        noisy_d = torch.from_numpy(np.stack([raw_noisy, raw_noisy])).to(DEV)
        aux_d = torch.from_numpy(np.stack([raw_aux, raw_aux])).to(DEV)
        imp = ops.importance_map(noisy_d, aux_d, p)
        ref = torch.from_numpy(g[f"{name}__imp"]).to(DEV)
        assert float((imp[0] - ref).abs().max()) < 2e-6, name
        assert torch.equal(imp[0], imp[1])
        # (not bit-identical: numpy's float32 power is a SIMD routine that is itself 1-2 ulp off correct rounding)
        assert float((imp[0] - ref).abs().mean()) < 1e-7, name
        for s0, s1 in ((SEEDS[0], SEEDS[1]), (SEEDS[2], SEEDS[3])):
            seeds = torch.tensor([s0, s1], dtype=torch.int64, device=DEV)
            centres, counts = ops.importance_sample(seeds, imp, p, n)
            for k, seed in enumerate((s0, s1)):
                want = g[f"{name}__seed{seed}"]
                assert int(counts[k]) == len(want), (name, seed)
                assert np.array_equal(centres[k, :len(want)].cpu().numpy(), want), (name, seed)
                assert bool((centres[k, len(want):] == -1).all())


def test_importance_map_cleans_nan_normals_and_negative_radiance():
    ops = _ops()
    from oracle import sampler_oracle as S
    rs = np.random.default_rng(7)
    h, w, p = 64, 96, 32
    noisy = rs.standard_normal((h, w, 3)).astype(np.float32)          # negative values: clipped at 0 by preprocess_data
    aux = rs.uniform(-1, 1, (h, w, 7)).astype(np.float32)
    aux[rs.uniform(size=(h, w, 7)) < 0.02] = np.nan
    imp = ops.importance_map(torch.from_numpy(noisy[None]).to(DEV), torch.from_numpy(aux[None]).to(DEV), p)
    ref = S.importance_map(np.clip(np.nan_to_num(noisy), 0, None), np.nan_to_num(aux[..., :3]), p)
    assert float((imp[0].cpu() - torch.from_numpy(ref)).abs().max()) < 2e-6


def test_validation_metrics_match_reference_golden():
    """pht_tonemap_u8 / pht_image_metrics_u8 / pht_mrse against the real reference's tensor2img, calculate_psnr,
    calculate_ssim, calculate_rmse (fixture: tests/golden/make_golden_metrics.py).  The uint8 images may differ by one
    level where powf rounds differently next to a truncation boundary (a handful of pixels); the integer MSE and the fp64
    SSIM are exact given the same images."""
    from pixel_heal_thyself_b200 import metrics as M
    from oracle import metrics_oracle as MO
    g, r = load_npz("metrics.npz"), load_json("metrics.json")
    out_log, gt, noisy_log = (torch.from_numpy(g[k]).to(DEV) for k in ("out_log", "gt", "noisy_log"))
    imgs = {}
    for name, t, spec in (("out_img", out_log, True), ("gt_img", gt, False), ("noisy_img", noisy_log, True)):
        img = M.tensor2img(t, post_spec=spec)
        d = (img.cpu().numpy().astype(int) - g[name].astype(int))
        assert np.abs(d).max() <= 1 and (d != 0).mean() < 1e-3, (name, np.abs(d).max(), (d != 0).mean())
        imgs[name] = img
    # metrics on the REFERENCE's own uint8 images: exact integer MSE, fp64 SSIM
    ref_out, ref_gt, ref_noisy = (torch.from_numpy(g[k]).to(DEV) for k in ("out_img", "gt_img", "noisy_img"))
    psnr, ssim = M.image_metrics(ref_out, ref_gt)
    assert abs(psnr - r["psnr_out"]) < 1e-9 and abs(ssim - r["ssim_out"]) < 1e-9
    psnr_n, ssim_n = M.image_metrics(ref_noisy, ref_gt)
    assert abs(psnr_n - r["psnr_noisy"]) < 1e-9 and abs(ssim_n - r["ssim_noisy"]) < 1e-9
    # end to end (our tone mapping): within what a one-level flip of a few pixels can move
    psnr2, ssim2 = M.image_metrics(imgs["out_img"], imgs["gt_img"])
    assert abs(psnr2 - r["psnr_out"]) < 1e-2 and abs(ssim2 - r["ssim_out"]) < 1e-4
    mrse = M.calculate_rmse(out_log, gt, output_is_log=True)
    assert abs(mrse - r["mrse_out"]) < 1e-5 * r["mrse_out"]
    assert abs(mrse - MO.rmse(np.exp(g["out_log"]) - 1, g["gt"])) < 1e-5 * r["mrse_out"]
    assert M.image_metrics(ref_gt, ref_gt)[0] == 0.0          # mse == 0 -> 0.0 (metric.py:22-23)
    with pytest.raises(ValueError):
        M.image_metrics(ref_out, ref_gt[:, :-1])


def test_attention_backward_accumulation_modes():
    """The tcgen05 attention backward has two ways of summing the <= 4 overlapping 14x14-window contributions of a key
    pixel into dK / dV (option "attn_bwd_direct"): 1 (default) = vector reductions straight into the NHWC gradient, in
    arrival order; 0 = window-major scratch + fold kernel, fixed order.  At a size where every CTA walks several blocks
    (8 x 128 x 128 = 2048 blocks on 148 CTAs), with dK living in a channel slice of a wider buffer (the engine's
    [dQ | dK] layout): the fold mode is bit-reproducible, the direct mode agrees with it to the bf16 rounding of the
    partial sums, and dQ / d rel (which do not go through the accumulation) are bit-identical in both."""
    ops = _ops()
    from pixel_heal_thyself_b200 import _lib
    B, C, H, W = 8, 256, 128, 128
    torch.manual_seed(9)
    mk = lambda s=1.0: (torch.randn(B, H, W, C, device=DEV) * s).to(torch.bfloat16)
    q, k, v, do = mk(0.3), mk(0.3), mk(), mk()
    rel_h, rel_w = torch.randn(1, 14, 1, 32, device=DEV), torch.randn(1, 1, 14, 32, device=DEV)
    out = torch.empty_like(q)
    lse = torch.empty(B, H, W, 4, device=DEV)
    ops.attn_fwd(q, k, v, rel_h, rel_w, out, lse=lse)

    def run():
        ws = torch.empty(max(ops.attn_bwd_workspace_bytes(q), 16) // 4, device=DEV)
        dqk = torch.full((B, H, W, 2 * C), float("nan"), device=DEV, dtype=torch.bfloat16)
        dv = torch.full((B, H, W, C), float("nan"), device=DEV, dtype=torch.bfloat16)
        drh, drw = torch.empty_like(rel_h), torch.empty_like(rel_w)
        ops.attn_bwd(q, k, v, rel_h, rel_w, lse, do, dqk[..., :C], dqk[..., C:], dv, drh, drw, ws)
        torch.cuda.synchronize()
        return dqk, dv, drh, drw

    direct = run()
    assert _lib.lib.pht_set_option(b"attn_bwd_direct", 0) == 0
    try:
        fold_a, fold_b = run(), run()
    finally:
        _lib.lib.pht_set_option(b"attn_bwd_direct", 1)
    for x, y in zip(fold_a, fold_b):
        assert torch.isfinite(x.float()).all() and torch.equal(x, y)
    for x, y in zip(direct, fold_a):
        assert torch.isfinite(x.float()).all() and rel_err(x, y) < 2e-2
    C2 = C
    assert torch.equal(direct[0][..., :C2], fold_a[0][..., :C2])        # dQ
    assert torch.equal(direct[2], fold_a[2]) and torch.equal(direct[3], fold_a[3])   # d rel_h, d rel_w


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (1, 40, 56), (3, 16, 24)])
def test_msssim_loss_matches_oracle(B, H, W):
    """pht_msssim_loss (SSIMLoss, losses.py:248-263 over kornia 0.8.0's MS_SSIMLoss -- parity unpinned, see the oracle's
    header) against oracle.msssim_oracle: loss value and the full gradient w.r.t. the network output; images smaller
    than the 33-tap window (16 x 24) exercise the zero padding on every side."""
    from oracle import msssim_oracle as M
    from pixel_heal_thyself_b200.models.losses import SSIMLoss
    torch.manual_seed(21)
    gt = torch.rand(B, 3, H, W) * 2.5                       # log-radiance-like, some pixels above 1 (scale > 1)
    out = (gt + 0.15 * torch.randn(B, 3, H, W)).requires_grad_(True)
    ref = M.ssim_loss(out.double(), gt.double())
    (gref,) = torch.autograd.grad(ref, out)
    od = out.detach().to(DEV).requires_grad_(True)
    crit = SSIMLoss(window_size=11)
    loss = crit(od, gt.to(DEV))
    (0.1 * loss).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_err(od.grad.cpu(), 0.1 * gref) < 1e-4
    # same images -> zero loss, finite gradient
    z = gt.to(DEV).clone().requires_grad_(True)
    l0 = crit(z, gt.to(DEV))
    l0.backward()
    assert abs(float(l0)) < 1e-4 and torch.isfinite(z.grad).all()


@pytest.mark.parametrize("B,C,H,W", [(8, 64, 32, 32), (4, 128, 16, 24), (8, 512, 4, 4), (2, 256, 8, 8)])
def test_bn_act_forward_backward_and_double_backward_match_torch(B, C, H, W):
    """pht_bn_act_fwd / _bwd / _bwd_bwd (the critic's BatchNorm2d(train) + LeakyReLU(0.2), model.py:52-83, 264-344) against
    stock torch modules: output, running statistics, first-order gradients, and the SECOND-order pass of a
    gradient-penalty-shaped objective (losses.py:12-57: a function of the input gradient, differentiated w.r.t. the input,
    gamma and the upstream gradient)."""
    from pixel_heal_thyself_b200 import ops
    from pixel_heal_thyself_b200.models.afgsa.discriminator import _BNActFn
    torch.manual_seed(31)
    x0 = (torch.randn(B, C, H, W, device=DEV) * 1.5 + 0.3).contiguous(memory_format=torch.channels_last)
    wgt = torch.randn(B, C, H, W, device=DEV).contiguous(memory_format=torch.channels_last)   # stands in for the layers above
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    act = torch.nn.LeakyReLU(0.2)
    ws = ops.bn_act_ws(C, DEV)

    def gp_objective(f, x, gamma, beta):
        z = f(x, gamma, beta)
        (gx,) = torch.autograd.grad((z * wgt).sum(), x, create_graph=True)
        pen = ((gx.reshape(B, -1).norm(2, dim=1) - 1) ** 2).mean()
        return z, gx, pen

    # reference: stock modules (fresh running stats)
    xr = x0.clone().requires_grad_(True)
    gr, br = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    cb = torch.randn(C, device=DEV) * 0.5     # bias of the convolution in front: added by the reference, folded by ours
    f_ref = lambda x, g, b: act(torch.nn.functional.batch_norm(x + cb.view(1, C, 1, 1), rm_r, rv_r, g, b, True, 0.1, 1e-5))
    z_r, gx_r, pen_r = gp_objective(f_ref, xr, gr, br)
    d_r = torch.autograd.grad(pen_r + 0.01 * z_r.pow(2).mean(), [xr, gr, br])
    # ours
    xo = x0.clone().requires_grad_(True)
    go, bo = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    rm_o, rv_o = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    f_our = lambda x, g, b: _BNActFn.apply(x, g, b, cb, rm_o, rv_o, 1e-5, 0.1, 0.2, ws)
    z_o, gx_o, pen_o = gp_objective(f_our, xo, go, bo)
    d_o = torch.autograd.grad(pen_o + 0.01 * z_o.pow(2).mean(), [xo, go, bo])
    assert rel_err(z_o, z_r) < 2e-6
    assert rel_err(rm_o, rm_r) < 1e-5 and rel_err(rv_o, rv_r) < 1e-5
    assert rel_err(gx_o, gx_r) < 2e-6
    assert abs(float(pen_o) - float(pen_r)) < 1e-4 * max(1.0, abs(float(pen_r)))
    print(f"bn_act {B}x{C}x{H}x{W}: z {rel_err(z_o, z_r):.1e} gx {rel_err(gx_o, gx_r):.1e} "
          + " ".join(f"{n} {rel_err(a, b):.1e}" for a, b, n in zip(d_o, d_r, ("d/dx", "d/dgamma", "d/dbeta"))))
    for a, b, name in zip(d_o, d_r, ("d/dx", "d/dgamma", "d/dbeta")):
        assert rel_err(a, b) < 1e-5, (name, rel_err(a, b))      # measured 1e-7 .. 5e-7


def test_critic_with_fused_bn_act_matches_stock_modules():
    """DiscriminatorVGG at the prod shape with the hand-written BatchNorm + LeakyReLU kernels == the same critic on stock
    torch modules: WGAN-GP critic loss (base_trainer.py:391-406) and every parameter gradient, same weights, same
    interpolation coefficients."""
    from pixel_heal_thyself_b200.models.afgsa.discriminator import DiscriminatorVGG
    from pixel_heal_thyself_b200.models.losses import GANLoss, GradientPenaltyLoss
    torch.manual_seed(5)
    D = DiscriminatorVGG(3, 64, 64).to(DEV)
    real, fake = torch.rand(4, 3, 64, 64, device=DEV), torch.rand(4, 3, 64, 64, device=DEV)
    gan, gp = GANLoss("wgan").to(DEV), GradientPenaltyLoss(torch.device(DEV))
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    res = {}
    for fused in (False, True):
        D.load_state_dict(sd)
        D.fused_bn_act = fused
        D.zero_grad(set_to_none=True)
        torch.manual_seed(77)                                   # the gradient penalty draws its alphas from the global RNG
        loss = (gan(D(fake), False) + gan(D(real), True)) / 2 + 10.0 * gp(D, real, fake)
        loss.backward()
        res[fused] = (float(loss), {n: p.grad.clone() for n, p in D.named_parameters()},
                      {n: b.clone() for n, b in D.named_buffers()})
    assert abs(res[True][0] - res[False][0]) < 1e-4 * max(1.0, abs(res[False][0])), (res[True][0], res[False][0])
    # (a conv bias in front of a BatchNorm has a mathematically zero gradient -- the batch mean removes it -- so both sides
    # hold only round-off there: errors are measured against the larger of the tensor's and the global gradient scale)
    # The kernels themselves agree with torch to fp32 rounding (5e-7, test above).  At the critic level the penalty
    # ((|grad| - 1)^2 with |grad| close to 1) cancels catastrophically: a 1e-6 relative difference of the input gradient
    # (last-bit differences, cuDNN's choice of algorithm with / without a bias operand) becomes a UNIFORM ~5e-4 relative
    # difference of every penalty gradient (tools/diag_critic.py: l2 5e-4 on all tensors, worst single element 5.8e-3 of
    # its tensor's max; two stock runs differ by 1e-6 .. 2e-5).
    gmax = max(float(g.abs().max()) for g in res[False][1].values())
    for n, g in res[False][1].items():
        d = res[True][1][n] - g
        err = float(d.abs().max()) / max(float(g.abs().max()), 1e-4 * gmax)
        l2 = float(d.norm()) / max(float(g.norm()), 1e-4 * gmax * g.numel() ** 0.5)
        assert err < 2e-2 and l2 < 5e-3, (n, err, l2)
    for n, b in res[False][2].items():
        assert rel_err(res[True][2][n].float(), b.float()) < 1e-4, n
