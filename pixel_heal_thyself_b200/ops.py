"""Tensor-level wrappers over the C ABI (one function per include/pht_b200.h op).

All tensors are CUDA tensors; activations are channels-last [B, H, W, C]
(possibly strided views of padded buffers).  Nothing here computes on the CPU
or through torch ops: each function marshals pointers and launches kernels on
the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import lib

# optional per-launch CUDA-event timing of conv_gemm (bench.py's roofline leg): a list that receives
# (tag, start_event, end_event) for every launch whose tag passes `_profile_filter`
_profile_sink = None
_profile_filter = None


def set_launch_profiler(sink, flt=None):
    global _profile_sink, _profile_filter
    _profile_sink, _profile_filter = sink, flt


def conv_gemm(srcs, w, N, *, ksize=1, out_domain=None, bias=None, slope=None, resid=None, resid_mode=None, mask=None,
              mslope=None, out1=None, out2=None, src_offsets=None, padfold=False, ring1=None, ring2=None):
    """pht_conv_gemm.  srcs: list of [B,H,W,C] tensors (virtual concat along C).
    out_domain: (B, Ho, Wo), default = shape of the first output.
    src_offsets: list of (oy, ox) per source (default 0,0).
    resid_mode: None | "pre" | "post".
    padfold: out_domain is the padded domain, resid/mask/outputs are interior views (PHT_EPI_PADFOLD).
    ring1 / ring2: None | "replicate" | "reflect": out1 / out2 is the interior view of a padded buffer whose 1-pixel frame
    is written too (PHT_EPI_RING1 / RING2; replaces a pht_border_fill launch)."""
    a = L.ConvGemmArgs()
    ref = out1 if out1 is not None else out2
    L.require_cuda(*srcs, w, ref)
    if out_domain is None:
        out_domain = (ref.shape[0], ref.shape[1], ref.shape[2])
    a.dtype = L.DTYPES[srcs[0].dtype]
    a.B, a.Ho, a.Wo = out_domain
    a.N, a.ksize, a.n_src = N, ksize, len(srcs)
    flags = 0
    for i, s in enumerate(srcs):
        oy, ox = src_offsets[i] if src_offsets else (0, 0)
        a.src[i] = L.view(s, oy, ox)
    a.w = w.data_ptr()
    a.bias, a.slope, a.mslope = L.ptr(bias), L.ptr(slope), L.ptr(mslope)
    if resid is not None:
        a.resid = L.view(resid)
        flags |= L.EPI_RESID_PRE if resid_mode == "pre" else L.EPI_RESID_POST
    if mask is not None:
        a.mask = L.view(mask)
        flags |= L.EPI_MASK
    if padfold:
        flags |= L.EPI_PADFOLD
    for ring, bit in ((ring1, L.EPI_RING1), (ring2, L.EPI_RING2)):
        if ring is not None:
            flags |= bit | (L.EPI_RING_REFLECT if ring == "reflect" else 0)
    a.flags = flags
    a.out1, a.out2 = L.view(out1), L.view(out2)
    if _profile_sink is not None:
        tag = (ksize, N, sum(s.shape[-1] for s in srcs), a.B * a.Ho * a.Wo)
        if _profile_filter is None or _profile_filter(tag):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.pht_conv_gemm(C.byref(a), L.stream_ptr()), "pht_conv_gemm")
            e1.record()
            _profile_sink.append((tag, e0, e1))
            return
    L.check(lib.pht_conv_gemm(C.byref(a), L.stream_ptr()), "pht_conv_gemm")


def _wgrad_args(dy, srcs, ksize, dw, dbias, src_offsets):
    a = L.WgradArgs()
    a.dtype = L.DTYPES[dy.dtype]
    a.B, a.Ho, a.Wo, a.N = dy.shape
    a.ksize, a.n_src = ksize, len(srcs)
    a.dy = L.view(dy)
    for i, s in enumerate(srcs):
        oy, ox = src_offsets[i] if src_offsets else (0, 0)
        a.src[i] = L.view(s, oy, ox)
    a.dw, a.dbias = L.ptr(dw), L.ptr(dbias)
    return a


def wgrad_workspace_bytes(dy, srcs, ksize=1, src_offsets=None) -> int:
    a = _wgrad_args(dy, srcs, ksize, None, None, src_offsets)
    return int(lib.pht_wgrad_workspace_bytes(C.byref(a)))


def wgrad(dy, srcs, dw, *, ksize=1, dbias=None, workspace=None, src_offsets=None):
    """pht_wgrad: dw fp32 [ksize*ksize][N][Ktot] (overwritten), dbias fp32 [N]."""
    L.require_cuda(dy, *srcs, dw)
    a = _wgrad_args(dy, srcs, ksize, dw, dbias, src_offsets)
    if workspace is not None:
        a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    L.check(lib.pht_wgrad(C.byref(a), L.stream_ptr()), "pht_wgrad")


class WgradBucket:
    """Deferred weight-gradients of one gradient bucket (decoder / one transformer block / encoders).

    ``wgrad()`` launches only the split GEMM (``pht_wgrad_partial``) into this bucket's private workspace;
    ``unpack()`` queues a packed -> OIHW scatter; ``flush()`` then finishes every pending reduction in ONE launch
    (``pht_wgrad_reduce_batched``) and every scatter in one more (``pht_unpack_wgrads_batched``).  The descriptor
    tables are cached on the device by content and only uploaded the first time a job list is seen."""

    def __init__(self, device, workspace_bytes=320 << 20):
        self.ws = torch.empty(workspace_bytes // 4, dtype=torch.float32, device=device)
        self.cursor = 0
        self.jobs, self.unpacks = [], []
        self._tables = {}

    def wgrad(self, dy, srcs, dw, *, ksize=1, dbias=None, src_offsets=None):
        L.require_cuda(dy, *srcs, dw)
        a = _wgrad_args(dy, srcs, ksize, dw, dbias, src_offsets)
        need = int(lib.pht_wgrad_workspace_bytes(C.byref(a)))
        need = (need + 255) // 256 * 256
        if self.cursor * 4 + need > self.ws.numel() * 4:
            raise RuntimeError("WgradBucket: workspace exhausted")
        a.workspace, a.workspace_bytes = self.ws.data_ptr() + self.cursor * 4, need
        job = L.WgradReduceJob()
        rc = lib.pht_wgrad_partial(C.byref(a), C.byref(job), L.stream_ptr())
        if rc == -3:     # PHT_ERR_UNSUPPORTED: not a split tensor-core shape -> immediate path
            L.check(lib.pht_wgrad(C.byref(a), L.stream_ptr()), "pht_wgrad")
            return
        L.check(rc, "pht_wgrad_partial")
        self.cursor += need // 4
        self.jobs.append(job)

    def unpack(self, w_grad, packed, **kw):
        L.require_cuda(w_grad, packed)
        assert w_grad.is_contiguous() and w_grad.dtype == torch.float32 and packed.dtype == torch.float32
        self.unpacks.append(_pack_args(w_grad, packed, **kw))

    def _table(self, kind, items, ctype, key, nbytes):
        """Descriptor table of this exact job list on the device: (host array, device table, upload?).  A step flushes
        several different buckets (and the gradient arenas alternate between steps), so tables are cached by content."""
        sig = (kind,) + tuple(key(x) for x in items)
        hit = self._tables.get(sig)
        if hit is not None:
            return hit[0], hit[1], 0
        if len(self._tables) > 256:
            self._tables.clear()
        arr = (ctype * len(items))(*items)
        tab = torch.empty(nbytes + 64, dtype=torch.uint8, device=self.ws.device)
        self._tables[sig] = (arr, tab)
        return arr, tab, 1

    def flush(self):
        if self.jobs:
            arr, tab, up = self._table(
                "r", self.jobs, L.WgradReduceJob,
                lambda j: (j.partials, j.dw, j.bias_partials, j.dbias, j.elems, j.splits, j.bias_rows, j.N),
                C.sizeof(L.WgradReduceJob) * len(self.jobs))
            L.check(lib.pht_wgrad_reduce_batched(arr, len(self.jobs), tab.data_ptr(), tab.numel(), up,
                                                 L.stream_ptr()), "pht_wgrad_reduce_batched")
        if self.unpacks:
            arr, tab, up = self._table(
                "u", self.unpacks, L.PackArgs,
                lambda a: (a.w, a.packed, a.O, a.I, a.ksize, a.Ntot, a.Ktot, a.n_off, a.k_off, a.transpose, a.grid,
                           a.i_begin, a.i_count, a.scale), int(lib.pht_pack_table_bytes(len(self.unpacks))))
            L.check(lib.pht_unpack_wgrads_batched(arr, len(self.unpacks), tab.data_ptr(), tab.numel(),
                                                  up, L.stream_ptr()), "pht_unpack_wgrads_batched")
        self.jobs, self.unpacks, self.cursor = [], [], 0


def border_fill(buf, mode):
    """buf: padded [B, H+2, W+2, C] contiguous."""
    L.require_cuda(buf)
    assert buf.is_contiguous()
    B, Hp, Wp, Cc = buf.shape
    L.check(lib.pht_border_fill(buf.data_ptr(), L.DTYPES[buf.dtype], B, Hp - 2, Wp - 2, Cc, mode, L.stream_ptr()),
            "pht_border_fill")


def pad_fold(gpad, mode, *, resid=None, mask=None, mslope=None, out1=None, out2=None):
    L.require_cuda(gpad)
    assert gpad.is_contiguous()
    B, Hp, Wp, Cc = gpad.shape
    vs = [L.view(t) if t is not None else None for t in (resid, mask, out1, out2)]
    refs = [C.byref(v) if v is not None else None for v in vs]
    L.check(lib.pht_pad_fold(gpad.data_ptr(), L.DTYPES[gpad.dtype], B, Hp - 2, Wp - 2, Cc, mode, refs[0], refs[1],
                             L.ptr(mslope), refs[2], refs[3], L.stream_ptr()), "pht_pad_fold")


def film_fwd(gb, x, out):
    """FiLM modulation (film.py:36-45): out = gb[..., :C] * x + gb[..., C:]; gb [B,H,W,2C], x / out [B,H,W,C] views."""
    L.require_cuda(gb, x, out)
    vg, vx, vo = L.view(gb), L.view(x), L.view(out)
    L.check(lib.pht_film_fwd(C.byref(vg), C.byref(vx), C.byref(vo), x.shape[0], x.shape[3], L.stream_ptr()), "pht_film_fwd")


def film_bwd(gb, x, dout, dgb, dx_in=None, dx=None):
    """dgb = [dout * x | dout]; dx = (dx_in or 0) + gamma * dout (dx may alias dx_in)."""
    L.require_cuda(gb, x, dout, dgb)
    vs = [L.view(t) for t in (gb, x, dout, dgb)]
    vi = L.view(dx_in) if dx_in is not None else None
    vd = L.view(dx) if dx is not None else None
    L.check(lib.pht_film_bwd(C.byref(vs[0]), C.byref(vs[1]), C.byref(vs[2]), C.byref(vs[3]),
                             C.byref(vi) if vi is not None else None, C.byref(vd) if vd is not None else None,
                             x.shape[0], x.shape[3], L.stream_ptr()), "pht_film_bwd")


def im2col5(x_nchw, col, mode):
    """x_nchw fp32 [B,Cin,H,W] -> col [B,H,W,Kpad] (dtype of col)."""
    L.require_cuda(x_nchw, col)
    assert x_nchw.is_contiguous() and col.is_contiguous() and x_nchw.dtype == torch.float32
    B, Cin, H, W = x_nchw.shape
    L.check(lib.pht_im2col5(x_nchw.data_ptr(), col.data_ptr(), L.DTYPES[col.dtype], B, Cin, H, W, col.shape[-1], mode,
                            L.stream_ptr()), "pht_im2col5")


def _attn_args(q, k, v, rel_h, rel_w, heads, block, halo, resid=None, out=None, lse=None):
    a = L.AttnArgs()
    a.dtype = L.DTYPES[q.dtype]
    a.B, a.H, a.W = q.shape[0], q.shape[1], q.shape[2]
    a.heads, a.head_dim, a.block, a.halo = heads, q.shape[3] // heads, block, halo
    a.q, a.k, a.v = L.view(q), L.view(k), L.view(v)
    a.rel_h, a.rel_w = rel_h.data_ptr(), rel_w.data_ptr()
    a.resid, a.out = L.view(resid), L.view(out)
    a.lse = L.ptr(lse)
    return a


def attn_fwd(q, k, v, rel_h, rel_w, out, *, heads=4, block=8, halo=3, resid=None, lse=None, ring=None):
    """ring: None | "replicate" | "reflect": ``out`` is the interior view of a padded buffer whose 1-pixel frame is written
    too (replaces a pht_border_fill launch; tcgen05 path only)."""
    L.require_cuda(q, k, v, rel_h, rel_w, out)
    a = _attn_args(q, k, v, rel_h, rel_w, heads, block, halo, resid, out, lse)
    a.ring = {None: 0, "replicate": 1, "reflect": 2}[ring]
    L.check(lib.pht_attn_fwd(C.byref(a), L.stream_ptr()), "pht_attn_fwd")


def attn_bwd_workspace_bytes(q, heads=4, block=8, halo=3) -> int:
    a = L.AttnBwdArgs()
    a.fwd.dtype = L.DTYPES[q.dtype]
    a.fwd.B, a.fwd.H, a.fwd.W = q.shape[0], q.shape[1], q.shape[2]
    a.fwd.heads, a.fwd.head_dim, a.fwd.block, a.fwd.halo = heads, q.shape[3] // heads, block, halo
    return int(lib.pht_attn_bwd_workspace_bytes(C.byref(a)))


def _attn_bwd_args(q, k, v, rel_h, rel_w, lse, d_out, dq, dk, dv, d_rel_h, d_rel_w, workspace, heads, block, halo):
    a = L.AttnBwdArgs()
    a.fwd = _attn_args(q, k, v, rel_h, rel_w, heads, block, halo, None, None, lse)
    a.d_out, a.dq = L.view(d_out), L.view(dq)
    a.dk, a.dv = L.view(dk), L.view(dv)
    a.d_rel_h, a.d_rel_w = d_rel_h.data_ptr(), d_rel_w.data_ptr()
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    return a


def attn_bwd_zero(q, k, v, rel_h, rel_w, lse, d_out, dq, dk, dv, d_rel_h, d_rel_w, workspace, *, heads=4, block=8, halo=3):
    """pht_attn_bwd_zero: zero dk / dv (and the ordering flags in the workspace) ahead of ``attn_bwd(..., prezeroed=True)``
    with the same arguments; may run on another stream, overlapped with earlier kernels (nothing else is touched)."""
    L.require_cuda(q, k, v, d_out, dq, dk, dv)
    a = _attn_bwd_args(q, k, v, rel_h, rel_w, lse, d_out, dq, dk, dv, d_rel_h, d_rel_w, workspace, heads, block, halo)
    L.check(lib.pht_attn_bwd_zero(C.byref(a), L.stream_ptr()), "pht_attn_bwd_zero")


def attn_bwd(q, k, v, rel_h, rel_w, lse, d_out, dq, dk, dv, d_rel_h, d_rel_w, workspace, *, heads=4, block=8,
             halo=3, prezeroed=False):
    """dq / dk / dv: [B,H,W,C] views (activation dtype), overwritten with the final gradients."""
    L.require_cuda(q, k, v, d_out, dq, dk, dv)
    a = _attn_bwd_args(q, k, v, rel_h, rel_w, lse, d_out, dq, dk, dv, d_rel_h, d_rel_w, workspace, heads, block, halo)
    a.prezeroed = 1 if prezeroed else 0
    L.check(lib.pht_attn_bwd(C.byref(a), L.stream_ptr()), "pht_attn_bwd")


def dec_tail_fwd(h, w, bias, x_nchw, out_nchw):
    L.require_cuda(h, w, bias, x_nchw, out_nchw)
    B, H, W, _ = h.shape
    hv = L.view(h)
    L.check(lib.pht_dec_tail_fwd(C.byref(hv), w.data_ptr(), bias.data_ptr(), x_nchw.data_ptr(), out_nchw.data_ptr(),
                                 B, H, W, L.stream_ptr()), "pht_dec_tail_fwd")


def dec_tail_bwd_data(dout_nchw, w, h, dh_pre):
    L.require_cuda(dout_nchw, w, h, dh_pre)
    B, H, W, _ = h.shape
    hv, dv = L.view(h), L.view(dh_pre)
    L.check(lib.pht_dec_tail_bwd_data(dout_nchw.data_ptr(), w.data_ptr(), C.byref(hv), C.byref(dv), B, H, W,
                                      L.stream_ptr()), "pht_dec_tail_bwd_data")


def dec_tail_ws_bytes(B, H, W, Cc) -> int:
    return int(lib.pht_dec_tail_ws_bytes(B, H, W, Cc))


def dec_tail_bwd_weight(dout_nchw, h, dw, dbias, workspace):
    L.require_cuda(dout_nchw, h, dw, dbias, workspace)
    B, H, W, _ = h.shape
    hv = L.view(h)
    L.check(lib.pht_dec_tail_bwd_weight(dout_nchw.data_ptr(), C.byref(hv), dw.data_ptr(), dbias.data_ptr(),
                                        workspace.data_ptr(), workspace.numel() * workspace.element_size(), B, H, W,
                                        L.stream_ptr()), "pht_dec_tail_bwd_weight")


def l1_loss(a, b, loss, grad=None, grad_scale=1.0):
    L.require_cuda(a, b, loss)
    assert a.is_contiguous() and b.is_contiguous() and a.dtype == torch.float32 and b.dtype == torch.float32
    L.check(lib.pht_l1_loss(a.data_ptr(), b.data_ptr(), a.numel(), float(grad_scale), loss.data_ptr(), L.ptr(grad),
                            L.stream_ptr()), "pht_l1_loss")


def _nhwc_rows(t):
    """logical NCHW channels-last fp32 tensor -> (data_ptr, m, C)"""
    assert t.dim() == 4 and t.dtype == torch.float32 and t.permute(0, 2, 3, 1).is_contiguous(), "need channels-last fp32 NCHW"
    return t.data_ptr(), t.shape[0] * t.shape[2] * t.shape[3], t.shape[1]


def bn_act_ws(C, device):
    """workspace of the pht_bn_act_* / pht_colsum_f32 calls; must be ZERO on first use (its last word is the
    "last block done" ticket, which every call leaves at zero again)"""
    return torch.zeros((int(lib.pht_bn_act_ws_bytes(C)) + 3) // 4, dtype=torch.float32, device=device)


def colsum_nhwc(x, out, ws):
    """out[c] = sum over (b, y, x) of a channels-last fp32 NCHW tensor (pht_colsum_f32)."""
    L.require_cuda(x, out, ws)
    xp, m, Cc = _nhwc_rows(x)
    L.check(lib.pht_colsum_f32(xp, out.data_ptr(), m, Cc, ws.data_ptr(), ws.numel() * 4, L.stream_ptr()), "pht_colsum_f32")


def bn_act_fwd(x, gamma, beta, run_mean, run_var, stat, z, ws, *, eps=1e-5, momentum=0.1, slope=0.2, pre_bias=None):
    L.require_cuda(x, gamma, beta, stat, z, ws)
    xp, m, Cc = _nhwc_rows(x)
    zp, _, _ = _nhwc_rows(z)
    L.check(lib.pht_bn_act_fwd(xp, gamma.data_ptr(), beta.data_ptr(), L.ptr(pre_bias), L.ptr(run_mean), L.ptr(run_var),
                               stat.data_ptr(), zp, m, Cc, eps, momentum, slope, ws.data_ptr(), ws.numel() * 4, L.stream_ptr()),
            "pht_bn_act_fwd")


def bn_act_bwd(x, gz, gamma, beta, stat, gx, g_gamma, g_beta, ws, *, slope=0.2):
    L.require_cuda(x, gz, gamma, beta, stat, gx, ws)
    xp, m, Cc = _nhwc_rows(x)
    L.check(lib.pht_bn_act_bwd(xp, _nhwc_rows(gz)[0], gamma.data_ptr(), beta.data_ptr(), stat.data_ptr(), _nhwc_rows(gx)[0],
                               L.ptr(g_gamma), L.ptr(g_beta), m, Cc, slope, ws.data_ptr(), ws.numel() * 4, L.stream_ptr()),
            "pht_bn_act_bwd")


def bn_act_bwd_bwd(x, gz, h, gamma, beta, stat, h_gz, h_x, h_gamma, ws, *, slope=0.2):
    L.require_cuda(x, gz, h, gamma, beta, stat, h_gz, h_x, ws)
    xp, m, Cc = _nhwc_rows(x)
    L.check(lib.pht_bn_act_bwd_bwd(xp, _nhwc_rows(gz)[0], _nhwc_rows(h)[0], gamma.data_ptr(), beta.data_ptr(), stat.data_ptr(),
                                   _nhwc_rows(h_gz)[0], _nhwc_rows(h_x)[0], L.ptr(h_gamma), m, Cc, slope, ws.data_ptr(),
                                   ws.numel() * 4, L.stream_ptr()), "pht_bn_act_bwd_bwd")


def msssim_ws_bytes(B, H, W) -> int:
    return int(lib.pht_msssim_ws_bytes(B, H, W))


def msssim_loss(out, gt, loss, grad, workspace, grad_scale=1.0):
    """pht_msssim_loss: out / gt fp32 NCHW [B,3,H,W]; loss fp32 [1]; grad fp32 like out or None."""
    L.require_cuda(out, gt, loss, workspace)
    assert out.is_contiguous() and gt.is_contiguous() and out.dtype == torch.float32 and gt.dtype == torch.float32
    B, Cc, H, W = out.shape
    if Cc != 3:
        raise ValueError("msssim_loss: 3-channel images only (the reference's grouped window bank is built for RGB)")
    L.check(lib.pht_msssim_loss(out.data_ptr(), gt.data_ptr(), B, H, W, float(grad_scale), loss.data_ptr(), L.ptr(grad),
                                workspace.data_ptr(), workspace.numel() * workspace.element_size(), L.stream_ptr()),
            "pht_msssim_loss")


def preprocess(noisy_nhwc, gt_nhwc, aux_nhwc, noisy_out, gt_out, aux_out):
    L.require_cuda(noisy_nhwc, aux_nhwc, noisy_out, aux_out)
    B, H, W, _ = noisy_nhwc.shape
    L.check(lib.pht_preprocess(noisy_nhwc.data_ptr(), L.ptr(gt_nhwc), aux_nhwc.data_ptr(), noisy_out.data_ptr(),
                               L.ptr(gt_out), aux_out.data_ptr(), B, H, W, L.stream_ptr()), "pht_preprocess")


def crop_preprocess(noisy_f, gt_f, aux_f, centres, P, noisy_out, gt_out, aux_out, img_idx=None):
    """frames [n_img, Hf, Wf, C] (or [Hf, Wf, C]) fp32 resident on the device; centres int32 [n, 2] = (x, y)."""
    L.require_cuda(noisy_f, aux_f, centres, noisy_out, aux_out)
    Hf, Wf = noisy_f.shape[-3], noisy_f.shape[-2]
    assert centres.dtype == torch.int32 and centres.is_contiguous()
    assert img_idx is None or (img_idx.dtype == torch.int32 and img_idx.is_contiguous())
    L.check(lib.pht_crop_preprocess(noisy_f.data_ptr(), L.ptr(gt_f), aux_f.data_ptr(), Hf, Wf, centres.data_ptr(),
                                    L.ptr(img_idx), centres.shape[0], P, noisy_out.data_ptr(), L.ptr(gt_out),
                                    aux_out.data_ptr(), L.stream_ptr()), "pht_crop_preprocess")


def adam(p, g, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, step=1, grad_scale=1.0):
    L.require_cuda(p, g, m, v)
    L.check(lib.pht_adam(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps,
                         step, grad_scale, L.stream_ptr()), "pht_adam")


def adam_dev(p, g, m, v, hyper, *, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """pht_adam_dev: hyper = fp32 [4] device tensor {lr, step count (int32 bits), -, -}; CUDA-graph replayable."""
    L.require_cuda(p, g, m, v, hyper)
    L.check(lib.pht_adam_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), beta1, beta2, eps, grad_scale,
                             hyper.data_ptr(), L.stream_ptr()), "pht_adam_dev")


def _pack_args(w, packed, *, ksize, Ntot, Ktot, n_off=0, k_off=0, transpose=0, grid=0, i_begin=0, i_count=0,
               scale=1.0):
    a = L.PackArgs()
    a.w, a.packed = w.data_ptr(), packed.data_ptr()
    a.dtype = L.DTYPES[packed.dtype]
    a.O, a.I, a.ksize = w.shape[0], w.shape[1], ksize
    a.Ntot, a.Ktot, a.n_off, a.k_off = Ntot, Ktot, n_off, k_off
    a.transpose, a.grid, a.i_begin, a.i_count, a.scale = transpose, grid, i_begin, i_count, scale
    return a


def pack_weight(w, packed, **kw):
    L.require_cuda(w, packed)
    assert w.is_contiguous() and w.dtype == torch.float32
    a = _pack_args(w, packed, **kw)
    L.check(lib.pht_pack_weight(C.byref(a), L.stream_ptr()), "pht_pack_weight")


class PackPlan:
    """A cached batch of weight-pack descriptors executed by ONE kernel launch (pht_pack_weights_batched)."""

    def __init__(self):
        self.jobs = []          # (w, packed, kwargs)
        self._arr = None
        self._table = None
        self._sig = None

    def add(self, w, packed, **kw):
        self.jobs.append((w, packed, kw))

    def run(self):
        if not self.jobs:
            return
        sig = tuple((w.data_ptr(), p.data_ptr(), p.dtype) for w, p, _ in self.jobs)
        upload = 0
        if sig != self._sig:
            n = len(self.jobs)
            arr = (L.PackArgs * n)()
            for i, (w, p, kw) in enumerate(self.jobs):
                L.require_cuda(w, p)
                assert w.is_contiguous() and w.dtype == torch.float32
                arr[i] = _pack_args(w, p, **kw)
            nbytes = int(lib.pht_pack_table_bytes(n))
            self._table = torch.empty(nbytes, dtype=torch.uint8, device=self.jobs[0][0].device)
            self._arr, self._sig, upload = arr, sig, 1
        L.check(lib.pht_pack_weights_batched(self._arr, len(self.jobs), self._table.data_ptr(), self._table.numel(), upload,
                                             L.stream_ptr()), "pht_pack_weights_batched")


def tail_finish(y, bias, x_nchw, out_nchw):
    """y fp32 [B,H,W,ldy] (channels 0..2 = conv result) -> out_nchw = y[..., :3] + bias + x_nchw."""
    L.require_cuda(y, bias, x_nchw, out_nchw)
    B, H, W, ldy = y.shape
    assert y.is_contiguous() and y.dtype == torch.float32
    L.check(lib.pht_tail_finish(y.data_ptr(), ldy, bias.data_ptr(), x_nchw.data_ptr(), out_nchw.data_ptr(), B, H, W,
                                L.stream_ptr()), "pht_tail_finish")


def tail_gather(y, bias, x_nchw, out_nchw):
    """y fp32 [B,H,W,ldy] with y[p][t*3+co] (1x1 GEMM of the decoder tail) -> out_nchw = 9-tap gather + bias + x_nchw."""
    L.require_cuda(y, bias, x_nchw, out_nchw)
    B, H, W, ldy = y.shape
    assert y.is_contiguous() and y.dtype == torch.float32
    L.check(lib.pht_tail_gather(y.data_ptr(), ldy, bias.data_ptr(), x_nchw.data_ptr(), out_nchw.data_ptr(), B, H, W,
                                L.stream_ptr()), "pht_tail_gather")


def tail_im2col_bwd(dout_nchw, a, dbias):
    """dout fp32 [B,3,H,W] -> a bf16 [B,H,W,64] (27 shifted copies), dbias fp32 [3]."""
    L.require_cuda(dout_nchw, a, dbias)
    B, _, H, W = dout_nchw.shape
    assert a.is_contiguous() and a.dtype == torch.bfloat16 and a.shape[-1] == 64
    L.check(lib.pht_tail_im2col_bwd(dout_nchw.data_ptr(), a.data_ptr(), dbias.data_ptr(), B, H, W, L.stream_ptr()),
            "pht_tail_im2col_bwd")


def unpack_wgrad(w_grad, packed, **kw):
    L.require_cuda(w_grad, packed)
    assert w_grad.is_contiguous() and w_grad.dtype == torch.float32 and packed.dtype == torch.float32
    a = _pack_args(w_grad, packed, **kw)
    L.check(lib.pht_unpack_wgrad(C.byref(a), L.stream_ptr()), "pht_unpack_wgrad")


def cast2d(src, dst):
    """src/dst: 2-D [rows, cols] with unit column stride (row stride free)."""
    L.require_cuda(src, dst)
    assert src.dim() == 2 and dst.shape == src.shape and src.stride(1) == 1 and dst.stride(1) == 1
    L.check(lib.pht_cast2d(src.data_ptr(), L.DTYPES[src.dtype], src.stride(0), dst.data_ptr(), L.DTYPES[dst.dtype],
                           dst.stride(0), src.shape[0], src.shape[1], L.stream_ptr()), "pht_cast2d")


def sample_patches(seeds, frame_hw, P, n, max_iter=5000):
    """seeds int64 [n_img] (CUDA) -> int32 [n_img, n, 2] top-left corners (x, y)."""
    L.require_cuda(seeds)
    assert seeds.dtype == torch.int64
    out = torch.empty(seeds.numel(), n, 2, dtype=torch.int32, device=seeds.device)
    L.check(lib.pht_sample_patches(seeds.data_ptr(), seeds.numel(), frame_hw[0], frame_hw[1], P, n, max_iter,
                                   out.data_ptr(), L.stream_ptr()), "pht_sample_patches")
    return out


def importance_map(noisy_f, aux_f, P):
    """frames [n_img, Hf, Wf, 3] / [n_img, Hf, Wf, 7] fp32 (raw, HBM-resident) -> importance map fp32 [n_img, Hf, Wf]."""
    L.require_cuda(noisy_f, aux_f)
    assert noisy_f.is_contiguous() and aux_f.is_contiguous() and noisy_f.dtype == torch.float32 and aux_f.dtype == torch.float32
    n_img, Hf, Wf, _ = noisy_f.shape
    imp = torch.empty(n_img, Hf, Wf, dtype=torch.float32, device=noisy_f.device)
    nbytes = int(lib.pht_importance_map_ws_bytes(n_img, Hf, Wf))
    ws = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=noisy_f.device)
    L.check(lib.pht_importance_map(noisy_f.data_ptr(), aux_f.data_ptr(), n_img, Hf, Wf, P, imp.data_ptr(), ws.data_ptr(),
                                   ws.numel() * 4, L.stream_ptr()), "pht_importance_map")
    return imp


def importance_sample(seeds, imp, P, n, max_iter=5000):
    """seeds int64 [n_img] (CUDA), imp fp32 [n_img, Hf, Wf] -> (centres int32 [n_img, n, 2] (-1 padded), counts int32 [n_img])."""
    L.require_cuda(seeds, imp)
    assert seeds.dtype == torch.int64 and imp.dtype == torch.float32 and imp.is_contiguous()
    n_img, Hf, Wf = imp.shape
    assert seeds.numel() == n_img
    out = torch.empty(n_img, n, 2, dtype=torch.int32, device=seeds.device)
    counts = torch.empty(n_img, dtype=torch.int32, device=seeds.device)
    L.check(lib.pht_importance_sample(seeds.data_ptr(), n_img, Hf, Wf, P, n, max_iter, imp.data_ptr(), out.data_ptr(),
                                      counts.data_ptr(), L.stream_ptr()), "pht_importance_sample")
    return out, counts
