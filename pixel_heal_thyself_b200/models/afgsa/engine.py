"""Kernel orchestration for the AFGSA generator (forward, backward, weight packing).

This is the host-side schedule of C-ABI launches that replaces the reference's
``AFGSANet.forward`` (pht/models/afgsa/model.py:717-733) and its autograd
backward.  Layout decisions:

* activations are channels-last, in the compute dtype (bf16 production / fp32
  parity); inputs of 3x3 convolutions live in padded buffers [B, H+2, W+2, C]
  whose 1-px border is filled by ``pht_border_fill`` (replicate / reflect), so
  the convolution itself is a plain shifted-window GEMM (TMA friendly);
* ``torch.cat`` never materialises: concatenated inputs are extra K-sources;
* the 1x1/3x3/5x5 encoder branches are one GEMM over a 5x5 im2col;
* data-gradients of padded convolutions are computed over the padded domain and
  folded back by ``pht_pad_fold`` (which also applies the residual add and the
  activation-derivative mask of the producing layer).
"""
from __future__ import annotations

import os

import torch

from ... import ops
from ..._lib import PAD_MODES

LEAKY = 0.2
ENC_KPAD = {3: 128, 7: 192}  # 25*Cin rounded up to a multiple of 64


class _Arena:
    """Named persistent device buffers for one (B, H, W, dtype, tag)."""

    def __init__(self, device):
        self.device = device
        self.bufs: dict[str, torch.Tensor] = {}
        self.slices: dict = {}

    def get(self, name, shape, dtype, zero=False):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())

    # cached slices: the same Python objects every step (slicing costs ~5 us, and ops memoise pht_view per object)
    def _slice(self, t, kind, fn):
        key = (id(t), kind)
        ent = self.slices.get(key)
        if ent is None or ent[0] is not t:
            if len(self.slices) > 1024:
                self.slices.clear()
            ent = self.slices[key] = (t, fn(t))
        return ent[1]

    def inner(self, t):
        """interior [:, 1:-1, 1:-1, :] of a padded buffer"""
        return self._slice(t, "inner", lambda u: u[:, 1:-1, 1:-1, :])

    def lo(self, t, c):
        return self._slice(t, ("lo", c), lambda u: u[..., :c])

    def hi(self, t, c):
        return self._slice(t, ("hi", c), lambda u: u[..., c:])


class AfgsaEngine:
    def __init__(self, net):
        self.net = net
        self.C = net.base_ch
        self.heads, self.block, self.halo = net.num_heads, net.block_size, net.halo_size
        self.num_sa = net.num_sa
        self.d = self.C // self.heads
        self.win = self.block + 2 * self.halo
        self.scale = float(self.d) ** -0.5
        self.pad_mode = PAD_MODES[net.padding_mode]
        self._arenas: dict[tuple, _Arena] = {}
        self._packed: dict[str, torch.Tensor] = {}
        self._consts: dict[str, torch.Tensor] = {}
        self._packed_key = None
        self._pack_plans = {}
        self._wg_bucket = None
        self._side_stream = None
        self._saved_gen = {}
        self._gen = 0
        # data-parallel hook: called with "decoder" / "block<i>" / "encoders" as soon as that group's
        # parameter gradients are final in net.flat_grad (parallel.GradBucketer.ready)
        self.grad_ready_hook = None

    def _dbg(self, name, t):
        if getattr(self, "debug_sink", None) is not None:
            self.debug_sink[name] = t.detach().clone()

    def _ready(self, tag):
        if self.grad_ready_hook is not None:
            self.grad_ready_hook(tag)

    # ------------------------------------------------------------------ helpers
    @property
    def dtype(self):
        return self.net.compute_dtype

    @property
    def device(self):
        return self.net.flat_param.device

    # at most this many activation arenas stay alive (least recently used first out): a training shape, an evaluation
    # shape and the odd partial batch / validation patch.  Tiled inference pads its tiles to ONE haloed shape.
    max_arenas = 4

    def _arena(self, B, H, W, tag):
        key = (B, H, W, self.dtype, tag)
        a = self._arenas.pop(key, None)
        if a is None:
            a = _Arena(self.device)
            while len(self._arenas) >= self.max_arenas:
                # never the arena of a forward whose backward is still pending
                victim = next((k for k in self._arenas if not (k[4] == "train" and k[:3] in self._saved_gen)), None)
                if victim is None:
                    break
                del self._arenas[victim]
        self._arenas[key] = a          # (re-)insert as most recently used
        return a

    def release(self):
        """Drop every cached device buffer (activation arenas, weight-gradient workspaces, packed weights and descriptor
        tables); they are rebuilt on the next forward.  For long-lived processes that change resolution."""
        self._arenas.clear()
        self._saved_gen.clear()
        self._packed.clear()
        self._pack_plans.clear()
        self._packed_key = None
        self._wg_bucket = None
        self._consts.clear()

    def arena_bytes(self) -> int:
        return sum(a.nbytes() for a in self._arenas.values())

    def _const(self, name, values):
        t = self._consts.get(name)
        if t is None or t.device != self.device:
            t = self._consts[name] = torch.tensor(values, dtype=torch.float32, device=self.device)
        return t

    def _pk(self, name, shape):
        t = self._packed.get(name)
        if t is None or t.dtype != self.dtype or t.device != self.device or tuple(t.shape) != tuple(shape):
            t = self._packed[name] = torch.zeros(shape, dtype=self.dtype, device=self.device)
        return t

    # ------------------------------------------------------------------ weights
    def pack_weights(self, backward: bool):
        """(Re)pack the fp32 OIHW master weights into the kernels' [tap][N][K]
        layout (and the flipped/transposed copies the data-gradients use): every
        pack of the step is one descriptor of ONE batched launch."""
        key = (backward, self.dtype, str(self.device), self.net.flat_param.data_ptr())
        plan = self._pack_plans.get(key)
        if plan is None:
            plan = self._pack_plans[key] = self._build_pack_plan(backward)
        plan.run()

    def _pk32(self, name, shape):
        t = self._packed.get(name)
        if t is None or t.dtype != torch.float32 or t.device != self.device or tuple(t.shape) != tuple(shape):
            t = self._packed[name] = torch.zeros(shape, dtype=torch.float32, device=self.device)
        return t

    def _build_pack_plan(self, backward: bool):
        net, C = self.net, self.C
        P = net.params_dict_cached()
        plan = ops.PackPlan()
        pw = plan.add
        tc_tail = self.dtype == torch.bfloat16
        # encoders: 1x1 / 3x3 / 5x5 kernels embedded in the 5x5 im2col K axis; biases concatenated to [768]
        for tag, names, cin in (("encN", ("conv1", "conv3", "conv5"), net.input_channels),
                                ("encA", ("conv_a1", "conv_a3", "conv_a5"), net.aux_input_channels)):
            kpad = ENC_KPAD.get(cin, ((25 * cin + 63) // 64) * 64)
            wp = self._pk(tag, (1, 768, kpad))
            bp = self._pk32(tag + ".bias", (768,))
            for j, (nm, ks) in enumerate(zip(names, (1, 3, 5))):
                pw(P[f"{nm}.0.weight"], wp, ksize=ks, Ntot=768, Ktot=kpad, n_off=256 * j, grid=5)
                pw(P[f"{nm}.0.bias"].view(256, 1), bp, ksize=1, Ntot=768, Ktot=1, n_off=256 * j)
        for nm, I in (("conv_map", 768), ("conv_aenc1", 768), ("conv_aenc2", C)):
            pw(P[f"{nm}.0.weight"], self._pk(nm, (1, C, I)), ksize=1, Ntot=C, Ktot=I)
            if backward:
                pw(P[f"{nm}.0.weight"], self._pk(nm + ".T", (1, I, C)), ksize=1, Ntot=I, Ktot=C, transpose=1)
        for i in range(self.num_sa):
            pre = f"transformer_blocks.{i}."
            film = getattr(net, "use_film", False)
            wq, wk, wv = (P[pre + "attention.q_conv.weight"], P[pre + "attention.k_conv.weight"],
                          P[pre + "attention.v_conv.weight"])
            if film:   # FiLM variant: gamma_beta = W2 relu(W0 a + b0) + b2 (film.py:28-38)
                w0, w2 = P[pre + "attention.film.affine.0.weight"], P[pre + "attention.film.affine.2.weight"]
                FH = w0.shape[0]
                pw(w0, self._pk(f"b{i}.film0", (1, FH, C)), ksize=1, Ntot=FH, Ktot=C)
                pw(w2, self._pk(f"b{i}.film2", (1, 2 * C, FH)), ksize=1, Ntot=2 * C, Ktot=FH)
            else:
                wmap = P[pre + "attention.conv_map.0.weight"]
                pw(wmap, self._pk(f"b{i}.map", (1, C, 2 * C)), ksize=1, Ntot=C, Ktot=2 * C)
            wqk = self._pk(f"b{i}.qk", (1, 2 * C, C))
            pw(wq, wqk, ksize=1, Ntot=2 * C, Ktot=C, n_off=0, scale=self.scale)
            pw(wk, wqk, ksize=1, Ntot=2 * C, Ktot=C, n_off=C)
            pw(wv, self._pk(f"b{i}.v", (1, C, C)), ksize=1, Ntot=C, Ktot=C)
            for j in (0, 1):
                w = P[pre + f"feed_forward.{j}.0.weight"]
                pw(w, self._pk(f"b{i}.ff{j}", (9, C, C)), ksize=3, Ntot=C, Ktot=C)
                if backward:
                    pw(w, self._pk(f"b{i}.ff{j}.T", (9, C, C)), ksize=3, Ntot=C, Ktot=C, transpose=1)
            if backward:
                wqkT = self._pk(f"b{i}.qk.T", (1, C, 2 * C))       # dM = [dQ | dK] @ [s*Wq ; Wk]
                pw(wq, wqkT, ksize=1, Ntot=C, Ktot=2 * C, k_off=0, transpose=1, scale=self.scale)
                pw(wk, wqkT, ksize=1, Ntot=C, Ktot=2 * C, k_off=C, transpose=1)
                if film:
                    pw(wv, self._pk(f"b{i}.v.T", (1, C, C)), ksize=1, Ntot=C, Ktot=C, transpose=1)          # dX += dV @ Wv
                    pw(w2, self._pk(f"b{i}.film2.T", (1, FH, 2 * C)), ksize=1, Ntot=FH, Ktot=2 * C, transpose=1)
                    pw(w0, self._pk(f"b{i}.a.T", (1, C, FH)), ksize=1, Ntot=C, Ktot=FH, transpose=1)        # dA += dH @ W0
                else:
                    wx = self._pk(f"b{i}.x.T", (1, C, 2 * C))          # dX = [dV | dMpre] @ [Wv ; Wmap[:, :C]]
                    pw(wv, wx, ksize=1, Ntot=C, Ktot=2 * C, k_off=0, transpose=1)
                    pw(wmap, wx, ksize=1, Ntot=C, Ktot=2 * C, k_off=C, transpose=1, i_begin=0, i_count=C)
                    pw(wmap, self._pk(f"b{i}.a.T", (1, C, C)), ksize=1, Ntot=C, Ktot=C, transpose=1, i_begin=C, i_count=C)
        for j in (0, 1):
            w = P[f"decoder.{j}.0.weight"]
            pw(w, self._pk(f"dec{j}", (9, C, C)), ksize=3, Ntot=C, Ktot=C)
            if backward:
                pw(w, self._pk(f"dec{j}.T", (9, C, C)), ksize=3, Ntot=C, Ktot=C, transpose=1)
        w2 = P["decoder.2.0.weight"]
        if tc_tail:
            # decoder tail on the tensor cores: 3 output channels padded to a 64-wide N tile (rows 3..63 stay zero);
            # backward operand [C][27 -> 64]: k = tap*3 + co
            # forward as a 1x1 GEMM [27 -> 64][C] (row t*3 + co; rows 27..63 stay zero) + a 9-tap gather of its output
            pw(w2, self._pk("dec2f", (1, 64, C)), ksize=3, Ntot=3, Ktot=C)
            if backward:
                pw(w2, self._pk("dec2g.T", (1, C, 64)), ksize=3, Ntot=C, Ktot=64, transpose=2)
        else:
            # fp32 parity path: CUDA-core tail kernels read [3][9][C] = dst[co][t*C + c]
            pw(w2, self._pk32("dec2", (3, 9, C)), ksize=3, Ntot=3, Ktot=9 * C, grid=3)
        return plan

    def _maybe_pack(self, backward: bool):
        key = (self.net.weights_version(), backward or (self._packed_key is not None and self._packed_key[1]),
               self.dtype, str(self.device))
        if key != self._packed_key:
            self.pack_weights(key[1])
            self._packed_key = key

    # ------------------------------------------------------------------ forward
    def forward(self, x, aux, save: bool):
        """x [B,3,H,W], aux [B,7,H,W] fp32 NCHW -> out [B,3,H,W] fp32 NCHW."""
        net, C, T = self.net, self.C, self.dtype
        B, _, H, W = x.shape
        if H % self.block or W % self.block:
            raise AssertionError("feature map dimensions must be divisible by the block size")  # model.py:469-471
        x = x.contiguous().float()
        aux = aux.contiguous().float()
        self._maybe_pack(backward=save)
        A = self._arena(B, H, W, "train" if save else "eval")
        pk, P = self._packed, net.params_dict_cached()
        relu = self._const("relu", [0.0] * C)
        leaky = self._const("leaky", [LEAKY] * C)
        slopeN = self._const("slopeN", [0.0] * 768)
        slopeA = self._const("slopeA", [0.0] * 256 + [LEAKY] * 512)
        mode = self.pad_mode
        # bf16: the padding frame of a 3x3 convolution's input is written by the epilogue of the kernel that produces it
        # (PHT_EPI_RING* / pht_attn_args.ring) instead of a pht_border_fill launch -- for the launch-bound small batches
        # only: measured +1.2 % at dev (8 x 32 x 32) but -0.6 % at prod (8 x 128 x 128), where a border_fill launch hides
        # under its neighbours' tails while the extra stores of the edge threads sit on the epilogue's critical path.
        # PHT_FUSED_RING=1 / 0 forces it on / off.
        want = os.environ.get("PHT_FUSED_RING", "auto")
        ring_on = want == "1" or (want not in ("0", "1") and B * H * W <= 65536)
        ring = ("reflect" if net.padding_mode == "reflect" else "replicate") if (
            T == torch.bfloat16 and H >= 4 and W >= 4 and ring_on and not getattr(self, "no_fused_ring", False)) else None
        nb = lambda i: i if save else 0          # per-block buffers only when saving for backward
        g = A.get

        colN = g("colN", (B, H, W, pk["encN"].shape[-1]), T)
        colA = g("colA", (B, H, W, pk["encA"].shape[-1]), T)
        catN = g("catN", (B, H, W, 768), T)
        catA = g("catA", (B, H, W, 768), T)
        ops.im2col5(x, colN, mode)
        ops.im2col5(aux, colA, mode)
        ops.conv_gemm([colN], pk["encN"], 768, bias=pk["encN.bias"], slope=slopeN, out1=catN)
        ops.conv_gemm([colA], pk["encA"], 768, bias=pk["encA.bias"], slope=slopeA, out1=catA)
        Xp = g("Xp0", (B, H + 2, W + 2, C), T)
        X = A.inner(Xp)
        ops.conv_gemm([catN], pk["conv_map"], C, bias=P["conv_map.0.bias"], slope=relu, out1=X)
        A1 = g("A1", (B, H, W, C), T)
        Af = g("A", (B, H, W, C), T)
        ops.conv_gemm([catA], pk["conv_aenc1"], C, bias=P["conv_aenc1.0.bias"], slope=leaky, out1=A1)
        ops.conv_gemm([A1], pk["conv_aenc2"], C, bias=P["conv_aenc2.0.bias"], slope=leaky, out1=Af)

        for i in range(self.num_sa):
            pre = f"transformer_blocks.{i}."
            j = nb(i)
            M = g(f"M{j}", (B, H, W, C), T)
            QK = g(f"QK{j}", (B, H, W, 2 * C), T)
            V = g(f"V{j}", (B, H, W, C), T)
            X1p = g(f"X1p{j}", (B, H + 2, W + 2, C), T)
            H1p = g(f"H1p{j}", (B, H + 2, W + 2, C), T)
            H2 = g(f"H2{j}", (B, H, W, C), T)
            lse = g(f"lse{j}", (B, H, W, self.heads), torch.float32)
            Xn = g(f"Xp{(i + 1) if save else (i + 1) % 2}", (B, H + 2, W + 2, C), T)
            X1 = A.inner(X1p)
            if net.use_film:
                # n_aux = gamma * x + beta with [gamma | beta] = W2 relu(W0 a + b0) + b2 (film.py:36-45, model.py:458-460)
                FH = pk[f"b{i}.film0"].shape[1]
                Fh = g(f"Fh{j}", (B, H, W, FH), T)
                GB = g(f"GB{j}", (B, H, W, 2 * C), T)
                ops.conv_gemm([Af], pk[f"b{i}.film0"], FH, bias=P[pre + "attention.film.affine.0.bias"],
                              slope=self._const("reluF", [0.0] * FH), out1=Fh)
                ops.conv_gemm([Fh], pk[f"b{i}.film2"], 2 * C, bias=P[pre + "attention.film.affine.2.bias"], out1=GB)
                ops.film_fwd(GB, X, M)
            else:
                ops.conv_gemm([X, Af], pk[f"b{i}.map"], C, bias=P[pre + "attention.conv_map.0.bias"], slope=relu, out1=M)
            ops.conv_gemm([M], pk[f"b{i}.qk"], 2 * C, out1=QK)
            ops.conv_gemm([X], pk[f"b{i}.v"], C, out1=V)
            ops.attn_fwd(A.lo(QK, C), A.hi(QK, C), V, P[pre + "attention.rel_h"], P[pre + "attention.rel_w"], X1,
                         heads=self.heads, block=self.block, halo=self.halo, resid=X, lse=lse, ring=ring)
            if ring is None:
                ops.border_fill(X1p, mode)
            ops.conv_gemm([X1p], pk[f"b{i}.ff0"], C, ksize=3, src_offsets=[(1, 1)],
                          bias=P[pre + "feed_forward.0.0.bias"], slope=relu, out1=A.inner(H1p), ring1=ring)
            if ring is None:
                ops.border_fill(H1p, mode)
            last = i == self.num_sa - 1          # (only the decoder reads a block output through a 3x3 window)
            ops.conv_gemm([H1p], pk[f"b{i}.ff1"], C, ksize=3, src_offsets=[(1, 1)],
                          bias=P[pre + "feed_forward.1.0.bias"], slope=relu, resid=X1, resid_mode="post",
                          out1=H2, out2=A.inner(Xn), ring2=ring if last else None)
            Xp, X = Xn, A.inner(Xn)

        if ring is None or self.num_sa == 0:
            ops.border_fill(Xp, mode)
        D1p = g("D1p", (B, H + 2, W + 2, C), T)
        D2 = g("D2", (B, H, W, C), T)
        ops.conv_gemm([Xp], pk["dec0"], C, ksize=3, src_offsets=[(1, 1)], bias=P["decoder.0.0.bias"], slope=relu,
                      out1=A.inner(D1p), ring1=ring)
        if ring is None:
            ops.border_fill(D1p, mode)
        ops.conv_gemm([D1p], pk["dec1"], C, ksize=3, src_offsets=[(1, 1)], bias=P["decoder.1.0.bias"], slope=relu,
                      out1=D2)
        out = torch.empty_like(x)
        if T == torch.bfloat16:
            # 256->3 zero-padded 3x3 conv: per-pixel products with all 27 (tap, channel) weight rows as ONE 64-wide 1x1
            # tensor-core GEMM (D2 is read once, not nine times), then the 9-tap gather + bias + residual + NCHW
            Y = g("tailY", (B, H, W, 64), torch.float32)
            ops.conv_gemm([D2], pk["dec2f"], 64, ksize=1, out1=Y)
            ops.tail_gather(Y, P["decoder.2.0.bias"], x, out)
        else:
            ops.dec_tail_fwd(D2, pk["dec2"], P["decoder.2.0.bias"], x, out)
        if save:
            self._gen += 1
            self._saved_gen[(B, H, W)] = self._gen
            return out, (B, H, W, self._gen)
        return out, None

    # ------------------------------------------------------------------ backward
    def backward(self, token, d_out):
        """d_out [B,3,H,W] fp32 -> writes every parameter gradient into
        ``net.flat_grad`` (views of it are returned by the autograd Function)."""
        B, H, W, gen = token
        if self._saved_gen.get((B, H, W)) != gen:
            raise RuntimeError("AFGSANet backward: the saved activations were overwritten by a later forward "
                               "of the same shape (run backward before the next grad-enabled forward)")
        net, C, T = self.net, self.C, self.dtype
        d_out = d_out.contiguous().float()
        A = self._arena(B, H, W, "train")
        g, pk, mode = A.get, self._packed, self.pad_mode
        G = net.grad_views()
        P = net.params_dict_cached()
        relu0 = self._const("relu", [0.0] * C)
        leaky = self._const("leaky", [LEAKY] * C)
        slopeN = self._const("slopeN", [0.0] * 768)
        slopeA = self._const("slopeA", [0.0] * 256 + [LEAKY] * 512)
        npx = B * H * W

        # gradient ping-pong buffers live in padded frames: the fused data-gradient (PHT_EPI_PADFOLD) stores whole
        # padded-domain tiles (frame = don't care), everything else addresses the interior views
        Gpad = {n: g(n + "p", (B, H + 2, W + 2, C), T) for n in ("G0", "G1", "G2", "GX")}
        G0, G1, G2, GX = (A.inner(Gpad[n]) for n in ("G0", "G1", "G2", "GX"))
        pad_of = {id(G0): Gpad["G0"], id(G1): Gpad["G1"], id(G2): Gpad["G2"], id(GX): Gpad["GX"]}
        GA = g("GA", (B, H, W, C), T)
        fused_fold = T == torch.bfloat16 and net.padding_mode == "replicate" and not getattr(self, "no_fused_fold", False)
        GP = None if fused_fold else g("GP", (B, H + 2, W + 2, C), T)
        dQK = g("dQK", (B, H, W, 2 * C), T)
        dV = g("dV", (B, H, W, C), T)
        dcat = g("dcat", (B, H, W, 768), T)
        wtmp = g("wtmp", (9 * C * 768,), torch.float32)
        btmp = g("btmp", (768,), torch.float32)
        attn_ws = g("attn_ws", (max(ops.attn_bwd_workspace_bytes(A.lo(dQK, C), self.heads, self.block, self.halo), 16) // 4,),
                    torch.float32)
        tail_ws = g("tail_ws", (max(ops.dec_tail_ws_bytes(B, H, W, C), 16) // 4,), torch.float32)
        wg_ws = g("wg_ws", (64 * 1024 * 1024 // 4,), torch.float32)
        # bf16: weight-gradients of a bucket are deferred and finished by one batched reduce + one batched unpack
        bucket = None
        if T == torch.bfloat16 and not getattr(self, "no_wgrad_batching", False):
            bucket = self._wg_bucket
            if bucket is None or bucket.ws.device != self.device:
                bucket = self._wg_bucket = ops.WgradBucket(self.device)
        tmp_off = [0]

        def tmp(numel):
            """fp32 staging for packed gradients: one slice per job of the current bucket"""
            o = tmp_off[0]
            tmp_off[0] = o + (numel + 63) // 64 * 64
            assert tmp_off[0] <= wtmp_b.numel()
            return wtmp_b[o:o + numel]

        wtmp_b = g("wtmp_b", (6 * 9 * C * C + 4 * 768 * 256,), torch.float32)

        def wgrad(dy, srcs, dw, **kw):
            if bucket is None:
                ops.wgrad(dy, srcs, dw, workspace=wg_ws, **kw)
            else:
                bucket.wgrad(dy, srcs, dw, **kw)

        # The 1x1 weight-gradient GEMMs run on half of the SMs each (74 CTAs: fewer fp32 partials to reduce).  Two that
        # are ready at the same time are launched side by side -- one on a second stream -- and joined right away, so
        # nothing else ever overlaps them and no buffer they read can be overwritten underneath them.
        pair_ok = bucket is not None and os.environ.get("PHT_PAIR_WGRADS", "1") != "0"

        def paired(main_fn, side_fns):
            if not pair_ok:
                main_fn()
                for f in side_fns:
                    f()
                return
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            cur = torch.cuda.current_stream()
            self._side_stream.wait_stream(cur)
            with torch.cuda.stream(self._side_stream):
                for f in side_fns:
                    f()
            main_fn()
            cur.wait_stream(self._side_stream)

        def unpack(wg, packed, **kw):
            if bucket is None:
                ops.unpack_wgrad(wg, packed, **kw)
            else:
                bucket.unpack(wg, packed, **kw)

        def ready(tag):
            if bucket is not None:
                bucket.flush()
            tmp_off[0] = 0
            self._ready(tag)
        pdom = (B, H + 2, W + 2)

        def conv3_wgrad(dy, src_pad, name):
            w = tmp(9 * C * C).view(9, C, C)
            wgrad(dy, [src_pad], w, ksize=3, dbias=G[name + ".bias"], src_offsets=[(1, 1)])
            unpack(G[name + ".weight"], w, ksize=3, Ntot=C, Ktot=C)

        def conv3_dgrad(dy, wT, *, resid=None, mask=None, out1=None, out2=None):
            """d(input of a padded 3x3 conv): data-gradient over the padded domain, border folded back into the
            interior, then out1 = fold (+ resid), out2 = out1 * relu'(mask)."""
            if fused_fold:   # one launch: the fold, residual and mask run in the GEMM epilogue
                ops.conv_gemm([dy], wT, C, ksize=3, out_domain=pdom, src_offsets=[(-1, -1)], padfold=True,
                              resid=resid, resid_mode="pre" if resid is not None else None,
                              mask=mask, mslope=relu0 if mask is not None else None,
                              out1=None if out1 is None else pad_of[id(out1)],
                              out2=None if out2 is None else pad_of[id(out2)])
            else:
                ops.conv_gemm([dy], wT, C, ksize=3, out_domain=pdom, src_offsets=[(-1, -1)], out1=GP)
                ops.pad_fold(GP, mode, resid=resid, mask=mask, mslope=relu0 if mask is not None else None,
                             out1=out1, out2=out2)

        # ---- decoder -------------------------------------------------------------------------
        D1p, D2 = g("D1p", (B, H + 2, W + 2, C), T), g("D2", (B, H, W, C), T)
        if T == torch.bfloat16:
            tailA = g("tailA", (B, H, W, 64), T)               # a[p][t*3+co] = d_out[p - tap_t][co]
            ops.tail_im2col_bwd(d_out, tailA, G["decoder.2.0.bias"])
            dw2 = tmp(64 * C).view(1, C, 64)                   # dw2[c][t*3+co] = sum_p D2[p][c] a[p][t*3+co]
            wgrad(D2, [tailA], dw2)
            unpack(G["decoder.2.0.weight"], dw2, ksize=3, Ntot=C, Ktot=64, transpose=2)
            ops.conv_gemm([tailA], pk["dec2g.T"], C, mask=D2, mslope=relu0, out2=G0)   # G0 = d(D2 pre-act)
        else:
            dw2 = wtmp[: 27 * C].view(3, 9, C)
            ops.dec_tail_bwd_weight(d_out, D2, dw2, G["decoder.2.0.bias"], tail_ws)
            ops.unpack_wgrad(G["decoder.2.0.weight"], dw2, ksize=3, Ntot=3, Ktot=9 * C, grid=3)
            ops.dec_tail_bwd_data(d_out, pk["dec2"], D2, G0)                       # G0 = d(D2 pre-act)
        self._dbg("dD2pre", G0); self._dbg("d_out", d_out); self._dbg("D2", D2); self._dbg("D1", D1p)
        conv3_wgrad(G0, D1p, "decoder.1.0")
        conv3_dgrad(G0, pk["dec1.T"], mask=A.inner(D1p), out2=G1)          # G1 = d(D1 pre-act)
        self._dbg("dD1pre", G1)
        Xlast = g(f"Xp{self.num_sa}", (B, H + 2, W + 2, C), T)
        conv3_wgrad(G1, Xlast, "decoder.0.0")
        ready("decoder")
        if self.num_sa > 0:
            conv3_dgrad(G1, pk["dec0.T"], mask=g(f"H2{self.num_sa - 1}", (B, H, W, C), T), out1=GX, out2=G0)
            self._dbg("dXlast", GX); self._dbg("dH2pre_last", G0)
        else:
            conv3_dgrad(G1, pk["dec0.T"], mask=A.inner(Xlast), out2=G0)

        Af, A1 = g("A", (B, H, W, C), T), g("A1", (B, H, W, C), T)
        # ---- transformer blocks, last to first ------------------------------------------------
        for i in reversed(range(self.num_sa)):
            pre = f"transformer_blocks.{i}."
            M, QK, V = g(f"M{i}", (B, H, W, C), T), g(f"QK{i}", (B, H, W, 2 * C), T), g(f"V{i}", (B, H, W, C), T)
            X1p, H1p = g(f"X1p{i}", (B, H + 2, W + 2, C), T), g(f"H1p{i}", (B, H + 2, W + 2, C), T)
            lse = g(f"lse{i}", (B, H, W, self.heads), torch.float32)
            X = A.inner(g(f"Xp{i}", (B, H + 2, W + 2, C), T))
            # GX = d(block output) raw, G0 = d(H2 pre-act)
            attn_args = (A.lo(QK, C), A.hi(QK, C), V, P[pre + "attention.rel_h"], P[pre + "attention.rel_w"], lse, G2,
                         A.lo(dQK, C), A.hi(dQK, C), dV, G[pre + "attention.rel_h"], G[pre + "attention.rel_w"], attn_ws)
            attn_kw = dict(heads=self.heads, block=self.block, halo=self.halo)
            prezero = T == torch.bfloat16 and pair_ok
            if prezero:
                # the tcgen05 attention backward ACCUMULATES the window contributions into dK / dV: zero them on the
                # second stream now, underneath the four feed-forward gradient GEMMs (tensor-bound, HBM mostly idle);
                # every reader of the previous block's dQK / dV is already enqueued on the main stream
                if self._side_stream is None:
                    self._side_stream = torch.cuda.Stream(device=self.device)
                self._side_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side_stream):
                    ops.attn_bwd_zero(*attn_args, **attn_kw)
            conv3_wgrad(G0, H1p, pre + "feed_forward.1.0")
            conv3_dgrad(G0, pk[f"b{i}.ff1.T"], mask=A.inner(H1p), out2=G1)       # G1 = d(H1 pre-act)
            conv3_wgrad(G1, X1p, pre + "feed_forward.0.0")
            conv3_dgrad(G1, pk[f"b{i}.ff0.T"], resid=GX, out1=G2)                          # G2 = dX1 = dO
            # attention
            if prezero:
                torch.cuda.current_stream().wait_stream(self._side_stream)
            ops.attn_bwd(*attn_args, **attn_kw, prezeroed=prezero)
            wqk = tmp(2 * C * C).view(1, 2 * C, C)
            paired(lambda: wgrad(dQK, [M], wqk),
                   [lambda: wgrad(dV, [X], G[pre + "attention.v_conv.weight"])])
            unpack(G[pre + "attention.q_conv.weight"], wqk, ksize=1, Ntot=2 * C, Ktot=C, n_off=0, scale=self.scale)
            unpack(G[pre + "attention.k_conv.weight"], wqk, ksize=1, Ntot=2 * C, Ktot=C, n_off=C)
            first = i == self.num_sa - 1
            if net.use_film:
                FH = pk[f"b{i}.film0"].shape[1]
                Fh, GB = g(f"Fh{i}", (B, H, W, FH), T), g(f"GB{i}", (B, H, W, 2 * C), T)
                dGB = g("dGB", (B, H, W, 2 * C), T)
                dFh = g("dFh", (B, H, W, FH), T)
                reluF = self._const("reluF", [0.0] * FH)
                ops.conv_gemm([dQK], pk[f"b{i}.qk.T"], C, out1=G1)                         # G1 = dM (no activation on M)
                # dGB = [dM * x | dM];  G2 = dX1 + gamma * dM
                ops.film_bwd(GB, X, G1, dGB, dx_in=G2, dx=G2)
                wgrad(dGB, [Fh], G[pre + "attention.film.affine.2.weight"], dbias=G[pre + "attention.film.affine.2.bias"])
                ops.conv_gemm([dGB], pk[f"b{i}.film2.T"], FH, mask=Fh, mslope=reluF, out2=dFh)   # d(hidden pre-act)
                wgrad(dFh, [Af], G[pre + "attention.film.affine.0.weight"], dbias=G[pre + "attention.film.affine.0.bias"])
                G[pre + "attention.alpha"].zero_()      # registered by the reference, unused by its forward
                ready(f"block{i}")
                # d(block input) = (dX1 + gamma dM) + dV Wv; also emit the next layer's masked gradient
                if i > 0:
                    ops.conv_gemm([dV], pk[f"b{i}.v.T"], C, resid=G2, resid_mode="pre",
                                  mask=g(f"H2{i - 1}", (B, H, W, C), T), mslope=relu0, out1=GX, out2=G0)
                else:
                    ops.conv_gemm([dV], pk[f"b{i}.v.T"], C, resid=G2, resid_mode="pre", mask=X, mslope=relu0, out2=G0)
                dA_src = dFh
            else:
                ops.conv_gemm([dQK], pk[f"b{i}.qk.T"], C, mask=M, mslope=relu0, out2=G1)       # G1 = d(M pre-act)
                wgrad(G1, [X, Af], G[pre + "attention.conv_map.0.weight"], dbias=G[pre + "attention.conv_map.0.bias"])
                ready(f"block{i}")
                # d(block input) = dX1 + dV Wv + dMpre Wmap[:, :C]; also emit the next layer's masked gradient
                if i > 0:
                    ops.conv_gemm([dV, G1], pk[f"b{i}.x.T"], C, resid=G2, resid_mode="pre",
                                  mask=g(f"H2{i - 1}", (B, H, W, C), T), mslope=relu0, out1=GX, out2=G0)
                else:
                    ops.conv_gemm([dV, G1], pk[f"b{i}.x.T"], C, resid=G2, resid_mode="pre", mask=X, mslope=relu0, out2=G0)
                dA_src = G1
            # dA accumulates over blocks; the last accumulation applies LeakyReLU' of conv_aenc2
            if i > 0:
                ops.conv_gemm([dA_src], pk[f"b{i}.a.T"], C, resid=None if first else GA, resid_mode="pre", out1=GA)
            else:
                ops.conv_gemm([dA_src], pk[f"b{i}.a.T"], C, resid=None if first else GA, resid_mode="pre",
                              mask=Af, mslope=leaky, out2=G2)                              # G2 = d(A pre-act)

        # ---- noisy encoder: G0 = d(conv_map pre-act) -----------------------------------------------
        catN, catA = g("catN", (B, H, W, 768), T), g("catA", (B, H, W, 768), T)
        colN = g("colN", (B, H, W, pk["encN"].shape[-1]), T)
        colA = g("colA", (B, H, W, pk["encA"].shape[-1]), T)
        ops.conv_gemm([G0], pk["conv_map.T"], 768, mask=catN, mslope=slopeN, out2=dcat)
        paired(lambda: wgrad(G0, [catN], G["conv_map.0.weight"], dbias=G["conv_map.0.bias"]),
               [lambda: self._encoder_wgrad(dcat, colN, ("conv1", "conv3", "conv5"), G, tmp, wgrad, unpack)])
        # ---- aux encoder: G2 = d(conv_aenc2 pre-act) ------------------------------------------------
        if self.num_sa > 0:
            ops.conv_gemm([G2], pk["conv_aenc2.T"], C, mask=A1, mslope=leaky, out2=G1)   # G1 = d(aenc1 pre-act)
            ops.conv_gemm([G1], pk["conv_aenc1.T"], 768, mask=catA, mslope=slopeA, out2=dcat)
            paired(lambda: wgrad(G1, [catA], G["conv_aenc1.0.weight"], dbias=G["conv_aenc1.0.bias"]),
                   [lambda: self._encoder_wgrad(dcat, colA, ("conv_a1", "conv_a3", "conv_a5"), G, tmp, wgrad, unpack),
                    lambda: wgrad(G2, [A1], G["conv_aenc2.0.weight"], dbias=G["conv_aenc2.0.bias"])])
        else:  # no attention block consumes the aux features: their gradient is zero
            for n in ("conv_a1", "conv_a3", "conv_a5", "conv_aenc1", "conv_aenc2"):
                G[n + ".0.weight"].zero_()
                G[n + ".0.bias"].zero_()
        ready("encoders")
        del self._saved_gen[(B, H, W)]

    def _encoder_wgrad(self, dcat, col, names, G, tmp, wgrad, unpack):
        kpad = col.shape[-1]
        w = tmp(768 * kpad).view(1, 768, kpad)
        b = tmp(768)
        wgrad(dcat, [col], w, dbias=b)
        for j, (nm, ks) in enumerate(zip(names, (1, 3, 5))):
            unpack(G[f"{nm}.0.weight"], w, ksize=ks, Ntot=768, Ktot=kpad, n_off=256 * j, grid=5)
            unpack(G[f"{nm}.0.bias"].view(256, 1), b, ksize=1, Ntot=768, Ktot=1, n_off=256 * j)
