"""AFGSANet on B200 kernels: drop-in for ``pht.models.afgsa.model.AFGSANet``.

Same constructor, same ``forward(x, aux) -> Tensor`` (NCHW fp32 in/out), same
parameter names / shapes / registration + initialisation order (so
``torch.manual_seed(s)`` reproduces the reference's random init and
``state_dict``s are interchangeable; reference: pht/models/afgsa/model.py:585-733).
Everything under ``forward`` and its backward is hand-written CUDA reached
through the C ABI (``engine.AfgsaEngine``); there is no PyTorch compute path.

The ``nn.Conv2d`` modules below are parameter containers only -- they are never
called.
"""
from __future__ import annotations

from enum import Enum

import torch
from torch import nn
from torch.nn import init

from .engine import AfgsaEngine


class CurveOrder(str, Enum):
    """Query ordering inside a block (reference: model.py:347-352).  The order is
    numerically irrelevant for softmax(QK^T)V (rows are independent), so the
    kernels ignore it; the key is kept for config compatibility."""

    RASTER = "raster"
    HILBERT = "hilbert"
    ZORDER = "zorder"


def _morton_order(block: int) -> torch.Tensor:
    def part1(v):
        v = (v | (v << 8)) & 0x00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F
        v = (v | (v << 2)) & 0x33333333
        return (v | (v << 1)) & 0x55555555

    codes = [(part1(i // block) << 1) | part1(i % block) for i in range(block * block)]
    return torch.tensor(codes).argsort().to(torch.long)


def _hilbert_order(block: int) -> torch.Tensor:
    p = block.bit_length() - 1
    assert block == 1 << p, "Hilbert: block_size must be power of two"

    def d_of(x, y):  # distance along a 2-D Hilbert curve of order p
        d, s = 0, block // 2
        while s > 0:
            rx, ry = int((x & s) > 0), int((y & s) > 0)
            d += s * s * ((3 * rx) ^ ry)
            if ry == 0:
                if rx == 1:
                    x, y = block - 1 - x, block - 1 - y
                x, y = y, x
            s //= 2
        return d

    return torch.tensor([d_of(i % block, i // block) for i in range(block * block)]).argsort().to(torch.long)


def make_curve_indices(block_size: int, mode: CurveOrder) -> torch.Tensor:
    if mode is CurveOrder.RASTER:
        return torch.arange(block_size * block_size)
    if mode is CurveOrder.ZORDER:
        return _morton_order(block_size)
    return _hilbert_order(block_size)


def _conv(cin, cout, k, act, padding_mode="zeros"):
    mods = [nn.Conv2d(cin, cout, kernel_size=k, padding=(k - 1) // 2, padding_mode=padding_mode if k > 1 else "zeros")]
    if act == "relu":
        mods.append(nn.ReLU(True))
    elif act == "leakyrelu":
        mods.append(nn.LeakyReLU(0.2, True))
    return nn.Sequential(*mods)


class _FiLMParams(nn.Module):
    """Parameters of FiLM (reference: film.py:20-34): affine = Conv1x1(cond_ch, hidden) - ReLU - Conv1x1(hidden, 2 C)."""

    def __init__(self, ch, cond_ch, hidden=128):
        super().__init__()
        self.affine = nn.Sequential(nn.Conv2d(cond_ch, hidden, 1), nn.ReLU(True), nn.Conv2d(hidden, 2 * ch, 1))


class _AttentionParams(nn.Module):
    """Parameters of one AFGSA layer (reference: model.py:401-454, 518-524)."""

    def __init__(self, ch, block_size, halo_size, num_heads, curve_order, use_film=False):
        super().__init__()
        assert ch % num_heads == 0, "ch should be divided by # heads"
        head_ch = ch // num_heads
        win = block_size + 2 * halo_size
        self.register_buffer("curve_indices", make_curve_indices(block_size, curve_order))
        self.register_buffer("inv_curve_indices", torch.argsort(self.curve_indices))
        self.rel_h = nn.Parameter(torch.randn(1, win, 1, head_ch // 2))
        self.rel_w = nn.Parameter(torch.randn(1, 1, win, head_ch // 2))
        self.use_film = use_film
        if use_film:   # model.py:443-449 (alpha is registered by the reference but not used by its forward)
            self.alpha = nn.Parameter(torch.zeros(1))
            self.film = _FiLMParams(ch, ch, hidden=128)
        else:
            self.conv_map = _conv(ch * 2, ch, 1, "relu")
        self.q_conv = nn.Conv2d(ch, ch, kernel_size=1, bias=False)
        self.k_conv = nn.Conv2d(ch, ch, kernel_size=1, bias=False)
        self.v_conv = nn.Conv2d(ch, ch, kernel_size=1, bias=False)
        for m in (self.q_conv, self.k_conv, self.v_conv):
            init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        init.normal_(self.rel_h, 0, 1)
        init.normal_(self.rel_w, 0, 1)


class _BlockParams(nn.Module):
    def __init__(self, ch, block_size, halo_size, num_heads, padding_mode, curve_order, use_film=False):
        super().__init__()
        self.attention = _AttentionParams(ch, block_size, halo_size, num_heads, curve_order, use_film)
        self.feed_forward = nn.Sequential(_conv(ch, ch, 3, "relu", padding_mode), _conv(ch, ch, 3, "relu", padding_mode))


class _GeneratorFn(torch.autograd.Function):
    """Whole-generator autograd node: forward/backward are kernel schedules."""

    @staticmethod
    def forward(ctx, net, x, aux, *params):
        out, token = net.engine.forward(x, aux, save=True)
        ctx.net, ctx.token = net, token
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_out):
        net = ctx.net
        net.select_grad_arena()
        net.engine.backward(ctx.token, d_out)
        views = net.grad_views()
        grads = tuple(views[n].detach() if p.requires_grad else None for n, p in net.named_params_cached())
        return (None, None, None) + grads


_DTYPES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32, "float32": torch.float32}


class AFGSANet(nn.Module):
    """AFGSANet (reference: pht/models/afgsa/model.py:585-733).

    Extra keyword (not in the reference): ``compute_dtype`` -- "bf16"
    (production: bf16 activations, fp32 accumulation, tcgen05 tensor cores) or
    "fp32" (parity mode).  ``num_gcp`` (gradient checkpointing) is accepted and
    ignored: activations are kept in bf16 and B200 has 180 GB.
    """

    def __init__(self, input_channels: int, aux_input_channels: int, base_ch: int, num_sa: int = 5,
                 block_size: int = 8, halo_size: int = 3, num_heads: int = 4, num_gcp: int = 2,
                 padding_mode: str = "reflect", curve_order: CurveOrder = CurveOrder.RASTER, use_film: bool = False,
                 compute_dtype: str | torch.dtype = "bf16") -> None:
        super().__init__()
        assert num_gcp <= num_sa
        self.use_film = bool(use_film)
        if base_ch != 256 or base_ch // num_heads != 64:
            # the reference hard-codes 256-wide encoder branches (model.py:606-652); smaller widths are
            # supported by the fp32/CUDA-core kernels only
            pass
        if padding_mode not in ("replicate", "reflect"):
            raise ValueError(f"padding_mode must be 'replicate' or 'reflect', got {padding_mode!r}")
        self.input_channels, self.aux_input_channels, self.base_ch = input_channels, aux_input_channels, base_ch
        self.num_sa, self.block_size, self.halo_size, self.num_heads = num_sa, block_size, halo_size, num_heads
        self.padding_mode = padding_mode
        self.curve_order = CurveOrder(curve_order)
        self.compute_dtype = _DTYPES[compute_dtype] if isinstance(compute_dtype, str) else compute_dtype

        # registration / init order == reference (model.py:606-715)
        self.conv1 = _conv(input_channels, 256, 1, "relu")
        self.conv3 = _conv(input_channels, 256, 3, "relu", padding_mode)
        self.conv5 = _conv(input_channels, 256, 5, "relu", padding_mode)
        self.conv_map = _conv(256 * 3, base_ch, 1, "relu")
        self.conv_a1 = _conv(aux_input_channels, 256, 1, "relu")
        self.conv_a3 = _conv(aux_input_channels, 256, 3, "leakyrelu", padding_mode)
        self.conv_a5 = _conv(aux_input_channels, 256, 5, "leakyrelu", padding_mode)
        self.conv_aenc1 = _conv(256 * 3, base_ch, 1, "leakyrelu")
        self.conv_aenc2 = _conv(base_ch, base_ch, 1, "leakyrelu")
        self.transformer_blocks = nn.Sequential(*[
            _BlockParams(base_ch, block_size, halo_size, num_heads, padding_mode, self.curve_order, self.use_film)
            for _ in range(num_sa)])
        self.decoder = nn.Sequential(_conv(base_ch, base_ch, 3, "relu", padding_mode),
                                     _conv(base_ch, base_ch, 3, "relu", padding_mode),
                                     _conv(base_ch, 3, 3, None, "zeros"))

        self.flat_param: torch.Tensor | None = None
        self.flat_grad: torch.Tensor | None = None
        self._offsets: dict[str, tuple[int, int]] = {}
        self._dirty = 0
        self.engine = AfgsaEngine(self)

    # ------------------------------------------------------------------ cached parameter list
    def named_params_cached(self) -> list[tuple[str, nn.Parameter]]:
        """``list(self.named_parameters())`` computed once: the module tree is fixed after construction (``.to()`` and
        ``load_state_dict`` keep the Parameter objects), and walking it costs ~0.4 ms -- seven times per training step."""
        c = self.__dict__.get("_np_cache")
        if c is None:
            c = self.__dict__["_np_cache"] = list(self.named_parameters())
            self.__dict__["_np_dict"] = dict(c)
        return c

    def params_dict_cached(self) -> dict[str, nn.Parameter]:
        self.named_params_cached()
        return self.__dict__["_np_dict"]

    # ------------------------------------------------------------------ flat parameter arena
    def _flatten(self) -> None:
        """Re-home every parameter into one flat fp32 CUDA buffer (fused Adam and
        the data-parallel all-reduce then work on two flat arrays)."""
        params = self.named_params_cached()
        dev = params[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("AFGSANet (B200) needs its parameters on a CUDA device: there is no CPU fallback")
        if self.flat_param is not None and self.flat_param.device == dev and all(
                p.data_ptr() == self.flat_param.data_ptr() + 4 * self._offsets[n][0] for n, p in params):
            return
        align, off = 64, 0
        offsets = {}
        for n, p in params:
            offsets[n] = (off, p.numel())
            off += (p.numel() + align - 1) // align * align
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        for n, p in params:
            o, k = offsets[n]
            flat[o:o + k].copy_(p.detach().reshape(-1).float())
            p.data = flat[o:o + k].view(p.shape)
        self.flat_param, self._offsets = flat, offsets
        # two gradient arenas, used alternately: autograd may still hold (and accumulate into) views of the
        # arena written by the previous backward (optimizer.zero_grad(set_to_none=False))
        self._grad_arenas = [torch.zeros_like(flat), torch.zeros_like(flat)]
        self._grad_view_cache = [None, None]
        self._cur = 0
        self.flat_grad = self._grad_arenas[0]
        self._dirty += 1

    def select_grad_arena(self) -> None:
        """Pick the gradient arena the next backward writes: never the one a live ``p.grad`` aliases."""
        p0 = self.named_params_cached()[0][1]
        if p0.grad is not None and p0.grad.data_ptr() == self._grad_arenas[self._cur].data_ptr() + 4 * 0:
            self._cur ^= 1
        self.flat_grad = self._grad_arenas[self._cur]

    def grad_views(self) -> dict[str, torch.Tensor]:
        if self._grad_view_cache[self._cur] is None:
            shapes = {n: p.shape for n, p in self.named_params_cached()}
            self._grad_view_cache[self._cur] = {n: self.flat_grad[o:o + k].view(shapes[n])
                                                for n, (o, k) in self._offsets.items()}
        return self._grad_view_cache[self._cur]

    def weights_version(self) -> int:
        return self._dirty + sum(p._version for _, p in self.named_params_cached())

    def mark_weights_dirty(self) -> None:
        """Call after updating ``flat_param`` outside autograd's version tracking (fused Adam)."""
        self._dirty += 1

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._dirty += 1

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, aux: torch.Tensor) -> torch.Tensor:
        self._flatten()
        params = [p for _, p in self.named_params_cached()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _GeneratorFn.apply(self, x, aux, *params)
        out, _ = self.engine.forward(x, aux, save=False)
        return out
