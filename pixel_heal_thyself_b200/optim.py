"""Fused Adam over AFGSANet's flat parameter arena (``pht_adam``).

Numerically the reference's ``optim.Adam(G.parameters(), lr, betas=(0.9, 0.999),
eps=1e-8)`` (pht/models/base_trainer.py:182-187): one kernel over 9.28 M
parameters instead of a foreach chain.  It is a ``torch.optim.Optimizer`` so
``lr_scheduler.MultiStepLR`` and ``zero_grad`` work unchanged.
"""
from __future__ import annotations

import torch

from . import ops


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, net, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        self.net = net
        super().__init__(list(net.parameters()), dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = grad_scale
        self._m = self._v = None
        # {lr, step count (int32 bits), -, -} on the device: nothing step-dependent is a kernel argument, so step() can be
        # captured into / replayed from a CUDA graph
        self._hyper = None
        self._hyper_lr = None

    @property
    def step_count(self) -> int:
        return 0 if self._hyper is None else int(self._hyper[1:2].view(torch.int32).item())

    def sync_lr(self) -> None:
        """Push the (scheduler-controlled) learning rate to the device copy if it changed.  Called by step(); callers that
        replay a captured step() must call it themselves before the replay."""
        lr = float(self.param_groups[0]["lr"])
        if self._hyper is None or self._hyper.device != self.net.flat_param.device:
            self._hyper = torch.zeros(4, dtype=torch.float32, device=self.net.flat_param.device)
            self._hyper_lr = None
        if lr != self._hyper_lr:
            self._hyper[0:1].fill_(lr)
            self._hyper_lr = lr

    def gather_grads(self) -> tuple[torch.Tensor, bool]:
        """(arena, aliased): the flat gradient arena this optimiser will read, made to hold every ``p.grad``.

        Normally autograd adopts the views of ``net.flat_grad`` that backward returns, so every ``p.grad`` already
        aliases the arena (aliased = True, nothing is copied).  When it does not -- ``zero_grad(set_to_none=False)`` or
        gradient accumulation (autograd adds into the previous ``p.grad``), a hook or clip that replaced the tensor, a
        parameter without gradient -- the values are gathered into the arena and ``p.grad`` is re-homed onto its slice,
        so that whatever happens to the arena next (the data-parallel all-reduce, ``step()``) acts on the gradients
        autograd produced and a second call is a no-op.  Data-parallel callers MUST call this BEFORE the all-reduce."""
        net = self.net
        net._flatten()
        arena = net.flat_grad
        base = arena.data_ptr()
        views = net.grad_views()
        aliased = True
        for n, p in net.named_params_cached():
            o, k = net._offsets[n]
            if p.grad is None:
                arena[o:o + k].zero_()
                aliased = False
            elif p.grad.data_ptr() != base + 4 * o:
                arena[o:o + k].copy_(p.grad.reshape(-1))
                p.grad = views[n].detach()
                aliased = False
        return arena, aliased

    def _flat_grad(self) -> torch.Tensor:
        return self.gather_grads()[0]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        net = self.net
        g = self._flat_grad()
        if self._m is None or self._m.shape != net.flat_param.shape or self._m.device != net.flat_param.device:
            self._m = torch.zeros_like(net.flat_param)
            self._v = torch.zeros_like(net.flat_param)
        self.sync_lr()
        grp = self.param_groups[0]
        ops.adam_dev(net.flat_param, g, self._m, self._v, self._hyper, beta1=grp["betas"][0], beta2=grp["betas"][1],
                     eps=grp["eps"], grad_scale=self.grad_scale)
        net.mark_weights_dirty()
        return loss
