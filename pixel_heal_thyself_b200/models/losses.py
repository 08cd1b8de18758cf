"""Losses of the PHT trainer (reference: pht/models/losses.py).

``L1ReconstructionLoss`` -- the image loss on the hot path -- runs the fused
forward+backward CUDA kernel ``pht_l1_loss``.  The adversarial terms (WGAN
critic loss, gradient penalty) belong to the PyTorch discriminator, which is
adjacent to the hot path and stays stock PyTorch.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target):
        a = inp.detach().contiguous().float()
        b = target.detach().contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        grad = torch.empty_like(a) if inp.requires_grad else None
        ops.l1_loss(a, b, loss, grad)          # loss = mean|a-b|, grad = sign(a-b)/n in one pass
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        return (ctx.grad * go if ctx.grad is not None else None), None


class L1ReconstructionLoss(nn.Module):
    """mean(|input - target|) (reference: losses.py:175-184)."""

    def forward(self, input_data: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _L1Fn.apply(input_data, target)


class _SSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, target, module):
        a = inp.detach().contiguous().float()
        b = target.detach().contiguous().float()
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        grad = torch.empty_like(a) if inp.requires_grad else None
        ops.msssim_loss(a, b, loss, grad, module._workspace(a))
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        return (ctx.grad * go if ctx.grad is not None else None), None, None


class SSIMLoss(nn.Module):
    """MS-SSIM + L1 loss on scale-normalised images (reference: losses.py:248-263, kornia 0.8.0 ``MS_SSIMLoss`` with its
    defaults; see ``pht_msssim_loss`` -- parity unpinned, kornia is not available to check against).  Fused
    forward + backward CUDA kernels; ``_window_size`` is accepted and ignored exactly like the reference's."""

    def __init__(self, _window_size: int = 11, window_size: int | None = None) -> None:
        super().__init__()
        self._ws = None

    def _workspace(self, a: torch.Tensor) -> torch.Tensor:
        B, _, H, W = a.shape
        n = (ops.msssim_ws_bytes(B, H, W) + 3) // 4
        if self._ws is None or self._ws.numel() < n or self._ws.device != a.device:
            self._ws = torch.empty(n, dtype=torch.float32, device=a.device)
        return self._ws

    def forward(self, input_data: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _SSIMFn.apply(input_data, target, self)


class GANLoss(nn.Module):
    """Adversarial loss on critic scores (reference: losses.py:103-172).  The AFGSA trainer uses "wgan"
    (base_trainer.py:141); "nsgan" (BCE on probabilities), "lsgan" (MSE) and "hinge" follow the reference's
    definitions so that the class is a drop-in for every ``loss_type`` it accepts.  Plain torch on the critic's output
    (a [B, 1] tensor) -- the critic is outside the hand-written path."""

    TYPES = ("nsgan", "wgan", "lsgan", "hinge")

    def __init__(self, loss_type: str = "nsgan", target_real_label: float = 1.0, target_fake_label: float = 0.0) -> None:
        super().__init__()
        if loss_type not in self.TYPES:
            raise NotImplementedError(f"GAN type {loss_type} is not found!")
        self.type = loss_type
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))

    def forward(self, input_data: torch.Tensor, target_is_real: bool,
                is_discriminator: bool | None = None) -> torch.Tensor:
        if self.type == "wgan":
            return -input_data.mean() if target_is_real else input_data.mean()
        if self.type == "hinge":
            if not is_discriminator:
                return (-input_data).mean()
            sign = -1.0 if target_is_real else 1.0
            return torch.relu(1.0 + sign * input_data).mean()
        label = (self.real_label if target_is_real else self.fake_label).expand_as(input_data)
        if self.type == "lsgan":
            return torch.nn.functional.mse_loss(input_data, label)
        return torch.nn.functional.binary_cross_entropy(input_data, label)


class GradientPenaltyLoss(nn.Module):
    """WGAN-GP penalty (reference: losses.py:12-57): ((||grad_x D(x_hat)||_2 - 1)^2).mean()."""

    def __init__(self, device: torch.device) -> None:
        super().__init__()
        self.device = device

    def forward(self, D: nn.Module, real_data: torch.Tensor, fake_data: torch.Tensor) -> torch.Tensor:  # noqa: N803
        alpha = torch.rand((real_data.shape[0], 1, 1, 1), dtype=torch.float32, device=self.device)
        x_hat = (alpha * fake_data.detach() + (1 - alpha) * real_data).requires_grad_(True)
        pred = D(x_hat)
        from .afgsa.discriminator import input_grad_only
        with input_grad_only():     # (this pass needs d pred / d x_hat only: skip the critic's parameter gradients in it)
            grad = torch.autograd.grad(pred, x_hat, torch.ones_like(pred), create_graph=True, retain_graph=True)[0]
        norm = grad.reshape(grad.size(0), -1).norm(2, dim=1)
        return ((norm - 1) ** 2).mean()
