"""VGG-style WGAN critic (reference: pht/models/afgsa/model.py:264-344).

Adjacent to the hot path: it stays stock PyTorch / cuDNN in
this round and exists so a full GAN training step (base_trainer.py:388-457)
can run and be timed.  Same parameter names and init order as the reference.
"""
from __future__ import annotations

import math

import torch
from torch import nn


def _block(cin, cout, k, stride, bn):
    mods: list[nn.Module] = [nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=1)]
    if bn:
        mods.append(nn.BatchNorm2d(cout, affine=True))
    mods.append(nn.LeakyReLU(0.2, True))
    return nn.Sequential(*mods)


class DiscriminatorVGG(nn.Module):
    def __init__(self, in_nc: int, base_nf: int, input_size: int) -> None:
        super().__init__()
        n_down = int(math.log2(input_size / 4))
        feats = [_block(in_nc, base_nf, 3, 1, bn=False)]
        nc = base_nf
        for i in range(n_down):
            nxt = min(base_nf * 2 ** (i + 1), base_nf * 8)
            feats.append(_block(nc, nxt, 3, 1, bn=True))
            feats.append(_block(nxt, nxt, 4, 2, bn=True))
            nc = nxt
        self.features = nn.Sequential(*feats)
        side = input_size // 2 ** n_down
        self.classifier = nn.Sequential(nn.Linear(nc * side * side, 100), nn.LeakyReLU(0.2, True), nn.Linear(100, 1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.features(x)
        return self.classifier(x.reshape(x.size(0), -1))
