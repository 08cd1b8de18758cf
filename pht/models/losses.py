"""Alias of the B200 losses under the reference's import path."""
from pixel_heal_thyself_b200.models.losses import GANLoss, GradientPenaltyLoss, L1ReconstructionLoss  # noqa: F401
