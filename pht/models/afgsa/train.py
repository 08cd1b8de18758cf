"""Alias of the B200 AFGSA trainer under the reference's import path."""
from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer  # noqa: F401
