"""Alias of the B200 generator under the reference's import path."""
from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet, CurveOrder, make_curve_indices  # noqa: F401
