"""AFGSA trainer: the plug-in point (reference: pht/models/afgsa/train.py:11-37)."""
from __future__ import annotations

from ...config import AFGSAModelConfig
from ..base_trainer import BaseTrainer
from .model import AFGSANet


class AFGSATrainer(BaseTrainer):
    def create_generator(self) -> AFGSANet:
        m = self.cfg.model
        assert isinstance(m, AFGSAModelConfig)
        if m.reproducible_attention_backward:       # (process-wide library option; the default is left as configured)
            from ... import _lib
            _lib.lib.pht_set_option(b"attn_bwd_direct", 0)
        return AFGSANet(
            m.input_channels,
            m.aux_input_channels,
            m.feature_map_channels,
            num_sa=m.self_attention.num_layers,
            block_size=m.self_attention.block_size,
            halo_size=m.self_attention.halo_size,
            num_heads=m.self_attention.num_heads,
            num_gcp=m.num_gradient_checkpoints,
            padding_mode=self.padding_mode,
            curve_order=m.curve_order,
            use_film=m.use_film,
            compute_dtype=m.compute_dtype,
        ).to(self.device)
