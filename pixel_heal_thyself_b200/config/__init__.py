"""Typed configuration (reference: pht/config/base.py, registry.py, config/*.yaml).

Same dataclass tree and preset names (``-cn ci|dev|stag|prod``) and the same
``key.sub=value`` override syntax as the reference's Hydra entry point.  Hydra
and OmegaConf are not available in this image, so presets are composed with
PyYAML: ``presets/default.yaml`` deep-merged with ``presets/<name>.yaml`` and
then the command-line overrides.
"""
from __future__ import annotations

import os
from dataclasses import MISSING, dataclass, field, fields, is_dataclass
from typing import Any, List

import yaml

from ..models.afgsa.model import CurveOrder

PRESET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "presets")


@dataclass
class PathConfig:
    root: str = "."
    output_dir: str = "outputs/runs"


@dataclass
class ImagesConfig:
    dir: str = "data/images"
    scale: float = 1.0


@dataclass
class PatchesConfig:
    patch_size: int = 128
    num_patches: int = 400
    dir: str = ""


@dataclass
class SyntheticConfig:
    num_images: int = 4
    height: int = 1024
    width: int = 1024


@dataclass
class DataConfig:
    images: ImagesConfig = field(default_factory=ImagesConfig)
    patches: PatchesConfig = field(default_factory=PatchesConfig)
    source: str = "synthetic"
    synthetic: SyntheticConfig = field(default_factory=SyntheticConfig)


@dataclass
class OptimizerConfig:
    _target_: str = "torch.optim.Adam"
    lr: float = 1e-4
    betas: List[float] = field(default_factory=lambda: [0.9, 0.999])


@dataclass
class SchedulerConfig:
    _target_: str = "torch.optim.lr_scheduler.MultiStepLR"
    milestones: List[int] = field(default_factory=lambda: [3, 6, 9])
    gamma: float = 0.5


@dataclass
class LossesConfig:
    l1_loss_w: float = 1.0
    gan_loss_w: float = 0.005
    gp_loss_w: float = 10
    use_lpips_loss: bool = False
    lpips_loss_w: float = 0.1
    use_ssim_loss: bool = False
    ssim_loss_w: float = 0.1


@dataclass
class TrainerConfig:
    batch_size: int = 8
    epochs: int = 12
    deterministic: bool = True
    save_interval: int = 1
    num_saved_imgs: int = 6
    optim: OptimizerConfig = field(default_factory=OptimizerConfig)
    scheduler: SchedulerConfig = field(default_factory=SchedulerConfig)
    lr_g: float = 1e-4
    lr_d: float = 1e-4
    lr_gamma: float = 0.5
    lr_milestone: int = 3
    load_model: bool = False
    # directory holding G.pt / D.pt for load_model (the reference reads cfg.trainer.model_path, base_trainer.py:341-347,
    # without declaring the field)
    model_path: str = ""


@dataclass
class SelfAttentionConfig:
    num_layers: int = 5
    block_size: int = 8
    halo_size: int = 3
    num_heads: int = 4


@dataclass
class DiscriminatorConfig:
    use_multiscale_discriminator: bool = False
    use_film: bool = False


@dataclass
class AFGSAModelConfig:
    name: str = "afgsa"
    input_channels: int = 3
    aux_input_channels: int = 7
    feature_map_channels: int = 256
    curve_order: CurveOrder = CurveOrder.RASTER
    use_film: bool = False
    num_gradient_checkpoints: int = 0
    discriminator: DiscriminatorConfig = field(default_factory=DiscriminatorConfig)
    losses: LossesConfig = field(default_factory=LossesConfig)
    self_attention: SelfAttentionConfig = field(default_factory=SelfAttentionConfig)
    compute_dtype: str = "bf16"
    # bit-reproducible attention backward (window-major scratch + fixed-order fold) instead of the default arrival-order
    # vector reductions into dK / dV (library option "attn_bwd_direct"; process-wide)
    reproducible_attention_backward: bool = False


@dataclass
class LoggingConfig:
    level: str = "INFO"


@dataclass
class Config:
    seed: int = 990819
    data_ratio: float = 0.95
    run_num: int = -1
    paths: PathConfig = field(default_factory=PathConfig)
    data: DataConfig = field(default_factory=DataConfig)
    trainer: TrainerConfig = field(default_factory=TrainerConfig)
    model: AFGSAModelConfig = field(default_factory=AFGSAModelConfig)
    logging: LoggingConfig = field(default_factory=LoggingConfig)


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v
    return dst


def _build(cls, d: dict):
    kw = {}
    names = {f.name: f for f in fields(cls)}
    for k, v in d.items():
        if k not in names:
            raise KeyError(f"unknown config key {k!r} for {cls.__name__}")
        f = names[k]
        proto = f.default_factory() if f.default_factory is not MISSING else None
        kw[k] = _build(type(proto), v) if is_dataclass(proto) and isinstance(v, dict) else v
    return cls(**kw)


def _set_dotted(d: dict, key: str, value: Any) -> None:
    parts = key.lstrip("+").split(".")
    for p in parts[:-1]:
        d = d.setdefault(p, {})
    d[parts[-1]] = value


def compose(config_name: str = "default", overrides: list[str] | None = None) -> dict:
    """default.yaml <- <config_name>.yaml <- overrides, as a plain dict."""
    with open(os.path.join(PRESET_DIR, "default.yaml")) as f:
        cfg = yaml.safe_load(f)
    if config_name != "default":
        path = os.path.join(PRESET_DIR, f"{config_name}.yaml")
        if not os.path.exists(path):
            raise FileNotFoundError(f"unknown config preset {config_name!r} (have: "
                                    f"{sorted(p[:-5] for p in os.listdir(PRESET_DIR))})")
        with open(path) as f:
            _merge(cfg, yaml.safe_load(f) or {})
    for ov in overrides or []:
        k, _, v = ov.partition("=")
        _set_dotted(cfg, k, yaml.safe_load(v))
    return cfg


def load_config(config_name: str = "default", overrides: list[str] | None = None) -> Config:
    """Typed Config.  As in the reference (base.py:187-188) the model dataclass is
    built from the ``model.<name>`` sub-tree only."""
    raw = compose(config_name, overrides)
    model_raw = raw.pop("model")
    if model_raw.get("name", "afgsa") != "afgsa":
        raise ValueError(f"Unsupported model: {model_raw.get('name')}")  # mamba is out of scope
    cfg = _build(Config, raw)
    cfg.model = _build(AFGSAModelConfig, dict(model_raw.get("afgsa", {})))
    cfg.model.curve_order = CurveOrder(cfg.model.curve_order)
    return cfg
