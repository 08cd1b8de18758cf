"""Training entry point: ``python -m pixel_heal_thyself_b200.train -cn <ci|dev|stag|prod> [key=value ...]``
(also reachable as ``python -m pht.train``; reference: pht/train.py:16-34).  Under
``torchrun --nproc-per-node N`` it trains data-parallel, one process per GPU."""
from __future__ import annotations

import argparse
import logging

from .config import load_config
from .models.afgsa.train import AFGSATrainer


def main(argv: list[str] | None = None) -> None:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("-cn", "--config-name", default="default")
    ap.add_argument("overrides", nargs="*", help="hydra-style key.sub=value overrides")
    args = ap.parse_args(argv)
    cfg = load_config(args.config_name, args.overrides)
    logging.basicConfig(level=getattr(logging, cfg.logging.level.upper(), logging.INFO),
                        format="%(asctime)s %(levelname)s %(message)s")
    if cfg.model.name != "afgsa":
        raise ValueError(f"Unsupported model: {cfg.model.name}")
    AFGSATrainer(cfg).train()


if __name__ == "__main__":
    main()
