"""``python -m pht.train -cn <cfg>`` -> pixel_heal_thyself_b200.train (reference entry: pht/train.py)."""
from pixel_heal_thyself_b200.train import main

if __name__ == "__main__":
    main()
