"""Golden fixtures for the importance-sampling path (preprocessing.py:119-168, 223-322) from the REAL reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_importance.py
Writes tests/golden/importance.npz: the reference importance maps and the kept patch centres for several seeds; also
asserts that the CPU oracle reproduces both bit for bit."""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402
from importance_inputs import CASES, SEEDS, frames  # noqa: E402
from oracle import sampler_oracle as S  # noqa: E402


def main():
    _, _, ref_pre = mg.import_reference()
    out = {}
    for name, (h, w, p, n) in CASES.items():
        noisy, normal, _ = frames(h, w)
        imp = ref_pre.get_importance_map([noisy, normal], ["relative", "variance"], [1.0, 1.0], p)
        assert imp.dtype == np.float32
        assert np.array_equal(imp, S.importance_map(noisy, normal, p)), "oracle importance map differs from the reference"
        out[f"{name}__imp"] = imp
        for seed in SEEDS:
            kept = ref_pre.importance_sampling({"noisy": noisy, "normal": normal}, p, n, random.Random(seed))
            o = S.importance_sampling(noisy, normal, p, n, S.MT19937(seed))
            assert np.array_equal(kept, o), "oracle importance sampling differs from the reference"
            out[f"{name}__seed{seed}"] = kept.astype(np.int32)
            print(name, seed, "kept", len(kept), "of", n)
    np.savez_compressed(os.path.join(HERE, "importance.npz"), **out)


if __name__ == "__main__":
    main()
