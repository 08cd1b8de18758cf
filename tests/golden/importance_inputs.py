"""Seeded synthetic frames for the importance-sampling fixtures (shared by the generator and the tests)."""
import numpy as np

CASES = {  # name -> (H, W, P, n)
    "h256_w320_p32_n100": (256, 320, 32, 100),
    "h200_w264_p64_n40": (200, 264, 64, 40),
}
SEEDS = (990819, 1, 2, 3)


def frames(h, w, seed=3):
    """(noisy [h,w,3], normal [h,w,3], aux [h,w,7]) float32, already 'cleaned' like preprocess_data's output
    (finite, radiance >= 0): heavy-tailed smooth radiance x gamma noise, normals with a flat region and an edge."""
    rs = np.random.default_rng(seed)
    base = np.exp(rs.standard_normal((h // 8 + 2, w // 8 + 2, 3))).astype(np.float32)
    base = np.kron(base, np.ones((8, 8, 1), np.float32))[:h, :w]
    noisy = (base * rs.gamma(2.0, 0.5, (h, w, 3))).astype(np.float32)
    normal = rs.uniform(-1, 1, (h, w, 3)).astype(np.float32)
    normal[h // 4:h // 2, w // 3:2 * w // 3] = np.float32(0.3)
    normal[:, : w // 5] *= np.float32(0.1)
    depth = rs.uniform(0, 1, (h, w, 1)).astype(np.float32)
    albedo = rs.uniform(0, 1, (h, w, 3)).astype(np.float32)
    aux = np.concatenate([normal, depth, albedo], axis=2)
    return noisy, normal, aux
