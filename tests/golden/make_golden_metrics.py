"""Golden fixture for the validation metrics, from the REAL reference (pht/models/afgsa/metric.py, util.py).

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_golden_metrics.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as MG  # noqa: E402


def main():
    from oracle import metrics_oracle as MO
    MG.import_reference()
    from pht.models.afgsa import metric as ref_metric
    from pht.models.afgsa import util as ref_util
    from pht.models.afgsa.preprocessing import postprocess_specular
    rng = np.random.default_rng(MG.SEED)
    B, H, W = 3, 40, 56
    gt = np.exp(rng.normal(0, 0.6, (B, 3, H, W))).astype(np.float32) * 0.4             # linear radiance
    out_log = (np.log(gt + 1) + rng.normal(0, 0.08, gt.shape)).astype(np.float32)       # denoiser output, log space
    out_log[0, :, 3, 4] = -0.2                                                          # negative radiance after exp-1 -> NaN -> 0
    noisy_log = (np.log(gt * rng.gamma(2.0, 0.5, gt.shape) + 1)).astype(np.float32)
    out_img = ref_util.tensor2img(out_log, post_spec=True)
    gt_img = ref_util.tensor2img(gt)
    noisy_img = ref_util.tensor2img(noisy_log, post_spec=True)
    out_lin = postprocess_specular(out_log)
    res = {"psnr_out": float(ref_metric.calculate_psnr(out_img.copy(), gt_img.copy())),
           "ssim_out": float(ref_metric.calculate_ssim(out_img.copy(), gt_img.copy())),
           "psnr_noisy": float(ref_metric.calculate_psnr(noisy_img.copy(), gt_img.copy())),
           "ssim_noisy": float(ref_metric.calculate_ssim(noisy_img.copy(), gt_img.copy())),
           "mrse_out": float(ref_metric.calculate_rmse(out_lin.copy(), gt.copy()))}
    # the oracle restatement agrees with the reference
    o_out, o_gt = MO.tensor2img(out_log, True), MO.tensor2img(gt)
    pins = {"tensor2img_maxdiff": int(np.abs(o_out.astype(int) - out_img.astype(int)).max()),
            "tensor2img_gt_maxdiff": int(np.abs(o_gt.astype(int) - gt_img.astype(int)).max()),
            "psnr_diff": abs(MO.psnr(out_img, gt_img) - res["psnr_out"]),
            "ssim_diff": abs(MO.ssim(out_img, gt_img) - res["ssim_out"]),
            "mrse_rel_diff": abs(MO.rmse(out_lin, gt) - res["mrse_out"]) / res["mrse_out"]}
    assert pins["tensor2img_maxdiff"] == 0 and pins["tensor2img_gt_maxdiff"] == 0, pins
    assert pins["psnr_diff"] < 1e-9 and pins["ssim_diff"] < 1e-9 and pins["mrse_rel_diff"] < 1e-6, pins
    res["pins"] = pins
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), out_log=out_log, gt=gt, noisy_log=noisy_log, out_img=out_img,
                        gt_img=gt_img, noisy_img=noisy_img)
    with open(os.path.join(HERE, "metrics.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
