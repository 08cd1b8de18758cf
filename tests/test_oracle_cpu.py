"""The CPU oracle against the golden fixtures generated from the real reference
(tests/golden/make_golden.py).  No GPU needed."""
import random

import numpy as np
import pytest
import torch

from conftest import load_json, load_npz
from oracle import afgsa_oracle as O
from oracle import sampler_oracle as S


def _ref_init_state_dict(mode="replicate"):
    """The reference's random init under the reference seed, reproduced by the product model's
    parameter containers (pinned by test_host_cpu.test_param_init_matches_reference)."""
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    torch.manual_seed(990819)
    net = AFGSANet(3, 7, 256, num_gcp=0, padding_mode=mode)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def test_pins_recorded(golden_meta):
    pins = golden_meta["pins"]
    assert pins["sampler_bit_exact"] is True
    assert pins["zorder_vs_raster_maxdiff"] == 0.0
    for k, v in pins.items():
        if k.endswith("maxdiff") or k.endswith("diff"):
            assert v < 2e-6, (k, v)


def test_preprocess_matches_reference():
    g = load_npz("preprocess.npz")
    n, t, a = O.preprocess_batch(torch.from_numpy(g["noisy_hwc"]), torch.from_numpy(g["gt_hwc"]),
                                 torch.from_numpy(g["aux_hwc"]))
    assert np.abs(n.numpy() - g["noisy"]).max() < 1e-6
    assert np.abs(t.numpy() - g["gt"]).max() < 1e-6
    assert np.abs(a.numpy() - g["aux"]).max() < 1e-6
    assert not np.isnan(a.numpy()).any()


def test_afgsa_module_matches_reference():
    g = load_npz("afgsa_module.npz")
    sd = {k.replace("__", "."): torch.from_numpy(v) for k, v in g.items() if k.startswith("p__")}
    out = O.afgsa(torch.from_numpy(g["noisy"]), torch.from_numpy(g["aux"]), sd, "p.", 8, 3, 4)
    assert (out - torch.from_numpy(g["out"])).abs().max() < 1e-5


@pytest.mark.parametrize("mode", ["replicate", "reflect"])
def test_net_forward_backward_matches_reference(mode):
    g = load_npz(f"net_{mode}.npz")
    grads_ref = load_json(f"net_{mode}_grads.json")
    sd = _ref_init_state_dict(mode)
    x, aux, gt = (torch.from_numpy(g[k]) for k in ("x", "aux", "gt"))
    out, loss, grads = O.g_only_train_step(x, aux, gt, sd, mode)
    assert (out - torch.from_numpy(g["out"])).abs().max() < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for name, ref in grads_ref.items():
        gr = grads[name].flatten()
        scale = ref["absmax"] + 1e-30
        probe = torch.tensor([float(gr[i]) for i in ref["probe_idx"]])
        assert (probe - torch.tensor(ref["probe"])).abs().max() / scale < 1e-4, name
        assert abs(float(gr.double().abs().sum()) - ref["abssum"]) / (ref["abssum"] + 1e-30) < 1e-4, name


def test_mt19937_matches_cpython(golden_meta):
    kat = golden_meta["mt_kat"]
    m = S.MT19937(990819)
    assert [m.getrandbits(32) for _ in range(3)] == kat["getrandbits32"]
    assert [m.randint(0, 479) for _ in range(4)] == kat["randint_0_479"]
    assert m.random() == kat["random"]
    for seed in (0, 1, 990819, 2 ** 32 + 5, 123456789012345):
        a, b = S.MT19937(seed), random.Random(seed)
        for n in (1, 2, 3, 100, 479, 895, 1 << 20):
            assert a.randint(0, n) == b.randint(0, n)
        assert a.random() == b.random()
        assert [a.getrandbits(32) for _ in range(700)] == [b.getrandbits(32) for _ in range(700)]


def test_sampler_matches_reference_golden():
    g = load_npz("sampler.npz")
    for key, ref in g.items():
        parts = key.split("_")
        h, w, p, n = int(parts[0][1:]), int(parts[1][1:]), int(parts[2][1:]), int(parts[3][1:])
        if n > 200:
            continue  # the 400-patch case is covered on the GPU; keep the CPU suite fast
        pts = S.dart_throwing((h, w), p, n, S.MT19937(990819))
        assert np.array_equal(pts, ref.astype(np.int64)), key


def test_adam_oracle_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(1000)
        ref.grad = g.clone()
        opt.step()
        O.adam_step(p, g, m, v, step, 1e-3)
    assert (p - ref.detach()).abs().max() < 1e-6


def test_importance_map_and_sampling_match_reference_golden():
    """oracle importance map / prune (scipy uniform_filter restated in numpy) == fixtures from the real reference, bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from importance_inputs import CASES, SEEDS, frames
    g = load_npz("importance.npz")
    for name, (h, w, p, n) in CASES.items():
        noisy, normal, _ = frames(h, w)
        imp = S.importance_map(noisy, normal, p)
        assert np.array_equal(imp, g[f"{name}__imp"]), name
        for seed in SEEDS[:2]:
            kept = S.importance_sampling(noisy, normal, p, n, S.MT19937(seed), imp=imp)
            assert np.array_equal(kept, g[f"{name}__seed{seed}"].astype(np.int64)), (name, seed)
            kept2 = S.importance_sampling(noisy, normal, p, n, random.Random(seed), imp=imp)
            assert np.array_equal(kept, kept2)


def test_uniform_filter_restatement_matches_scipy():
    ndimage = pytest.importorskip("scipy.ndimage")
    rs = np.random.default_rng(0)
    a = rs.standard_normal((37, 53, 3)).astype(np.float32)
    for size in (1, 2, 5, 8, 32):
        assert np.array_equal(S.uniform_filter_pp1(a, size), ndimage.uniform_filter(a, size=(size, size, 1))), size


def test_film_variant_matches_reference_golden():
    """use_film=True (AFGSA.forward's FiLM branch, model.py:458-460, film.py:36-45): parameter names / shapes / seeded
    init of the product containers and the oracle's forward + gradients against the fixture generated from the real
    reference (tests/golden/make_golden_film.py)."""
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    meta = load_json("net_film_meta.json")
    g = load_npz("net_film.npz")
    gref = load_json("net_film_grads.json")
    torch.manual_seed(990819)
    net = AFGSANet(3, 7, 256, num_sa=2, num_gcp=0, padding_mode="replicate", use_film=True)
    assert [k for k, _ in net.named_parameters()] == meta["param_order"]
    for k, p in net.named_parameters():
        assert list(p.shape) == meta["param_shapes"][k]
        assert abs(float(p.detach().double().sum()) - meta["param_checksums"][k][0]) <= 1e-6 * max(1.0, abs(meta["param_checksums"][k][0]))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x, aux, gt = (torch.from_numpy(g[k]) for k in ("x", "aux", "gt"))
    out, loss, grads = O.g_only_train_step(x, aux, gt, sd, "replicate", num_sa=2)
    assert float((out - torch.from_numpy(g["out"])).abs().max()) < 2e-6
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    for name, r in gref.items():
        gr = grads[name].flatten()
        probe = torch.tensor([float(gr[i]) for i in r["probe_idx"]])
        assert float((probe - torch.tensor(r["probe"])).abs().max()) <= 1e-5 * (r["absmax"] + 1e-30) + 1e-12, name
        assert abs(float(gr.double().abs().sum()) - r["abssum"]) <= 1e-5 * (r["abssum"] + 1e-30) + 1e-12, name


def test_metrics_oracle_matches_reference_golden():
    """tensor2img / PSNR / SSIM / MRSE restatements (oracle/metrics_oracle.py) against the values the real reference
    functions produced (tests/golden/make_golden_metrics.py)."""
    from oracle import metrics_oracle as MO
    g, r = load_npz("metrics.npz"), load_json("metrics.json")
    out_img, gt_img = MO.tensor2img(g["out_log"], True), MO.tensor2img(g["gt"])
    assert np.array_equal(out_img, g["out_img"]) and np.array_equal(gt_img, g["gt_img"])
    assert np.array_equal(MO.tensor2img(g["noisy_log"], True), g["noisy_img"])
    assert abs(MO.psnr(out_img, gt_img) - r["psnr_out"]) < 1e-9
    assert abs(MO.ssim(out_img, gt_img) - r["ssim_out"]) < 1e-9
    assert abs(MO.ssim(g["noisy_img"], gt_img) - r["ssim_noisy"]) < 1e-9
    assert abs(MO.rmse(np.exp(g["out_log"]) - 1, g["gt"]) - r["mrse_out"]) < 1e-6 * r["mrse_out"]


def test_oracle_matches_the_reference_at_the_dev_baseline_shape():
    """tests/golden/make_golden_shapes.py: the REAL reference's output / loss / gradient norms at dev (8 x 32 x 32);
    stag / prod are pinned the same way by the generating script (``pins`` in the fixture) and re-checked on the GPU box."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    from make_golden_shapes import checksum, shape_inputs
    meta = load_json("net_shapes_meta.json")
    for k, v in meta["pins"].items():
        assert v < (1e-5 if "out" in k or "loss" in k else 5e-3), (k, v)
    x, gt, aux = shape_inputs("dev")
    assert abs(checksum(x, gt, aux) - meta["inputs_checksum"]["dev"]) <= 1e-9 * meta["inputs_checksum"]["dev"]
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    torch.manual_seed(990819)
    sd = {k: v.detach().clone() for k, v in AFGSANet(3, 7, 256, num_gcp=0, padding_mode="replicate").state_dict().items()}
    out, loss, grads = O.g_only_train_step(x, aux, gt, sd, "replicate")
    ref = torch.from_numpy(load_npz("net_shapes.npz")["dev_out"])
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-5
    assert abs(float(loss) - meta["dev_loss"]) / meta["dev_loss"] < 1e-5
    for n, r in meta["grads"]["dev"].items():
        assert abs(float(grads[n].double().norm()) - r["l2"]) / (r["l2"] + 1e-30) < 5e-3, n


def test_oracle_reference_init_reproduces_the_reference_parameters(golden_meta):
    """oracle.reference_init_state_dict (plain torch.nn modules in the reference's creation order) == the reference's own
    random init under seed 990819 (checksums recorded from the real module) == the product model's parameter containers."""
    sd = O.reference_init_state_dict(golden_meta["seed"])
    assert list(sd) and set(sd) == set(golden_meta["param_checksums"])
    for k, (s1, s2, first) in golden_meta["param_checksums"].items():
        v = sd[k]
        assert list(v.shape) == golden_meta["param_shapes"][k], k
        assert abs(float(v.double().sum()) - s1) < 1e-9 + 1e-12 * abs(s1) and abs(float(v.double().abs().sum()) - s2) < 1e-9 + 1e-12 * s2, k
        assert float(v.flatten()[0]) == first, k
    prod = _ref_init_state_dict()
    for k, v in sd.items():
        assert torch.equal(v, prod[k]), k


def test_msssim_oracle_basic_properties():
    """oracle.msssim_oracle (parity unpinned: kornia is not importable): identical images give zero loss, the window bank
    is normalised and ordered [sigma_i x 3], the reference wrapper's scale only depends on the target and is >= 1, and
    the loss grows with the distortion."""
    from oracle import msssim_oracle as M
    g = M.window_bank()
    assert g.shape == (15, 1, 33, 33)
    assert torch.allclose(g.sum((1, 2, 3)), torch.ones(15), atol=1e-6)
    assert torch.equal(g[0], g[2]) and not torch.equal(g[2], g[3])
    torch.manual_seed(0)
    y = torch.rand(2, 3, 40, 48) * 3.0
    assert float(M.ssim_loss(y.clone(), y)) == pytest.approx(0.0, abs=1e-5)
    small = float(M.ssim_loss(y + 0.01 * torch.randn_like(y), y))
    big = float(M.ssim_loss(y + 0.3 * torch.randn_like(y), y))
    assert 0.0 < small < big
    x = (y + 0.1 * torch.randn_like(y)).requires_grad_(True)
    M.ssim_loss(x, y).backward()
    assert torch.isfinite(x.grad).all() and float(x.grad.abs().max()) > 0
