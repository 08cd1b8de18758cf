"""End-to-end parity of the B200 generator (through the C ABI) against the reference golden vectors
and the CPU oracle: forward, backward, optimiser trajectory, tiled inference, trainer step."""
import math

import numpy as np
import pytest
import torch

from conftest import load_json, load_npz
from oracle import afgsa_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
# Forward outputs and losses are held to north_star's 1e-5 relative.
# Gradients pass through ~45 ReLU / LeakyReLU kinks and the L1 sign.  A pre-activation that lies within fp32
# rounding of zero can fall on the other side of the kink under a different (equally valid) fp32 summation
# order; ONE such flip moves the downstream gradient by ~2e-3 in relative L2 norm (measured on the B200 with
# tools/diag_decoder.py: 1 flip among 491,520 decoder activations -> 2.1e-3).  Hence two kinds of checks:
#   * STRICT (no flip occurs for these seeded inputs): every gradient tensor within 2e-5 of the oracle;
#   * FLIP-TOLERANT (golden inputs from the reference run): 5e-3 in relative L2 / of the tensor's max.
STRICT_TOL = 2e-5
FLIP_TOL = 5e-3


@pytest.fixture
def reproducible_attention_backward():
    """Bit-identity tests: the attention backward's fixed-order (scratch + fold) accumulation instead of the default
    arrival-order vector reductions, whose last bf16 bits depend on the order (option "attn_bwd_direct")."""
    from pixel_heal_thyself_b200 import _lib
    assert _lib.lib.pht_set_option(b"attn_bwd_direct", 0) == 0
    yield
    _lib.lib.pht_set_option(b"attn_bwd_direct", 1)


def make_net(mode="replicate", dtype="fp32", num_sa=5, seed=990819):
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    torch.manual_seed(seed)
    return AFGSANet(3, 7, 256, num_sa=num_sa, num_gcp=0, padding_mode=mode, compute_dtype=dtype).to(DEV)


def psnr(a, b):
    mse = float(((a - b) ** 2).mean())
    return 10 * math.log10(float(b.max() - b.min()) ** 2 / max(mse, 1e-20))


@pytest.mark.parametrize("mode", ["replicate", "reflect"])
def test_fp32_forward_backward_matches_reference_golden(mode):
    """fp32 parity mode vs the REAL reference's output / loss / gradients on the same inputs and the same
    random-init weights (north_star: within 1e-5 relative in fp32)."""
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    g = load_npz(f"net_{mode}.npz")
    gref = load_json(f"net_{mode}_grads.json")
    net = make_net(mode, "fp32")
    x, aux, gt = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "aux", "gt"))
    out = net(x, aux)
    ref = torch.from_numpy(g["out"]).to(DEV)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-5
    loss = L1ReconstructionLoss()(out, gt)
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-5
    loss.backward()
    worst = 0.0
    for name, p in net.named_parameters():
        r = gref[name]
        gr = p.grad.flatten()
        probe = torch.tensor([float(gr[i]) for i in r["probe_idx"]])
        e = float((probe - torch.tensor(r["probe"])).abs().max() / (r["absmax"] + 1e-30))
        e = max(e, abs(float(gr.double().abs().sum()) - r["abssum"]) / (r["abssum"] + 1e-30))
        e = max(e, abs(float(gr.abs().max()) - r["absmax"]) / (r["absmax"] + 1e-30))
        worst = max(worst, e)
        assert e < FLIP_TOL, (name, e)
    print(f"worst relative gradient deviation vs reference ({mode}): {worst:.2e}")


def _grad_errors(num_sa, B, H, W, mode, seed):
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    net = make_net(mode, "fp32", num_sa=num_sa)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    torch.manual_seed(seed)
    x, aux, gt = torch.randn(B, 3, H, W) * 0.5, torch.rand(B, 7, H, W), torch.randn(B, 3, H, W) * 0.5
    # fp64 oracle = ground truth (a CPU fp32 run has its own, different, kink flips)
    sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
    o_out, o_loss, o_grads = O.g_only_train_step(x.double(), aux.double(), gt.double(), sd64, mode, num_sa=num_sa)
    out = net(x.to(DEV), aux.to(DEV))
    loss = L1ReconstructionLoss()(out, gt.to(DEV))
    loss.backward()
    assert float((out.detach().cpu() - o_out).abs().max() / o_out.abs().max()) < 1e-5
    assert abs(float(loss) - float(o_loss)) / float(o_loss) < 1e-5
    errs = {}
    for n, p in net.named_parameters():
        ref = o_grads[n]
        got = p.grad.cpu().double()
        errs[n] = (float((got - ref).abs().max() / (ref.abs().max() + 1e-300)),
                   float((got - ref).norm() / (ref.norm() + 1e-300)))
    return errs


@pytest.mark.parametrize("num_sa,B,H,W,mode", [(1, 2, 24, 40, "replicate"), (2, 1, 16, 16, "replicate"),
                                                (1, 1, 16, 24, "reflect")])
def test_fp32_all_gradients_strict(num_sa, B, H, W, mode):
    """Ragged batches (non-square, edge blocks on every side) against the CPU oracle: EVERY gradient tensor in
    full, at fp32-rounding tolerance (these seeded inputs have no activation within rounding of a ReLU kink)."""
    errs = _grad_errors(num_sa, B, H, W, mode, seed=3)
    worst = max(errs.items(), key=lambda kv: kv[1][0])
    assert worst[1][0] < STRICT_TOL, worst


def test_fp32_gradients_deep_ragged_flip_tolerant():
    """Two blocks, batch 2, 24x40: one decoder activation sits within fp32 rounding of zero for this input
    (see the note at the top of this file), so the bound is the flip-tolerant one."""
    errs = _grad_errors(2, 2, 24, 40, "replicate", seed=3)
    for n, (emax, el2) in errs.items():
        assert el2 < FLIP_TOL and emax < 10 * FLIP_TOL, (n, emax, el2)


def test_bf16_forward_close_to_reference():
    """Production dtype: PSNR of the bf16 output against the reference fp32 output (north_star: PSNR delta
    < 0.05 dB on the denoised image <=> the bf16-vs-fp32 error is far below the denoising error)."""
    g = load_npz("net_replicate.npz")
    net = make_net("replicate", "bf16")
    x, aux, gt = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "aux", "gt"))
    out = net(x, aux)
    ref = torch.from_numpy(g["out"]).to(DEV)
    p_impl = psnr(out, ref)
    d_ref, d_out = psnr(ref, gt), psnr(out, gt)
    print(f"bf16 vs fp32-reference PSNR {p_impl:.1f} dB; PSNR vs gt: reference {d_ref:.3f} dB, bf16 {d_out:.3f} dB")
    assert p_impl > 40.0
    assert abs(d_ref - d_out) < 0.05


def test_bf16_gradients_close_to_fp32():
    g = load_npz("net_replicate.npz")
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    x, aux, gt = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "aux", "gt"))
    grads = {}
    for dt in ("fp32", "bf16"):
        net = make_net("replicate", dt)
        L1ReconstructionLoss()(net(x, aux), gt).backward()
        grads[dt] = torch.cat([p.grad.flatten() for p in net.parameters()])
    cos = torch.nn.functional.cosine_similarity(grads["fp32"], grads["bf16"], dim=0)
    print(f"bf16 vs fp32 full-gradient cosine similarity {float(cos):.5f}")
    assert float(cos) > 0.98


def test_three_adam_steps_match_reference_losses(golden_meta):
    """G-only training trajectory (L1 + Adam 1e-4) vs the reference's own three losses."""
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    from pixel_heal_thyself_b200.optim import FlatAdam
    g = load_npz("net_replicate.npz")
    net = make_net("replicate", "fp32")
    opt = FlatAdam(net, lr=1e-4)
    x, aux, gt = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "aux", "gt"))
    l1 = L1ReconstructionLoss()
    losses = []
    for _ in range(3):
        opt.zero_grad()
        ls = l1(net(x, aux), gt)
        ls.backward()
        opt.step()
        losses.append(float(ls))
    ref = golden_meta["adam_l1_losses_3steps"]
    assert max(abs(a - b) / b for a, b in zip(losses, ref)) < 1e-4, (losses, ref)
    w = dict(net.named_parameters())["decoder.2.0.weight"].detach().flatten().cpu()
    idx = [(i * 2654435761 + 12345) % w.numel() for i in range(8)]
    probe = torch.tensor([float(w[i]) for i in idx])
    assert (probe - torch.tensor(golden_meta["adam_decoder2_weight_probe"])).abs().max() < 1e-5


def test_grad_accumulation_and_zero_grad_modes():
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    net = make_net("replicate", "fp32", num_sa=1)
    torch.manual_seed(0)
    x, aux, gt = torch.randn(1, 3, 16, 16, device=DEV), torch.rand(1, 7, 16, 16, device=DEV), torch.randn(1, 3, 16, 16, device=DEV)
    l1 = L1ReconstructionLoss()
    l1(net(x, aux), gt).backward()
    g1 = torch.cat([p.grad.flatten().clone() for p in net.parameters()])
    l1(net(x, aux), gt).backward()          # accumulates: 2x
    g2 = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(g2, 2 * g1, rtol=1e-5, atol=1e-8)
    net.zero_grad(set_to_none=False)
    l1(net(x, aux), gt).backward()
    g3 = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(g3, g1, rtol=1e-5, atol=1e-8)


def test_state_dict_roundtrip_and_eval_matches_train_forward():
    net = make_net("replicate", "fp32", num_sa=2)
    torch.manual_seed(1)
    x, aux = torch.randn(1, 3, 16, 24, device=DEV), torch.rand(1, 7, 16, 24, device=DEV)
    y_train = net(x, aux).detach()
    with torch.no_grad():
        y_eval = net.eval()(x, aux)
    assert torch.equal(y_train, y_eval)
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    net2 = make_net("replicate", "fp32", num_sa=2, seed=1)
    assert not torch.equal(net2(x, aux).detach(), y_train)
    net2.load_state_dict(sd)
    assert torch.equal(net2(x, aux).detach(), y_train)
    with pytest.raises(AssertionError, match="divisible by the block size"):
        net(torch.zeros(1, 3, 12, 16, device=DEV), torch.zeros(1, 7, 12, 16, device=DEV))


def test_tiled_inference_is_exact_with_48px_halo():
    from pixel_heal_thyself_b200.inference import denoise_frame
    net = make_net("replicate", "fp32").eval()
    torch.manual_seed(2)
    x, aux = torch.randn(1, 3, 160, 224, device=DEV) * 0.5, torch.rand(1, 7, 160, 224, device=DEV)
    with torch.no_grad():
        full = net(x, aux)
    tiled = denoise_frame(net, x, aux, rows=2, cols=2, halo=48)
    assert float((full - tiled).abs().max()) < 2e-5 * float(full.abs().max())
    loose = denoise_frame(net, x, aux, rows=2, cols=2, halo=8)
    assert float((full - loose).abs().max()) > float((full - tiled).abs().max())


def test_trainer_gan_step_runs_and_learns():
    """One full GAN iteration (base_trainer.py:388-457) through the trainer API, then G-only steps reduce L1."""
    from pixel_heal_thyself_b200.config import load_config
    from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer
    cfg = load_config("ci", ["data.synthetic.num_images=1", "data.synthetic.height=128", "data.synthetic.width=128",
                             "data.patches.num_patches=16"])
    tr = AFGSATrainer(cfg)
    tr.setup()
    ds = tr.setup_data()
    assert 8 <= len(ds) <= 16 and len(ds) == int(ds.counts.sum())   # importance pruning keeps a subset of the 16 darts
    idx = torch.arange(2, device=tr.device)
    noisy, gt, aux = ds.batch_device(idx)
    assert noisy.shape == (2, 3, 32, 32) and aux.shape == (2, 7, 32, 32) and not torch.isnan(aux).any()
    g0, d0 = tr.train_step(noisy, gt, aux)
    assert torch.isfinite(g0) and torch.isfinite(d0)
    # the host path (pinned NHWC patches -> H2D -> pht_preprocess) yields the same batch
    from pixel_heal_thyself_b200.data import preprocess_host_batch
    hp = ds.host_patches()
    n2, g2, a2 = preprocess_host_batch({k: v[:2] for k, v in hp.items()}, tr.device)
    assert torch.equal(n2, noisy) and torch.equal(g2, gt) and torch.equal(a2, aux)
    tr2 = AFGSATrainer(cfg)
    tr2.setup(g_only=True)
    # (Adam's first updates overshoot on this 2-patch batch: 0.344 -> 0.736 -> 0.358 ... -> 0.330 at step 7; the sixth
    # step sits within 5e-4 of the first, so compare the best of the last steps with a margin instead of one marginal pair)
    losses = [float(tr2.train_step(noisy, gt, aux)[0]) for _ in range(8)]
    assert min(losses[-3:]) < losses[0] - 5e-3, losses


def test_launch_options_are_bitwise_neutral(reproducible_attention_backward):
    """Programmatic dependent launch, the serpentine tile order and the strip tiles of the fused pad-fold
    data-gradient only change WHEN and WHERE a tile is computed, never its arithmetic: output, loss and every
    gradient must be bit-identical with the options on and off (bf16 production path, 2 blocks, 64x64)."""
    from pixel_heal_thyself_b200 import _lib
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    torch.manual_seed(7)
    x = (torch.randn(2, 3, 64, 64) * 0.5).to(DEV)
    aux = torch.rand(2, 7, 64, 64).to(DEV)
    gt = (torch.randn(2, 3, 64, 64) * 0.5).to(DEV)
    results = []
    try:
        for on in (1, 0, 1):
            for name in (b"pdl", b"serpentine", b"strips"):
                assert _lib.lib.pht_set_option(name, on) == 0
            # (the CTA-pair GEMMs -- tcgen05 cta_group::2, default off -- ride along with the "off" pass: same arithmetic)
            assert _lib.lib.pht_set_option(b"cta_pairs", 1 - on) == 0
            net = make_net("replicate", "bf16", num_sa=2)
            out = net(x, aux)
            loss = L1ReconstructionLoss()(out, gt)
            loss.backward()
            torch.cuda.synchronize()
            results.append((out.detach().clone(), float(loss), [p.grad.detach().clone() for p in net.parameters()]))
    finally:
        for name in (b"pdl", b"serpentine", b"strips"):
            _lib.lib.pht_set_option(name, 1)
        _lib.lib.pht_set_option(b"cta_pairs", 0)
    for other in results[1:]:
        assert torch.equal(results[0][0], other[0])
        assert results[0][1] == other[1]
        for a, b in zip(results[0][2], other[2]):
            assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["replicate", "reflect"])
def test_fused_padding_frames_are_bitwise_neutral(mode, reproducible_attention_backward):
    """bf16 path: the padding frames written by the producing GEMM's epilogue (PHT_EPI_RING*) instead of pht_border_fill
    launches -- output, loss and every gradient bit-identical with and without (2 blocks, 32 x 48, both padding modes)."""
    from pixel_heal_thyself_b200 import _lib
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    torch.manual_seed(11)
    x = (torch.randn(2, 3, 32, 48) * 0.5).to(DEV)
    aux = torch.rand(2, 7, 32, 48).to(DEV)
    gt = (torch.randn(2, 3, 32, 48) * 0.5).to(DEV)
    results = []
    for fused in (True, False):
        net = make_net(mode, "bf16", num_sa=2)
        net.engine.no_fused_ring = not fused
        before = _lib.counters()["other"]
        out = net(x, aux)
        launches = _lib.counters()["other"] - before
        loss = L1ReconstructionLoss()(out, gt)
        loss.backward()
        torch.cuda.synchronize()
        results.append((out.detach().clone(), float(loss), [p.grad.detach().clone() for p in net.parameters()], launches))
    assert results[0][3] == results[1][3] - 6        # 2 blocks: X1p x 2, H1p x 2, the last block's output, D1p
    assert torch.equal(results[0][0], results[1][0]) and results[0][1] == results[1][1]
    for a, b in zip(results[0][2], results[1][2]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_film_variant_matches_reference_golden(dtype):
    """use_film=True: the FiLM branch of the AFGSA layer (two 1x1 GEMMs + pht_film_fwd / pht_film_bwd) against the real
    reference's output / loss / gradients (tests/golden/make_golden_film.py).  fp32 mode: 1e-5 output, flip-tolerant
    gradients; bf16: the production tolerances."""
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    from pixel_heal_thyself_b200.models.losses import L1ReconstructionLoss
    g = load_npz("net_film.npz")
    gref = load_json("net_film_grads.json")
    torch.manual_seed(990819)
    net = AFGSANet(3, 7, 256, num_sa=2, num_gcp=0, padding_mode="replicate", use_film=True, compute_dtype=dtype).to(DEV)
    x, aux, gt = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "aux", "gt"))
    out = net(x, aux)
    ref = torch.from_numpy(g["out"]).to(DEV)
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < (1e-5 if dtype == "fp32" else 3e-2), err
    loss = L1ReconstructionLoss()(out, gt)
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < (1e-5 if dtype == "fp32" else 2e-2)
    loss.backward()
    tol = FLIP_TOL if dtype == "fp32" else 8e-2
    for name, p in net.named_parameters():
        r = gref[name]
        gr = p.grad.flatten()
        if r["absmax"] == 0.0:                     # alpha: registered by the reference, unused by its forward
            assert float(gr.abs().max()) == 0.0, name
            continue
        e = abs(float(gr.double().abs().sum()) - r["abssum"]) / (r["abssum"] + 1e-30)
        if dtype == "fp32":
            probe = torch.tensor([float(gr[i]) for i in r["probe_idx"]])
            e = max(e, float((probe - torch.tensor(r["probe"])).abs().max() / (r["absmax"] + 1e-30)))
        assert e < tol, (name, e)


def test_validation_pass_writes_the_reference_evaluation_line(tmp_path):
    """BaseTrainer._validate_and_save: checkpoint names (base_trainer.py:521-533) and the evaluation.txt line of
    base_trainer.py:591-595 with MRSE / PSNR / 1-SSIM from the GPU metrics."""
    import os
    import re
    from pixel_heal_thyself_b200.config import load_config
    from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer
    cfg = load_config("ci", ["data.synthetic.num_images=1", "data.synthetic.height=128", "data.synthetic.width=128",
                             "data.patches.num_patches=16"])
    tr = AFGSATrainer(cfg)
    tr.setup(g_only=True)
    ds = tr.setup_data()
    n = len(ds)
    tr._validate_and_save(0, ds, n - 2, 2, str(tmp_path))
    assert os.path.exists(tmp_path / "model_epoch1" / "G.pt")
    line = open(tmp_path / "evaluation.txt").read()
    m = re.fullmatch(r"Validation: 1 \tAvg MRSE: (\d+\.\d{4}) \tAvg PSNR: (\d+\.\d{4}) \tAvg 1-SSIM: (-?\d+\.\d{4})\n", line)
    assert m, line
    mrse, psnr, one_minus_ssim = (float(v) for v in m.groups())
    assert mrse > 0 and 0 < psnr < 100 and 0 <= one_minus_ssim <= 2


def test_critic_step_graph_replay_trains_like_eager(monkeypatch):
    """The critic step (base_trainer.py:391-412, stays PyTorch) replayed as a CUDA graph: the same update rule as the
    eager step -- with the gradient-penalty weight at 0 (no RNG in the step) the two d_loss trajectories must agree to
    fp32 round-off; with the penalty on they stay finite and the critic's weights move."""
    from pixel_heal_thyself_b200.config import load_config
    from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer
    torch.manual_seed(3)
    fake = torch.rand(4, 3, 32, 32, device=DEV)
    gt = torch.rand(4, 3, 32, 32, device=DEV)

    def run(graph, gp_w):
        monkeypatch.setenv("PHT_CRITIC_GRAPH", "1" if graph else "0")
        cfg = load_config("ci", ["data.synthetic.num_images=1", "data.synthetic.height=128", "data.synthetic.width=128",
                                 "data.patches.num_patches=16", f"model.losses.gp_loss_w={gp_w}"])
        tr = AFGSATrainer(cfg)
        tr.setup()
        w0 = next(tr.D.parameters()).detach().clone()
        losses = [float(tr._critic_step(fake, gt)) for _ in range(5)]
        return losses, float((next(tr.D.parameters()).detach() - w0).abs().max()), getattr(tr, "_critic_graph", None)

    eager, moved_e, _ = run(False, 0.0)
    graph, moved_g, st = run(True, 0.0)
    assert st is not None, "the critic step did not take the CUDA-graph path"
    assert max(abs(a - b) for a, b in zip(eager, graph)) < 1e-4 * max(1.0, max(abs(a) for a in eager)), (eager, graph)
    assert moved_e > 0 and abs(moved_e - moved_g) < 1e-3 * moved_e + 1e-7
    with_gp, moved, st = run(True, 10.0)
    assert st is not None and all(math.isfinite(v) for v in with_gp) and moved > 0


def test_device_prefetcher_yields_the_same_batches():
    """DevicePrefetcher (H2D + pht_preprocess of batch i+1 on a side stream behind step i) == preprocess_host_batch."""
    from pixel_heal_thyself_b200.data import DevicePrefetcher, preprocess_host_batch
    g = torch.Generator().manual_seed(5)
    host = [{"noisy": torch.rand(2, 16, 16, 3, generator=g).pin_memory(), "gt": torch.rand(2, 16, 16, 3, generator=g).pin_memory(),
             "aux": (torch.rand(2, 16, 16, 7, generator=g) * 2 - 1).pin_memory()} for _ in range(4)]
    got = []
    for batch in DevicePrefetcher(host, torch.device(DEV)):
        got.append([t.clone() for t in batch])
        torch.cuda.synchronize()
    assert len(got) == 4
    for hb, gb in zip(host, got):
        ref = preprocess_host_batch(hb, torch.device(DEV))
        assert all(torch.equal(a, b) for a, b in zip(ref, gb))
    assert list(DevicePrefetcher([], torch.device(DEV))) == []


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_graph_replayed_train_step_is_bit_identical_to_eager(dtype, reproducible_attention_backward):
    """BaseTrainer.train_step replayed from CUDA graphs (two eager warm-up steps, capture, replays) == the same steps
    launched eagerly: same kernels in the same order.  The bf16 production path is deterministic (fixed-order reductions
    everywhere), so losses and weights are bit-identical; the fp32 parity path's CUDA-core weight-gradient / attention
    kernels accumulate with atomics, so it is held to fp32 round-off.  The launch counters account for the replays."""
    from pixel_heal_thyself_b200 import _lib
    from pixel_heal_thyself_b200.config import load_config
    from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer
    cfg = load_config("dev", ["trainer.batch_size=4", f"model.afgsa.compute_dtype={dtype}"])
    g = torch.Generator().manual_seed(11)
    batches = [((torch.randn(4, 3, 32, 32, generator=g) * 0.5).to(DEV), (torch.randn(4, 3, 32, 32, generator=g) * 0.5).to(DEV),
                torch.rand(4, 7, 32, 32, generator=g).to(DEV)) for _ in range(6)]
    runs = {}
    for graph in (False, True):
        tr = AFGSATrainer(cfg)
        tr.use_step_graph = graph
        tr.setup(g_only=True)
        _lib.lib.pht_reset_counters()
        losses = [float(tr.train_step(n, gt, a)[0]) for (n, gt, a) in batches]
        torch.cuda.synchronize()
        if graph:
            assert tr._step_graph is not None and tr._step_graph["rec"] is not None, "the step was not captured"
            assert tr.opt_g.step_count == 6
        runs[graph] = (losses, tr.G.flat_param.clone(), sum(_lib.counters().values()))
        with torch.no_grad():                                  # an eager forward after replays sees the current weights
            y = tr.G.eval()(batches[0][0], batches[0][2])
        runs[graph] += (y,)
    if dtype == "bf16":
        assert runs[False][0] == runs[True][0], (runs[False][0], runs[True][0])
        assert torch.equal(runs[False][1], runs[True][1])
        assert torch.equal(runs[False][3], runs[True][3])
    else:
        # (two runs of the fp32 path differ by atomic-accumulation round-off, which Adam's first steps amplify: the step-6
        # loss of two EAGER runs already differs by ~2e-5 relative)
        assert max(abs(a - b) / b for a, b in zip(runs[False][0], runs[True][0])) < 2e-4, (runs[False][0], runs[True][0])
        assert float((runs[False][1] - runs[True][1]).norm() / runs[True][1].norm()) < 2e-3
        assert float((runs[False][3] - runs[True][3]).abs().max() / runs[True][3].abs().max()) < 5e-3
    assert runs[False][2] == runs[True][2] > 0, (runs[False][2], runs[True][2])


def test_graph_replayed_full_gan_step_trains():
    """The full iteration (generator + PyTorch critic with gradient penalty) captured as one graph: finite losses, both
    networks move, and a G-only L1 evaluation improves over the steps like the eager trainer's."""
    from pixel_heal_thyself_b200.config import load_config
    from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer
    cfg = load_config("dev", ["trainer.batch_size=4"])
    g = torch.Generator().manual_seed(12)
    n, gt, a = ((torch.randn(4, 3, 32, 32, generator=g) * 0.5).to(DEV), (torch.randn(4, 3, 32, 32, generator=g) * 0.5).to(DEV),
                torch.rand(4, 7, 32, 32, generator=g).to(DEV))
    tr = AFGSATrainer(cfg)
    tr.setup()
    tr.G._flatten()
    w0, d0 = tr.G.flat_param.clone(), next(tr.D.parameters()).detach().clone()
    out = [tr.train_step(n, gt, a) for _ in range(8)]
    torch.cuda.synchronize()
    assert tr._step_graph is not None and tr._step_graph["rec"] is not None, "the GAN step was not captured"
    assert all(math.isfinite(float(gl)) and math.isfinite(float(dl)) for gl, dl in out)
    assert not torch.equal(w0, tr.G.flat_param) and not torch.equal(d0, next(tr.D.parameters()).detach())
    assert float(out[-1][0]) < float(out[0][0]) + 0.5
