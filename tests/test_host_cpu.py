"""Host-side logic that needs no GPU: C-ABI surface, config, model containers, tiling plan, DP plumbing."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from pixel_heal_thyself_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pht_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pht_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in pht_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert raw.pht_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header_sizes(tmp_path):
    """sizeof of every ABI struct as gcc sees the header == sizeof of the ctypes mirror."""
    from pixel_heal_thyself_b200 import _lib
    names = ["pht_view", "pht_conv_gemm_args", "pht_wgrad_args", "pht_attn_args", "pht_attn_bwd_args", "pht_pack_args",
             "pht_wgrad_reduce_job"]
    mirrors = [_lib.PhtView, _lib.ConvGemmArgs, _lib.WgradArgs, _lib.AttnArgs, _lib.AttnBwdArgs, _lib.PackArgs,
               _lib.WgradReduceJob]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pht_b200.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_ops_refuse_cpu_tensors():
    from pixel_heal_thyself_b200 import ops
    a = torch.zeros(16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.l1_loss(a, a, torch.zeros(1))


def test_invalid_arguments_return_error_not_crash():
    from pixel_heal_thyself_b200 import _lib
    rc = _lib.lib.pht_conv_gemm(None, None)
    assert rc == -1 and b"null" in _lib.lib.pht_last_error()
    with pytest.raises(RuntimeError, match="pht_conv_gemm failed"):
        _lib.check(rc, "pht_conv_gemm")


def test_every_documented_option_is_accepted_and_unknown_ones_are_refused():
    """pht_set_option: every option name the header documents is accepted (host-side state only, no GPU needed), an
    unknown name returns PHT_ERR_INVALID with a message."""
    from pixel_heal_thyself_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pht_b200.h")).read()
    doc = hdr[hdr.index("Runtime options"):hdr.index("int pht_set_option")] if "Runtime options" in hdr else hdr
    names = set(re.findall(r'\* "([a-z0-9_]+)" =', doc))
    assert {"pdl", "serpentine", "strips", "cta_pairs", "half_ring", "attn_bwd_direct", "bf16_fallback"} <= names, names
    defaults = {"pdl": 1, "serpentine": 1, "strips": 1, "attn_bwd_direct": 1, "wgrad_split_div": 2}   # (leave the defaults set)
    for n in sorted(names):
        assert _lib.lib.pht_set_option(n.encode(), defaults.get(n, 0)) == 0, n
    assert _lib.lib.pht_set_option(b"no_such_option", 1) == -1
    assert b"unknown option" in _lib.lib.pht_last_error()


def test_view_memo_follows_the_tensor_not_its_id():
    """pht_view structs are memoised per tensor object; a new tensor (even at a recycled id), a different slice or a
    different origin must get its own struct."""
    from pixel_heal_thyself_b200 import _lib
    t = torch.zeros(2, 6, 6, 8)
    v = _lib.view(t)
    assert _lib.view(t) is v and (v.H, v.W, v.C, v.sb, v.sy, v.sx) == (6, 6, 8, 288, 48, 8)
    assert _lib.view(t, 1, 1) is not v and _lib.view(t, 1, 1).oy == 1
    inner = t[:, 1:-1, 1:-1, :]
    vi = _lib.view(inner)
    assert vi.ptr == t.data_ptr() + 4 * (48 + 8) and (vi.H, vi.W, vi.sy) == (4, 4, 48)
    for _ in range(64):                     # short-lived tensors recycle ids: the weak reference must catch it
        u = torch.zeros(1, 2, 2, 4)
        vu = _lib.view(u)
        assert vu.ptr == u.data_ptr() and (vu.H, vu.W, vu.C) == (2, 2, 4)
        del u
    w = torch.zeros(1, 3, 3, 4)
    vw = _lib.view(w)
    w.set_(torch.zeros(1, 5, 5, 4))          # same object, new storage: the pointer guard
    assert _lib.view(w) is not vw and _lib.view(w).H == 5


def test_gan_loss_types_match_reference_golden():
    """GANLoss for every loss_type the reference accepts (losses.py:103-172) against values from the real class."""
    import json
    import importlib.util
    from pixel_heal_thyself_b200.models.losses import GANLoss
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_golden_ganloss.py"))
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ganloss.json")))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert len(gold) == 10
    for key, want in gold.items():
        t, real, disc = key.split("/")
        x = mk.scores(t).requires_grad_(True)
        crit = GANLoss(t)
        loss = crit(x, bool(int(real)), disc == "True") if t == "hinge" else crit(x, bool(int(real)))
        (grad,) = torch.autograd.grad(loss, x)
        assert abs(float(loss.detach()) - want["loss"]) <= 1e-6 * max(1.0, abs(want["loss"])), key
        assert torch.allclose(grad.reshape(-1), torch.tensor(want["grad"]), rtol=1e-5, atol=1e-7), key
    with pytest.raises(NotImplementedError):
        GANLoss("ragan")


def test_param_init_matches_reference(golden_meta):
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    torch.manual_seed(golden_meta["seed"])
    net = AFGSANet(3, 7, 256, num_gcp=0, padding_mode="replicate")
    names = [n for n, _ in net.named_parameters()]
    assert names == golden_meta["param_order"]
    assert sum(p.numel() for p in net.parameters()) == golden_meta["num_params"] == 9282691
    for n, p in net.named_parameters():
        assert list(p.shape) == golden_meta["param_shapes"][n]
        s, a, first = golden_meta["param_checksums"][n]
        assert float(p.detach().flatten()[0]) == first, n
        assert abs(float(p.detach().double().sum()) - s) < 1e-9 * max(1.0, abs(s)), n
    bufs = {k: v.tolist() for k, v in net.named_buffers()}
    assert bufs == golden_meta["buffers"]


def test_curve_indices_are_permutations():
    from pixel_heal_thyself_b200.models.afgsa.model import CurveOrder, make_curve_indices
    for mode in CurveOrder:
        idx = make_curve_indices(8, mode)
        assert sorted(idx.tolist()) == list(range(64))
    assert make_curve_indices(8, CurveOrder.ZORDER)[:4].tolist() == [0, 1, 8, 9]
    assert make_curve_indices(8, CurveOrder.HILBERT)[0].item() == 0


def test_generator_needs_cuda():
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    net = AFGSANet(3, 7, 256, num_sa=1, num_gcp=0, padding_mode="replicate")
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, 7, 8, 8))
    film = AFGSANet(3, 7, 256, num_sa=1, num_gcp=0, padding_mode="replicate", use_film=True)   # FiLM variant: same rule
    assert "transformer_blocks.0.attention.film.affine.2.weight" in dict(film.named_parameters())
    with pytest.raises(RuntimeError, match="CUDA"):
        film(torch.zeros(1, 3, 8, 8), torch.zeros(1, 7, 8, 8))


def test_config_presets_and_overrides():
    from pixel_heal_thyself_b200.config import load_config
    prod, stag, dev, ci = (load_config(n) for n in ("prod", "stag", "dev", "ci"))
    assert (prod.data.patches.patch_size, prod.data.patches.num_patches) == (128, 400)
    assert (stag.data.patches.patch_size, stag.data.patches.num_patches, stag.data.images.scale) == (64, 200, 0.5)
    assert (dev.data.patches.patch_size, dev.data.patches.num_patches) == (32, 100)
    assert (ci.trainer.batch_size, ci.trainer.epochs) == (2, 2)
    assert prod.seed == 990819 and prod.trainer.batch_size == 8 and prod.trainer.deterministic
    assert prod.model.losses.l1_loss_w == 1.0 and prod.model.losses.gan_loss_w == 0.005
    c = load_config("dev", ["trainer.batch_size=4", "+model.afgsa.compute_dtype=fp32", "seed=7"])
    assert c.trainer.batch_size == 4 and c.model.compute_dtype == "fp32" and c.seed == 7
    with pytest.raises(FileNotFoundError):
        load_config("nope")
    with pytest.raises(KeyError):
        load_config("dev", ["trainer.bogus=1"])


def test_tile_plan_covers_frame_exactly():
    from pixel_heal_thyself_b200.inference import plan_tiles
    for (H, W, r, c) in ((2048, 2048, 2, 4), (256, 320, 3, 2), (64, 64, 1, 1), (72, 40, 4, 4)):
        for uniform in (False, True):
            tiles = plan_tiles(H, W, r, c, uniform=uniform)
            cover = torch.zeros(H, W, dtype=torch.int32)
            for t in tiles:
                cover[t.y0:t.y1, t.x0:t.x1] += 1
                assert t.y0 % 8 == 0 and t.x0 % 8 == 0 and t.ty0 % 8 == 0 and t.tx0 % 8 == 0
                assert (t.ty1 - t.ty0) % 8 == 0 and (t.tx1 - t.tx0) % 8 == 0
                # >= 48 px of halo on every side that is not the frame border, never outside the frame
                assert 0 <= t.ty0 <= max(0, t.y0 - 48) and min(H, t.y1 + 48) <= t.ty1 <= H
                assert 0 <= t.tx0 <= max(0, t.x0 - 48) and min(W, t.x1 + 48) <= t.tx1 <= W
                if not uniform:
                    assert t.ty0 == max(0, t.y0 - 48) and t.ty1 == min(H, t.y1 + 48)
            assert int(cover.min()) == 1 and int(cover.max()) == 1
            if uniform:   # one haloed shape for all tiles: the generator keeps a single activation arena
                assert len({(t.ty1 - t.ty0, t.tx1 - t.tx0) for t in tiles}) == 1
    with pytest.raises(AssertionError):
        plan_tiles(100, 64, 2, 2)


def test_tiled_stitching_is_exact_for_bounded_receptive_field():
    from pixel_heal_thyself_b200.inference import denoise_frame
    torch.manual_seed(0)
    w = torch.randn(3, 10, 31, 31, dtype=torch.float64) * 0.01     # receptive radius 15 << halo 48

    def net(x, aux):
        return torch.nn.functional.conv2d(torch.cat([x, aux], 1), w, padding=15)

    x = torch.randn(1, 3, 128, 192, dtype=torch.float64)
    aux = torch.randn(1, 7, 128, 192, dtype=torch.float64)
    full = net(x, aux)
    tiled = denoise_frame(net, x, aux, rows=2, cols=3)
    assert (full - tiled).abs().max() < 1e-12


def test_bucket_ranges_follow_backward_order():
    from pixel_heal_thyself_b200.models.afgsa.model import AFGSANet
    from pixel_heal_thyself_b200.parallel import bucket_ranges
    net = AFGSANet(3, 7, 256, num_sa=2, num_gcp=0, padding_mode="replicate")
    order, offsets, off = [], {}, 0
    for n, p in net.named_parameters():
        order.append(n)
        offsets[n] = (off, p.numel())
        off += (p.numel() + 63) // 64 * 64
    r = bucket_ranges(offsets, order, off)
    assert [k for k, _, _ in r] == ["decoder", "block1", "block0", "encoders"]
    spans = sorted((lo, hi) for _, lo, hi in r)
    assert spans[0][0] == 0 and spans[-1][1] == off
    assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))


_DP_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pixel_heal_thyself_b200 import parallel
rank, local, world = parallel.init_distributed()
assert world == 2 and dist.get_backend() == "gloo"
offsets = {"conv1.0.weight": (0, 100), "transformer_blocks.0.attention.rel_h": (128, 60),
           "transformer_blocks.1.attention.rel_h": (192, 60), "decoder.0.0.weight": (256, 200)}
order = list(offsets)
os.environ["PHT_GRAD_ALLREDUCE"] = "overlap"        # bucket by bucket as backward completes them
flat = torch.arange(512, dtype=torch.float32) * (rank + 1)
b = parallel.GradBucketer(lambda: flat, offsets, order, 512)
for tag in ("decoder", "block1", "block0", "encoders"):
    b.ready(tag)
assert b.launched == ["decoder", "block1", "block0", "encoders"]
b.finish()
assert torch.equal(flat, torch.arange(512, dtype=torch.float32) * 3), "all-reduce(sum) over 2 ranks"
os.environ["PHT_GRAD_ALLREDUCE"] = "end"            # the default: one all-reduce of the whole arena in finish()
flat2 = torch.arange(512, dtype=torch.float32) * (rank + 1)
b2 = parallel.GradBucketer(lambda: flat2, offsets, order, 512)
for tag in ("decoder", "block1", "block0", "encoders"):
    b2.ready(tag)
assert torch.equal(flat2, torch.arange(512, dtype=torch.float32) * (rank + 1)), "nothing reduced before finish()"
b2.finish()
assert torch.equal(flat2, torch.arange(512, dtype=torch.float32) * 3) and b2.launched == []
perm = torch.randperm(1000, generator=torch.Generator().manual_seed(5))
mine = parallel.shard_indices(1000, rank, world, 8, perm)
allidx = [torch.zeros_like(mine) for _ in range(2)]
dist.all_gather(allidx, mine)
both = torch.cat(allidx)
assert mine.numel() == 496 and both.unique().numel() == 992, "disjoint equal shards"
assert parallel.shard_indices(1000, 0, 1, 8, perm).numel() == 1000, "one process keeps the final partial batch"
# finish(aliased=False): the caller has just gathered p.grad into the arena -> the whole arena is reduced at the end
flat3 = torch.arange(512, dtype=torch.float32) * (rank + 1)
b3 = parallel.GradBucketer(lambda: flat3, offsets, order, 512)
b3.finish(aliased=False)
assert torch.equal(flat3, torch.arange(512, dtype=torch.float32) * 3)
os.environ["PHT_GRAD_ALLREDUCE"] = "overlap"
b4 = parallel.GradBucketer(lambda: flat3, offsets, order, 512)
try:
    b4.finish(aliased=False)
    raise SystemExit("overlap + non-aliased gradients must raise")
except RuntimeError:
    pass
os.environ["PHT_GRAD_ALLREDUCE"] = "end"
lin = torch.nn.Linear(4, 4)
torch.manual_seed(rank); lin(torch.randn(3, 4)).sum().backward()
g0 = lin.weight.grad.clone(); parallel.allreduce_module_grads(lin, world)
gs = [torch.zeros_like(g0) for _ in range(2)]; dist.all_gather(gs, g0)
assert torch.allclose(lin.weight.grad, (gs[0] + gs[1]) / 2)
dist.barrier(); dist.destroy_process_group()
print("DP_OK", rank)
"""


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"DP_OK {r}" in o, o[-2000:]


def test_bench_clock_sampler_selects_samples_inside_the_timed_region(tmp_path, monkeypatch):
    """bench.ClockSampler against a stand-in nvidia-smi that needs 0.3 s for its first sample: started early, only the
    samples stamped between mark_begin() and mark_end() count; started late, it falls back to the samples written after
    the region began."""
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.3\nwhile true; do\n"
                    "echo \"$(date '+%Y/%m/%d %H:%M:%S.%3N'), 0, 1875, 1965, 950.1, 0x4, Not Active, Not Active, Not Active, Active\"\n"
                    "sleep 0.02\ndone\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    sys.path.insert(0, ROOT)
    import bench
    c = bench.ClockSampler(0)
    c.start()
    time.sleep(0.6)
    c.mark_begin()
    time.sleep(0.15)
    c.mark_end()
    got = c.stop()
    assert got["window"] == "timed region" and 2 <= got["samples"] <= 9, got
    assert got["sm_mhz"] == 1875.0 and got["sm_max_mhz"] == 1965.0 and got["reasons"] == ["sw_power_cap"]
    c = bench.ClockSampler(0)
    c.start()
    c.mark_begin()
    time.sleep(0.1)
    c.mark_end()
    time.sleep(0.35)
    got = c.stop()
    assert got["samples"] >= 1 and got["window"].startswith("from the start of the timed region"), got


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port arm): one JSON line with the contract's keys; dev shape to stay short."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "dev",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "patches/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "config"):
        assert key in line, key


def test_gan_fixture_inputs_and_critic_init_reproduce_the_reference():
    """tests/golden/gan_step.json (one full GAN iteration of the real reference): its inputs regenerate from the seed and
    this package's DiscriminatorVGG draws the reference critic's random init (same modules in the same order) -- what the
    GPU test of the iteration starts from."""
    from conftest import GOLDEN, load_json
    sys.path.insert(0, GOLDEN)
    from make_golden_gan import P, gan_inputs, gp_alphas
    from pixel_heal_thyself_b200.models.afgsa.discriminator import DiscriminatorVGG
    m = load_json("gan_step.json")
    x, gt, aux = gan_inputs()
    assert abs(float(sum(t.double().abs().sum() for t in (x, gt, aux))) - m["inputs_checksum"]) < 1e-9 * m["inputs_checksum"]
    torch.manual_seed(m["seed"] + 1)
    D = DiscriminatorVGG(3, 64, P)
    assert abs(float(sum(p.detach().double().abs().sum() for p in D.parameters())) - m["d_init_checksum"]) < 1e-9 * m["d_init_checksum"]
    # the convolution weights live in channels-last memory: same values, names and shapes as the reference's
    w = D.features[1][0].weight
    assert w.shape == (128, 64, 3, 3) and w.is_contiguous(memory_format=torch.channels_last)
    assert set(D.state_dict()) >= {"features.0.0.weight", "features.1.1.running_mean", "classifier.2.bias"}
    a = gp_alphas()
    assert a.shape == (m["B"], 1, 1, 1) and 0.0 <= float(a.min()) and float(a.max()) < 1.0
