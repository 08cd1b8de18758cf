#!/usr/bin/env python
"""Benchmark of the AFGSA hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload prod|stag|dev]

Own arm: one "step" is one pass of the hot path over one patch batch -- generator
forward, L1 image loss, generator backward and the fused Adam update -- on the
prod configuration (128x128 patches, batch 8 per GPU, bf16), weak scaling over N
GPUs (one process per GPU under torchrun, one NCCL all-reduce of the flat
gradient arena per step).  `value` is timed with inputs resident in HBM; `e2e` is the same
metric through the trainer's public API starting from pinned host buffers.
Reference arm (--impl reference): the CPU restatement of the reference's path
(oracle/) on the host cores, on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import atexit
import datetime
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name -> (patch, per-GPU batch, preset)
    "prod": (128, 8, "prod"),
    "stag": (64, 8, "stag"),
    "dev": (32, 8, "dev"),
}
TRAIN_FLOP_PER_PX = 58_639_872      # fwd + dgrad + wgrad, BASELINE.md section 4
CONV3_FLOP_PER_PX = 2 * 9 * 256 * 256


def _traffic():
    """Measured DRAM bytes per launch of the roofline kernel (ncu --set full capture, profiles/traffic.json):
    launch-weighted mean over the two variants of the 3x3 256->256 conv_gemm that run in a step."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    n = t["launches_per_step"]
    mean = (t["conv3x3_deep_bytes"] * n["deep"] + t["conv3x3_fused_bytes"] * n["fused"]) / (n["deep"] + n["fused"])
    return mean, {"deep": t["conv3x3_deep_bytes"], "fused_epilogue": t["conv3x3_fused_bytes"], "unit": "bytes/launch",
                  "source": t.get("source", "profiles/r2_ncu_full_conv.txt")}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md clocks line).

    nvidia-smi needs 0.1-0.5 s before its first sample (longer on multi-GPU boxes), which is longer than a 10-step timed
    region, so it is started well before the region and the samples are selected by their timestamps:
    ``mark_begin()`` / ``mark_end()`` bracket the region on the host clock (the device is idle at both marks)."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index
        self.t0 = self.t1 = None
        self.pos0 = 0
        self.regions = {}               # name -> [t0, t1, pos0]: further timed regions sampled by the same poller

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
            atexit.register(self._kill)             # never leave the poller behind if the bench dies early
        except OSError:
            self.proc = None

    def _kill(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.kill()

    def mark_begin(self, name=None):
        try:
            pos = os.path.getsize(self.path)            # fallback selector if the stamps cannot be used
        except OSError:
            pos = 0
        if name is None:
            self.t0, self.pos0 = datetime.datetime.now(), pos
        else:
            self.regions[name] = [datetime.datetime.now(), None, pos]

    def mark_end(self, name=None):
        if name is None:
            self.t1 = datetime.datetime.now()
        else:
            self.regions[name][1] = datetime.datetime.now()

    @staticmethod
    def _stamp(text: str):
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f")
        except ValueError:
            return None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        self.proc.wait()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []                                   # (timestamp | None, sm MHz, max MHz, {reasons})
        try:
            text = open(self.path).read()
        except OSError:
            text = ""
        pos = 0
        for line in text.splitlines(keepends=True):
            at, pos = pos, pos + len(line)
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                sm, mx = float(f[2]), float(f[3])
            except ValueError:
                continue
            rows.append((self._stamp(f[0]), sm, mx, {n for n, v in zip(names, f[6:10]) if v.lower().startswith("active")}, at))
        try:
            os.unlink(self.path)
        except OSError:
            pass

        def select(t0, t1, pos0):
            window = "timed region"
            inside = [r for r in rows if r[0] is not None and t0 is not None and t1 is not None and t0 <= r[0] <= t1]
            if not inside:                              # unparsable stamps / region shorter than one period: samples written
                late = [r for r in rows if r[4] >= pos0]   # after the region began (the region + 0.1 s), else everything
                inside, window = (late, "from the start of the timed region to 0.1 s after it") if late else \
                                 (rows, "all samples since the data set-up (none could be placed inside the timed region)")
            sm = [r[1] for r in inside]
            reasons = set().union(*[r[3] for r in inside]) if inside else set()
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": inside[-1][2] if inside else None,
                    "reasons": sorted(reasons), "samples": len(sm), "window": window}

        self.extra = {name: select(t0, t1, pos) for name, (t0, t1, pos) in self.regions.items()}
        return select(self.t0, self.t1, self.pos0)


def cpu_reference_step_time(patch: int, steps: int, warmup: int, budget_s: float | None = None, batch: int = 1):
    """Times the CPU restatement of the reference path (oracle/) on all host cores: one G-only training step
    (forward, L1, backward, Adam) on ``batch`` patches of the workload per step.  With ``budget_s`` the timed loop stops
    early once the budget is spent (at least one timed step always runs).
    Returns (seconds per step, cores, timed steps)."""
    import torch
    from oracle import afgsa_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # (the reference's random init, restated inside the oracle: this arm never imports the product package)
    sd = O.reference_init_state_dict(990819)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 3, patch, patch, generator=g) * 0.5
    aux = torch.rand(batch, 7, patch, patch, generator=g)
    gt = torch.randn(batch, 3, patch, patch, generator=g) * 0.5
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, grads = O.g_only_train_step(x, aux, gt, sd, "replicate")
        for k in sd:
            O.adam_step(sd[k], grads[k], m[k], v2[k], it + 1, 1e-4)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
            if budget_s is not None and time.perf_counter() - t_start > budget_s:
                break
    return sum(times) / len(times), cores, len(times)


def stock_torch_step_time(device: str, patch: int, batch: int, steps: int, warmup: int, autocast_bf16: bool):
    """SURVEY 8d's "bar to beat": the reference's algorithm (the oracle port: plain torch ops + autograd + torch's fused
    Adam) on ONE GPU through stock cuDNN / cuBLAS, TF32 allowed, optionally under ``torch.autocast(bfloat16)``.
    Same G-only step and batch as our arm.  Returns seconds per step."""
    import torch
    from oracle import afgsa_oracle as O
    dev = torch.device(device)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    params = {k: v.to(dev).requires_grad_(True) for k, v in O.reference_init_state_dict(990819).items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4, betas=(0.9, 0.999), **({"fused": True} if dev.type == "cuda" else {}))
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(batch, 3, patch, patch, generator=g) * 0.5).to(dev)
    aux = torch.rand(batch, 7, patch, patch, generator=g).to(dev)
    gt = (torch.randn(batch, 3, patch, patch, generator=g) * 0.5).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast_bf16):
            out = O.afgsa_net_forward(x, aux, params, "replicate")
        loss = O.l1_loss(out.float(), gt)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / steps
    else:
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step()
        sec = (time.perf_counter() - t0) / steps
    assert bool(torch.isfinite(loss)), "stock torch step produced a non-finite loss"
    return sec


def stock_torch_inference_mpix(device: str, side: int, steps: int, autocast_bf16: bool) -> float:
    """Forward-only (no_grad) pass of the oracle port over one side x side frame through stock cuDNN / cuBLAS -> MPix/s."""
    import torch
    from oracle import afgsa_oracle as O
    dev = torch.device(device)
    params = {k: v.to(dev) for k, v in O.reference_init_state_dict(990819).items()}
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(1, 3, side, side, generator=g) * 0.5).to(dev)
    aux = torch.rand(1, 7, side, side, generator=g).to(dev)
    times = []
    with torch.no_grad():
        for it in range(steps + 2):
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast_bf16):
                out = O.afgsa_net_forward(x, aux, params, "replicate")
            if dev.type == "cuda":
                torch.cuda.synchronize()
            if it >= 2:
                times.append(time.perf_counter() - t0)
    assert bool(torch.isfinite(out.float()).all())
    return side * side / (sum(times) / len(times)) / 1e6


def run_reference_stock_gpu(args):
    """``--impl reference --ref-device cuda``: the oracle port on one B200 through stock cuDNN / cuBLAS (not the
    driver's reference arm, which is the CPU path; this is the extra comparison SURVEY 8d asks for)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    patch, batch, preset = WORKLOADS[args.workload]
    res = {}
    for name, ac in (("tf32", False), ("bf16_autocast", True)):
        sec = stock_torch_step_time("cuda:0", patch, batch, max(args.steps, 3), 3, ac)
        res[name] = {"patches_per_s": batch / sec, "ms_per_step": sec * 1e3}
    best = max(res.values(), key=lambda r: r["patches_per_s"])
    infer = {}
    try:
        for name, ac in (("tf32", False), ("bf16_autocast", True)):
            infer[name] = stock_torch_inference_mpix("cuda:0", 1024, 3, ac)
        infer["frame"] = "1024x1024, whole frame in one forward pass"
    except Exception as e:                                   # e.g. out of memory: report, do not fail the training number
        infer["error"] = f"{type(e).__name__}: {e}"[:200]
    print(json.dumps({
        "impl": "reference", "kind": "port on GPU: torch ops + autograd + fused torch Adam via stock cuDNN/cuBLAS",
        "metric": "AFGSA train patches/sec", "value": best["patches_per_s"], "unit": "patches/s", "n_gpus": 1,
        "steps": max(args.steps, 3), "warmup": 3, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
        "dtype": "tf32 / bf16 autocast", "data": "synthetic",
        "config": {"workload": f"{preset}: AFGSA G-only (hot path) training step, {patch}x{patch} patches, batch {batch}"},
        "stock": res, "stock_inference_mpix_per_s": infer, "gpu_launches": 0}), flush=True)


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is pure Python with
    third-party imports that are absent from this image and from the GPU box, so this is the
    oracle port of it (`kind: port`) on all host threads, on the own arm's configuration: every step is one full
    per-GPU batch of the workload (8 patches), time-bounded to ~2.5 min of timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    patch, batch, preset = WORKLOADS[args.workload]
    sec, cores, done = cpu_reference_step_time(patch, args.steps, min(args.warmup, 1), budget_s=150.0, batch=batch)
    val = batch / sec
    sample = (f"{batch} patches {patch}x{patch} per step (the own arm's per-GPU batch: G fwd + L1 + G bwd + Adam), oracle "
              f"port of the reference on {cores} host threads, {min(args.warmup, 1)} warm-up + {done} timed steps "
              f"(requested {args.steps}, time-bounded)")
    line = {
        "impl": "reference", "metric": "AFGSA train patches/sec", "value": val, "unit": "patches/s",
        "n_gpus": args.gpus, "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{preset}: AFGSA G-only (hot path) training step, {patch}x{patch} patches, batch {batch}",
                   "global_batch": batch, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "patches/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _timed(fn, n, sync_all):
    """n calls of fn between two CUDA events on the current stream, barrier + synchronize on both sides -> ms total."""
    import torch
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    sync_all()
    return e0.elapsed_time(e1)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pixel_heal_thyself_b200 import _lib, ops
    from pixel_heal_thyself_b200.config import load_config
    from pixel_heal_thyself_b200.data import DevicePrefetcher
    from pixel_heal_thyself_b200.models.afgsa.train import AFGSATrainer

    patch, batch, preset = WORKLOADS[args.workload]
    n_img = 2
    cfg = load_config(preset, [f"trainer.batch_size={batch}", f"data.synthetic.num_images={n_img}",
                               "data.synthetic.height=1024", "data.synthetic.width=1024",
                               f"model.afgsa.compute_dtype={args.dtype}"])
    tr = AFGSATrainer(cfg)
    rank, world, dev = tr.rank, tr.world, tr.device
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    tr.setup(g_only=not args.gan)
    clocks = ClockSampler(tr.local_rank)
    if rank == 0:
        clocks.start()                       # early: nvidia-smi takes a few hundred ms to deliver its first sample
    ds = tr.setup_data()
    gen = torch.Generator().manual_seed(cfg.seed + rank)
    total = args.warmup + args.steps
    order = torch.randperm(len(ds), generator=gen)
    need = 3 * total * batch
    order = order.repeat((need + len(ds) - 1) // len(ds))[:need].to(dev)
    npx = batch * patch * patch

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def assert_ranks_identical(net, what):
        """data parallel: every rank must hold bit-identical generator weights after the timed steps"""
        if world == 1:
            return
        ref = net.flat_param.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1.0 if torch.equal(ref, net.flat_param) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if float(same) != 1.0:
            raise SystemExit(f"bench: the ranks' generator weights diverged during {what}")

    # ---------------- device-resident arm ("value") ----------------
    batches = [ds.batch_device(order[i * batch:(i + 1) * batch]) for i in range(total)]
    for i in range(args.warmup):
        tr.train_step(*batches[i])
    sync_all()
    clocks.mark_begin()
    _lib.lib.pht_reset_counters()
    ms = _timed(lambda i: tr.train_step(*batches[args.warmup + i]), args.steps, sync_all)
    clocks.mark_end()
    counters = _lib.counters()
    assert_ranks_identical(tr.G, "the timed steps")

    # ---------------- roofline kernel, timed in its own short pass (per-launch CUDA events perturb the step: they are
    # kept out of the `value` loop) ----------------
    sink = []
    ops.set_launch_profiler(sink, lambda tag: tag[0] == 3 and tag[1] == 256 and tag[2] == 256)
    prof_steps = 3
    graphed, tr.use_step_graph = tr.use_step_graph, False      # eager launches: the events bracket individual kernels
    ms_prof = _timed(lambda i: tr.train_step(*batches[i % total]), prof_steps, sync_all)
    tr.use_step_graph = graphed
    ops.set_launch_profiler(None)
    conv_ms = [a.elapsed_time(b) for _, a, b in sink]

    # ---------------- sustained: the same step back to back for >= 2 s (power management settles) ----------------
    sus = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2200.0 / max(ms / args.steps, 1e-3)) + 1)
        clocks.mark_begin("sustained")
        ms_sus = _timed(lambda i: tr.train_step(*batches[i % total]), n_sus, sync_all)
        clocks.mark_end("sustained")
        sus = (ms_sus, n_sus)

    # ---------------- end-to-end arm ("e2e"): pinned host NHWC patches -> H2D -> preprocess -> step -> D2H loss
    host = ds.host_patches()
    hidx = order.cpu()
    pf_stream = torch.cuda.Stream(device=dev)

    def e2e_leg(trainer, first):
        warm = [{k: v[hidx[i * batch:(i + 1) * batch]].pin_memory() for k, v in host.items()} for i in range(args.warmup)]
        for dev_batch in DevicePrefetcher(warm, dev, pf_stream):   # warm-up through the same path (side stream + its pool)
            float(trainer.train_step(*dev_batch)[0])
        staged = [{k: v[hidx[(first + i) * batch:(first + i + 1) * batch]].pin_memory() for k, v in host.items()}
                  for i in range(args.steps)]
        prefetcher = DevicePrefetcher(staged, dev, pf_stream)
        sync_all()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        last = 0.0
        # every step copies ITS batch from pinned host memory and reads ITS loss back; the copy + preprocess of batch i+1
        # is enqueued on a side stream before step i's loss is read (the reference's DataLoader prefetches the same way).
        # The loss of step i travels to pinned host memory by an asynchronous 4-byte copy enqueued right behind the step and
        # is consumed on the host one step later (after step i+1 has been enqueued), so the device never waits for the
        # host's round trip; the last step's loss is read before the region ends.
        slots = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        events = [torch.cuda.Event() for _ in range(2)]
        pending = None
        for i, dev_batch in enumerate(prefetcher):
            g_loss, _ = trainer.train_step(*dev_batch)
            slots[i % 2].copy_(g_loss.reshape(1), non_blocking=True)   # device -> host read of the step's result, every step
            events[i % 2].record()
            if pending is not None:
                events[pending].synchronize()
                last = float(slots[pending])
            pending = i % 2
        if pending is not None:
            events[pending].synchronize()
            last = float(slots[pending])
        h1.record()
        sync_all()
        return h0.elapsed_time(h1), last, sum(v.numel() * v.element_size() for v in staged[0].values())

    ms_e2e, last, h2d = e2e_leg(tr, total)

    # ---------------- full-frame tiled inference (the metric's second half: MPix/s on a 2048x2048 frame) ----------
    ms_inf, inf_cfg = None, None
    if not args.no_inference:
        from pixel_heal_thyself_b200.data import synthetic_frames
        from pixel_heal_thyself_b200.inference import EXACT_HALO, denoise_frame
        # one tile per GPU: the fewest tiles that keep every GPU busy (halo recompute is pure overhead: +19 % executed
        # FLOPs at 2x4, +2 % at 1x2, none for the single whole-frame "tile" of one GPU)
        side, n_frames = 2048, 3
        rows, cols = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}.get(world, (2, 4))
        fr = synthetic_frames(1, side, side, cfg.seed + 7, dev)
        fx = torch.empty(1, 3, side, side, device=dev)
        fa = torch.empty(1, 7, side, side, device=dev)
        ops.preprocess(fr["noisy"], None, fr["aux"], fx, None, fa)
        del fr
        tr.G.eval()
        denoise_frame(tr.G, fx, fa, rows, cols, EXACT_HALO, rank, world)          # warm-up (allocates the eval arena)
        ms_inf = _timed(lambda i: denoise_frame(tr.G, fx, fa, rows, cols, EXACT_HALO, rank, world), n_frames, sync_all) / n_frames
        tr.G.train()
        inf_cfg = {"frame": f"{side}x{side}", "tiles": f"{rows}x{cols} 8-aligned, {EXACT_HALO}-px halo (exact)", "frames_timed": n_frames}
        del fx, fa

    # ---------------- the FULL iteration of base_trainer.py:388-457 (G step + critic step), at every N ----------------
    full = None
    if not args.gan and not args.no_gan_extra:
        trg = AFGSATrainer(cfg)
        trg.setup(g_only=False)
        for i in range(args.warmup + 1):
            trg.train_step(*batches[i % total])
        _lib.lib.pht_reset_counters()
        ms_full = _timed(lambda i: trg.train_step(*batches[i % total]), args.steps, sync_all)
        full_launches = sum(_lib.counters().values())
        assert_ranks_identical(trg.G, "the full-iteration steps")
        ms_full_e2e, full_last, _ = e2e_leg(trg, 2 * total)
        full = (ms_full, ms_full_e2e, full_last, full_launches)
        del trg

    t = torch.tensor([ms, ms_e2e, ms_inf or 0.0, sus[0] if sus else 0.0, full[0] if full else 0.0, full[1] if full else 0.0,
                      ms_prof], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_prof = float(t[0]), float(t[1]), float(t[6])
    ms_inf = float(t[2]) if ms_inf is not None else None
    if sus:
        sus = (float(t[3]), sus[1])
    if full:
        full = (float(t[4]), float(t[5]), full[2], full[3])
    if world > 1:
        # tear the process group down BEFORE the CPU baseline: ranks > 0 exit here, so rank 0's host threads are not
        # fighting seven ranks spinning in a barrier
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    clk = clocks.stop()

    # ---------------- comparison arms on rank 0: stock cuDNN on this GPU, the CPU port on the host cores ----------------
    stock = None
    if world == 1 and not args.no_stock:
        try:
            torch.cuda.empty_cache()
            sclk = ClockSampler(tr.local_rank)
            sclk.start()
            time.sleep(0.6)
            sclk.mark_begin()
            sec = stock_torch_step_time(str(dev), patch, batch, 5, 3, True)
            sclk.mark_end()
            stock = {"value": batch / sec, "unit": "patches/s", "ms_per_step": sec * 1e3, "clocks": sclk.stop(),
                     "what": "the reference's algorithm (oracle port: torch ops + autograd + torch's fused Adam) on this GPU "
                             "through stock cuDNN/cuBLAS under torch.autocast(bfloat16), same G-only step and batch, "
                             "3 warm-up + 5 timed steps (the clock window includes the warm-up)"}
        except Exception as e:       # e.g. out of memory: report, do not fail the run
            stock = {"error": f"{type(e).__name__}: {e}"[:200]}
    cpu_sec, cores, _ = (cpu_reference_step_time(patch, 1, 1) if not args.no_cpu_baseline else (None, os.cpu_count(), 0))

    peaks = _peaks()
    patches = batch * world * args.steps
    per_step = ms / args.steps
    conv_avg_ms = sum(conv_ms) / max(len(conv_ms), 1)
    conv_flop = CONV3_FLOP_PER_PX * npx
    ach = conv_flop / (conv_avg_ms * 1e-3) / 1e12 if conv_ms else None
    # denominator: the burst peak when the kernel is timed in a sub-second region, the sustained one otherwise
    short = ms_prof < 1000.0
    peak = peaks["tf_burst"] if short else peaks["tf_sustained"]
    flop_step = TRAIN_FLOP_PER_PX * npx * world
    line = {
        "metric": "AFGSA train patches/sec", "value": patches / (ms * 1e-3), "unit": "patches/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.dtype == "bf16" else "f32",
        "data": "synthetic",
        "config": {"workload": f"{preset}: AFGSA {'GAN' if args.gan else 'G-only (hot path)'} training step, "
                               f"{patch}x{patch} patches, batch {batch}/GPU, {n_img} synthetic 1024x1024 frames/GPU",
                   "global_batch": batch * world, "parallelism": f"dp{world}",
                   "l2": f"per-step working set ~{3.3 * npx / 131072:.1f} GB >> 126 MB L2 (no explicit flush)",
                   "step": "G forward + L1 + G backward + fused Adam" + (" + critic step (PyTorch)" if args.gan else ""),
                   "launch": ("captured once, replayed from CUDA graphs" if getattr(tr, "_step_graph", None) is not None
                              and tr._step_graph.get("rec") is not None else "eager launches")},
        "clocks": clk,
        "e2e": {"value": patches / (ms_e2e * 1e-3), "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last},
        "sustained": ({"value": batch * world * sus[1] / (sus[0] * 1e-3), "unit": "patches/s", "steps": sus[1],
                       "seconds": sus[0] * 1e-3, "ms_per_step": sus[0] / sus[1], "clocks": clocks.extra.get("sustained"),
                       "step_tflops": flop_step / (sus[0] / sus[1] * 1e-3) / 1e12,
                       "frac_of_sustained_peak": flop_step / world / (sus[0] / sus[1] * 1e-3) / 1e12 / peaks["tf_sustained"]}
                      if sus else None),
        "full_step": ({"value": patches / (full[0] * 1e-3), "unit": "patches/s", "ms_per_step": full[0] / args.steps,
                       "e2e": {"value": patches / (full[1] * 1e-3), "unit": "patches/s", "h2d_bytes_per_step": h2d,
                               "d2h_bytes_per_step": 4, "ms_per_step": full[1] / args.steps, "last_loss": full[2]},
                       "gpu_launches": full[3],
                       "note": "full iteration of base_trainer.py:388-457 = the G step above + the critic step; the critic "
                               "(DiscriminatorVGG + WGAN-GP) stays PyTorch/cuDNN, replayed from CUDA graphs (data parallel: "
                               "two graphs around one flat NCCL all-reduce of its gradients)"}
                      if full else None),
        "gpu_launches": sum(counters.values()),
        "launch_counters": counters,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm 3x3 256->256 (forward + data-grad launches)",
                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": (ach / peak) if ach else None,
                     "peak_source": f"{peaks['src']} {'burst' if short else 'sustained'} bf16 (kernel timed over a "
                                    f"{ms_prof * 1e-3:.2f} s pass of {prof_steps} steps)",
                     "frac_burst": (ach / peaks["tf_burst"]) if ach else None,
                     "frac_sustained": (ach / peaks["tf_sustained"]) if ach else None,
                     "traffic": _traffic()[0], "traffic_detail": _traffic()[1],
                     "launches_timed": len(conv_ms), "avg_launch_ms": conv_avg_ms, "algorithmic_flop_per_launch": conv_flop,
                     "share_of_step": (sum(conv_ms) / ms_prof) if conv_ms else None,
                     "timed_in": "a separate 3-step pass (per-launch CUDA events are kept out of the `value` loop)"},
        "step_tflops": flop_step / (per_step * 1e-3) / 1e12,
        "inference": ({"value": 2048 * 2048 / (ms_inf * 1e-3) / 1e6, "unit": "MPix/s", "ms_per_frame": ms_inf, "n_gpus": world,
                       **inf_cfg, "dtype": "bf16", "note": "G.eval() forward, frame resident in HBM, stitched on rank 0"}
                      if ms_inf else None),
        "stock_cudnn": stock,
        "cpu_baseline": ({"value": 1.0 / cpu_sec, "unit": "patches/s", "cores": cores, "kind": "port",
                          "sample": f"1 patch {patch}x{patch} of the {batch}-patch batch, 1 warm-up + 1 timed G-only step of "
                                    f"the oracle port on rank 0's host cores"}
                         if cpu_sec else None),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cpu = the reference arm (default); cuda = the same port on one GPU "
                         "through stock cuDNN/cuBLAS (extra comparison, SURVEY 8d)")
    ap.add_argument("--workload", default="prod", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--gan", action="store_true", help="time the full GAN iteration (adds the PyTorch critic step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gan-extra", action="store_true", help="skip the informational full-GAN-iteration timing")
    ap.add_argument("--no-inference", action="store_true", help="skip the full-frame tiled inference measurement")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s back-to-back run")
    ap.add_argument("--no-stock", action="store_true", help="skip the stock cuDNN/cuBLAS comparison on the same GPU (N=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        (run_reference_stock_gpu if args.ref_device == "cuda" else run_reference)(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
