"""GAN trainer driving the B200 generator (reference: pht/models/base_trainer.py).

Same template-method API as the reference's ``BaseTrainer`` -- abstract
``create_generator()`` (the plug-in point, base_trainer.py:102-110), overridable
``create_discriminator / create_losses / create_optimizers`` and ``train()`` --
and the same per-iteration arithmetic (base_trainer.py:371-457), restated as
``train_step`` so the benchmark can time exactly one step.  Differences, all
B200-side: batches are preprocessed on the device by ``pht_preprocess`` /
``pht_crop_preprocess`` instead of numpy on the host; the generator optimiser is
the fused flat Adam; loss scalars are read back once per epoch instead of every
iteration.  Under torchrun every rank runs ``trainer.batch_size`` patches per
step (per-rank batch: the global batch is world x batch_size, learning rates
unscaled -- see ``parallel``), the generator's flat gradient arena is SUM
all-reduced once after backward (``PHT_GRAD_ALLREDUCE=overlap`` issues it
bucket by bucket during backward instead) and averaged inside the Adam kernel;
the critic's gradients are averaged with one flat all-reduce; validation is
sharded over the ranks and checkpoints are written by rank 0.
"""
from __future__ import annotations

import logging
import math
import os
import random
import time
from abc import ABC, abstractmethod

import numpy as np
import torch
from torch import optim
from torch.optim import lr_scheduler

from .. import parallel
from ..config import Config
from ..data import PatchDataset, synthetic_frames
from ..optim import FlatAdam
from .afgsa.discriminator import DiscriminatorVGG
from .losses import GANLoss, GradientPenaltyLoss, L1ReconstructionLoss, SSIMLoss

logger = logging.getLogger("pht")


class _StepRecorder:
    """One training step as a replayable sequence: CUDA graphs separated by the NCCL all-reduces of a data-parallel step
    (collectives stay eager; everything between them is captured).  All graphs share one memory pool and are captured on
    the stream the warm-up steps ran on."""

    def __init__(self, stream):
        self.stream = stream
        self.items = []          # ("graph", CUDAGraph) | ("allreduce", tensor)
        self.pool = None
        self._ctx = self._g = None

    def begin(self):
        self._g = torch.cuda.CUDAGraph()
        kw = {} if self.pool is None else {"pool": self.pool}
        self._ctx = torch.cuda.graph(self._g, stream=self.stream, **kw)
        self._ctx.__enter__()

    def end(self, exc=(None, None, None)):
        ctx, self._ctx = self._ctx, None
        ctx.__exit__(*exc)
        if exc[0] is None:
            self.items.append(("graph", self._g))
            if self.pool is None:
                self.pool = self._g.pool()

    def abort(self, e):
        if self._ctx is not None:
            try:
                self.end((type(e), e, e.__traceback__))
            except Exception:  # noqa: BLE001
                pass

    def allreduce(self, t):
        """Called where the eager step all-reduces ``t``: close the running capture, remember the collective, open the
        next capture.  (Nothing executes during capture, so the collective itself is only issued at replay.)"""
        self.end()
        self.items.append(("allreduce", t))
        self.begin()

    def replay(self):
        for kind, x in self.items:
            if kind == "graph":
                x.replay()
            else:
                torch.distributed.all_reduce(x, op=torch.distributed.ReduceOp.SUM)


def set_determinism(seed: int, deterministic: bool = True) -> None:
    """reference: base_trainer.py:50-67."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


class BaseTrainer(ABC):
    def __init__(self, cfg: Config) -> None:
        self.cfg = cfg
        self.deterministic = cfg.trainer.deterministic
        self.model_name = self.__class__.__name__.replace("Trainer", "")
        self.rank, self.local_rank, self.world = parallel.init_distributed()
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 trainer needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", self.local_rank)
        set_determinism(cfg.seed, self.deterministic)
        self.padding_mode = "replicate" if self.deterministic else "reflect"  # base_trainer.py:334
        self.G = self.D = None
        self.g_only = False
        # Replay the whole training step from CUDA graphs (PHT_STEP_GRAPH=0 disables): a prod step is ~155 launches at
        # ~22 us of Python/ctypes each, which bounds the small presets (dev / stag) and leaves gaps at prod
        self.use_step_graph = os.environ.get("PHT_STEP_GRAPH", "1") != "0"
        self._step_graph = None
        self._capturing = None

    # ------------------------------------------------------------------ factories (reference API)
    @abstractmethod
    def create_generator(self) -> torch.nn.Module:
        """Create and return the generator (reference: base_trainer.py:102-110)."""

    def create_discriminator(self) -> torch.nn.Module:
        if self.cfg.model.discriminator.use_multiscale_discriminator:
            raise NotImplementedError("multiscale discriminator is outside the AFGSA hot path")
        return DiscriminatorVGG(3, 64, self.cfg.data.patches.patch_size).to(self.device)

    def create_losses(self):
        """(l1_loss, gan_loss, gp_loss, lpips_loss, ssim_loss) as in base_trainer.py:127-154."""
        lc = self.cfg.model.losses
        if lc.use_lpips_loss:
            raise NotImplementedError("the LPIPS loss needs the lpips package and its downloaded VGG weights, which are "
                                      "not available offline (SURVEY 8c: out of scope)")
        ssim = SSIMLoss(window_size=11).to(self.device) if lc.use_ssim_loss else None     # base_trainer.py:149-153
        return (L1ReconstructionLoss().to(self.device), GANLoss("wgan").to(self.device),
                GradientPenaltyLoss(self.device).to(self.device), None, ssim)

    def create_optimizers(self, G, D):  # noqa: N803
        """Adam x2 + MultiStepLR(gamma 0.5) with the reference's milestones (base_trainer.py:177-204)."""
        t = self.cfg.trainer
        milestones = [i * t.lr_milestone - 1 for i in range(1, t.epochs // t.lr_milestone)]
        opt_g = FlatAdam(G, lr=t.lr_g, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0 / self.world)
        sch_g = lr_scheduler.MultiStepLR(opt_g, milestones=milestones, gamma=0.5)
        # capturable: the critic step is replayed as CUDA graphs (same update rule, device-side step count)
        opt_d = optim.Adam(D.parameters(), lr=t.lr_d, betas=(0.9, 0.999), eps=1e-8, capturable=self.device.type == "cuda")
        sch_d = lr_scheduler.MultiStepLR(opt_d, milestones=milestones, gamma=0.5)
        return opt_g, sch_g, opt_d, sch_d

    # ------------------------------------------------------------------ setup
    def setup(self, g_only: bool = False) -> None:
        self.g_only = g_only
        self.G = self.create_generator()
        self.D = None if g_only else self.create_discriminator()
        self.l1_loss, self.gan_loss, self.gp_loss, _, self.ssim_loss = self.create_losses()
        if g_only:
            t = self.cfg.trainer
            self.opt_g = FlatAdam(self.G, lr=t.lr_g, grad_scale=1.0 / self.world)
            self.sch_g = self.opt_d = self.sch_d = None
        else:
            self.opt_g, self.sch_g, self.opt_d, self.sch_d = self.create_optimizers(self.G, self.D)
        if self.cfg.trainer.load_model:                     # base_trainer.py:341-347
            mp = self.cfg.trainer.model_path
            if not mp:
                raise ValueError("trainer.load_model=true needs trainer.model_path=<directory with G.pt / D.pt>")
            self.G.load_state_dict(torch.load(os.path.join(mp, "G.pt"), map_location=self.device))
            if self.D is not None:
                self.D.load_state_dict(torch.load(os.path.join(mp, "D.pt"), map_location=self.device))
        self.bucketer = None
        if self.world > 1:
            self.G._flatten()
            self.bucketer = parallel.GradBucketer(lambda: self.G.flat_grad, self.G._offsets,
                                                  [n for n, _ in self.G.named_parameters()], self.G.flat_param.numel(),
                                                  allreduce=self._allreduce)
            self.G.engine.grad_ready_hook = self.bucketer.ready
            # identical initial weights on every rank
            torch.distributed.broadcast(self.G.flat_param, 0)
            if self.D is not None:
                for p in list(self.D.parameters()) + list(self.D.buffers()):
                    torch.distributed.broadcast(p.data, 0)
            # the gradient penalty draws torch.rand from the global generator (losses.py:35-39): one stream per rank
            torch.cuda.manual_seed(self.cfg.seed + self.rank)

    def setup_data(self) -> PatchDataset:
        d = self.cfg.data
        if d.source != "synthetic":
            raise NotImplementedError("only data.source=synthetic is available (no EXR / HDF5 readers in this image)")
        s = d.synthetic
        frames = synthetic_frames(s.num_images, s.height, s.width, self.cfg.seed, self.device)
        return PatchDataset(frames, d.patches.patch_size, d.patches.num_patches, self.cfg.seed)

    # ------------------------------------------------------------------ one iteration
    def _allreduce(self, t):
        """SUM all-reduce of a persistent tensor; under step capture the recorder splits its graphs around it."""
        if self._capturing is not None:
            self._capturing.allreduce(t)
        else:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)

    def train_step(self, noisy, gt, aux):
        """One iteration of base_trainer.py:388-457 on preprocessed NCHW device tensors.
        Returns (g_loss, d_loss) as 0-dim device tensors (no host sync).  After two eager warm-up steps of a given
        batch shape the step is captured once and replayed from CUDA graphs (``use_step_graph``)."""
        if (self.use_step_graph and noisy.is_cuda and not getattr(self, "_step_graph_failed", False)
                and not (self.bucketer is not None and self.bucketer.mode == "overlap")):
            return self._train_step_graphed(noisy, gt, aux)
        return self._train_step_eager(noisy, gt, aux)

    def _train_step_graphed(self, noisy, gt, aux):
        from .. import _lib
        lrs = (float(self.opt_g.param_groups[0]["lr"]),) + (tuple(float(g["lr"]) for g in self.opt_d.param_groups)
                                                             if self.opt_d is not None else ())
        key = (tuple(noisy.shape), lrs, self.g_only)
        st = self._step_graph
        if st is None or st["key"] != key:
            st = self._step_graph = {"key": key, "warm": 0, "rec": None, "stream": torch.cuda.Stream(device=self.device)}
        cur = torch.cuda.current_stream()
        if st["warm"] < 2:
            # eager warm-up steps ON THE CAPTURE STREAM (allocator pools, per-stream library scratch, descriptor tables,
            # cuDNN plans, optimizer state are all created here, never under capture)
            st["warm"] += 1
            st["stream"].wait_stream(cur)
            with torch.cuda.stream(st["stream"]):
                out = self._train_step_eager(noisy, gt, aux)
            cur.wait_stream(st["stream"])
            for t in (noisy, gt, aux) + tuple(o for o in out if o is not None):
                t.record_stream(st["stream"])
            return out
        if st["rec"] is None:
            st["noisy"], st["gt"], st["aux"] = noisy.clone(), gt.clone(), aux.clone()
            for attempt in (0, 1):
                rec = _StepRecorder(st["stream"])
                before = list(_lib.raw_counters())
                try:
                    self.opt_g.sync_lr()
                    rec.begin()
                    self._capturing = rec
                    try:
                        st["g_loss"], st["d_loss"] = self._train_step_eager(st["noisy"], st["gt"], st["aux"])
                    finally:
                        self._capturing = None
                    rec.end()
                except Exception as e:  # noqa: BLE001 -- capture is an optimisation: fall back to the eager step
                    rec.abort(e)
                    torch.cuda.synchronize()
                    if attempt == 0 and _lib.lib.pht_set_option(b"pdl", 0) == 0:
                        logger.warning(f"step capture failed ({e!r}); retrying without programmatic dependent launch")
                        continue
                    logger.warning(f"step CUDA-graph capture failed ({e!r}); running the training step eagerly")
                    self._step_graph_failed = True
                    self._step_graph = None
                    return self._train_step_eager(noisy, gt, aux)
                after = list(_lib.raw_counters())
                st["launches"] = (_lib.C.c_uint64 * 8)(*[a - b for a, b in zip(after, before)])
                st["rec"] = rec
                st["fresh"] = True          # (the capture itself already bumped the host-side launch counters once)
                break
        st["noisy"].copy_(noisy)
        st["gt"].copy_(gt)
        st["aux"].copy_(aux)
        self.opt_g.sync_lr()
        st["rec"].replay()
        self.G.mark_weights_dirty()                     # (the replayed Adam changed the weights behind Python's back)
        if st.pop("fresh", False) is False:
            _lib.lib.pht_add_counters(st["launches"])  # launch counters are bumped host-side: account for the replay
        return st["g_loss"].clone(), (st["d_loss"].clone() if st["d_loss"] is not None else None)

    def _train_step_eager(self, noisy, gt, aux):
        lw = self.cfg.model.losses
        output = self.G(noisy, aux)
        d_loss = None
        if not self.g_only:
            d_loss = self._critic_step(output.detach(), gt)
        self.opt_g.zero_grad()
        if (self.g_only and self.ssim_loss is None and output.is_cuda and output.dtype == torch.float32
                and type(self.l1_loss) is L1ReconstructionLoss):
            # G-only step, L1 term only (the BASELINE configuration): the fused loss kernel already produces
            # d(l1_loss_w * L1) / d(output), so the generator's backward starts from it directly -- no scalar
            # ones-fill / multiply launches between the loss kernel and the first gradient kernel
            from .. import ops
            a, b = output.detach().contiguous(), gt.contiguous().float()
            loss = torch.empty(1, dtype=torch.float32, device=a.device)
            d_out = torch.empty_like(a)
            ops.l1_loss(a, b, loss, d_out, grad_scale=float(lw.l1_loss_w))
            g_loss = loss.reshape(()) if float(lw.l1_loss_w) == 1.0 else loss.reshape(()) * float(lw.l1_loss_w)
            output.backward(d_out)
            if self.bucketer is not None:
                _, aliased = self.opt_g.gather_grads()
                self.bucketer.finish(aliased)
            self.opt_g.step()
            return g_loss.detach(), None
        g_loss = lw.l1_loss_w * self.l1_loss(output, gt)
        if not self.g_only:
            # the generator's adversarial term only needs d D / d input: skip the critic's weight gradients (the reference
            # computes and then discards them at the next opt_d.zero_grad())
            d_params = list(self.D.parameters())
            for p in d_params:
                p.requires_grad_(False)
            g_loss = lw.gan_loss_w * self.gan_loss(self.D(output), True) + g_loss
            for p in d_params:
                p.requires_grad_(True)
        if self.ssim_loss is not None:                      # base_trainer.py:450-452
            g_loss = g_loss + lw.ssim_loss_w * self.ssim_loss(output, gt)
        g_loss.backward()
        if self.bucketer is not None:
            # gather first: the all-reduce must act on the arena the optimiser reads (p.grad may not alias the arena
            # backward wrote: zero_grad(set_to_none=False), accumulation, hooks)
            _, aliased = self.opt_g.gather_grads()
            self.bucketer.finish(aliased)
        self.opt_g.step()
        return g_loss.detach(), (d_loss.detach() if d_loss is not None else None)

    # ------------------------------------------------------------------ critic step (base_trainer.py:391-412)
    def _critic_eager(self, fake, gt):
        lw = self.cfg.model.losses
        self.opt_d.zero_grad()
        pred_fake = self.D(fake)
        pred_real = self.D(gt)
        d_loss = (self.gan_loss(pred_fake, False) + self.gan_loss(pred_real, True)) / 2 \
            + lw.gp_loss_w * self.gp_loss(self.D, gt, fake)
        d_loss.backward()
        parallel.allreduce_module_grads(self.D, self.world, allreduce=self._allreduce)
        self.opt_d.step()
        return d_loss.detach()

    def _critic_step(self, fake, gt):
        """The PyTorch critic step, replayed from CUDA graphs captured once per (shape, learning rate): the step is
        ~2,000 small eager launches (BatchNorm / LeakyReLU double backward of the gradient penalty), i.e. bound by the
        host's launch rate, not by the GPU.  One GPU: one graph.  Data parallel: two graphs around the one eager NCCL
        call -- [zero grads, 3 x D forward, gradient penalty, backward into ONE flat gradient buffer] -> all-reduce(flat)
        -> [average, Adam step] -- so every rank still replays instead of launching."""
        use_graph = (fake.is_cuda and os.environ.get("PHT_CRITIC_GRAPH", "1") != "0" and self._capturing is None
                     and not getattr(self, "_critic_graph_failed", False))
        if not use_graph:
            return self._critic_eager(fake, gt)
        key = (tuple(fake.shape), tuple(float(g["lr"]) for g in self.opt_d.param_groups))
        st = getattr(self, "_critic_graph", None)
        if st is None or st["key"] != key:
            self._critic_warm_loss = None
            try:
                st = self._capture_critic(fake, gt, key) if self.world == 1 else self._capture_critic_dp(fake, gt, key)
            except Exception as e:  # noqa: BLE001 -- capture is an optimisation: fall back to the eager step
                logger.warning(f"critic CUDA-graph capture failed ({e!r}); running the critic step eagerly")
                self._critic_graph_failed = True
                self._critic_graph = None
                if self.world > 1:
                    for p in self.D.parameters():           # un-home the gradients from the flat buffer
                        p.grad = None
                if self._critic_warm_loss is not None:     # the warm-up already took this iteration's critic step
                    return self._critic_warm_loss
                return self._critic_eager(fake, gt)
            self._critic_graph = st
            return st["warm_loss"]          # (the capture warm-up already took this iteration's step)
        st["fake"].copy_(fake)
        st["gt"].copy_(gt)
        st["graph"].replay()
        if self.world > 1:
            torch.distributed.all_reduce(st["flat"], op=torch.distributed.ReduceOp.SUM)
            st["graph_step"].replay()
        return st["loss"].clone()

    def _capture_critic(self, fake, gt, key):
        s_fake, s_gt = fake.clone(), gt.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up on a side stream (allocator / cuDNN / optimizer state), 1 real step
            warm_loss = self._critic_eager(s_fake, s_gt)
        torch.cuda.current_stream().wait_stream(side)
        self._critic_warm_loss = warm_loss
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self._critic_eager(s_fake, s_gt)
        return {"key": key, "graph": graph, "fake": s_fake, "gt": s_gt, "loss": loss, "warm_loss": warm_loss}

    def _capture_critic_dp(self, fake, gt, key):
        lw = self.cfg.model.losses
        s_fake, s_gt = fake.clone(), gt.clone()
        params = [p for p in self.D.parameters() if p.requires_grad]
        flat = torch.zeros(sum(p.numel() for p in params), device=fake.device)
        off = 0
        for p in params:                    # every gradient is a view of ONE flat buffer: one all-reduce, no copies
            g = flat[off:off + p.numel()]
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                p.grad = g.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)   # the parameter's strides
            else:
                p.grad = g.view_as(p)
            off += p.numel()

        def fwd_bwd():
            flat.zero_()                    # (== zero_grad(set_to_none=False): backward accumulates into the views)
            d_loss = (self.gan_loss(self.D(s_fake), False) + self.gan_loss(self.D(s_gt), True)) / 2 \
                + lw.gp_loss_w * self.gp_loss(self.D, s_gt, s_fake)
            d_loss.backward()
            return d_loss.detach()

        def step():
            flat.div_(self.world)
            self.opt_d.step()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up = this iteration's real step
            warm_loss = fwd_bwd()
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
            step()
        torch.cuda.current_stream().wait_stream(side)
        self._critic_warm_loss = warm_loss
        g_a, g_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_a):
            loss = fwd_bwd()
        with torch.cuda.graph(g_b, pool=g_a.pool()):
            step()
        return {"key": key, "graph": g_a, "graph_step": g_b, "flat": flat, "fake": s_fake, "gt": s_gt, "loss": loss,
                "warm_loss": warm_loss}

    # ------------------------------------------------------------------ full loop
    def train(self) -> None:
        cfg = self.cfg
        if self.G is None:
            self.setup()
        ds = self.setup_data()
        bs = cfg.trainer.batch_size
        n = len(ds)
        n_val = max(1, int(n * (1 - cfg.data_ratio)))
        n_train = n - n_val
        gen = torch.Generator().manual_seed(cfg.seed)
        out_dir = cfg.paths.output_dir
        if self.rank == 0:
            os.makedirs(out_dir, exist_ok=True)
        logger.info(f"Starting training: model={self.model_name}, seed={cfg.seed}, batch_size={bs}, "
                    f"epochs={cfg.trainer.epochs}, world={self.world}, patches={n_train}")
        for epoch in range(cfg.trainer.epochs):
            start = time.time()
            perm = torch.randperm(n_train, generator=gen)
            idx = parallel.shard_indices(n_train, self.rank, self.world, bs, perm).to(self.device)
            iters = (idx.numel() + bs - 1) // bs            # one process: the last batch may be partial (base_trainer.py:363)
            acc_g = torch.zeros((), device=self.device)
            acc_d = torch.zeros((), device=self.device)
            for it in range(iters):
                noisy, gt, aux = ds.batch_device(idx[it * bs:(it + 1) * bs])
                g_loss, d_loss = self.train_step(noisy, gt, aux)
                acc_g += g_loss / bs
                if d_loss is not None:
                    acc_d += d_loss / bs
                if it % 10 == 0 or it == iters - 1:
                    logger.debug(f"[Train] epoch={epoch + 1} iter={it + 1}/{iters}")
            g_avg, d_avg = float(acc_g) / max(iters, 1), float(acc_d) / max(iters, 1)
            logger.info(f"[Train] epoch={epoch + 1} summary: g_loss={g_avg:.4f} d_loss={d_avg:.4f} "
                        f"time={int(time.time() - start)}s")
            if self.rank == 0:
                with open(os.path.join(out_dir, "train_loss.txt"), "a") as f:  # format: base_trainer.py:475-479
                    f.write(f"Epoch: {epoch + 1} \tG loss: {g_avg:.4f} \tD Loss: {d_avg:.4f}\n")
            if self.sch_g is not None:
                self.sch_d.step()
                self.sch_g.step()
            if epoch % cfg.trainer.save_interval == 0:      # every rank validates its share; rank 0 writes
                self._validate_and_save(epoch, ds, n_train, n_val, out_dir)

    def _validate_and_save(self, epoch, ds, n_train, n_val, out_dir) -> None:
        """Checkpoint (same file names / state_dict keys as base_trainer.py:521-533) and the validation pass of
        base_trainer.py:535-595 on the GPU: tone-mapped uint8 images (tensor2img), MRSE on the linear radiance, PSNR and
        SSIM on the images (``pixel_heal_thyself_b200.metrics``), and the reference's ``evaluation.txt`` line."""
        from .. import metrics as M
        dp = self.world > 1 and torch.distributed.is_initialized()
        if dp and self.D is not None:
            # the critic's BatchNorm running statistics are rank-local: save their average, keep the ranks identical
            for b in self.D.buffers():
                if b.dtype.is_floating_point:
                    torch.distributed.all_reduce(b, op=torch.distributed.ReduceOp.SUM)
                    b.div_(self.world)
        if self.rank == 0:
            path = os.path.join(out_dir, f"model_epoch{epoch + 1}")
            os.makedirs(path, exist_ok=True)
            torch.save(self.G.state_dict(), os.path.join(path, "G.pt"))
            if self.D is not None:
                torch.save(self.D.state_dict(), os.path.join(path, "D.pt"))
        self.G.eval()
        avg_mrse = avg_psnr = avg_ssim = 0.0
        cnt = 0
        with torch.no_grad():
            # validation patches are strided over the ranks (no rank idles inside a collective while rank 0 validates)
            for k in range(n_train + (self.rank if dp else 0), n_train + n_val, self.world if dp else 1):
                noisy, gt_log, aux = ds.batch_device(torch.tensor([k], device=self.device))
                out = self.G(noisy, aux)
                gt = torch.expm1(gt_log)                     # the reference validates against the un-preprocessed gt
                avg_mrse += M.calculate_rmse(out, gt, output_is_log=True)                    # base_trainer.py:566
                psnr, ssim = M.image_metrics(M.tensor2img(out, post_spec=True), M.tensor2img(gt))   # :567-568
                avg_psnr += psnr
                avg_ssim += ssim
                cnt += 1
        self.G.train()
        if dp:
            t = torch.tensor([avg_mrse, avg_psnr, avg_ssim, float(cnt)], dtype=torch.float64, device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
            avg_mrse, avg_psnr, avg_ssim, cnt = float(t[0]), float(t[1]), float(t[2]), int(t[3])
        cnt = max(cnt, 1)
        avg_mrse, avg_psnr, avg_ssim = avg_mrse / cnt, avg_psnr / cnt, avg_ssim / cnt
        if self.rank != 0:
            return
        logger.info(f"[Val] epoch={epoch + 1} summary: avg_mrse={avg_mrse:.4f} avg_psnr={avg_psnr:.4f} "
                    f"avg_1-ssim={1 - avg_ssim:.4f}")
        with open(os.path.join(out_dir, "evaluation.txt"), "a") as f:   # format: base_trainer.py:591-595
            f.write(f"Validation: {epoch + 1} \tAvg MRSE: {avg_mrse:.4f} \tAvg PSNR: {avg_psnr:.4f} "
                    f"\tAvg 1-SSIM: {1 - avg_ssim:.4f}\n")
