"""Synthetic EXR-shaped data and the patch pipeline feeding the trainer.

The reference reads EXR frames with pyexr, samples patch positions on the CPU
(preprocessing.py:179-213), crops them into an HDF5 file (gen_hdf5.py:87-195)
and preprocesses every batch with numpy before the H2D copy
(base_trainer.py:373-383).  None of pyexr / h5py exist in this image and
BASELINE.json asks for synthetic EXR-shaped tensors, so here:

* ``synthetic_frames`` makes radiance + aux frames of the reference's layout
  (NHWC fp32: noisy/gt 3 ch, aux = normal(3) | depth(1) | albedo(3));
* patch positions come from the CUDA importance sampler (importance map + bit-exact dart throwing and
  error-diffusion pruning, preprocessing.py:119-322);
* ``PatchDataset.batch_device`` gathers + preprocesses a batch straight from
  the HBM-resident frames in one fused pass (``pht_crop_preprocess``);
* ``PatchDataset.batch_host`` returns the raw NHWC patches in pinned host
  memory -- what the reference's DataLoader yields -- for the end-to-end path
  (H2D copy + ``pht_preprocess`` per step).
"""
from __future__ import annotations

import torch

from . import ops


def synthetic_frames(num_images: int, height: int, width: int, seed: int, device) -> dict[str, torch.Tensor]:
    """Seeded synthetic frames: gt = smoothed exp(N(0,1)) radiance, noisy = gt * Gamma(2, 0.5)
    multiplicative Monte-Carlo noise, unit normals with 1% NaNs, depth and albedo in [0, 1]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n, h, w = num_images, height, width
    gt = torch.exp(torch.randn(n, 3, h, w, generator=g))
    gt = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(gt, (2, 2, 2, 2), mode="replicate"), 5, stride=1)
    u1 = torch.rand(n, 3, h, w, generator=g).clamp_min(1e-7)
    u2 = torch.rand(n, 3, h, w, generator=g).clamp_min(1e-7)
    noisy = gt * (-(u1.log() + u2.log()) * 0.5)          # Gamma(k=2, theta=0.5) = -0.5 * (ln u1 + ln u2)
    normal = torch.rand(n, 3, h, w, generator=g) * 2 - 1
    normal = normal / normal.norm(dim=1, keepdim=True).clamp_min(1e-6)
    normal[torch.rand(n, 3, h, w, generator=g) < 0.01] = float("nan")
    depth = torch.rand(n, 1, h, w, generator=g)
    albedo = torch.rand(n, 3, h, w, generator=g)
    aux = torch.cat([normal, depth, albedo], 1)
    to = lambda t: t.permute(0, 2, 3, 1).contiguous().to(device)
    return {"noisy": to(noisy), "gt": to(gt), "aux": to(aux)}


class PatchDataset:
    """Frames resident in HBM + sampled patch centres."""

    def __init__(self, frames: dict[str, torch.Tensor], patch_size: int, num_patches: int, seed: int,
                 importance: bool = True):
        """``importance=True`` is the reference's get_cropped_patches (preprocessing.py:347-359): dart throwing
        followed by importance pruning, so every image yields <= num_patches patches; ``False`` keeps every dart."""
        self.frames, self.P = frames, patch_size
        dev = frames["noisy"].device
        n_img, hf, wf, _ = frames["noisy"].shape
        seeds = torch.arange(n_img, dtype=torch.int64, device=dev) + seed
        img = torch.arange(n_img, dtype=torch.int32, device=dev).repeat_interleave(num_patches)
        if importance:
            self.importance_map = ops.importance_map(frames["noisy"], frames["aux"], patch_size)
            centres, counts = ops.importance_sample(seeds, self.importance_map, patch_size, num_patches)
            if bool((counts < 0).any()):
                raise RuntimeError("patch sampler: dart throwing did not converge")
            keep = (centres[:, :, 0] >= 0).reshape(-1)
            self.centres = centres.reshape(-1, 2)[keep].contiguous()
            self.img_idx = img[keep].contiguous()
            self.counts = counts
        else:
            corners = ops.sample_patches(seeds, (hf, wf), patch_size, num_patches)       # [n_img, n, 2] (x, y)
            # importance_sampling() shifts corners by P/2 to centres (preprocessing.py:309-322)
            self.centres = (corners + patch_size // 2).reshape(-1, 2).contiguous()
            self.img_idx = img.contiguous()
            self.counts = torch.full((n_img,), num_patches, dtype=torch.int32, device=dev)
        self._host = None

    def __len__(self) -> int:
        return self.centres.shape[0]

    def batch_device(self, idx: torch.Tensor):
        """idx: int64 [B] (device).  -> (noisy, gt, aux) NCHW fp32, preprocessed, on the device."""
        dev, P, B = self.centres.device, self.P, idx.numel()
        c = self.centres.index_select(0, idx).contiguous()
        ii = self.img_idx.index_select(0, idx).contiguous()
        noisy = torch.empty(B, 3, P, P, device=dev)
        gt = torch.empty(B, 3, P, P, device=dev)
        aux = torch.empty(B, 7, P, P, device=dev)
        ops.crop_preprocess(self.frames["noisy"], self.frames["gt"], self.frames["aux"], c, P, noisy, gt, aux, ii)
        return noisy, gt, aux

    def host_patches(self) -> dict[str, torch.Tensor]:
        """Raw NHWC fp32 patches in pinned host memory (the reference's HDF5 rows, gen_hdf5.py:135-139)."""
        if self._host is None:
            P, h = self.P, self.P // 2
            c, ii = self.centres.cpu(), self.img_idx.cpu()
            out = {}
            for key, fr in self.frames.items():
                frc = fr.cpu()
                t = torch.stack([frc[ii[k], c[k, 1] - h:c[k, 1] + h, c[k, 0] - h:c[k, 0] + h, :]
                                 for k in range(len(c))])
                out[key] = t.pin_memory() if torch.cuda.is_available() else t
            self._host = out
        return self._host


def preprocess_host_batch(batch: dict[str, torch.Tensor], device):
    """The per-step input path of base_trainer.py:373-383 on the device: pinned NHWC fp32 host patches ->
    H2D copy -> ``pht_preprocess`` -> NCHW fp32 (noisy, gt, aux)."""
    n = batch["noisy"].to(device, non_blocking=True)
    g = batch["gt"].to(device, non_blocking=True)
    a = batch["aux"].to(device, non_blocking=True)
    B, P = n.shape[0], n.shape[1]
    noisy = torch.empty(B, 3, P, P, device=device)
    gt = torch.empty(B, 3, P, P, device=device)
    aux = torch.empty(B, 7, P, P, device=device)
    ops.preprocess(n, g, a, noisy, gt, aux)
    return noisy, gt, aux


class DevicePrefetcher:
    """Iterates over pinned host batches and yields preprocessed device batches, with the H2D copy + ``pht_preprocess`` of
    batch i+1 enqueued on a side stream while step i runs (the reference hides its input path the same way:
    DataLoader workers + ``BackgroundGenerator`` prefetch thread, prefetch_dataloader.py:7-12).  Usage::

        for noisy, gt, aux in DevicePrefetcher(host_batches, device):
            trainer.train_step(noisy, gt, aux)
    """

    def __init__(self, host_batches, device, stream=None):
        """``stream``: reuse one side stream across prefetchers (the caching allocator keeps a pool per stream: a fresh
        stream's first copies cudaMalloc, which synchronises the device)."""
        self.batches, self.device = host_batches, device
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)

    def _stage(self, batch):
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            return preprocess_host_batch(batch, self.device)

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            torch.cuda.current_stream().wait_stream(self.stream)    # batch i is on the device
            cur = nxt
            for t in cur:
                t.record_stream(torch.cuda.current_stream())
            try:
                nxt = self._stage(next(it))                           # batch i+1: copy + preprocess behind step i
            except StopIteration:
                nxt = None
            yield cur
