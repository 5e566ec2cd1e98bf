"""Drop-in for the reference's COO -> dense densification.

``sparse_to_dense(features, coordinates, image_size)`` has the signature, argument meaning and
result layout of transformercvn/network/trainers/neutrino_full_dense_trainer.py:15-24;
``densify`` is the fused form (value scaling inside the kernel, image count supplied by the
caller so there is no device->host sync).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import lib as _lib


def densify(values: torch.Tensor, coords: torch.Tensor, image_size: Sequence[int], num_images: Optional[int] = None,
            divisor: float = 0.0, out: Optional[torch.Tensor] = None, noise_std: float = 0.0, seed: int = 0) -> torch.Tensor:
    """(nnz,C) values [f32|u8] + (nnz,3) int32 [image,y,x] (sorted by image, as the reference's file format stores them)
    -> (N,C,H,W) fp32, value/divisor fused.  ``noise_std`` > 0 also fuses the training-time multiplicative pixel noise of
    ``preprocess_pixels`` (neutrino_full_dense_trainer.py:62-65), drawn from a counter hash of (seed, hit, channel)."""
    _lib.require_cuda(values, "densify(values)")
    _lib.require_cuda(coords, "densify(coords)")
    L = _lib.load()
    h, w = int(image_size[0]), int(image_size[1])
    if coords.dtype != torch.int32:
        coords = coords.to(torch.int32)
    coords = coords.contiguous()
    if values.dtype == torch.uint8:
        vd = _lib.TCVN_VAL_U8
    else:
        vd = _lib.TCVN_VAL_F32
        if values.dtype != torch.float32:
            values = values.float()
    values = values.contiguous()
    nnz = coords.shape[0]
    if num_images is None:
        # the reference does exactly this host sync (neutrino_full_dense_trainer.py:19)
        num_images = int(coords[-1, 0].item()) + 1 if nnz else 0
    c = values.shape[1] if values.dim() == 2 else 1
    if out is None:
        out = torch.empty((num_images, c, h, w), dtype=torch.float32, device=values.device)
    with torch.cuda.device(values.device):
        _lib.check(L.tcvn_densify_noise(_lib.ptr(coords), _lib.ptr(values), vd, nnz, c, num_images, h, w, float(divisor),
                                        float(noise_std), int(seed) & 0xFFFFFFFFFFFFFFFF, _lib.ptr(out), _lib.TCVN_NCHW_F32,
                                        _lib.stream_ptr(values.device)), "tcvn_densify_noise")
    return out


def sparse_to_dense(features: torch.Tensor, coordinates: torch.Tensor, image_size) -> torch.Tensor:
    """Same contract as the reference function of this name (features already scaled by the caller)."""
    return densify(features, coordinates, image_size)


def collate_sparse(coordinates, values, masks):
    """Drop-in for ``MinkowskiCollection.collate_sparse`` (transformercvn/dataset/minkowski_dataset.py:34-47) for samples
    that already live on the GPU: lists of per-event ``(nnz_e,3)`` int32 coordinates ``[image-in-event, y, x]``,
    ``(nnz_e,C)`` values and ``(L,)`` bool masks -> ``(coords (nnz,3)`` with batch-global image indices, ``values
    (nnz,C))``.  One scan + one pass over the hits on the device; no per-event ``.item()`` sync."""
    import ctypes as C
    if len(coordinates) == 0:
        raise _lib.TcvnError("collate_sparse: empty batch")
    _lib.require_cuda(coordinates[0], "collate_sparse(coordinates)")
    L = _lib.load()
    dev = coordinates[0].device
    coords = torch.cat([c.to(torch.int32) for c in coordinates]).contiguous()
    vals = torch.cat(list(values))
    hits = torch.tensor([int(c.shape[0]) for c in coordinates], dtype=torch.int32).to(dev, non_blocking=True)  # shapes: host facts
    m = torch.stack([mk.to(torch.uint8) for mk in masks]).contiguous()
    b, slots = m.shape
    nbytes = L.tcvn_collate_workspace_bytes(b)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty_like(coords)
    _lib.check(L.tcvn_collate_coords(_lib.ptr(coords), coords.shape[0], _lib.ptr(hits), None, _lib.ptr(m), b, slots, _lib.ptr(out),
                                     _lib.ptr(ws), nbytes, _lib.stream_ptr(dev)), "tcvn_collate_coords")
    return out, vals


class Prefetcher:
    """Host -> device staging of the collated hit lists on a copy stream, one batch ahead of the compute stream.

    The reference's DataLoader hands ``training_step`` pageable host tensors that Lightning copies synchronously
    (trainers/neutrino_full_base_trainer.py:85: no ``pin_memory``); at B200 speeds that copy (14.5 MB for 256 events) is
    3-4 % of an inference step.  ``submit(batch)`` starts the copy of a pinned batch, ``take()`` makes the compute stream
    wait for the oldest submitted batch and returns it (its memory is tied to the compute stream with
    ``record_stream``, so the caching allocator cannot hand it out again while kernels still read it)."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TcvnError("Prefetcher stages batches for the CUDA path (there is no CPU implementation)")
        self.stream = torch.cuda.Stream(device=self.device)
        self.queue = []

    def submit(self, host_batch) -> None:
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            dev_batch = host_batch.to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self.queue.append((dev_batch, ready, compute))

    def take(self):
        if not self.queue:
            raise _lib.TcvnError("Prefetcher.take() without a submitted batch")
        dev_batch, ready, _ = self.queue.pop(0)
        compute = torch.cuda.current_stream(self.device)
        compute.wait_event(ready)
        for t in dev_batch.tensors():
            t.record_stream(compute)
        return dev_batch
