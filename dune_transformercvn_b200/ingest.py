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
            divisor: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(nnz,C) values [f32|u8] + (nnz,3) int32 [image,y,x] -> (N,C,H,W) fp32, value/divisor fused."""
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
    _lib.check(L.tcvn_densify(_lib.ptr(coords), _lib.ptr(values), vd, nnz, c, num_images, h, w, float(divisor),
                              _lib.ptr(out), _lib.TCVN_NCHW_F32, _lib.stream_ptr(values.device)), "tcvn_densify")
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
