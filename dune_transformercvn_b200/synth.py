"""Seeded synthetic inputs and weights for the hot path (no dataset, no checkpoint is reachable).

Input format = what the reference's collate hands to the trainer
(transformercvn/dataset/minkowski_dataset.py:34-86): Minkowski-style COO hit lists
``coords (nnz,3) int32 [image, y, x]`` sorted by (image, y, x) with batch-global image
indices, ``values (nnz,3)``, prefix-true ``prong_mask (B,L)``, ``event_mask (B,1)``.
Recipe and occupancies follow SURVEY.md §8(d).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .config import PIXEL_CHANNELS, PIXEL_H, PIXEL_W
from .params import TensorSpec


@dataclass
class SparseBatch:
    features: torch.Tensor        # (B, L, F) fp32 zeros
    extra: torch.Tensor           # (B, E) fp32 zeros
    event_coords: torch.Tensor    # (nnz_e, 3) int32
    event_values: torch.Tensor    # (nnz_e, 3) fp32 (or uint8)
    event_mask: torch.Tensor      # (B, 1) bool
    prong_coords: torch.Tensor    # (nnz_p, 3) int32
    prong_values: torch.Tensor    # (nnz_p, 3)
    prong_mask: torch.Tensor      # (B, L) bool, prefix-true
    prongs_per_event: List[int]

    @property
    def num_events(self) -> int:
        return self.prong_mask.shape[0]

    @property
    def num_prongs(self) -> int:
        return int(sum(self.prongs_per_event))

    def to(self, device, non_blocking: bool = False) -> "SparseBatch":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return SparseBatch(mv(self.features), mv(self.extra), mv(self.event_coords), mv(self.event_values),
                           mv(self.event_mask), mv(self.prong_coords), mv(self.prong_values), mv(self.prong_mask),
                           list(self.prongs_per_event))

    def pin(self) -> "SparseBatch":
        mv = lambda t: t.pin_memory()
        return SparseBatch(mv(self.features), mv(self.extra), mv(self.event_coords), mv(self.event_values),
                           mv(self.event_mask), mv(self.prong_coords), mv(self.prong_values), mv(self.prong_mask),
                           list(self.prongs_per_event))

    def tensors(self):
        return (self.features, self.extra, self.event_coords, self.event_values, self.event_mask,
                self.prong_coords, self.prong_values, self.prong_mask)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    def select_events(self, lo: int, hi: int) -> "SparseBatch":
        """Events [lo, hi) as a batch of their own - the shard a data-parallel rank receives (hit lists re-based to the
        shard's first image, masks trimmed to the shard's longest event, like the reference's collate of those events)."""
        ppe = list(self.prongs_per_event)
        p_lo, p_hi = sum(ppe[:lo]), sum(ppe[:hi])
        width = max(ppe[lo:hi])

        def cut(coords, values, a, b):
            img = coords[:, 0]
            keep = (img >= a) & (img < b)
            c = coords[keep].clone()
            c[:, 0] -= a
            return c, values[keep].clone()

        ec, ev = cut(self.event_coords, self.event_values, lo, hi)
        pc, pv = cut(self.prong_coords, self.prong_values, p_lo, p_hi)
        return SparseBatch(self.features[lo:hi, :width].clone(), self.extra[lo:hi].clone(), ec, ev, self.event_mask[lo:hi].clone(),
                           pc, pv, self.prong_mask[lo:hi, :width].clone(), ppe[lo:hi])


def _hits(rng: np.random.Generator, n_images: int, occupancy: float, h: int, w: int, value_dtype):
    n_hit = max(1, int(round(occupancy * h * w)))
    coords = np.empty((n_images * n_hit, 3), dtype=np.int32)
    for i in range(n_images):
        pos = np.sort(rng.choice(h * w, size=n_hit, replace=False))
        sl = slice(i * n_hit, (i + 1) * n_hit)
        coords[sl, 0] = i
        coords[sl, 1] = pos // w
        coords[sl, 2] = pos % w
    vals = rng.integers(1, 256, size=(n_images * n_hit, PIXEL_CHANNELS)).astype(value_dtype)
    return coords, vals


def make_batch(num_events: int, seed: int = 1234, max_prongs: int = 10, fixed_prongs: Optional[int] = None,
               event_occupancy: float = 0.01, prong_occupancy: float = 0.002,
               h: int = PIXEL_H, w: int = PIXEL_W, value_dtype=np.float32,
               features_dim: int = 1, extra_dim: int = 1,
               prongs_per_event: Optional[Sequence[int]] = None) -> SparseBatch:
    """One collated batch.  ``fixed_prongs`` = P for every event (config 5 uses 20)."""
    rng = np.random.default_rng(seed)
    if prongs_per_event is None:
        if fixed_prongs is not None:
            prongs_per_event = [int(fixed_prongs)] * num_events
        else:
            prongs_per_event = [int(p) for p in rng.integers(1, max_prongs + 1, size=num_events)]
    prongs_per_event = list(prongs_per_event)
    L = max(prongs_per_event)
    T = sum(prongs_per_event)
    ec, ev = _hits(rng, num_events, event_occupancy, h, w, value_dtype)
    pc, pv = _hits(rng, T, prong_occupancy, h, w, value_dtype)
    mask = np.zeros((num_events, L), dtype=bool)
    for b, p in enumerate(prongs_per_event):
        mask[b, :p] = True
    return SparseBatch(
        features=torch.zeros(num_events, L, features_dim),
        extra=torch.zeros(num_events, extra_dim),
        event_coords=torch.from_numpy(ec), event_values=torch.from_numpy(ev),
        event_mask=torch.ones(num_events, 1, dtype=torch.bool),
        prong_coords=torch.from_numpy(pc), prong_values=torch.from_numpy(pv),
        prong_mask=torch.from_numpy(mask), prongs_per_event=prongs_per_event)


def balanced_prongs(num_events: int, seed: int, max_prongs: int = 10, total: Optional[int] = None) -> List[int]:
    """Prong counts in [1, max_prongs] whose sum is forced to ``total`` (equal work per rank)."""
    rng = np.random.default_rng(seed)
    p = rng.integers(1, max_prongs + 1, size=num_events).astype(np.int64)
    if total is None:
        total = int(round(num_events * (1 + max_prongs) / 2))
    total = min(max(total, num_events), num_events * max_prongs)
    i = 0
    while p.sum() != total:
        j = i % num_events
        if p.sum() < total and p[j] < max_prongs:
            p[j] += 1
        elif p.sum() > total and p[j] > 1:
            p[j] -= 1
        i += 1
    return [int(x) for x in p]


def init_state(specs: Sequence[TensorSpec], seed: int = 0, perturb: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic CPU state_dict for ``specs``.

    ``perturb=False`` mimics a freshly constructed reference network (BN gamma 1 / beta 0 /
    mean 0 / var 1, PReLU 0.25, LayerNorm 1/0, fan-in uniform conv/linear weights).
    ``perturb=True`` additionally randomises every normalisation/activation tensor so that
    parity tests are sensitive to each of them.
    """
    g = torch.Generator().manual_seed(seed)
    u = lambda shape, a: (torch.rand(shape, generator=g) * 2 - 1) * a
    state: Dict[str, torch.Tensor] = {}
    last_fan_in = 1
    for s in specs:
        r = s.role
        if r in ("conv_w", "lin_w"):
            fan_in = 1
            for d in s.shape[1:]:
                fan_in *= d
            last_fan_in = fan_in
            t = u(s.shape, 1.0 / math.sqrt(fan_in))
        elif r in ("conv_b", "lin_b"):
            t = u(s.shape, 1.0 / math.sqrt(last_fan_in))
        elif r == "attn_w":
            fan_out, fan_in = s.shape
            t = u(s.shape, math.sqrt(6.0 / (fan_in + fan_out)))
        elif r == "attn_b":
            t = u(s.shape, 0.05) if perturb else torch.zeros(s.shape)
        elif r in ("bn_w", "ln_w"):
            t = 1.0 + u(s.shape, 0.3) if perturb else torch.ones(s.shape)
        elif r in ("bn_b", "ln_b"):
            t = u(s.shape, 0.2) if perturb else torch.zeros(s.shape)
        elif r == "bn_rm":
            t = u(s.shape, 0.05) if perturb else torch.zeros(s.shape)
        elif r == "bn_rv":
            t = 0.6 + torch.rand(s.shape, generator=g) * 0.8 if perturb else torch.ones(s.shape)
        elif r == "bn_nbt":
            t = torch.zeros((), dtype=torch.int64)
        elif r == "prelu":
            t = 0.25 + u(s.shape, 0.15) if perturb else torch.full(s.shape, 0.25)
        elif r == "pos":
            t = torch.randn(s.shape, generator=g)
        else:
            raise ValueError(r)
        if "out_proj.bias" in s.name and not perturb:
            t = torch.zeros(s.shape)
        state[s.name] = t.contiguous()
    return state


def state_checksum(state: Dict[str, torch.Tensor]) -> float:
    """Order-independent fingerprint used to detect RNG drift between images."""
    acc = 0.0
    for k in sorted(state):
        t = state[k]
        if t.dtype.is_floating_point:
            acc += float(t.double().abs().sum()) + 0.5 * float(t.double().sum())
    return acc
