"""Single-event export surface: what the reference ships to LArSoft.

Mirrors the three wrapper modules of the reference's ``CreateCompiled.ipynb`` (cells 6-8: ``DynamicSimplifedNetwork``,
``DynamicEmbeddingNetwork``, ``DynamicCombinedNetwork``; README.md:70-78): ONE event given as ``(1 + Npng, 3, 400, 280)``
pixel maps with 0..255 values (image 0 = the event map, the rest its prongs) ->

  ``pid``         (event probabilities (E,), prong probabilities (Npng, P))
  ``embeddings``  (event hidden vector (128,), prong hidden vectors (Npng, 128))     - encoder output, before the heads
  ``combined``    all four

The reference freezes these with ``torch.jit.script`` (a tracing compiler over ~700 ATen calls per event).  Here the
batch-1 plan is a **CUDA graph**: the ~310 kernel launches of one event (two DenseNets on two streams, the fused
token/encoder/heads kernel, the softmaxes) are captured once per prong count and replayed with one ``cudaGraphLaunch``,
so single-event latency is bound by the kernels, not by launch overhead.  ``torch.ops.tcvn.classify_event`` registers the
same call as a ``torch.library`` custom op (with a shape-only fake implementation), so ``torch.export`` / ``torch.compile``
of a module that calls it work.  There is no CPU implementation: CPU tensors raise.
"""
from __future__ import annotations

import itertools
import weakref
from typing import Dict, Tuple

import torch
from torch import nn

from . import lib as _lib
from .network import _PRECISIONS, NeutrinoDenseNetwork

_OUTPUTS = ("pid", "embeddings", "combined")


class _Plan:
    """Static buffers + the captured graph of one prong count."""

    def __init__(self):
        self.pixels = None
        self.out: Tuple[torch.Tensor, ...] = ()
        self.graph = None
        self.generation = -1   # engine.generation at capture: the graph bakes in workspace / packed-block pointers


class EventClassifier(nn.Module):
    """Batch-1 inference module with the I/O of the reference's exported TorchScript files."""

    def __init__(self, network: NeutrinoDenseNetwork, outputs: str = "combined", log_pixels: bool = False,
                 use_graph: bool = True):
        super().__init__()
        if outputs not in _OUTPUTS:
            raise ValueError(f"outputs must be one of {_OUTPUTS}")
        if network.training:
            raise _lib.TcvnError("EventClassifier wraps an eval-mode network (call .eval() first, as CreateCompiled cell 2 does)")
        self.network = network
        self.outputs = outputs
        self.log_pixels = bool(log_pixels)
        self.use_graph = bool(use_graph)
        self._plans: Dict[Tuple[int, str], _Plan] = {}
        self._handle = _register(self)

    # ---- the un-captured computation (what the graph records) --------------------------------------------------
    def _run(self, pixels: torch.Tensor):
        net = self.network
        eng = net.engine
        prec = _PRECISIONS[net.precision]
        n = pixels.shape[0]
        dev = pixels.device
        # CreateCompiled cell 6: log(pixels.float() + 1) or pixels.float() / 255 (true division, like the trainer's :60)
        px = torch.log(pixels.float() + 1) if self.log_pixels else pixels.float() / 255
        eng.ensure_packed(prec)
        main = torch.cuda.current_stream(dev)
        if net._side is None or net._side.device != dev:
            net._side = torch.cuda.Stream(device=dev)
        side = net._side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ev = eng.cnn("event", px[:1], prec, ws_kind="cnn_event")
        pr = eng.cnn("prong", px[1:], prec)
        main.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():
            ev.record_stream(main)
        event_mask = torch.ones((1, 1), dtype=torch.bool, device=dev)
        prong_mask = torch.ones((1, n - 1), dtype=torch.bool, device=dev)
        _, hidden, ev_logits, pr_logits = eng.seq(_lib.SEQ_TOKENS | _lib.SEQ_ENCODER | _lib.SEQ_HEADS, ev, pr, event_mask,
                                                  prong_mask)
        event = torch.softmax(ev_logits[0], 0)
        prongs = torch.softmax(pr_logits[0], 1)
        if event.shape[-1] > 4:   # CreateCompiled cell 6: flavour x interaction-type classes folded to 4 flavours
            event = torch.stack((event[:4].sum(), event[4:8].sum(), event[8], event[9]), dim=0)
        return event, prongs, hidden[0, 0], hidden[1:, 0]

    def _select(self, full):
        if self.outputs == "pid":
            return full[0], full[1]
        if self.outputs == "embeddings":
            return full[2], full[3]
        return full

    def classify(self, pixels: torch.Tensor):
        """All four outputs (event probabilities, prong probabilities, event vector, prong vectors)."""
        _lib.require_cuda(pixels, "pixels")
        h, w = self.network.image_size
        pixels = pixels.reshape(-1, self.network.pixel_dim, h, w)
        n = pixels.shape[0]
        if n < 2:
            raise _lib.TcvnError("EventClassifier needs the event map and at least one prong map (the dataset forces "
                                 "prong_mask[:, 0] = True, minkowski_dataset.py:181)")
        if not self.use_graph:
            with torch.no_grad():
                ev, pr, h_ev, h_pr = self._run(pixels)
            return ev, pr, h_ev.clone(), h_pr.clone()   # the two hidden vectors are views of one buffer
        key = (n, str(pixels.dtype), str(pixels.device))
        plan = self._plans.get(key)
        if plan is not None and plan.generation != self.network.engine.generation:
            # a workspace grew (a plan for more prongs, a batched forward on the same network) or the parameters were
            # re-packed since this graph was captured: its pointers may be dangling.  Capture again over the current buffers.
            plan = None
        if plan is None:
            plan = _Plan()
            plan.pixels = torch.empty_like(pixels, memory_format=torch.contiguous_format)
            plan.pixels.copy_(pixels)
            self.network.freeze_packed(False)
            with torch.no_grad():
                # warm-up on a side stream (allocates the workspaces, packs the parameters, sets kernel attributes),
                # then record the same call sequence
                s = torch.cuda.Stream(device=pixels.device)
                s.wait_stream(torch.cuda.current_stream(pixels.device))
                with torch.cuda.stream(s):
                    for _ in range(2):
                        self._run(plan.pixels)
                torch.cuda.current_stream(pixels.device).wait_stream(s)
                self.network.freeze_packed(True)   # serving: the packed parameter block is not re-checked per call
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    plan.out = self._run(plan.pixels)
                plan.graph = g
                plan.generation = self.network.engine.generation
            self._plans[key] = plan
        else:
            plan.pixels.copy_(pixels)
        plan.graph.replay()
        return tuple(t.clone() for t in plan.out)

    def forward(self, pixels: torch.Tensor):
        return self._select(self.classify(pixels))

    def invalidate(self) -> None:
        """Drop the captured graphs (after the network's parameters changed)."""
        self._plans.clear()
        self.network.freeze_packed(False)


# ---- torch.library custom op: the traceable / exportable form of the same call ----------------------------------
# handle -> classifier; weak, so that a dropped classifier (network, workspaces, captured graphs) can be collected
_REGISTRY: "weakref.WeakValueDictionary[int, EventClassifier]" = weakref.WeakValueDictionary()
_handles = itertools.count(1)
_op_defined = False


def _register(mod: "EventClassifier") -> int:
    _define_op()
    handle = next(_handles)
    _REGISTRY[handle] = mod
    return handle


def _lookup(handle: int) -> "EventClassifier":
    mod = _REGISTRY.get(handle)
    if mod is None:
        raise _lib.TcvnError(f"tcvn::classify_event: classifier handle {handle} is gone (the EventClassifier was deleted)")
    return mod


def _define_op() -> None:
    global _op_defined
    if _op_defined:
        return
    _op_defined = True

    @torch.library.custom_op("tcvn::classify_event", mutates_args=())
    def classify_event(pixels: torch.Tensor, handle: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        return _lookup(handle).classify(pixels)

    @classify_event.register_fake
    def _(pixels, handle):
        mod = _lookup(handle)
        net = mod.network
        h, w = net.image_size
        n = pixels.numel() // (net.pixel_dim * h * w)
        e = 4 if net.num_event_classes > 4 else net.num_event_classes
        f32 = dict(dtype=torch.float32, device=pixels.device)
        return (torch.empty((e,), **f32), torch.empty((n - 1, net.num_prong_classes), **f32),
                torch.empty((net.options.hidden_dim,), **f32), torch.empty((n - 1, net.options.hidden_dim), **f32))


class ExportableEventClassifier(nn.Module):
    """``torch.export`` / ``torch.compile``-friendly wrapper: its forward is the single custom-op call."""

    def __init__(self, classifier: EventClassifier):
        super().__init__()
        self.handle = classifier._handle
        self.outputs = classifier.outputs
        self._keep = (classifier,)

    def forward(self, pixels: torch.Tensor):
        full = torch.ops.tcvn.classify_event(pixels, self.handle)
        if self.outputs == "pid":
            return full[0], full[1]
        if self.outputs == "embeddings":
            return full[2], full[3]
        return full
