"""Drop-in for ``transformercvn.network``'s dense network, running on libtcvn's CUDA kernels.

``NeutrinoDenseNetwork`` keeps the reference's constructor, ``forward`` signature, sub-module
surface and ``state_dict`` names/shapes/order
(transformercvn/network/networks/neutrino_full_dense_network.py:19-21,
neutrino_full_base_network.py:128-188), so ``load_state_dict`` of a reference checkpoint, the
weight-decay grouping by parameter name (trainers/neutrino_base.py:116-128) and
``Evaluate.ipynb`` work unchanged.  The fp32 master parameters are ordinary ``nn.Parameter``s;
the kernel-ready (BN-folded, re-laid, bf16) copies are a cache rebuilt when a parameter changes.

All arithmetic happens in libtcvn.so; without it, or on a CPU tensor, every call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import lib as _lib
from . import synth
from .config import PIXEL_H, PIXEL_W
from .params import TensorSpec, embedding_dims, network_specs, prong_decoder_widths

_PRECISIONS = {"fp32": _lib.TCVN_FP32, "bf16": _lib.TCVN_BF16}


class _Node(nn.Module):
    """Name-only container: the reference's module tree is reproduced for its parameter names."""


def _attach(root: nn.Module, spec: TensorSpec, value: torch.Tensor) -> None:
    parts = spec.name.split(".")
    mod = root
    for p in parts[:-1]:
        child = mod._modules.get(p)
        if child is None:
            child = _Node()
            mod.add_module(p, child)
        mod = child
    if spec.is_param:
        mod.register_parameter(parts[-1], nn.Parameter(value))
    else:
        mod.register_buffer(parts[-1], value)


class _Engine:
    """Packed-parameter cache, workspaces and the calls into the C ABI for one network."""

    def __init__(self, owner: "NeutrinoDenseNetwork"):
        self.owner = (owner,)  # tuple: keeps nn.Module.__setattr__ from registering a cycle
        self.key = None
        self.packed: Dict[str, torch.Tensor] = {}
        self.ws: Dict[Tuple, torch.Tensor] = {}
        self.frozen = False
        # bumped whenever a workspace or a packed parameter block is (re)allocated: captured CUDA graphs bake those
        # pointers in, so every plan records the generation it was captured under and is re-captured when it is stale
        self.generation = 0
        self._tensors: Optional[List[torch.Tensor]] = None

    # ---- descriptors ------------------------------------------------------------------------
    def cnn_desc(self, out_features: int) -> _lib.CnnDesc:
        o = self.owner[0].options
        d = _lib.CnnDesc()
        d.in_channels = self.owner[0].pixel_dim
        d.init_features = o.initial_pixel_dim
        d.growth = o.densenet_growth_rate
        d.bn_size = o.densenet_batch_norm_size
        blocks = list(o.densenet_structure)
        d.num_blocks = len(blocks)
        for i, n in enumerate(blocks):
            d.block_layers[i] = n
        d.out_features = out_features
        d.height, d.width = self.owner[0].image_size
        d.bn_eps = 1e-5
        return d

    def seq_desc(self) -> _lib.SeqDesc:
        net = self.owner[0]
        o = net.options
        pix, feat, pos = embedding_dims(o)
        d = _lib.SeqDesc()
        d.hidden, d.heads, d.layers, d.ffn = o.hidden_dim, o.num_attention_heads, o.num_encoder_layers, o.hidden_dim
        d.pixel_dim, d.feature_dim, d.position_dim = pix, feat, pos
        d.num_event_classes, d.num_prong_classes = net.num_event_classes, net.num_prong_classes
        widths = prong_decoder_widths(o)
        d.num_decoder_layers = len(widths)
        for i, w in enumerate(widths):
            d.decoder_widths[i] = w
        d.bn_eps = 1e-5
        d.ln_eps = 1e-5
        return d

    # ---- packing -----------------------------------------------------------------------------
    def _arena(self, tensors: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
        specs = [s for s in self.owner[0].specs if s.name.startswith(prefix) and s.in_arena]
        return torch.cat([tensors[s.name].detach().reshape(-1).float() for s in specs])

    def _state_key(self, prec: int):
        """(device, precision, sum of the tensors' in-place version counters, xor of their addresses): changes whenever
        an optimizer step, ``load_state_dict`` or ``.to()`` touched any parameter or buffer.  The tensor objects are
        listed once (the module tree never changes after construction); walking ``named_parameters()`` on every call
        cost more host time than launching the step."""
        if self._tensors is None:
            net = self.owner[0]
            self._tensors = [t for _, t in net.named_parameters()] + [t for _, t in net.named_buffers()]
        v = 0
        p = 0
        for t in self._tensors:
            v += t._version
            p ^= t.data_ptr()
        return (str(self._tensors[0].device), prec, v, p)

    def ensure_packed(self, prec: int) -> None:
        net = self.owner[0]
        if self.frozen and self.key is not None and self.key[1] == prec:
            return
        key = self._state_key(prec)
        if key == self.key:
            return
        tensors = dict(net.named_parameters())
        tensors.update(dict(net.named_buffers()))
        self.generation += 1
        L = _lib.load()
        dev = next(iter(tensors.values())).device
        if dev.type != "cuda":
            raise _lib.TcvnError("NeutrinoDenseNetwork: parameters are on the CPU; move the module to a CUDA device "
                                 "(this path has no CPU implementation)")
        st = _lib.stream_ptr(dev)
        pix, feat, pos = embedding_dims(net.options)
        pe = "prong_embedding."
        for tag, prefix, width in (("prong", pe + "prong_pixel_embedding.", pix),
                                   ("event", pe + "event_pixel_embedding.", pix + feat)):
            d = self.cnn_desc(width)
            arena = self._arena(tensors, prefix)
            expect = L.tcvn_cnn_arena_floats(C.byref(d))
            if arena.numel() != expect:
                raise _lib.TcvnError(f"{tag} CNN arena has {arena.numel()} floats, library expects {expect}")
            nbytes = L.tcvn_cnn_packed_bytes(C.byref(d), prec)
            buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.tcvn_cnn_pack(C.byref(d), prec, _lib.ptr(arena), _lib.ptr(buf), nbytes, st), "tcvn_cnn_pack")
            self.packed[tag] = buf
        sd = self.seq_desc()
        nbytes = L.tcvn_seq_packed_bytes(C.byref(sd))
        if nbytes == 0:
            raise _lib.TcvnError("sequence descriptor rejected: " + L.tcvn_last_error().decode())
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        position = tensors[pe + "event_position_embedding"].detach().reshape(-1).float().contiguous()
        arenas = [self._arena(tensors, p) for p in (pe + "combined_embedding.", "encoder.", "event_decoder.",
                                                    "prong_decoder.")]
        _lib.check(L.tcvn_seq_pack(C.byref(sd), _lib.ptr(position), *[_lib.ptr(a) for a in arenas], _lib.ptr(buf),
                                   nbytes, st), "tcvn_seq_pack")
        self.packed["seq"] = buf
        self.key = key

    # ---- workspaces --------------------------------------------------------------------------
    def workspace(self, kind: str, nbytes: int, dev) -> torch.Tensor:
        k = (kind, str(dev))
        buf = self.ws.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)  # zero-filled once: ring rows stay finite
            self.ws[k] = buf
            self.generation += 1   # graphs captured over the old buffer must not be replayed
        return buf

    # ---- kernels -----------------------------------------------------------------------------
    def cnn(self, tag: str, pixels: torch.Tensor, prec: int, ws_kind: str = "cnn") -> torch.Tensor:
        L = _lib.load()
        net = self.owner[0]
        pix, feat, _ = embedding_dims(net.options)
        width = pix if tag == "prong" else pix + feat
        d = self.cnn_desc(width)
        n = pixels.shape[0]
        out = torch.empty((n, width), dtype=torch.float32, device=pixels.device)
        if n == 0:
            return out
        if tuple(pixels.shape[1:]) != (d.in_channels, d.height, d.width):
            raise _lib.TcvnError(f"{tag} pixels have shape {tuple(pixels.shape)}, expected (N,{d.in_channels},{d.height},{d.width})")
        pixels = pixels.contiguous().float()
        nbytes = L.tcvn_cnn_workspace_bytes(C.byref(d), prec, n)
        ws = self.workspace(ws_kind, nbytes, pixels.device)
        _lib.check(L.tcvn_cnn_forward(C.byref(d), prec, _lib.ptr(self.packed[tag]), _lib.ptr(pixels), n, _lib.ptr(out),
                                      _lib.ptr(ws), ws.numel(), _lib.stream_ptr(pixels.device)), "tcvn_cnn_forward")
        return out

    def cnn_sparse(self, tag: str, values: torch.Tensor, coords: torch.Tensor, n: int, prec: int,
                   divisor: float = 255.0, ws_kind: str = "cnn") -> torch.Tensor:
        """Pixel-map embedding straight from the COO hit list (the dense map is never built)."""
        L = _lib.load()
        net = self.owner[0]
        pix, feat, _ = embedding_dims(net.options)
        width = pix if tag == "prong" else pix + feat
        d = self.cnn_desc(width)
        out = torch.empty((n, width), dtype=torch.float32, device=values.device)
        if n == 0:
            return out
        if coords.dtype != torch.int32:
            coords = coords.to(torch.int32)
        coords = coords.contiguous()
        if values.dtype == torch.uint8:
            vd = _lib.TCVN_VAL_U8
        else:
            vd = _lib.TCVN_VAL_F32
            values = values.float()
        values = values.contiguous()
        if values.dim() != 2 or values.shape[1] != d.in_channels:
            raise _lib.TcvnError(f"{tag} hit values have shape {tuple(values.shape)}, expected (nnz,{d.in_channels})")
        nbytes = L.tcvn_cnn_workspace_bytes_sparse(C.byref(d), prec, n, coords.shape[0])
        ws = self.workspace(ws_kind, nbytes, values.device)
        _lib.check(L.tcvn_cnn_forward_sparse(C.byref(d), prec, _lib.ptr(self.packed[tag]), _lib.ptr(coords),
                                             _lib.ptr(values), vd, coords.shape[0], float(divisor), n, _lib.ptr(out),
                                             _lib.ptr(ws), ws.numel(), _lib.stream_ptr(values.device)),
                   "tcvn_cnn_forward_sparse")
        return out

    def read_stage(self, tag: str, n: int, stage: int, prec: int, dev) -> torch.Tensor:
        """Test hook: a feature map of the last ``cnn`` call as (n,C,H,W) fp32."""
        L = _lib.load()
        net = self.owner[0]
        pix, feat, _ = embedding_dims(net.options)
        d = self.cnn_desc(pix if tag == "prong" else pix + feat)
        ws = self.ws[("cnn", str(dev))]
        c, h, w = C.c_int32(), C.c_int32(), C.c_int32()
        st = _lib.stream_ptr(dev)
        _lib.check(L.tcvn_cnn_read_stage(C.byref(d), prec, _lib.ptr(ws), n, stage, None, C.byref(c), C.byref(h),
                                         C.byref(w), st), "tcvn_cnn_read_stage")
        out = torch.empty((n, c.value, h.value, w.value), dtype=torch.float32, device=dev)
        _lib.check(L.tcvn_cnn_read_stage(C.byref(d), prec, _lib.ptr(ws), n, stage, _lib.ptr(out), C.byref(c),
                                         C.byref(h), C.byref(w), st), "tcvn_cnn_read_stage")
        return out

    def seq(self, stages: int, event_emb, prong_emb, event_mask, prong_mask, tokens=None, hidden=None):
        L = _lib.load()
        net = self.owner[0]
        sd = self.seq_desc()
        b, l = prong_mask.shape
        dev = prong_mask.device
        s = 1 + l
        pm = prong_mask.contiguous().to(torch.uint8)
        em = None if event_mask is None else event_mask.contiguous().to(torch.uint8)
        f32 = dict(dtype=torch.float32, device=dev)
        if stages & _lib.SEQ_TOKENS and tokens is None:
            tokens = torch.empty((b, s, sd.hidden), **f32)
        if stages & _lib.SEQ_ENCODER and hidden is None:
            hidden = torch.empty((s, b, sd.hidden), **f32)
        ev_logits = pr_logits = None
        if stages & _lib.SEQ_HEADS:
            ev_logits = torch.empty((b, sd.num_event_classes), **f32)
            pr_logits = torch.empty((b, l, sd.num_prong_classes), **f32)
        nbytes = L.tcvn_seq_workspace_bytes(C.byref(sd), b, l)
        ws = self.workspace("seq", nbytes, dev)
        _lib.check(L.tcvn_seq_forward(C.byref(sd), _lib.ptr(self.packed["seq"]), stages, _lib.ptr(event_emb),
                                      _lib.ptr(prong_emb), _lib.ptr(em), _lib.ptr(pm), b, l, _lib.ptr(tokens),
                                      _lib.ptr(hidden), _lib.ptr(ev_logits), _lib.ptr(pr_logits), _lib.ptr(ws),
                                      ws.numel(), _lib.stream_ptr(dev)), "tcvn_seq_forward")
        return tokens, hidden, ev_logits, pr_logits


def _check_eval(mod: nn.Module) -> None:
    if mod.training:
        raise NotImplementedError(
            "dune_transformercvn_b200: in train mode only the whole-network forward (NeutrinoDenseNetwork.forward / "
            "forward_sparse, what the reference's training_step calls) is built; the sub-module surface "
            "(prong_embedding / encoder / decoders called on their own) is eval-only — call .eval(). There is "
            "deliberately no PyTorch fallback")


class ProngEmbedding(_Node):
    """``BaseProngEmbedding`` surface (neutrino_full_base_network.py:87-125): pixels -> tokens, mask."""

    def forward(self, features, extra, event_pixels, event_mask, prong_pixels, prong_mask):
        eng: _Engine = self._engine[0]
        net = eng.owner[0]
        _check_eval(net)
        prec = _PRECISIONS[net.precision]
        _lib.require_cuda(event_pixels, "event_pixels")
        eng.ensure_packed(prec)
        ev = eng.cnn("event", event_pixels, prec)
        pr = eng.cnn("prong", prong_pixels, prec)
        tokens, _, _, _ = eng.seq(_lib.SEQ_TOKENS, ev, pr, event_mask, prong_mask)
        return tokens, torch.cat((event_mask, prong_mask), dim=1)


class ProngEncoder(_Node):
    """``ProngCustomBertEncoder`` surface (prong_custom_bert_encoder.py:57-75)."""

    def forward(self, hidden, sequence_mask):
        eng: _Engine = self._engine[0]
        net = eng.owner[0]
        _check_eval(net)
        _lib.require_cuda(hidden, "hidden")
        eng.ensure_packed(_PRECISIONS[net.precision])
        _, out, _, _ = eng.seq(_lib.SEQ_ENCODER, None, None, sequence_mask[:, :1], sequence_mask[:, 1:],
                               tokens=hidden.contiguous().float())
        return out, ~sequence_mask, sequence_mask


class _HeadBase(_Node):
    def _run(self, x_event: Optional[torch.Tensor], x_prong: Optional[torch.Tensor]):
        """Heads over explicit hidden vectors: builds the (S,B,D) layout the fused kernel reads."""
        eng: _Engine = self._engine[0]
        net = eng.owner[0]
        _check_eval(net)
        eng.ensure_packed(_PRECISIONS[net.precision])
        d = net.options.hidden_dim
        if x_prong is None:
            b, l = x_event.shape[0], 0
            dev = x_event.device
        else:
            l, b = x_prong.shape[0], x_prong.shape[1]
            dev = x_prong.device
        hidden = torch.zeros((1 + l, b, d), dtype=torch.float32, device=dev)
        if x_event is not None:
            hidden[0] = x_event
        if x_prong is not None:
            hidden[1:] = x_prong
        mask = torch.ones((b, max(l, 0)), dtype=torch.bool, device=dev)
        _, _, ev, pr = eng.seq(_lib.SEQ_HEADS, None, None, None, mask, hidden=hidden)
        return ev, pr


class EventDecoder(_HeadBase):
    """``ProngDecoder`` surface (prong_decoder.py:13-16): (B,128) -> (B,num_event_classes)."""

    def forward(self, x):
        _lib.require_cuda(x, "event_decoder input")
        return self._run(x.float(), None)[0]


class ProngDecoder(_HeadBase):
    """``ProngTargetDecoder`` surface (prong_target_decoder.py:35-41): (L,B,128) -> (L,B,num_prong_classes)."""

    def forward(self, x):
        _lib.require_cuda(x, "prong_decoder input")
        return self._run(None, x.float())[1].transpose(0, 1)


class NeutrinoDenseNetwork(nn.Module):
    """Same constructor and forward as the reference class of this name."""

    cnn_kind = "dense"   # "sdxl" in NeutrinoSDXLNetwork (sdxl.py)

    def _make_engine(self) -> "_Engine":
        return _Engine(self)

    def __init__(self, options, features_dim: int, extra_dim: int, pixel_dim: int, num_prong_classes: int,
                 num_event_classes: int, image_size=(PIXEL_H, PIXEL_W), precision: str = "fp32", seed: int = 0):
        super().__init__()
        if not (bool(options.linear_batch_norm) and bool(options.linear_prelu_activation)):
            raise _lib.TcvnError("only linear_batch_norm=True, linear_prelu_activation=True (both shipped configs) are built")
        if not bool(options.disable_smart_features):
            raise _lib.TcvnError("disable_smart_features=False is not built (both shipped configs disable it)")
        if bool(options.one_hot_pixels) or bool(getattr(options, "transformer_norm_first", False)):
            raise _lib.TcvnError("one_hot_pixels / transformer_norm_first are not built (off in both shipped configs)")
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.options = options
        self.features_dim, self.extra_dim, self.pixel_dim = features_dim, extra_dim, pixel_dim
        self.num_prong_classes, self.num_event_classes = num_prong_classes, num_event_classes
        self.image_size = (int(image_size[0]), int(image_size[1]))
        self.precision = precision
        self.specs: List[TensorSpec] = network_specs(options, features_dim, extra_dim, pixel_dim, num_prong_classes,
                                                     num_event_classes, cnn=self.cnn_kind)
        engine = self._make_engine()
        self._engine = (engine,)
        self.overlap_cnns = True     # eval forward_sparse: event CNN on a side stream (+1.7 % at 256 events)
        # ... optionally on its own share of the SMs (tcvn_set_sm_limit).  Measured at 256 events: 21.7 ms with disjoint
        # budgets against 21.2 ms without (the un-partitioned stem / pooling kernels of one stream delay the fixed-grid
        # persistent kernels of the other), so it is off by default
        self.partition_sms = False
        self._side = None
        self._plans: Dict[Tuple, dict] = {}   # CUDA-graph plans of forward_sparse(graph=True), one per batch shape
        self.graph_launches_per_replay = 0
        for name, cls in (("prong_embedding", ProngEmbedding), ("encoder", ProngEncoder),
                          ("event_decoder", EventDecoder), ("prong_decoder", ProngDecoder)):
            m = cls()
            m._engine = (engine,)
            self.add_module(name, m)
        state = synth.init_state(self.specs, seed=seed, perturb=False)
        for s in self.specs:
            _attach(self, s, state[s.name])
        from .training import TrainEngine
        self._train_engine = (TrainEngine(self),)
        for p in self.parameters():
            p._tcvn_engine = self._train_engine

    @property
    def engine(self) -> _Engine:
        return self._engine[0]

    @property
    def train_engine(self):
        return self._train_engine[0]

    def freeze_packed(self, frozen: bool = True) -> None:
        """Serving mode: skip the per-call check for changed parameters."""
        self.engine.frozen = frozen

    def forward(self, features, extra, event_pixels, event_mask, prong_pixels, prong_mask):
        """(B,L,F), (B,E), (B,3,H,W), (B,1) bool, (T,3,H,W), (B,L) bool -> (B,E_cls), (B,L,P_cls).

        neutrino_full_base_network.py:166-188.  ``features``/``extra`` only feed the smart-feature
        embedding, which the shipped configs disable (it contributes zeros).

        In ``.train()`` mode this is the train-mode forward (batch-statistics BatchNorm with running-buffer
        updates, dropout) recorded as ONE autograd node whose backward is the hand-written CUDA backward
        (fp32; training.py)."""
        if self.training:
            from .training import train_forward
            return train_forward(self.train_engine, event_pixels, event_mask, prong_pixels, prong_mask)
        eng = self.engine
        prec = _PRECISIONS[self.precision]
        _lib.require_cuda(event_pixels, "event_pixels")
        _lib.require_cuda(prong_pixels, "prong_pixels")
        eng.ensure_packed(prec)
        ev = eng.cnn("event", event_pixels, prec)
        pr = eng.cnn("prong", prong_pixels, prec)
        _, _, ev_logits, pr_logits = eng.seq(_lib.SEQ_TOKENS | _lib.SEQ_ENCODER | _lib.SEQ_HEADS, ev, pr, event_mask,
                                             prong_mask)
        return ev_logits, pr_logits

    def forward_sparse(self, batch: "synth.SparseBatch", materialize: bool = False, graph: bool = False):
        """Trainer-level path (neutrino_full_base_trainer.py:113-116): /255 + sparse_to_dense + network.

        By default the stem consumes the hit lists directly (no dense map, no host sync);
        ``materialize=True`` runs the literal sequence densify kernel -> dense forward instead.
        ``graph=True`` (eval only): the ~200 kernel launches of the step are captured once per batch SHAPE (events,
        prong slots, images, hit counts) into a CUDA graph and replayed with one launch - the serving loop is then bound
        by the kernels, not by the host issuing them.  A batch of another shape is captured on first sight (pad hit
        lists with zero-valued hits / bucket the batch size upstream to bound the number of plans)."""
        if materialize:
            from .ingest import densify
            # preprocess_pixels (neutrino_full_dense_trainer.py:59-66): /255, and in training the multiplicative Gaussian
            # pixel noise, both fused into the densify kernel (a fresh counter-hash stream per step and per map kind)
            std = float(getattr(self.options, "pixel_noise_std", 0.0)) if self.training else 0.0
            seed, seed_off = self.train_engine.seed_args()      # (value, device-resident per-step offset or None)
            L = _lib.load()
            L.tcvn_set_seed_offset(seed_off if std > 0.0 else None)
            try:
                ev = densify(batch.event_values, batch.event_coords, self.image_size, batch.num_events, 255.0, noise_std=std,
                             seed=seed ^ 0xA5A5A5A500000000)
                pr = densify(batch.prong_values, batch.prong_coords, self.image_size, batch.num_prongs, 255.0, noise_std=std,
                             seed=seed ^ 0x5A5A5A5A00000000)
            finally:
                L.tcvn_set_seed_offset(None)
            return self.forward(batch.features, batch.extra, ev, batch.event_mask, pr, batch.prong_mask)
        if self.training:
            return self.forward_sparse(batch, materialize=True)
        _lib.require_cuda(batch.event_values, "hit values")
        with torch.cuda.device(batch.event_values.device):
            if graph:
                return self._forward_sparse_graph(batch)
            return self._forward_sparse_eager(batch)

    def _forward_sparse_graph(self, batch: "synth.SparseBatch"):
        eng = self.engine
        prec = _PRECISIONS[self.precision]
        eng.ensure_packed(prec)     # a changed parameter re-packs and bumps eng.generation: stale plans are dropped below
        key = (batch.num_events, batch.prong_mask.shape[1], batch.num_prongs, batch.event_coords.shape[0],
               batch.prong_coords.shape[0], str(batch.event_values.dtype), str(batch.event_values.device), prec)
        plan = self._plans.get(key)
        if plan is not None and plan["generation"] != eng.generation:
            plan = None
        if plan is None:
            dev = batch.event_values.device
            static = synth.SparseBatch(*[t.clone().contiguous() for t in batch.tensors()], list(batch.prongs_per_event))
            with torch.no_grad():
                warm = torch.cuda.Stream(device=dev)
                warm.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(warm):          # allocates the workspaces, sets the kernel attributes
                    for _ in range(2):
                        self._forward_sparse_eager(static)
                torch.cuda.current_stream(dev).wait_stream(warm)
                gen = eng.generation
                was_frozen = eng.frozen
                eng.frozen = True                       # no re-pack inside the capture
                l0 = _lib.load().tcvn_launch_count()
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        out = self._forward_sparse_eager(static)
                finally:
                    eng.frozen = was_frozen
                launches = _lib.load().tcvn_launch_count() - l0   # kernels of this library one replay executes
            if len(self._plans) >= 16:                  # bounded: drop the oldest plan
                self._plans.pop(next(iter(self._plans)))
            plan = {"graph": g, "static": static, "out": out, "generation": gen, "launches": launches}
            self._plans[key] = plan
        else:
            for dst, src in zip(plan["static"].tensors(), batch.tensors()):
                dst.copy_(src, non_blocking=True)
        plan["graph"].replay()
        self.graph_launches_per_replay = plan["launches"]
        return plan["out"][0].clone(), plan["out"][1].clone()

    def _forward_sparse_eager(self, batch: "synth.SparseBatch"):
        eng = self.engine
        prec = _PRECISIONS[self.precision]
        _lib.require_cuda(batch.event_values, "event hit values")
        eng.ensure_packed(prec)
        if self.overlap_cnns:
            # the two CNNs are independent: the event CNN runs on a side stream (own workspace) so that its launch
            # prologues and wave tails overlap the prong CNN
            dev = batch.event_values.device
            main = torch.cuda.current_stream(dev)
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
            side.wait_stream(main)
            # disjoint SM budgets in proportion to the image counts, so that the two walks run side by side
            L = _lib.load()
            n_ev, n_pr = batch.num_events, batch.num_prongs
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            ev_sms = 0
            if self.partition_sms and n_ev > 0 and n_pr > 0:
                ev_sms = min(sms - 1, max(1, round(sms * n_ev / (n_ev + n_pr))))
            try:
                L.tcvn_set_sm_limit(ev_sms)
                with torch.cuda.stream(side):
                    ev = eng.cnn_sparse("event", batch.event_values, batch.event_coords, n_ev, prec, ws_kind="cnn_event")
                L.tcvn_set_sm_limit(sms - ev_sms if ev_sms else 0)
                pr = eng.cnn_sparse("prong", batch.prong_values, batch.prong_coords, n_pr, prec)
            finally:
                L.tcvn_set_sm_limit(0)
            main.wait_stream(side)
            if not torch.cuda.is_current_stream_capturing():
                ev.record_stream(main)
        else:
            ev = eng.cnn_sparse("event", batch.event_values, batch.event_coords, batch.num_events, prec)
            pr = eng.cnn_sparse("prong", batch.prong_values, batch.prong_coords, batch.num_prongs, prec)
        _, _, ev_logits, pr_logits = eng.seq(_lib.SEQ_TOKENS | _lib.SEQ_ENCODER | _lib.SEQ_HEADS, ev, pr,
                                             batch.event_mask, batch.prong_mask)
        return ev_logits, pr_logits
