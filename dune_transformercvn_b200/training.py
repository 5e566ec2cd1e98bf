"""Training side of the drop-in network: flat parameter / gradient arenas, the autograd bridge to the
hand-written forward+backward in libtcvn, a fused AdamW, and the data-parallel gradient exchange.

What it replaces in the reference (paths relative to /root/reference/transformercvn):
  * autograd over ``NeutrinoDenseNetwork.forward`` in ``training_step``
    (network/trainers/neutrino_full_base_trainer.py:162-192) -> ``tcvn_cnn_train_*`` / ``tcvn_seq_train_*``;
  * ``torch.optim.AdamW`` built by ``configure_optimizers`` (network/trainers/neutrino_base.py:109-130) and
    Lightning's global-norm clip (train.py:140) -> :class:`TcvnAdamW` (``tcvn_sumsq`` + ``tcvn_adamw_step``);
  * DDP's gradient averaging (train.py:123-127) -> :class:`GradientExchange`: NCCL all-reduce over slices of the
    flat gradient arena, issued from inside backward so the exchange of one sub-network overlaps the backward
    of the next.
The loss itself (focal loss, 0.9/0.1 mix) stays in the caller, as in the reference.

Every float tensor of the network (parameters and BatchNorm running buffers, reference state_dict order) lives
in ONE fp32 arena; the ``nn.Parameter``s are views into it, their ``.grad``s views into a second arena of the
same layout.  There is no CPU path: everything here raises on CPU tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import lib as _lib
from .params import embedding_dims

BN_MOMENTUM = 0.1  # torch.nn.BatchNorm default, used by every BatchNorm of the reference

_PREFIX = {
    "prong": "prong_embedding.prong_pixel_embedding.",
    "event": "prong_embedding.event_pixel_embedding.",
    "position": "prong_embedding.event_position_embedding",
    "combined": "prong_embedding.combined_embedding.",
    "encoder": "encoder.",
    "event_decoder": "event_decoder.",
    "prong_decoder": "prong_decoder.",
}
# parameters the reference never reaches in forward (disable_smart_features=True; neutrino_full_base_network.py:107
# reads the EVENT position vector for prongs too): they keep grad None, exactly like under torch autograd
_NO_GRAD_PREFIXES = ("prong_embedding.feature_embedding.", "prong_embedding.prong_position_embedding")


def _tensors_of(net) -> Dict[str, torch.Tensor]:
    out = dict(net.named_parameters())
    out.update(dict(net.named_buffers()))
    return out


class FlatArena:
    """One fp32 buffer holding every float tensor of the network + a gradient buffer of the same layout."""

    def __init__(self, net):
        self.net = (net,)
        self.specs = [s for s in net.specs if s.in_arena]
        self.offset: Dict[str, int] = {}
        cur = 0
        for s in self.specs:
            self.offset[s.name] = cur
            cur += s.numel
        self.total = cur
        self.flat: Optional[torch.Tensor] = None
        self.gflat: Optional[torch.Tensor] = None
        self.gviews: Dict[str, torch.Tensor] = {}
        self.select: Optional[torch.Tensor] = None   # per-element optimizer group id (0 = not optimised)
        self.select_key = None
        self._probe = None
        self.seg: Dict[str, Tuple[int, int]] = {}
        for tag, prefix in _PREFIX.items():
            names = [s for s in self.specs if s.name.startswith(prefix)]
            lo = self.offset[names[0].name]
            self.seg[tag] = (lo, self.offset[names[-1].name] + names[-1].numel)

    # ---- layout ---------------------------------------------------------------------------------
    def has_grad(self, name: str) -> bool:
        return not name.startswith(_NO_GRAD_PREFIXES)

    def bound(self) -> bool:
        """Are the module's tensors still the views into the arena?  `.to()` / `.cuda()` move every tensor, so a sample
        (first, last and a few in between) decides; a full walk over 778 tensors per step would cost ~1 ms of host time."""
        if self.flat is None or self._probe is None:
            return False
        base = self.flat.data_ptr()
        return all(t.data_ptr() == base + off and t.dtype == torch.float32 for t, off in self._probe)

    def bind(self) -> None:
        """Move every float tensor into the arena (values preserved) and make the module's tensors views."""
        net = self.net[0]
        tensors = _tensors_of(net)
        dev = tensors[self.specs[0].name].device
        if dev.type != "cuda":
            raise _lib.TcvnError("training needs the module on a CUDA device (there is no CPU implementation)")
        flat = torch.empty(self.total, dtype=torch.float32, device=dev)
        for s in self.specs:
            o = self.offset[s.name]
            flat[o:o + s.numel].copy_(tensors[s.name].detach().reshape(-1).float())
        for s in self.specs:
            o = self.offset[s.name]
            tensors[s.name].data = flat[o:o + s.numel].view(s.shape)
        self.flat = flat
        step = max(1, len(self.specs) // 8)
        self._probe = [(tensors[sp.name], 4 * self.offset[sp.name]) for sp in (self.specs[::step] + self.specs[-1:])]
        self.gflat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.gviews = {s.name: self.gflat[self.offset[s.name]:self.offset[s.name] + s.numel].view(s.shape)
                       for s in self.specs if s.is_param and self.has_grad(s.name)}
        self.select = None

    def ensure(self) -> None:
        if not self.bound():
            self.bind()

    def attach_grads(self) -> None:
        """Make every parameter's .grad the view into the gradient arena; zero what was reset to None."""
        params = dict(self.net[0].named_parameters())
        missing = [n for n, v in self.gviews.items() if params[n].grad is not v]
        if not missing:
            return
        if len(missing) == len(self.gviews):
            self.gflat.zero_()
        else:
            for n in missing:
                self.gviews[n].zero_()
        for n in missing:
            params[n].grad = self.gviews[n]

    def ptr(self, tag: str, grad: bool = False) -> C.c_void_p:
        buf = self.gflat if grad else self.flat
        return C.c_void_p(buf.data_ptr() + 4 * self.seg[tag][0])

    def grad_slice(self, tag: str) -> torch.Tensor:
        lo, hi = self.seg[tag]
        return self.gflat[lo:hi]


class GradientExchange:
    """Data-parallel gradient averaging over NCCL (the reference's plain DDP, train.py:123-127: rank-local
    BatchNorm statistics, one averaged gradient).  ``reduce(slice)`` is called from inside backward as soon as
    a sub-network's gradients are final; ``wait()`` joins the collectives before the optimizer reads them."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pending: List = []
        self.bytes = 0
        self._buf_index: Optional[torch.Tensor] = None

    def broadcast_buffers(self, arena: "FlatArena") -> None:
        """torch DDP's ``broadcast_buffers=True`` default (what ``DDPStrategy()`` gives the reference, train.py:125): before
        every forward rank 0's BatchNorm running statistics replace the other ranks'.  The float buffers live scattered in
        the parameter arena, so they are gathered into one staging vector, broadcast (0.22 MB) and scattered back."""
        if self.world == 1:
            return
        if self._buf_index is None or self._buf_index.device != arena.flat.device:
            idx = [torch.arange(arena.offset[s.name], arena.offset[s.name] + s.numel) for s in arena.specs if not s.is_param]
            self._buf_index = torch.cat(idx).to(arena.flat.device)
        staging = arena.flat.index_select(0, self._buf_index)
        self.dist.broadcast(staging, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                            group=self.group)
        arena.flat.index_copy_(0, self._buf_index, staging)

    def reduce(self, grad_slice: torch.Tensor) -> None:
        if self.world == 1:
            return
        op = self.dist.ReduceOp.AVG if grad_slice.is_cuda else self.dist.ReduceOp.SUM
        work = self.dist.all_reduce(grad_slice, op=op, group=self.group, async_op=True)
        self.pending.append((work, grad_slice, op))
        self.bytes += grad_slice.numel() * grad_slice.element_size()

    def wait(self) -> None:
        for work, sl, op in self.pending:
            work.wait()
            if op == self.dist.ReduceOp.SUM:   # gloo (CPU tests) has no AVG
                sl.div_(self.world)
        self.pending.clear()


class TrainEngine:
    """Train-mode forward and backward of the whole network through the C ABI."""

    def __init__(self, net):
        self.net = (net,)
        self.arena = FlatArena(net)
        self.ws: Dict[str, torch.Tensor] = {}
        self.step_index = 0
        # data parallelism (the reference's DDPStrategy, train.py:123-127).  "auto": as soon as torch.distributed is
        # initialised with more than one rank, backward averages the gradient arena over the ranks (and forward takes
        # rank 0's BatchNorm running buffers) - the hand-written backward fills .grad directly, so torch DDP's autograd
        # hooks never fire for these parameters and wrapping the module in DDP alone would NOT average anything.
        # None disables it; an explicit GradientExchange pins the process group.
        self.exchange = "auto"
        self.broadcast_buffers = True
        # CUDA-graph mode (GraphedTrainStep): the per-step part of every random seed lives in this device int64 counter
        # instead of the kernel arguments, so a replayed graph draws fresh dropout masks / pixel noise
        self.seed_counter: Optional[torch.Tensor] = None
        self.saved = None
        self._nbt: Optional[List[torch.Tensor]] = None
        # the two pixel-map CNNs are independent until the token assembly: the (small, launch-bound) event CNN runs on
        # a side stream under the (large) prong CNN, forward and backward
        self.overlap_cnns = True
        self._side: Optional[torch.cuda.Stream] = None

    def workspace(self, kind: str, nbytes: int, dev) -> torch.Tensor:
        buf = self.ws.get(kind)
        if buf is None or buf.numel() < nbytes or buf.device != dev:
            self.ws.pop(kind, None)
            buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self.ws[kind] = buf
        return buf

    def _cnn_desc(self, tag: str) -> _lib.CnnDesc:
        net = self.net[0]
        pix, feat, _ = embedding_dims(net.options)
        return net.engine.cnn_desc(pix if tag == "prong" else pix + feat)

    def _issue(self, dev, fn, what: str, calls) -> None:
        """Run the C walks of the two pixel-map CNNs, each ~400-800 kernel launches issued by ONE ctypes call.
        Measured (round 2, scripts/gpu_train_host_bound.py): issuing the two walks from two host threads does NOT help -
        launches into one CUDA context serialise on the driver's lock (4 events: 10.8 ms per step either way, 6.3 us of
        host time per launch), so they are issued one after the other."""
        for tag, _, args, _ in calls:
            _lib.check(fn(*args), f"{what}({tag})")

    def seed_args(self) -> Tuple[int, Optional[int]]:
        """(seed value, device pointer of the seed offset or None) of the NEXT training step."""
        base = torch.initial_seed() * 1000003
        if self.seed_counter is not None:
            return base & 0xFFFFFFFFFFFF, self.seed_counter.data_ptr()
        return (base + self.step_index + 1) & 0xFFFFFFFFFFFF, None

    def _exchange(self) -> Optional[GradientExchange]:
        if self.exchange == "auto":
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                self.exchange = GradientExchange()
            else:
                return None
        return self.exchange

    def forward(self, event_pixels, event_mask, prong_pixels, prong_mask):
        for t, what in ((event_pixels, "event_pixels"), (prong_pixels, "prong_pixels"), (prong_mask, "prong_mask")):
            _lib.require_cuda(t, what)
        with torch.cuda.device(event_pixels.device):   # the C ABI launches on the current device
            seed, off = self.seed_args()
            L = _lib.load()
            L.tcvn_set_seed_offset(off)
            try:
                return self._forward(event_pixels, event_mask, prong_pixels, prong_mask, seed, off)
            finally:
                L.tcvn_set_seed_offset(None)

    def _forward(self, event_pixels, event_mask, prong_pixels, prong_mask, seed, seed_off):
        net = self.net[0]
        L = _lib.load()
        dev = event_pixels.device
        self.arena.ensure()
        ex = self._exchange()
        if ex is not None and self.broadcast_buffers:
            ex.broadcast_buffers(self.arena)
        net.engine.key = None  # running buffers (and soon the weights) change under the eval-path cache
        st = _lib.stream_ptr(dev)
        p_drop = float(net.options.dropout)
        prec = _lib.TCVN_BF16 if net.precision == "bf16" else _lib.TCVN_FP32
        self.step_index += 1
        b, l = prong_mask.shape
        t = prong_pixels.shape[0]
        pix, feat, _ = embedding_dims(net.options)
        ev_px = event_pixels.contiguous().float()
        pr_px = prong_pixels.contiguous().float()
        f32 = dict(dtype=torch.float32, device=dev)
        emb = {"event": torch.empty((b, pix + feat), **f32), "prong": torch.empty((t, pix), **f32)}
        main = torch.cuda.current_stream(dev)
        side = None
        if self.overlap_cnns:
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
        calls = []
        for site, (tag, px, n) in enumerate((("event", ev_px, b), ("prong", pr_px, t)), start=1):
            d = self._cnn_desc(tag)
            if tuple(px.shape[1:]) != (d.in_channels, d.height, d.width):
                raise _lib.TcvnError(f"{tag} pixels have shape {tuple(px.shape)}, expected (N,{d.in_channels},{d.height},{d.width})")
            nbytes = L.tcvn_cnn_train_workspace_bytes(C.byref(d), prec, n)
            if nbytes == 0:
                raise _lib.TcvnError("tcvn_cnn_train_workspace_bytes: " + L.tcvn_last_error().decode())
            ws = self.workspace("cnn_" + tag, nbytes, dev)
            use = side if (side is not None and tag == "event") else main
            calls.append((tag, use, (C.byref(d), prec, self.arena.ptr(tag), _lib.ptr(px), n, p_drop, BN_MOMENTUM, seed, site,
                                     _lib.ptr(emb[tag]), _lib.ptr(ws), ws.numel(), C.c_void_p(use.cuda_stream)), d))
        if side is not None:
            side.wait_stream(main)            # inputs, parameters and the workspace zero-fill were issued on main
        self._issue(dev, L.tcvn_cnn_train_forward, "tcvn_cnn_train_forward", calls)
        if side is not None:
            main.wait_stream(side)
        sd = net.engine.seq_desc()
        pm = prong_mask.contiguous().to(torch.uint8)
        em = None if event_mask is None else event_mask.contiguous().to(torch.uint8)
        nbytes = L.tcvn_seq_train_workspace_bytes(C.byref(sd), b, l, t)
        if nbytes == 0:
            raise _lib.TcvnError("tcvn_seq_train_workspace_bytes: " + L.tcvn_last_error().decode())
        ws = self.workspace("seq", nbytes, dev)
        ev_logits = torch.empty((b, sd.num_event_classes), **f32)
        pr_logits = torch.empty((l * b, sd.num_prong_classes), **f32)
        a = self.arena
        _lib.check(L.tcvn_seq_train_forward(C.byref(sd), a.ptr("position"), a.ptr("combined"), a.ptr("encoder"),
                                            a.ptr("event_decoder"), a.ptr("prong_decoder"), _lib.ptr(emb["event"]),
                                            _lib.ptr(emb["prong"]), _lib.ptr(em), _lib.ptr(pm), b, l, t, p_drop, BN_MOMENTUM,
                                            seed, _lib.ptr(ev_logits), _lib.ptr(pr_logits), _lib.ptr(ws), ws.numel(), st),
                   "tcvn_seq_train_forward")
        if self._nbt is None:
            self._nbt = [buf for name, buf in net.named_buffers()
                         if name.endswith("num_batches_tracked") and not name.startswith(_NO_GRAD_PREFIXES)]
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)
        self.saved = dict(ev_px=ev_px, pr_px=pr_px, pm=pm, b=b, l=l, t=t, seed=seed, seed_off=seed_off, p_drop=p_drop, dev=dev,
                          prec=prec)
        return ev_logits, pr_logits

    def backward(self, d_ev_logits: Optional[torch.Tensor], d_pr_logits: Optional[torch.Tensor]) -> None:
        if self.saved is None:
            raise _lib.TcvnError("backward without a train-mode forward (or called twice)")
        with torch.cuda.device(self.saved["dev"]):
            L = _lib.load()
            L.tcvn_set_seed_offset(self.saved["seed_off"])   # backward runs on autograd's thread: the offset is per thread
            try:
                self._backward(d_ev_logits, d_pr_logits)
            finally:
                L.tcvn_set_seed_offset(None)

    def _backward(self, d_ev_logits: Optional[torch.Tensor], d_pr_logits: Optional[torch.Tensor]) -> None:
        s, self.saved = self.saved, None
        net = self.net[0]
        L = _lib.load()
        dev = s["dev"]
        st = _lib.stream_ptr(dev)
        a = self.arena
        a.attach_grads()
        sd = net.engine.seq_desc()
        b, l, t = s["b"], s["l"], s["t"]
        f32 = dict(dtype=torch.float32, device=dev)
        pix, feat, _ = embedding_dims(net.options)
        d_ev = torch.zeros((b, sd.num_event_classes), **f32) if d_ev_logits is None else d_ev_logits.contiguous().float().clone()
        d_pr = torch.zeros((l * b, sd.num_prong_classes), **f32) if d_pr_logits is None else d_pr_logits.contiguous().float().clone()
        d_emb = {"event": torch.empty((b, pix + feat), **f32), "prong": torch.empty((t, pix), **f32)}
        ws = self.ws["seq"]
        _lib.check(L.tcvn_seq_train_backward(C.byref(sd), a.ptr("position"), a.ptr("combined"), a.ptr("encoder"),
                                             a.ptr("event_decoder"), a.ptr("prong_decoder"), a.ptr("position", True),
                                             a.ptr("combined", True), a.ptr("encoder", True), a.ptr("event_decoder", True),
                                             a.ptr("prong_decoder", True), _lib.ptr(s["pm"]), b, l, t, s["p_drop"], s["seed"],
                                             _lib.ptr(d_ev), _lib.ptr(d_pr), _lib.ptr(d_emb["event"]), _lib.ptr(d_emb["prong"]),
                                             _lib.ptr(ws), ws.numel(), st), "tcvn_seq_train_backward")
        ex = self._exchange()
        if ex is not None:
            if ex.pending:
                raise _lib.TcvnError("backward while a gradient exchange is in flight")
            lo = a.seg["combined"][0]
            ex.reduce(a.gflat[lo:])                       # combined embedding, encoder, both heads
            ex.reduce(a.grad_slice("position"))
        # event CNN on the side stream (its small exchange then overlaps the long prong-CNN backward as well)
        main = torch.cuda.current_stream(dev)
        side = self._side if self.overlap_cnns else None
        calls = []
        for site, tag, px, n in ((1, "event", s["ev_px"], b), (2, "prong", s["pr_px"], t)):
            d = self._cnn_desc(tag)
            ws = self.ws["cnn_" + tag]
            use = side if (side is not None and tag == "event") else main
            if use is side:
                d_emb[tag].record_stream(side)
            calls.append((tag, use, (C.byref(d), s["prec"], a.ptr(tag), a.ptr(tag, True), _lib.ptr(px), n, s["p_drop"], s["seed"],
                                     site, _lib.ptr(d_emb[tag]), _lib.ptr(ws), ws.numel(), C.c_void_p(use.cuda_stream)), d))
        if side is not None:
            side.wait_stream(main)
        self._issue(dev, L.tcvn_cnn_train_backward, "tcvn_cnn_train_backward", calls)
        if ex is not None:
            for tag, use, _, _ in calls:
                with torch.cuda.stream(use):
                    ex.reduce(a.grad_slice(tag))
        if side is not None:
            main.wait_stream(side)
        if ex is not None:
            # joined here, not in the optimizer: whoever reads .grad after backward() (Lightning's clip_grad_norm_, a stock
            # torch optimizer, a test) sees the averaged gradient.  The collectives were issued as each sub-network's
            # gradients became final, so they have been running under the rest of the backward pass.
            ex.wait()


class _TrainFn(torch.autograd.Function):
    """Autograd node of the whole network: the hand-written backward writes parameter gradients straight into the
    gradient arena (the parameters' .grad views), so no per-parameter tensors cross the autograd boundary."""

    @staticmethod
    def forward(ctx, anchor, engine, event_pixels, event_mask, prong_pixels, prong_mask):
        ctx.engine = engine
        ev, pr = engine.forward(event_pixels, event_mask, prong_pixels, prong_mask)
        return ev, pr

    @staticmethod
    def backward(ctx, d_ev, d_pr):
        ctx.engine.backward(d_ev, d_pr)
        return None, None, None, None, None, None


def train_forward(engine: TrainEngine, event_pixels, event_mask, prong_pixels, prong_mask):
    """(B,E) event logits and (B,L,C) prong logits, differentiable w.r.t. the network's parameters."""
    dev = event_pixels.device
    anchor = torch.zeros((), device=dev, requires_grad=True)
    ev, pr = _TrainFn.apply(anchor, engine, event_pixels, event_mask, prong_pixels, prong_mask)
    b, l = prong_mask.shape
    return ev, pr.view(l, b, -1).transpose(0, 1)


class TcvnAdamW(torch.optim.Optimizer):
    """Fused AdamW over the flat arenas of a dune_transformercvn_b200 network (torch.optim.AdamW semantics).

    Selected the way the reference selects its optimizer — ``getattr(torch.optim, options.optimizer)``
    (network/trainers/neutrino_base.py:109) — after ``register()`` publishes it as ``torch.optim.TcvnAdamW``;
    instantiated as ``cls(param_groups, lr=...)`` with the reference's decay / no-decay groups (:116-130).
    ``max_grad_norm`` fuses Lightning's ``gradient_clip_val`` (train.py:140) into the step (no host sync)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm: float = 0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = float(max_grad_norm)
        self._arena: Optional[FlatArena] = None
        self._m = self._v = self._ws = None
        self.device_state: Optional[Tuple[torch.Tensor, torch.Tensor]] = None   # (step int64[1], lr float32[groups]): graph mode
        self._steps = [0 for _ in self.param_groups]

    def attach(self, net) -> "TcvnAdamW":
        """Bind to the network whose parameters this optimizer was given."""
        eng = net.train_engine
        eng.arena.ensure()
        self._arena = eng.arena
        self._engine = eng
        dev = self._arena.flat.device
        self._m = torch.zeros_like(self._arena.flat)
        self._v = torch.zeros_like(self._arena.flat)
        self._ws = torch.zeros(_lib.load().tcvn_adamw_workspace_bytes(), dtype=torch.uint8, device=dev)
        return self

    @property
    def grad_norm_sq(self) -> Optional[torch.Tensor]:
        """Squared global gradient norm of the last step (device scalar, only when ``max_grad_norm`` > 0)."""
        return None if self._ws is None else self._ws.view(torch.float64)[-1]

    def _find_arena(self) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                eng = getattr(p, "_tcvn_engine", None)
                if eng is not None:
                    self.attach(eng[0].net[0])
                    return
        raise _lib.TcvnError("TcvnAdamW drives only the parameters of a dune_transformercvn_b200 network "
                             "(no generic / CPU fallback): none of the given parameters belongs to one")

    def _selector(self) -> torch.Tensor:
        a = self._arena
        # cheap per-step key: which parameters currently carry a gradient (group membership is fixed after construction)
        key = (a.flat.data_ptr(), tuple(p.grad is None for g in self.param_groups for p in g["params"]))
        if a.select is not None and a.select_key == key:
            return a.select
        by_ptr = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                by_ptr[p.data_ptr()] = (gi + 1, p)
        sel = torch.zeros(a.total, dtype=torch.uint8)
        base = a.flat.data_ptr()
        for s in a.specs:
            if not s.is_param:
                continue
            hit = by_ptr.get(base + 4 * a.offset[s.name])
            if hit is None or hit[1].grad is None:
                continue
            sel[a.offset[s.name]:a.offset[s.name] + s.numel] = hit[0]
        a.select = sel.to(a.flat.device)
        a.select_key = key
        return a.select

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._arena is None or not self._arena.bound():
            self._find_arena()
        a = self._arena
        L = _lib.load()
        st = _lib.stream_ptr(a.flat.device)
        sel = self._selector()
        ng = len(self.param_groups)
        dbl, i64 = C.c_double * ng, C.c_int64 * ng
        for gi in range(ng):
            self._steps[gi] += 1
        g = self.param_groups
        # every group in one launch (+ one for the gradient norm when clipping)
        with torch.cuda.device(a.flat.device):
            _lib.check(L.tcvn_adamw_fused(
                _lib.ptr(a.flat), _lib.ptr(a.gflat), _lib.ptr(self._m), _lib.ptr(self._v), a.total, _lib.ptr(sel), ng,
                dbl(*[float(x["lr"]) for x in g]), dbl(*[float(x["betas"][0]) for x in g]),
                dbl(*[float(x["betas"][1]) for x in g]), dbl(*[float(x["eps"]) for x in g]),
                dbl(*[float(x["weight_decay"]) for x in g]), i64(*self._steps), self.max_grad_norm, 1.0,
                _lib.ptr(self._ws), self._ws.numel(),
                None if self.device_state is None else _lib.ptr(self.device_state[0]),
                None if self.device_state is None else _lib.ptr(self.device_state[1]), st), "tcvn_adamw_fused")
        self._engine.net[0].engine.key = None
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        """One memset of the gradient arena; the .grad views stay attached (set_to_none would only force the next
        backward to re-attach them)."""
        if self._arena is not None and self._arena.gflat is not None:
            self._arena.gflat.zero_()
        else:
            super().zero_grad(set_to_none)

    # checkpointing: moments are exposed per parameter like torch.optim.AdamW's state
    def state_dict(self):
        if self._arena is not None:
            a = self._arena
            base = a.flat.data_ptr()
            off = {base + 4 * a.offset[s.name]: (a.offset[s.name], s) for s in a.specs if s.is_param}
            for gi, g in enumerate(self.param_groups):
                for p in g["params"]:
                    hit = off.get(p.data_ptr())
                    if hit is None:
                        continue
                    o, s = hit
                    self.state[p] = {"step": torch.tensor(float(self._steps[gi])),
                                     "exp_avg": self._m[o:o + s.numel].view(s.shape),
                                     "exp_avg_sq": self._v[o:o + s.numel].view(s.shape)}
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        if self._arena is None:
            self._find_arena()
        a = self._arena
        base = a.flat.data_ptr()
        off = {base + 4 * a.offset[s.name]: (a.offset[s.name], s) for s in a.specs if s.is_param}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                hit, stt = off.get(p.data_ptr()), self.state.get(p)
                if hit is None or not stt:
                    continue
                o, s = hit
                self._m[o:o + s.numel].copy_(stt["exp_avg"].reshape(-1))
                self._v[o:o + s.numel].copy_(stt["exp_avg_sq"].reshape(-1))
                self._steps[gi] = int(stt["step"])


class GraphedTrainStep:
    """One whole training step - zero_grad, densify (+ pixel noise), train-mode forward, fused focal loss, hand-written
    backward (+ gradient exchange), clip + AdamW - captured ONCE per batch shape into a CUDA graph and replayed with a
    single launch.  The step is ~1600 kernel launches; at the reference's batch sizes (2-16 events per GPU) issuing them
    costs more host time (6 us each) than the GPU needs to run them.

    What changes from step to step lives in device memory, not in kernel arguments: the step counter (dropout / noise seed
    offset, AdamW bias corrections) and the scheduler's learning rates (copied in before every replay).  Results are the
    same arithmetic as the eager step; the seed sequence differs (base + counter instead of base + host step index).

    A graph is keyed by the batch SHAPE (events, prong slots, images, hit counts).  Training data with a different shape in
    every step (variable prong counts) re-captures every time - use the eager step there (``graph=False`` in bench.py)."""

    def __init__(self, net, optimizer: "TcvnAdamW", options, max_plans: int = 8):
        from . import loss as _loss
        self.net, self.opt, self.options, self._loss = net, optimizer, options, _loss
        self.plans: Dict[Tuple, dict] = {}
        self.max_plans = max_plans
        self.stream: Optional[torch.cuda.Stream] = None
        self.step_dev = self.lr_dev = self.lr_host = None
        self._mirror = 0                 # host copy of step_dev
        self.launches_per_replay = 0

    def _body(self, batch, ev_t, pr_t):
        self.step_dev.add_(1)
        self.opt.zero_grad()
        ev, pr = self.net.forward_sparse(batch)
        loss, _ = self._loss.training_loss(ev, pr, ev_t, pr_t, self.options)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def _setup(self, dev) -> None:
        if self.opt._arena is None or not self.opt._arena.bound():
            self.opt._find_arena()
        ng = len(self.opt.param_groups)
        self.step_dev = torch.tensor([self.opt._steps[0]], dtype=torch.int64, device=dev)
        self._mirror = self.opt._steps[0]
        self.lr_dev = torch.zeros(ng, dtype=torch.float32, device=dev)
        self.lr_host = torch.zeros(ng, dtype=torch.float32).pin_memory()
        self.stream = torch.cuda.Stream(device=dev)

    def __call__(self, batch, ev_targets, pr_targets) -> torch.Tensor:
        dev = batch.event_values.device
        if self.step_dev is None:
            self._setup(dev)
        eng = self.net.train_engine
        key = (batch.num_events, batch.prong_mask.shape[1], batch.num_prongs, batch.event_coords.shape[0], batch.prong_coords.shape[0],
               str(batch.event_values.dtype), tuple(pr_targets.shape))
        for gi, g in enumerate(self.opt.param_groups):
            self.lr_host[gi] = float(g["lr"])
        self.lr_dev.copy_(self.lr_host, non_blocking=True)
        if self._mirror != self.opt._steps[0]:       # eager optimizer steps were taken in between: resynchronise the counter
            self.step_dev.fill_(self.opt._steps[0])
            self._mirror = self.opt._steps[0]
        plan = self.plans.get(key)
        if plan is None:
            plan = self._capture(key, batch, ev_targets, pr_targets)
        else:
            for dst, src in zip(plan["inputs"], list(batch.tensors()) + [ev_targets, pr_targets]):
                dst.copy_(src, non_blocking=True)
        plan["graph"].replay()
        # host mirrors of what the graph advanced on the device
        eng.step_index += 1
        self.opt._steps = [s + 1 for s in self.opt._steps]
        self._mirror += 1
        self.net.engine.key = None
        self.launches_per_replay = plan["launches"]
        return plan["loss"]

    def _capture(self, key, batch, ev_t, pr_t) -> dict:
        from . import synth
        dev = batch.event_values.device
        eng, opt = self.net.train_engine, self.opt
        static = synth.SparseBatch(*[t.clone().contiguous() for t in batch.tensors()], list(batch.prongs_per_event))
        t0, t1 = ev_t.clone(), pr_t.clone()
        cur = torch.cuda.current_stream(dev)
        # device-resident mode for the engine and the optimizer; snapshot everything a step mutates, run ONE real step to
        # allocate the workspaces / build the caches, restore, then record (recording executes nothing)
        eng.arena.ensure()
        eng.seed_counter = self.step_dev
        opt.device_state = (self.step_dev, self.lr_dev)
        try:
            return self._capture_inner(key, static, t0, t1, dev, cur)
        finally:       # eager steps outside this class keep their host-side seeds / step counts
            eng.seed_counter = None
            opt.device_state = None

    def _capture_inner(self, key, static, t0, t1, dev, cur) -> dict:
        eng, opt = self.net.train_engine, self.opt
        nbt = [b for n, b in self.net.named_buffers() if n.endswith("num_batches_tracked")]
        snap = (eng.arena.flat.clone(), opt._m.clone(), opt._v.clone(), self.step_dev.clone(), [b.clone() for b in nbt],
                eng.step_index, list(opt._steps))
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self._body(static, t0, t1)
        cur.wait_stream(self.stream)

        def restore():
            eng.arena.flat.copy_(snap[0]); opt._m.copy_(snap[1]); opt._v.copy_(snap[2]); self.step_dev.copy_(snap[3])
            for b, v in zip(nbt, snap[4]):
                b.copy_(v)
            eng.step_index, opt._steps = snap[5], list(snap[6])

        restore()
        torch.cuda.synchronize(dev)
        l0 = _lib.load().tcvn_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream, capture_error_mode="relaxed"):
            loss = self._body(static, t0, t1)
        launches = _lib.load().tcvn_launch_count() - l0
        eng.step_index, opt._steps = snap[5], list(snap[6])    # recording advanced the host mirrors only
        if len(self.plans) >= self.max_plans:
            self.plans.pop(next(iter(self.plans)))
        plan = {"graph": g, "inputs": list(static.tensors()) + [t0, t1], "loss": loss, "launches": launches}
        self.plans[key] = plan
        return plan


def register() -> None:
    """Publish the optimizer under ``torch.optim.TcvnAdamW`` so ``"optimizer": "TcvnAdamW"`` in an options JSON
    selects it through the reference's own lookup (network/trainers/neutrino_base.py:109)."""
    torch.optim.TcvnAdamW = TcvnAdamW


def reference_param_groups(net, weight_decay: float) -> List[dict]:
    """The reference's two groups (network/trainers/neutrino_base.py:116-128): names containing neither "bias" nor
    "LayerNorm.weight" are decayed."""
    no_decay = ("bias", "LayerNorm.weight")
    named = list(net.named_parameters())
    return [
        {"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": weight_decay},
        {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
    ]
