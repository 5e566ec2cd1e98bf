#!/usr/bin/env python
"""Headline benchmark of the TransformerCVN hot path (BASELINE.json: events/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--events B] [--precision bf16|fp32]

A "step" = one pass of the hot path over one synthetic batch: COO hits -> densify (/255 fused) ->
DenseNet(event maps) + DenseNet(prong maps) -> tokens -> 6-layer encoder -> event/prong logits.
N=1 workload = BASELINE.json configs[1]: eval-mode inference over 256 events (P ~ U[1,10] prongs each,
3x400x280 maps at 1 % / 0.2 % occupancy).  N>1: every rank runs its own 256-event shard
(events are independent in inference: no data-path collective, weak scaling).

  value  events/s with the COO hit lists already resident in HBM (CUDA events, max over ranks)
  e2e    the same step driven from pinned HOST buffers: H2D of the hit lists and D2H of the logits inside
         the timed region
  roofline      dominant kernel (the 3x3 bottleneck convolution of dense block 1) timed alone with CUDA events
  cpu_baseline  the oracle port of the reference's PyTorch CPU path on a bounded sample of the same workload
--impl reference times that CPU path alone (rank 0 only), with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "events/sec (event + prong 3x400x280 maps), inference"
UNIT = "events/s"
H, W = 400, 280
GFLOP_PER_IMAGE_FWD = 4.971  # SURVEY.md §8(d), forward hooks on the reference DenseNet


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_inputs(events: int, seed: int):
    from dune_transformercvn_b200 import synth
    return synth.make_batch(events, seed=seed, max_prongs=10)


def oracle_state_and_opts():
    from dune_transformercvn_b200 import synth
    from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
    from dune_transformercvn_b200.params import network_specs
    opts = PathOptions.tutorial()
    specs = network_specs(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    return synth.init_state(specs, seed=0, perturb=False), opts


def time_cpu_reference(sample_events: int, steps: int, warmup: int, seed: int):
    """The reference's CPU path (oracle port: the reference tree itself is absent on the GPU box)."""
    from oracle import restate
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state, opts = oracle_state_and_opts()
    batch = make_inputs(sample_events, seed)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            restate.sparse_forward(state, opts, batch)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    mean = sum(times) / len(times)
    sample = (f"{sample_events} events / {sample_events + batch.num_prongs} images per step, eval forward incl. densify, "
              f"fp32 torch CPU, {steps} timed steps")
    return sample_events / mean, mean * 1e3, cores, sample


def workload_config(args, images: int, world: int) -> dict:
    """The `config` object of the JSON line: identical in both arms (the reference arm times a bounded sample of it)."""
    return {"workload": f"BASELINE configs[1]: eval inference, {args.events} events/GPU "
                        f"({images} images of 3x400x280 on rank 0), P~U[1,10], occupancy 1%/0.2%",
            "precision": args.precision, "parallelism": f"event-sharded x{world}, no collective",
            "ingest": "densify kernel -> dense maps" if args.materialize else "hit lists consumed by the stem (no dense map)",
            "l2": "activations touched per step (%.1f GB at ~16 MB/image) exceed the 126 MB L2" % (images * 16e6 / 1e9)}


REFERENCE_SAMPLE_EVENTS = 8   # events of the workload one CPU step covers (~1.5 s on 16 cores): K + W steps stay within minutes


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port: the reference tree is absent on the GPU box), all
    host threads, EXACTLY --steps timed steps after --warmup untimed ones; every step is a bounded sample
    (REFERENCE_SAMPLE_EVENTS events, drawn with the workload's own generator and seed) of the configured workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    value, ms, cores, sample = time_cpu_reference(REFERENCE_SAMPLE_EVENTS, steps, warmup, 1234)
    full = make_inputs(args.events, 1234)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": workload_config(args, full.num_events + full.num_prongs, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def time_cpu_train(sample_events: int = 8, steps: int = 3, warmup: int = 1, seed: int = 1234):
    """BASELINE configs[0] / SURVEY 8(d): the reference's CPU training step (oracle port, fp32 torch autograd) on the
    tutorial config - 8 events x <= 10 prongs, train mode, dropout 0.1, forward + focal loss + backward, all host threads,
    1 warm-up + 3 timed iterations, median."""
    from oracle import restate
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state, opts = oracle_state_and_opts()
    batch = make_inputs(sample_events, seed)
    g = torch.Generator().manual_seed(seed)
    ev_t = torch.randint(0, 4, (sample_events,), generator=g)
    pr_t = torch.randint(0, 8, tuple(batch.prong_mask.shape), generator=g)
    pr_t[~batch.prong_mask] = -1
    st = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in state.items()}
    times = []
    torch.manual_seed(seed)
    for i in range(warmup + steps):
        for v in st.values():
            if v.is_floating_point():
                v.grad = None
        t0 = time.perf_counter()
        ev, pr = restate.sparse_forward(st, opts, batch, train=True, p_drop=float(opts.dropout))
        restate.training_loss(ev, pr, ev_t, pr_t, opts).backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": sample_events / med, "unit": "events/s", "cores": cores, "kind": "port", "ms_per_step": med * 1e3,
            "sample": f"BASELINE configs[0]: {sample_events} events / {sample_events + batch.num_prongs} images per step, train-mode "
                      f"forward + focal loss + backward, dropout {opts.dropout}, fp32 torch CPU autograd, {warmup} warm-up + "
                      f"{steps} timed steps, median"}


def time_torch_eager_b200(dev, infer_events: int, train_events: int):
    """the reference arithmetic run eagerly by PyTorch ON THE SAME B200 (ATen / cuDNN kernels; the
    oracle port, since the reference tree is absent on the box) - the baseline the hand-written kernels have to beat.
    fp32 and bf16 autocast; inference over `infer_events` events (eval), training forward + loss + backward (train-mode
    BatchNorm, dropout = identity) over `train_events` events.  Bounded: a few seconds."""
    from oracle import restate
    restate.FUSED_ATEN = True     # BatchNorm / PReLU through the fused ATen ops the reference's modules call
    state, opts = oracle_state_and_opts()
    dstate = {k: v.to(dev) for k, v in state.items()}
    out = {}

    def timed_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    b_inf = make_inputs(infer_events, 1234).to(dev)
    b_tr = make_inputs(train_events, 4321).to(dev)
    g = torch.Generator().manual_seed(99)
    ev_t = torch.randint(0, 4, (train_events,), generator=g).to(dev)
    pr_t = torch.randint(0, 8, tuple(b_tr.prong_mask.shape), generator=g)
    pr_t[~b_tr.prong_mask.cpu()] = -1
    pr_t = pr_t.to(dev)
    tstate = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in dstate.items()}
    for tag, ctx in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        def infer():
            with torch.no_grad(), torch.autocast("cuda", dtype=ctx, enabled=ctx is not None):
                restate.sparse_forward(dstate, opts, b_inf)

        def train():
            for v in tstate.values():
                if v.is_floating_point():
                    v.grad = None
            with torch.autocast("cuda", dtype=ctx, enabled=ctx is not None):
                ev, pr = restate.sparse_forward(tstate, opts, b_tr, train=True)
            restate.training_loss(ev.float(), pr.float(), ev_t, pr_t, opts).backward()

        ms_i = timed_ms(infer, 3)
        ms_t = timed_ms(train, 3)
        out[tag] = {"inference_events_per_s": infer_events / (ms_i / 1e3), "inference_sample_events": infer_events,
                    "train_events_per_s": train_events / (ms_t / 1e3), "train_sample_events": train_events}
    out["what"] = ("oracle port of the reference modules (F.conv2d / F.batch_norm / F.prelu / torch.cat, i.e. the ATen + cuDNN kernels "
                   "the reference's nn modules call) run eagerly on this GPU; train = forward + focal loss + autograd backward, no "
                   "optimizer step")
    restate.FUSED_ATEN = False
    torch.cuda.empty_cache()
    return out


def single_event_latency(dev, precision, prongs=6, reps=50):
    """SURVEY 8f rank 3 (the LArSoft export surface, CreateCompiled.ipynb): one event = (1 + prongs) dense uint8 maps ->
    probabilities + 128-d embeddings.  Latency of the captured CUDA-graph plan against the same call sequence launched
    kernel by kernel, CUDA events, pixels resident."""
    from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
    from dune_transformercvn_b200.export import EventClassifier
    from dune_transformercvn_b200.ingest import densify
    from dune_transformercvn_b200.network import NeutrinoDenseNetwork
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=precision).to(dev).eval()
    from dune_transformercvn_b200 import synth
    b = synth.make_batch(1, seed=77, prongs_per_event=[prongs]).to(dev)
    n_pr = int(b.num_prongs)
    ev = densify(b.event_values, b.event_coords, (H, W), 1, 1.0)
    pr = densify(b.prong_values, b.prong_coords, (H, W), n_pr, 1.0)
    px = torch.cat((ev, pr)).to(torch.uint8)
    out = {"images": 1 + n_pr, "precision": precision}
    for tag, graph in (("graph_us", True), ("launch_by_launch_us", False)):
        clf = EventClassifier(net, "combined", use_graph=graph)
        for _ in range(3):
            clf(px)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            clf(px)
        e1.record()
        torch.cuda.synchronize()
        out[tag] = e0.elapsed_time(e1) / reps * 1e3
        clf.invalidate()
    del net
    torch.cuda.empty_cache()
    return out


def run_train_leg(args, dev, rank, world, timed, train_events, with_cpu, fixed_prongs=None, label="BASELINE configs[2]"):
    """BASELINE configs[2]: DenseNet TransformerCVN training, event-sharded data parallel, NCCL gradient all-reduce.
    One step = densify (+ pixel noise) -> train-mode forward -> fused focal loss (tcvn_loss_forward) -> hand-written backward
    (gradient exchange issued from inside it, joined at its end) -> fused clip + AdamW.  Every rank holds `--train-events`
    events (weak scaling) with the SAME total prong count (synth.balanced_prongs: the slowest rank sets the step, so
    unequal work would measure the draw, not the path); `fixed_prongs` = every event has that many (configs[4]: 20)."""
    import torch.distributed as dist
    from dune_transformercvn_b200 import lib as tl
    from dune_transformercvn_b200 import loss as tloss
    from dune_transformercvn_b200 import training
    from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
    from dune_transformercvn_b200.network import NeutrinoDenseNetwork
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=args.train_precision).to(dev).train()
    if world > 1:   # same initial weights everywhere (the constructor is seeded, this is DDP's broadcast)
        net.train_engine.arena.ensure()
        dist.broadcast(net.train_engine.arena.flat, src=0)
        # nothing to install: TrainEngine averages the gradient arena over the ranks as soon as torch.distributed is up
    opt = training.TcvnAdamW(training.reference_param_groups(net, opts.l2_penalty), lr=opts.learning_rate,
                             max_grad_norm=opts.gradient_clip)
    from dune_transformercvn_b200 import synth
    if fixed_prongs is not None:
        batch = synth.make_batch(train_events, seed=4321 + rank, fixed_prongs=fixed_prongs)
    elif args.unbalanced:
        batch = make_inputs(train_events, 4321 + rank)
    else:
        ppe = synth.balanced_prongs(train_events, 4321 + rank, max_prongs=10, total=int(round(train_events * 5.5)))
        batch = synth.make_batch(train_events, seed=4321 + rank, prongs_per_event=ppe)
    g = torch.Generator().manual_seed(99 + rank)
    ev_t = torch.randint(0, NUM_EVENT_CLASSES, (train_events,), generator=g)
    pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
    pr_t[~batch.prong_mask] = -1
    host = batch.pin()
    host_t = (ev_t.pin_memory(), pr_t.pin_memory())
    resident = batch.to(dev)
    res_t = (ev_t.to(dev), pr_t.to(dev))
    loss_host = torch.empty(1).pin_memory()

    def eager_step(b, t):
        opt.zero_grad()
        ev, pr = net.forward_sparse(b)
        loss, _ = tloss.training_loss(ev, pr, t[0], t[1], opts)   # one kernel: focal loss mix + d loss / d logits
        loss.backward()
        opt.step()
        return loss

    # the whole step as ONE CUDA graph per batch shape (training.GraphedTrainStep); the synthetic batch has a fixed shape
    mode = "kernel by kernel"
    step = eager_step
    if not args.no_train_graph:
        try:
            graphed = training.GraphedTrainStep(net, opt, opts)
            graphed(resident, res_t[0], res_t[1])
            torch.cuda.synchronize()
            step = lambda b, t: graphed(b, t[0], t[1])   # noqa: E731
            mode = "CUDA graph replay per batch shape"
        except Exception as e:   # capture not possible here (e.g. a collective that cannot be captured): launch eagerly
            mode = f"kernel by kernel (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            step = eager_step

    def step_e2e():
        b = host.to(dev, non_blocking=True)
        t = (host_t[0].to(dev, non_blocking=True), host_t[1].to(dev, non_blocking=True))
        loss_host.copy_(step(b, t).detach().reshape(1), non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        step(resident, res_t)
    l0 = tl.load().tcvn_launch_count()
    ms = timed(lambda: step(resident, res_t), args.train_steps)
    launches = tl.load().tcvn_launch_count() - l0
    if launches == 0 and step is not eager_step:     # graph replay: the library's host-side launch counter does not tick
        launches = graphed.launches_per_replay * args.train_steps
    step_e2e()
    ms_e2e = timed(step_e2e, args.train_steps)
    loss_val = float(loss_host[0])
    images = batch.num_events + batch.num_prongs
    ex = net.train_engine.exchange if isinstance(net.train_engine.exchange, training.GradientExchange) else None
    dp = None
    if world > 1:
        # data-parallel correctness on the hardware: after a step every rank must hold the SAME averaged gradient and the
        # SAME parameters (fp64 checksums of both arenas, all-gathered), and the averaged gradient must equal the mean of
        # the ranks' local gradients (one extra step with the exchange suspended)
        a = net.train_engine.arena
        opt.zero_grad()
        ev, pr = net.forward_sparse(resident)
        loss, _ = tloss.training_loss(ev, pr, res_t[0], res_t[1], opts)
        loss.backward()
        g_avg = a.gflat.clone()
        saved_ex, net.train_engine.exchange = net.train_engine.exchange, None
        opt.zero_grad()
        net.train_engine.step_index -= 1                   # same dropout masks as the step above
        ev, pr = net.forward_sparse(resident)
        loss, _ = tloss.training_loss(ev, pr, res_t[0], res_t[1], opts)
        loss.backward()
        net.train_engine.exchange = saved_ex
        g_mean = a.gflat.clone()
        dist.all_reduce(g_mean, op=dist.ReduceOp.SUM)
        g_mean /= world
        # parameters only: the arena also holds the BatchNorm running buffers, which are rank-local between two forwards
        # by design (each rank updates them from its own shard; rank 0's are broadcast before the next forward, like DDP)
        params = a.flat[opt._selector() > 0].double()
        sums = torch.stack((g_avg.double().sum(), g_avg.double().abs().sum(), params.sum(), params.abs().sum()))
        gathered = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)
        gathered = torch.stack(gathered)
        dp = {"dp_grad_checksum_equal": bool((gathered[:, :2] == gathered[0, :2]).all()),
              "dp_param_checksum_equal": bool((gathered[:, 2:] == gathered[0, 2:]).all()),
              "dp_grad_vs_mean_of_local_rel_err": float((g_avg - g_mean).norm() / g_mean.norm().clamp_min(1e-30)),
              "grad_checksum": float(gathered[0, 0]), "param_checksum": float(gathered[0, 2])}
    out = {"metric": "events/sec, training step (forward + loss + backward + gradient all-reduce + AdamW)",
           "value": world * train_events * args.train_steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
           "ms_per_step": ms / args.train_steps, "steps": args.train_steps, "scaling": "weak",
           "dtype": "f32" if args.train_precision == "fp32" else "bf16 (activations + tcgen05 GEMM operands; fp32 accumulation, parameters, statistics, gradients of parameters)",
           "config": {"workload": f"{label}: tutorial DenseNet TransformerCVN, {train_events} events/GPU "
                                  f"({images} images on rank 0; {'P = %d for every event' % fixed_prongs if fixed_prongs else ('P~U[1,10]' if args.unbalanced else 'P~U[1,10] with equal total per rank')}), "
                                  f"dropout {opts.dropout}, pixel noise {opts.pixel_noise_std}, AdamW + clip {opts.gradient_clip}",
                      "parallelism": f"event-sharded x{world}, rank-local BatchNorm (rank 0's running buffers broadcast before every "
                                     f"forward like DDP), NCCL all-reduce (AVG) of the flat fp32 gradient arena in 4 slices issued from "
                                     f"inside backward"},
           "e2e": {"value": world * train_events * args.train_steps / (ms_e2e / 1e3), "unit": UNIT,
                   "h2d_bytes_per_step": host.nbytes() + sum(t.numel() * t.element_size() for t in host_t),
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.train_steps},
           "allreduce_bytes_per_step": 0 if ex is None else ex.bytes // max(1, (max(args.warmup, 3) + 2 * args.train_steps + 1)),
           "gpu_launches": launches, "images_per_s": world * images * args.train_steps / (ms / 1e3),
           "train_gflop_per_image": 14.39, "whole_net_tflops": images * 14.39e9 * args.train_steps / (ms / 1e3) / 1e12,
           "final_loss": loss_val, "data_parallel_check": dp, "launch_mode": mode}
    if rank == 0 and world == 1 and with_cpu and not args.no_cpu_baseline:
        out["cpu_baseline"] = time_cpu_train()
    del net, opt
    torch.cuda.empty_cache()
    return out


SDXL_GFLOP_PER_IMAGE = 57.23   # 3x3 / 1x1 convolutions of the --sdxl encoder at 400x280 (block 0: 35.1, block 1: 8.8, block 2: 7.9, ...)


def run_sdxl_leg(args, dev, timed, pk):
    """BASELINE configs[3]: the --sdxl CNN variant (SDXL-style VAE encoder as the pixel-map embedding), eval inference over
    `--sdxl-events` events, bf16 tensor-core walk (dune_transformercvn_b200/sdxl.py::_SdxlCnn16).  PARITY UNPINNED: the
    reference's arithmetic lives in un-vendored, unpinned `diffusers`; held to oracle/restate_sdxl.py only."""
    from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
    from dune_transformercvn_b200.sdxl import NeutrinoSDXLNetwork
    opts = PathOptions.tutorial()
    net = NeutrinoSDXLNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16").to(dev).eval()
    batch = make_inputs(args.sdxl_events, 1234)
    resident = batch.to(dev)
    images = batch.num_events + batch.num_prongs
    steps = 3
    with torch.no_grad():
        net.forward_sparse(resident)
        net.freeze_packed(True)
        net.forward_sparse(resident)
        ms = timed(lambda: net.forward_sparse(resident), steps)
    tf = images * SDXL_GFLOP_PER_IMAGE * 1e9 * steps / (ms / 1e3) / 1e12
    out = {"metric": METRIC, "value": args.sdxl_events * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
           "dtype": "bf16", "images_per_s": images * steps / (ms / 1e3),
           "config": {"workload": f"BASELINE configs[3]: --sdxl CNN variant, eval inference, {args.sdxl_events} events "
                                  f"({images} images of 3x400x280), hit lists resident -> densify -> logits",
                      "parity": "unpinned (diffusers absent; oracle/restate_sdxl.py restates its published layout)"},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": tf / pk["bf16_sustained"], "traffic": None,
                        "kernel": "whole --sdxl network (64-channel 3x3 convolutions: umma_conv2d_c64_kernel, fused 2-D tiles; wider stages: umma_gemm_kernel<false> in shifted-GEMM form)",
                        "algorithmic_gflop_per_image": SDXL_GFLOP_PER_IMAGE,
                        "peak_source": pk["source"] + " cuBLAS bf16 sustained (kernel timed inside a long step)"}}
    del net, resident
    torch.cuda.empty_cache()
    return out


def kernel_rooflines(net, resident, batch, dev, pk, args):
    """Times the two layer kernels of dense block 1 alone (CUDA events on the launching stream, inputs ~0.8 GB
    per launch, i.e. larger than L2).  Dominant kernel of the step = the fused-activation 1x1 GEMM family
    (umma_gemm_kernel<true>, 35-40 % of the step in profiles/): HBM-bound.  Algorithmic bytes per image-layer
    (DESIGN.md): conv1 reads K*2 B and writes 256 B per ringed pixel row, conv2 reads 256 B and writes 64 B."""
    import ctypes as C
    from dune_transformercvn_b200 import lib as tl
    L = tl.load()
    eng = net.engine
    n = min(batch.num_prongs, 194)
    if n < 8:
        return None, {}
    nnz = int((resident.prong_coords[:, 0] < n).sum().item())
    coords, values = resident.prong_coords[:nnz].contiguous(), resident.prong_values[:nnz].contiguous()
    eng.cnn_sparse("prong", values, coords, n, tl.TCVN_BF16)   # leaves this chunk's feature maps in the workspace
    d = eng.cnn_desc(256)
    ws = eng.ws[("cnn", str(dev))]
    rows = n * 101 * 71          # ringed rows of block 1 (99x69 map + ring)
    pixels = n * 99 * 69
    layer, k = 2, 128            # third bottleneck of block 1: 128 input channels
    out = {}
    for which, name in ((1, "conv1"), (2, "conv2")):
        def launch():
            tl.check(L.tcvn_cnn_run_layer(C.byref(d), tl.TCVN_BF16, tl.ptr(eng.packed["prong"]), tl.ptr(ws), ws.numel(), n,
                                          0, layer, which, tl.stream_ptr(dev)), "tcvn_cnn_run_layer")
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        if which == 1:
            # algorithmic bytes = real pixels x (K in + 128 out) bf16 + the weights; the zero ring rows the layout adds are
            # overhead, not algorithm (they show up in `traffic`)
            nbytes = pixels * k * 2 + 128 * k * 2 + pixels * 128 * 2
            gbs = nbytes / (us * 1e-6) / 1e9
            out[name] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "traffic": 653.8e6 * n / 194, "kernel": "umma_gemm_kernel<true> (BN+PReLU -> 1x1 conv -> BN+PReLU), "
                         f"dense1 layer 3, {n} images", "us_per_launch": us, "algorithmic_bytes": nbytes,
                         "peak_source": pk["source"] + " copy bandwidth",
                         "traffic_source": "ncu dram__bytes_read+write (356.3 + 297.5 MB at 194 images; part of the output is still "
                                           "dirty in L2 when the kernel ends), profiles/r2_conv1_dram_bytes_wres.csv, scaled by images"}
        else:
            flops = pixels * 2 * 9 * 128 * 32
            tf = flops / (us * 1e-6) / 1e12
            out[name] = {"bound": "tensor", "achieved": tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                         "frac": tf / pk["bf16_burst"], "traffic": None, "kernel": "umma_conv2_kernel (3x3 conv, N=96 MMAs), "
                         f"dense1 layer 3, {n} images", "us_per_launch": us, "algorithmic_flops": flops,
                         "peak_source": pk["source"] + " cuBLAS bf16 burst (kernel timed alone)"}
    return out["conv1"], {"conv2": out["conv2"]}


def run_ours(args):
    import torch.distributed as dist
    from dune_transformercvn_b200 import lib as tl
    from dune_transformercvn_b200 import synth
    from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
    from dune_transformercvn_b200.ingest import densify
    from dune_transformercvn_b200.network import NeutrinoDenseNetwork

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tl.load()

    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=args.precision)
    net = net.to(dev).eval()
    net.overlap_cnns = not args.no_overlap_cnns
    net.partition_sms = args.partition_sms
    batch = make_inputs(args.events, 1234 + rank)
    images = batch.num_events + batch.num_prongs
    host = batch.pin()
    resident = batch.to(dev)
    ev_buf = pr_buf = None
    if args.materialize:
        ev_buf = torch.empty((batch.num_events, 3, H, W), dtype=torch.float32, device=dev)
        pr_buf = torch.empty((batch.num_prongs, 3, H, W), dtype=torch.float32, device=dev)
    out_host = (torch.empty((batch.num_events, NUM_EVENT_CLASSES)).pin_memory(),
                torch.empty((batch.num_events, batch.prong_mask.shape[1], NUM_PRONG_CLASSES)).pin_memory())

    def step(b):
        if args.materialize:  # literal reference sequence: densify kernel -> dense maps -> network.forward
            ev = densify(b.event_values, b.event_coords, (H, W), batch.num_events, 255.0, out=ev_buf)
            pr = densify(b.prong_values, b.prong_coords, (H, W), batch.num_prongs, 255.0, out=pr_buf)
            return net(b.features, b.extra, ev, b.event_mask, pr, b.prong_mask)
        # same arithmetic, the stem reads the hit lists directly; the ~200 launches of the step replayed as one CUDA graph
        return net.forward_sparse(b, graph=not args.no_graph)

    from dune_transformercvn_b200.ingest import Prefetcher
    prefetch = Prefetcher(dev)

    def step_e2e():
        # every step copies ITS inputs from pinned host memory and reads its logits back; the copy of the next step's
        # hit lists runs on the copy stream under this step's kernels (ingest.Prefetcher), as a serving loop would
        prefetch.submit(host)
        b = prefetch.take() if len(prefetch.queue) > 1 else None
        if b is None:          # first call: nothing staged yet
            return
        ev, pr = step(b)
        out_host[0].copy_(ev, non_blocking=True)
        out_host[1].copy_(pr, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    with torch.no_grad():
        net.freeze_packed(False)
        for _ in range(max(args.warmup, 3)):
            step(resident)
        net.freeze_packed(True)
        sampler = ClockSampler(local)
        sampler.start()
        l0 = tl.load().tcvn_launch_count()
        ms = timed(lambda: step(resident), args.steps)
        launches_timed = tl.load().tcvn_launch_count() - l0
        if launches_timed == 0:    # graph replay: the library's host-side counter does not tick; one replay = the captured launches
            launches_timed = net.graph_launches_per_replay * args.steps
        sampler.stop_flag.set()
        sampler.join(timeout=6)   # no nvidia-smi query in flight while the host-driven e2e loop is timed
        # the first call only stages a batch; then W untimed steps like the device-timed loop.  (With 2 calls a one-off 65-150 ms
        # fell inside the timed region in about one run out of three - the copy stream's first allocations are the suspect,
        # the cause was not isolated.)
        for _ in range(max(args.warmup, 3) + 1):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        config5 = None
        if not args.no_config5:
            # BASELINE configs[4]: the 2023_08_07 JSON (batch 16 per GPU) at the maximum prong count (20), timed from the COO
            # hit lists resident on the GPU: ingest (hit lists -> stem) + forward end to end
            b5 = synth.make_batch(16, seed=555 + rank, fixed_prongs=20)
            r5 = b5.to(dev)
            for _ in range(3):
                net.forward_sparse(r5, graph=not args.no_graph)
            ms5 = timed(lambda: net.forward_sparse(r5, graph=not args.no_graph), args.steps)
            config5 = {"workload": "BASELINE configs[4]: fdhd_beam_2018prod_2023_08_07.json, 16 events/GPU x 20 prongs "
                                   f"({b5.num_events + b5.num_prongs} images), COO hits resident -> logits",
                       "inference": {"value": world * 16 * args.steps / (ms5 / 1e3), "unit": UNIT, "ms_per_step": ms5 / args.steps,
                                     "images_per_s": world * (b5.num_events + b5.num_prongs) * args.steps / (ms5 / 1e3)}}
            del r5
    sampler.join(timeout=2)
    value = world * args.events * args.steps / (ms / 1e3)
    e2e_value = world * args.events * args.steps / (ms_e2e / 1e3)
    train = train_large = None
    if not args.no_train:
        train = run_train_leg(args, dev, rank, world, timed, args.train_events, True)
        if args.train_events_large > 0:
            train_large = run_train_leg(args, dev, rank, world, timed, args.train_events_large, False)
        if config5 is not None:
            t5 = run_train_leg(args, dev, rank, world, timed, 16, False, fixed_prongs=20, label="BASELINE configs[4]")
            config5["training"] = {k: t5[k] for k in ("value", "unit", "ms_per_step", "images_per_s", "whole_net_tflops", "gpu_launches")}
    h2d = host.nbytes()
    d2h = sum(t.numel() * t.element_size() for t in out_host)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    total_images = images  # per rank; every rank has the same expected count
    tflops = total_images * GFLOP_PER_IMAGE_FWD * 1e9 * args.steps / (ms / 1e3) / 1e12
    roofline, extra_rooflines = (kernel_rooflines(net, resident, batch, dev, pk, args)
                                 if args.precision == "bf16" and not args.no_roofline else (None, {}))
    if roofline is None:
        roofline = {"bound": "tensor", "achieved": tflops, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": tflops / pk["bf16_sustained"], "traffic": None, "kernel": "whole DenseNet forward (fp32 path)"}
    cpu = None
    eager = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N = 1 only
        v, cms, cores, sample = time_cpu_reference(REFERENCE_SAMPLE_EVENTS, 3, 1, 1234)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        if world == 1:
            try:
                eager = time_torch_eager_b200(dev, 64, args.train_events)
            except Exception as e:   # a baseline must never take the benchmark down
                eager = {"error": f"{type(e).__name__}: {e}"[:300]}
    sdxl = None
    if world == 1 and not args.no_sdxl:
        try:
            sdxl = run_sdxl_leg(args, dev, timed, pk)
        except Exception as e:
            sdxl = {"error": f"{type(e).__name__}: {e}"[:300]}
    single = None
    if world == 1 and not args.no_train:
        try:
            single = single_event_latency(dev, args.precision)
        except Exception as e:
            single = {"error": f"{type(e).__name__}: {e}"[:300]}
    launches = launches_timed  # counted by the library itself (tcvn_launch_count) around the timed region
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, images, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "every timed step copies one batch of hit lists from pinned host memory (copy stream, one batch "
                                "ahead: ingest.Prefetcher), runs one batch and copies its logits to pinned host memory"},
            "gpu_launches": launches, "images_per_s": value * images / args.events,
            "whole_net_tflops": tflops, "roofline": roofline, "rooflines_other": extra_rooflines,
            "cpu_baseline": cpu, "clocks": sampler.summary(), "train": train, "train_large_batch": train_large,
            "config5_max_prongs": config5, "sdxl_variant": sdxl, "launch_mode": "kernel by kernel" if args.no_graph else "CUDA graph replay per batch shape",
            "torch_eager_b200": eager, "single_event_latency": single}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner
    included) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=int, default=256)
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap-cnns", action="store_true", help="inference: run the event CNN and the prong CNN on one stream")
    ap.add_argument("--partition-sms", action="store_true", help="inference: disjoint SM budgets for the two CNN streams (measured slower: 21.7 vs 21.2 ms)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated layer-kernel timings (chunk-size sweeps)")
    ap.add_argument("--no-train", action="store_true", help="skip the training leg (BASELINE configs[2]) reported under \"train\"")
    ap.add_argument("--train-events", type=int, default=16, help="events per GPU per training step (2023_08_07 JSON batch_size)")
    ap.add_argument("--train-events-large", type=int, default=64, help="second training measurement at a larger per-GPU batch "
                    "(SURVEY 8d: config 3 is also run at 64 events/GPU); 0 = skip")
    ap.add_argument("--train-steps", type=int, default=20)
    ap.add_argument("--unbalanced", action="store_true", help="training: independent prong draws per rank (unequal work) "
                    "instead of an equal total per rank")
    ap.add_argument("--no-sdxl", action="store_true", help="skip the BASELINE configs[3] leg (--sdxl CNN variant)")
    ap.add_argument("--sdxl-events", type=int, default=256)
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE configs[4] leg (16 events x 20 prongs)")
    ap.add_argument("--no-train-graph", action="store_true", help="training: launch the step kernel by kernel instead of replaying "
                    "its CUDA graph")
    ap.add_argument("--no-graph", action="store_true", help="inference: launch the step kernel by kernel instead of replaying "
                    "its CUDA graph")
    ap.add_argument("--train-precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--materialize", action="store_true", help="build the dense pixel maps (densify kernel) instead "
                    "of feeding the stem from the hit lists")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
