"""Is the training step host-bound?  Per step: host time to ISSUE the step (no sync) against the device time between
CUDA events, for a few batch sizes.  python scripts/gpu_train_host_bound.py [events ...]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dune_transformercvn_b200 import lib as tl
from dune_transformercvn_b200 import loss as tloss
from dune_transformercvn_b200 import synth, training
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork

dev = torch.device("cuda:0")
opts = PathOptions.tutorial()
for events in [int(a) for a in sys.argv[1:]] or [4, 16, 64]:
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16").to(dev).train()
    opt = training.TcvnAdamW(training.reference_param_groups(net, opts.l2_penalty), lr=opts.learning_rate, max_grad_norm=opts.gradient_clip)
    ppe = synth.balanced_prongs(events, 4321, max_prongs=10, total=int(round(events * 5.5)))
    batch = synth.make_batch(events, seed=4321, prongs_per_event=ppe).to(dev)
    g = torch.Generator().manual_seed(99)
    ev_t = torch.randint(0, NUM_EVENT_CLASSES, (events,), generator=g).to(dev)
    pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
    pr_t[~batch.prong_mask.cpu()] = -1
    pr_t = pr_t.to(dev)

    def step():
        opt.zero_grad()
        ev, pr = net.forward_sparse(batch)
        loss, _ = tloss.training_loss(ev, pr, ev_t, pr_t, opts)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n = 10
    l0 = tl.load().tcvn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    host = (time.perf_counter() - t0) / n * 1e3
    e1.record()
    torch.cuda.synchronize()
    print(f"{events} events ({batch.num_events + batch.num_prongs} images): device {e0.elapsed_time(e1) / n:.2f} ms/step, host issue "
          f"{host:.2f} ms/step, {(tl.load().tcvn_launch_count() - l0) // n} launches/step")
    del net, opt
    torch.cuda.empty_cache()
