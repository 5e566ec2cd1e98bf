"""Host (CPU) time vs device time of one training step: is the step launch-bound on the host?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dune_transformercvn_b200 import training
from dune_transformercvn_b200 import loss as tloss
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
events = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
opts = PathOptions.tutorial()
net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16").to(dev).train()
opt = training.TcvnAdamW(training.reference_param_groups(net, opts.l2_penalty), lr=opts.learning_rate, max_grad_norm=opts.gradient_clip)
batch = bench.make_inputs(events, 4321).to(dev)
g = torch.Generator().manual_seed(99)
ev_t = torch.randint(0, NUM_EVENT_CLASSES, (events,), generator=g).to(dev)
pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
pr_t[~batch.prong_mask.cpu()] = -1
pr_t = pr_t.to(dev)
for overlap in (True, False):
    net.train_engine.overlap_cnns = overlap
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        ev, pr = net.forward_sparse(batch)
        t1 = time.perf_counter()
        loss, _ = tloss.training_loss(ev, pr, ev_t, pr_t, opts)
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        opt.step()
        t4 = time.perf_counter()
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        if rep >= 2:
            print(f"overlap={overlap} host ms: fwd {1e3*(t1-t0):.2f} loss {1e3*(t2-t1):.2f} bwd {1e3*(t3-t2):.2f} opt {1e3*(t4-t3):.2f} | drain {1e3*(t5-t4):.2f} | total {1e3*(t5-t0):.2f}")
