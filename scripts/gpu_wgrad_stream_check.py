"""A/B check of the side-stream weight-gradient branch (TCVN_WGRAD_STREAM): one bf16 training step from identical state,
gradient arena dumped to a file; run twice (env 0 / 1) and compare with `--compare a b`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "--compare":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    d = (a - b).abs()
    print(f"max |a-b| = {float(d.max()):.3e}, max |a| = {float(a.abs().max()):.3e}, "
          f"differing elements {int((d > 0).sum())} of {a.numel()}, rel fro {float(d.norm() / a.norm()):.3e}")
    sys.exit(0)
import bench
from dune_transformercvn_b200 import loss as tloss, training
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
events = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
opts = PathOptions.tutorial()
net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16").to(dev).train()
batch = bench.make_inputs(events, 4321).to(dev)
g = torch.Generator().manual_seed(99)
ev_t = torch.randint(0, NUM_EVENT_CLASSES, (events,), generator=g).to(dev)
pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
pr_t[~batch.prong_mask.cpu()] = -1
pr_t = pr_t.to(dev)
for rep in range(3):     # same state every time: no optimizer step; gradients zeroed
    net.zero_grad()
    net.train_engine.step_index = 0
    ev, pr = net.forward_sparse(batch)
    loss, _ = tloss.training_loss(ev, pr, ev_t, pr_t, opts)
    loss.backward()
    torch.cuda.synchronize()
torch.save(net.train_engine.arena.gflat.detach().cpu().clone(), sys.argv[1])
print("loss", float(loss.detach()), "grad norm", float(net.train_engine.arena.gflat.norm()))
