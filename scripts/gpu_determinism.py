"""Run-to-run determinism of the training step from identical state: loss and gradient-arena differences between repeats."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dune_transformercvn_b200 import loss as tloss
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
events = int(sys.argv[2]) if len(sys.argv) > 2 else 16
overlap = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
drop = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
dev = torch.device("cuda:0")
torch.manual_seed(0)
opts = PathOptions.tutorial()
opts.dropout = drop
net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=prec).to(dev).train()
net.train_engine.overlap_cnns = overlap
batch = bench.make_inputs(events, 4321).to(dev)
g = torch.Generator().manual_seed(99)
ev_t = torch.randint(0, NUM_EVENT_CLASSES, (events,), generator=g).to(dev)
pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
pr_t[~batch.prong_mask.cpu()] = -1
pr_t = pr_t.to(dev)
prev = None
for rep in range(4):
    net.zero_grad()
    net.train_engine.step_index = 0
    ev, pr = net.forward_sparse(batch)
    loss, _ = tloss.training_loss(ev, pr, ev_t, pr_t, opts)
    loss.backward()
    torch.cuda.synchronize()
    gcur = net.train_engine.arena.gflat.detach().clone()
    evc = ev.detach().clone()
    msg = f"{prec} overlap={overlap} drop={drop} rep {rep}: loss {float(loss.detach()):.7f}"
    if prev is not None:
        msg += f"  logits max diff {float((evc - prev[1]).abs().max()):.3e}  grad rel diff {float((gcur - prev[0]).norm() / prev[0].norm()):.3e}"
    print(msg)
    prev = (gcur, evc)
