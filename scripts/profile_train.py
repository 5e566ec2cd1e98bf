"""One training step of BASELINE configs[2] (16 events) bracketed by cudaProfilerStart/Stop, for
ncu --profile-from-start off --metrics gpu__time_duration.sum (launch list) or --set full -k <kernel>."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dune_transformercvn_b200 import training
from dune_transformercvn_b200 import loss as tloss
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork

events = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda:0")
opts = PathOptions.tutorial()
net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=prec).to(dev).train()
opt = training.TcvnAdamW(training.reference_param_groups(net, opts.l2_penalty), lr=opts.learning_rate, max_grad_norm=opts.gradient_clip)
batch = bench.make_inputs(events, 4321).to(dev)
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

for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("images", batch.num_events + batch.num_prongs)
