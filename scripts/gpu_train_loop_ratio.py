"""Repeats the 25-step training loop of tests/test_gpu_train.py::test_training_loop_reduces_the_loss and prints the ratio
mean(last 5 losses) / mean(first 3 losses) per repeat - the spread of that ratio (bf16 training is not bit-reproducible)
and a check that the side-stream weight-gradient branch (TCVN_WGRAD_STREAM) does not change its distribution."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import loss as tloss, synth, training
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
opts = PathOptions.tutorial()
batch = synth.make_batch(8, seed=3, max_prongs=6).to(dev)
g = torch.Generator().manual_seed(1)
ev_t = torch.randint(0, NUM_EVENT_CLASSES, (8,), generator=g).to(dev)
pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
pr_t[~batch.prong_mask.cpu()] = -1
pr_t = pr_t.to(dev)
ratios = []
for rep in range(reps):
    torch.manual_seed(0)
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=precision).to(dev).train()
    opt = training.TcvnAdamW(training.reference_param_groups(net, opts.l2_penalty), lr=2e-3, max_grad_norm=opts.gradient_clip)
    losses = []
    for _ in range(25):
        opt.zero_grad()
        ev, pr = net.forward_sparse(batch)
        loss, _ = tloss.training_loss(ev, pr, ev_t, pr_t, opts)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    ratios.append(sum(losses[-5:]) / 5 / (sum(losses[:3]) / 3))
print(f"{precision} WGRAD_STREAM={os.environ.get('TCVN_WGRAD_STREAM', '1')}: ratios " + " ".join(f"{r:.3f}" for r in ratios))
