"""bf16 training path vs the fp32 training path of this library on a well-conditioned batch: per-tensor cosine and norm
ratio of the gradients, grouped by tensor kind and dense block.  A wiring bug shows as cosine ~ 0 for a whole kind."""
import os, sys, time, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.ingest import densify
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
from oracle import restate

dev = torch.device("cuda:0")
events = int(sys.argv[1]) if len(sys.argv) > 1 else 12
opts = PathOptions.tutorial(); opts.dropout = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
batch = synth.make_batch(events, seed=77, max_prongs=10)
db = batch.to(dev)
ev_px = densify(db.event_values, db.event_coords, (400, 280), db.num_events, 255.0)
pr_px = densify(db.prong_values, db.prong_coords, (400, 280), db.num_prongs, 255.0)
g = torch.Generator().manual_seed(5)
ev_t = torch.randint(0, NUM_EVENT_CLASSES, (events,), generator=g).to(dev)
pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
pr_t[~batch.prong_mask] = -1
pr_t = pr_t.to(dev)
grads, logits = {}, {}
for prec in ("fp32", "bf16"):
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=prec)
    net.load_state_dict(synth.init_state(net.specs, seed=2, perturb=True))
    net = net.to(dev).train()
    net.train_engine.step_index = 10
    ev, pr = net(db.features, db.extra, ev_px, db.event_mask, pr_px, db.prong_mask)
    loss = restate.training_loss(ev, pr, ev_t, pr_t, opts)
    loss.backward()
    torch.cuda.synchronize()
    grads[prec] = {n: p.grad.detach().double().clone() for n, p in net.named_parameters() if p.grad is not None}
    logits[prec] = (ev.detach().double(), pr.detach().double(), float(loss))
    # timing
    for _ in range(2):
        net.zero_grad(); ev, pr = net(db.features, db.extra, ev_px, db.event_mask, pr_px, db.prong_mask)
        restate.training_loss(ev, pr, ev_t, pr_t, opts).backward()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(3):
        net.zero_grad(); ev, pr = net(db.features, db.extra, ev_px, db.event_mask, pr_px, db.prong_mask)
        restate.training_loss(ev, pr, ev_t, pr_t, opts).backward()
    torch.cuda.synchronize()
    print(f"{prec}: loss {logits[prec][2]:.6f}  fwd+bwd {(time.time()-t0)/3*1e3:.1f} ms for {events} events / {batch.num_prongs} prongs")
    del net
a, b = logits["fp32"], logits["bf16"]
print("logits bf16 vs fp32: ev rel", float((a[0]-b[0]).abs().max()/a[0].abs().max()), "pr rel", float((a[1]-b[1]).abs().max()/a[1].abs().max()))
tot_dot = tot_a = tot_b = 0.0
groups = collections.defaultdict(list)
for n, ga in grads["fp32"].items():
    gb = grads["bf16"][n]
    dot, na, nb = float((ga*gb).sum()), float(ga.norm()), float(gb.norm())
    tot_dot += dot; tot_a += na*na; tot_b += nb*nb
    if n.endswith(("conv1.bias", "conv2.bias", "conv.bias", "conv0.bias")) or "position" in n:
        continue   # true gradient is zero
    parts = n.split(".")
    if "features" in parts:
        i = parts.index("features")
        blk = parts[i+1]
        kind = ".".join(parts[-2:]) if blk.startswith(("dense",)) else ".".join(parts[i+1:])
        key = (parts[1][:5], blk if blk.startswith("dense") else "-", kind)
    else:
        key = (parts[0][:12], "-", ".".join(parts[-2:]))
    groups[key].append((dot/(na*nb+1e-300), nb/(na+1e-300)))
print("whole gradient: cosine", tot_dot/(tot_a*tot_b)**0.5, "norm ratio", (tot_b/tot_a)**0.5)
for key in sorted(groups):
    v = groups[key]
    cs = [x[0] for x in v]; rs = [x[1] for x in v]
    print(f"  {key[0]:12s} {key[1]:8s} {key[2]:34s} n={len(v):3d} cos min {min(cs):.4f} mean {sum(cs)/len(cs):.4f}   ratio {min(rs):.3f}..{max(rs):.3f}")
