"""Where does a graphed training step differ from an eager one?  One step each from identical state; prints max |diff| of the
gradient arena, the moments and the parameters.  argv: [overlap 0/1]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import loss as tloss
from dune_transformercvn_b200 import synth, training
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
overlap = (sys.argv[1] != "0") if len(sys.argv) > 1 else True
dev = torch.device("cuda:0")
opts = PathOptions.tutorial()
opts.dropout = 0.0
opts.pixel_noise_std = 0.0
batch = synth.make_batch(6, seed=3, max_prongs=5).to(dev)
g = torch.Generator().manual_seed(1)
ev_t = torch.randint(0, NUM_EVENT_CLASSES, (6,), generator=g).to(dev)
pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
pr_t[~batch.prong_mask.cpu()] = -1
pr_t = pr_t.to(dev)
res = {}
for mode in ("eager", "graph", "graph2"):
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    net.load_state_dict(synth.init_state(net.specs, seed=2, perturb=True))
    net = net.to(dev).train()
    net.train_engine.overlap_cnns = overlap
    opt = training.TcvnAdamW(training.reference_param_groups(net, opts.l2_penalty), lr=1e-3, max_grad_norm=opts.gradient_clip)
    if mode == "eager":
        opt.zero_grad()
        ev, pr = net.forward_sparse(batch)
        loss, _ = tloss.training_loss(ev, pr, ev_t, pr_t, opts)
        loss.backward()
        opt.step()
    else:
        st = training.GraphedTrainStep(net, opt, opts)
        loss = st(batch, ev_t, pr_t)
    torch.cuda.synchronize()
    a = net.train_engine.arena
    res[mode] = (float(loss), a.gflat.clone(), opt._m.clone(), opt._v.clone(), a.flat.clone(), opt.grad_norm_sq.clone())
for mode in ("graph", "graph2"):
    e, gph = res["eager"], res[mode]
    print(mode, "overlap", overlap, "loss", e[0], gph[0], "| max abs diff: grad %.3e  m %.3e  v %.3e  params %.3e  norm^2 %s %s" % (
        float((e[1] - gph[1]).abs().max()), float((e[2] - gph[2]).abs().max()), float((e[3] - gph[3]).abs().max()),
        float((e[4] - gph[4]).abs().max()), float(e[5]), float(gph[5])))
    d = (e[1] - gph[1]).abs()
    if float(d.max()) > 0:
        idx = int(d.argmax())
        off = sorted((o, n) for n, o in a.offset.items())
        name = [n for o, n in off if o <= idx][-1]
        print("   largest gradient difference in", name, "count of differing elements", int((d > 0).sum()))
