"""bf16 tcgen05 inference path vs the fp32 path of this library at the full BASELINE configs[1] size: logit error and
top-1 agreement of the event / prong classes (north_star: >= 99.9 %)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
dev = torch.device("cuda:0")
opts = PathOptions.tutorial()
for perturb in (False, True):
    outs = {}
    batch = synth.make_batch(256, seed=1234, max_prongs=10).to(dev)
    for prec in ("fp32", "bf16"):
        net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=prec)
        net.load_state_dict(synth.init_state(net.specs, seed=1, perturb=perturb))
        net = net.to(dev).eval()
        with torch.no_grad():
            outs[prec] = net.forward_sparse(batch)
    ev32, pr32 = outs["fp32"]; ev16, pr16 = outs["bf16"]
    m = batch.prong_mask
    e_rel = float((ev16 - ev32).abs().max() / ev32.abs().max()); p_rel = float((pr16 - pr32)[m].abs().max() / pr32[m].abs().max())
    e_agree = float((ev16.argmax(-1) == ev32.argmax(-1)).float().mean())
    p_agree = float((pr16.argmax(-1) == pr32.argmax(-1))[m].float().mean())
    # margin of the fp32 decision: top1 - top2
    t = ev32.topk(2, -1).values; em = (t[:, 0] - t[:, 1])
    t = pr32[m].topk(2, -1).values; pm = (t[:, 0] - t[:, 1])
    print(f"perturb={perturb}: logits rel err event {e_rel:.2e} prong {p_rel:.2e}; top-1 agreement event {e_agree:.4f} ({ev32.shape[0]}) "
          f"prong {p_agree:.4f} ({int(m.sum())}); fp32 margins: event min {float(em.min()):.2e} median {float(em.median()):.2e}, "
          f"prong min {float(pm.min()):.2e} median {float(pm.median()):.2e}; max |logit| {float(ev32.abs().max()):.2e}")
