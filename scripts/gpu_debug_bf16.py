"""GPU debug aid: bf16 tcgen05 path vs the CPU oracle, stage by stage (run under gpurun)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import lib as tl, synth  # noqa: E402
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions  # noqa: E402
from dune_transformercvn_b200.ingest import densify  # noqa: E402
from dune_transformercvn_b200.network import NeutrinoDenseNetwork  # noqa: E402
from oracle import restate  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    dev = torch.device("cuda:0")
    opts = PathOptions.tutorial()
    perturb = "--default" not in sys.argv
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    state = synth.init_state(net.specs, seed=1, perturb=perturb)
    net.load_state_dict(state)
    net = net.to(dev).eval()
    batch = synth.make_batch(2, seed=21, prongs_per_event=[3, 1])
    gb = batch.to(dev)
    taps = {}
    with torch.no_grad():
        want_ev, want_pr = restate.sparse_forward(state, opts, batch, taps=taps)
    pr = densify(gb.prong_values, gb.prong_coords, (400, 280), batch.num_prongs, 255.0)
    eng = net.engine
    eng.ensure_packed(tl.TCVN_BF16)
    emb = eng.cnn("prong", pr, tl.TCVN_BF16)
    torch.cuda.synchronize()
    names = ["stem_pool"]
    for b in range(5):
        names.append(f"dense{b + 1}")
        if b < 4:
            names.append(f"transition{b + 1}")
    for stage, name in enumerate(names):
        got = eng.read_stage("prong", batch.num_prongs, stage, tl.TCVN_BF16, dev).cpu()
        want = taps["prong_cnn"][name]
        line = f"{name:12s} rel={rel(got, want):.3e} nan={int(torch.isnan(got).sum())}"
        if name.startswith("dense"):
            c0 = want.shape[1] - 32 * {"dense1": 3, "dense2": 6, "dense3": 12, "dense4": 6, "dense5": 3}[name]
            parts = [rel(got[:, :c0], want[:, :c0])]
            for i in range((want.shape[1] - c0) // 32):
                sl = slice(c0 + 32 * i, c0 + 32 * i + 32)
                parts.append(rel(got[:, sl], want[:, sl]))
            line += " slices=" + " ".join(f"{p:.1e}" for p in parts)
        print(line, flush=True)
    print("embedding rel", rel(emb.cpu(), taps["prong_embedding"]))
    with torch.no_grad():
        ev, prl = net.forward_sparse(gb)
    print("event logits rel", rel(ev.cpu(), want_ev), "prong logits rel", rel(prl.cpu(), want_pr))
    print("argmax agree", bool((ev.argmax(-1).cpu() == want_ev.argmax(-1)).all()))


if __name__ == "__main__":
    main()
