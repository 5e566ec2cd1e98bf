"""One bf16 pass of the prong CNN over N images (default: one block-0 chunk) — the ncu target."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import lib as tl, synth  # noqa: E402
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions  # noqa: E402
from dune_transformercvn_b200.ingest import densify  # noqa: E402
from dune_transformercvn_b200.network import NeutrinoDenseNetwork  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    net = NeutrinoDenseNetwork(PathOptions.tutorial(), 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    net = net.to(dev).eval()
    batch = synth.make_batch(1, seed=3, prongs_per_event=[n]).to(dev)
    px = densify(batch.prong_values, batch.prong_coords, (400, 280), n, 255.0)
    eng = net.engine
    eng.ensure_packed(tl.TCVN_BF16)
    sparse = "--sparse" in sys.argv
    for _ in range(reps):
        if sparse:
            emb = eng.cnn_sparse("prong", batch.prong_values, batch.prong_coords, n, tl.TCVN_BF16)
        else:
            emb = eng.cnn("prong", px, tl.TCVN_BF16)
    torch.cuda.synchronize()
    print("ok", float(emb.float().abs().mean()))


if __name__ == "__main__":
    main()
