"""bf16 TRAINING step of this library against the fp64 oracle, next to PyTorch's own bf16-autocast run of the same oracle.

    python scripts/gpu_bf16_oracle_gate.py [events] [out.json]

Everything but the library's own step is the checker (oracle/restate.py run with CUDA tensors: fp64 = the reference
arithmetic, bf16 autocast = what `precision=16` training of the reference would compute on this GPU).  Prints, for both
bf16 runs, the error of the logits / loss and the cosine / norm ratio of the gradient against fp64 autograd - whole
network and per parameter tensor.  tests/test_gpu_train.py::test_bf16_training_step_against_fp64_oracle asserts on the
same quantities; this script is how its thresholds were measured."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import synth  # noqa: E402
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions  # noqa: E402
from dune_transformercvn_b200.ingest import densify  # noqa: E402
from dune_transformercvn_b200.network import NeutrinoDenseNetwork  # noqa: E402
from oracle import restate  # noqa: E402


def compare(grads, ref):
    dot = na = nb = 0.0
    per = []
    gtot = sum(float((g * g).sum()) for g in ref.values())
    for n, r in ref.items():
        g = grads[n].double()
        d, a, b = float((g * r).sum()), float((g * g).sum()), float((r * r).sum())
        dot += d; na += a; nb += b
        if b > 1e-8 * gtot:
            per.append((d / ((a * b) ** 0.5 + 1e-300), n))
    per.sort()
    return {"cosine": dot / ((na * nb) ** 0.5 + 1e-300), "norm_ratio": (na / nb) ** 0.5, "worst_tensor": per[0],
            "median_tensor_cosine": per[len(per) // 2][0], "tensors": len(per)}


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max())


def run(events: int, seed: int = 77):
    dev = torch.device("cuda:0")
    opts = PathOptions.tutorial()
    opts.dropout = 0.0          # the oracle has no dropout stream to share with the kernels
    batch = synth.make_batch(events, seed=seed, max_prongs=10)
    db = batch.to(dev)
    g = torch.Generator().manual_seed(5)
    ev_t = torch.randint(0, NUM_EVENT_CLASSES, (events,), generator=g).to(dev)
    pr_t = torch.randint(0, NUM_PRONG_CLASSES, tuple(batch.prong_mask.shape), generator=g)
    pr_t[~batch.prong_mask] = -1
    pr_t = pr_t.to(dev)
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    state = synth.init_state(net.specs, seed=2, perturb=True)
    net.load_state_dict(state)
    net = net.to(dev).train()

    def oracle(dtype, autocast):
        st = {k: (v.detach().to(dev).to(dtype).requires_grad_(True) if v.is_floating_point() else v.to(dev)) for k, v in state.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            ev, pr = restate.sparse_forward(st, opts, db, train=True, dtype=dtype)
            loss = restate.training_loss(ev.float() if autocast else ev, pr.float() if autocast else pr, ev_t, pr_t, opts)
        loss.backward()
        grads = {k: v.grad.detach().double() for k, v in st.items() if v.is_floating_point() and v.grad is not None}
        return ev.detach().double(), pr.detach().double(), float(loss.detach()), grads

    restate.FUSED_ATEN = False
    r_ev, r_pr, r_loss, r_grads = oracle(torch.float64, False)
    restate.FUSED_ATEN = True      # F.batch_norm / F.prelu: the ATen kernels the reference's modules dispatch to
    a_ev, a_pr, a_loss, a_grads = oracle(torch.float32, True)
    restate.FUSED_ATEN = False
    ev_px = densify(db.event_values, db.event_coords, (400, 280), db.num_events, 255.0)
    pr_px = densify(db.prong_values, db.prong_coords, (400, 280), db.num_prongs, 255.0)
    ev, pr = net(db.features, db.extra, ev_px, db.event_mask, pr_px, db.prong_mask)
    loss = restate.training_loss(ev, pr, ev_t, pr_t, opts)
    loss.backward()
    grads = {n: p.grad.detach() for n, p in net.named_parameters() if p.grad is not None}
    valid = db.prong_mask
    out = {
        "events": events, "images": batch.num_events + batch.num_prongs,
        "tcvn_bf16": {"event_logits": rel(ev.detach(), r_ev), "prong_logits": rel(pr.detach()[valid], r_pr[valid]),
                      "loss": abs(float(loss.detach()) - r_loss) / abs(r_loss), **compare(grads, r_grads)},
        "torch_bf16_autocast": {"event_logits": rel(a_ev, r_ev), "prong_logits": rel(a_pr[valid], r_pr[valid]),
                                "loss": abs(a_loss - r_loss) / abs(r_loss), **compare(a_grads, r_grads)},
    }
    return out


if __name__ == "__main__":
    events = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    res = run(events)
    print(json.dumps(res, indent=1))
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            json.dump(res, f, indent=1)
