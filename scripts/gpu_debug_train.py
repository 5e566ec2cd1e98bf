"""GPU box debug run of the training path: every parameter gradient, the running buffers, the optimizer step and the
dropout mask consistency against the CPU oracle (fp64 autograd over oracle/restate.py).  Prints a table; asserts nothing."""
import copy
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dune_transformercvn_b200 import synth, training  # noqa: E402
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions  # noqa: E402
from dune_transformercvn_b200.ingest import densify  # noqa: E402
from dune_transformercvn_b200.network import NeutrinoDenseNetwork  # noqa: E402
from oracle import restate  # noqa: E402

H, W = 400, 280
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def oracle_run(state, opts, batch, ev_t, pr_t, dtype=torch.float64):
    st = {k: (v.detach().clone().to(dtype).requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in state.items()}
    stats = restate.Stats()
    ev, pr = restate.sparse_forward(st, opts, batch, train=True, stats=stats, dtype=dtype)
    loss = restate.training_loss(ev, pr, ev_t, pr_t, opts)
    loss.backward()
    grads = {k: v.grad for k, v in st.items() if v.is_floating_point() and v.grad is not None}
    return ev.detach(), pr.detach(), float(loss), grads, stats.updated


def main():
    opts = PathOptions.tutorial()
    opts.dropout = 0.0
    prongs = [2, 3] if "--big" not in sys.argv else [3, 1, 4, 2]
    batch = synth.make_batch(len(prongs), seed=31, prongs_per_event=prongs)
    precision = "bf16" if "--bf16" in sys.argv else "fp32"
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=precision)
    state = synth.init_state(net.specs, seed=2, perturb=True)
    net.load_state_dict(state)
    net = net.to(dev).train()
    g = torch.Generator().manual_seed(5)
    ev_t = torch.randint(0, NUM_EVENT_CLASSES, (len(prongs),), generator=g)
    pr_t = torch.randint(0, NUM_PRONG_CLASSES, (len(prongs), max(prongs)), generator=g)
    pr_t[~batch.prong_mask] = -1
    if prongs == [2, 3]:
        ev_t = torch.tensor([1, 3])
        pr_t = torch.tensor([[0, 5, -1], [7, 2, 4]])

    t0 = time.time()
    o_ev, o_pr, o_loss, o_grads, o_stats = oracle_run(state, opts, batch, ev_t, pr_t)
    print(f"oracle fp64 fwd+bwd: {time.time() - t0:.1f}s loss {o_loss:.9f}")
    gold_path = os.path.join(ROOT, "tests", "golden", "forward_train.pt")
    if prongs == [2, 3] and os.path.exists(gold_path):
        gold = torch.load(gold_path)
        print("oracle vs golden(reference fp64): loss", abs(o_loss - gold["loss"]), "ev", rel(o_ev, gold["event_logits"]))

    db = batch.to(dev)
    ev_px = densify(db.event_values, db.event_coords, (H, W), db.num_events, 255.0)
    pr_px = densify(db.prong_values, db.prong_coords, (H, W), db.num_prongs, 255.0)
    ev, pr = net(db.features, db.extra, ev_px, db.event_mask, pr_px, db.prong_mask)
    torch.cuda.synchronize()
    print("fwd: ev rel", rel(ev, o_ev), "pr rel", rel(pr, o_pr))
    loss = restate.training_loss(ev, pr, ev_t.to(dev), pr_t.to(dev), opts)
    print("loss", float(loss), "oracle", o_loss)
    loss.backward()
    torch.cuda.synchronize()
    named = dict(net.named_parameters())
    worst = []
    tot = 0.0
    for n, p in named.items():
        if p.grad is None:
            if n in o_grads:
                print("MISSING grad", n)
            continue
        if n not in o_grads:
            print("EXTRA grad", n, float(p.grad.abs().max()))
            continue
        worst.append((float((p.grad.double().cpu() - o_grads[n]).abs().max()), n, float(o_grads[n].abs().max())))
        tot += float((p.grad.double() ** 2).sum())
    gmax = max(w[2] for w in worst)
    # error of a tensor relative to max(its own scale, 1e-4 of the largest gradient entry of the network): gradients
    # that are mathematically zero (conv biases in front of a train-mode BatchNorm) are pure rounding noise
    worst = [(w[0] / max(w[2], 1e-4 * gmax), w[1], w[2]) for w in worst]
    worst.sort(reverse=True)
    print("largest gradient entry", gmax)
    print("grad norm", tot ** 0.5, "oracle", float(sum((v.double() ** 2).sum() for v in o_grads.values())) ** 0.5)
    print("worst 25 parameter gradients (rel err, name, max|ref|):")
    for w in worst[:25]:
        print(f"  {w[0]:.3e}  {w[1]}  {w[2]:.3e}")
    import statistics
    print("median rel err", statistics.median(w[0] for w in worst), "n", len(worst))
    bad = [w for w in worst if w[0] > 1e-2]
    print("count > 1e-2:", len(bad))
    # group summary by component
    groups = {}
    for e, n, _ in worst:
        key = ".".join(n.split(".")[:4])
        groups[key] = max(groups.get(key, 0.0), e)
    for k in sorted(groups):
        print(f"   {groups[k]:.3e}  {k}")
    # running buffers
    sd = net.state_dict()
    rs = max(rel(sd[k], v) for k, v in o_stats.items())
    print("running buffers worst rel", rs, "over", len(o_stats))
    nbt = [v for k, v in sd.items() if k.endswith("num_batches_tracked")]
    print("num_batches_tracked values:", sorted(set(int(v) for v in nbt)))

    # ---- optimizer step vs torch.optim.AdamW on the oracle gradients (fp32)
    ref_params = {k: v.clone().float().requires_grad_(True) for k, v in state.items() if k in named}
    for k, v in ref_params.items():
        if k in o_grads:
            v.grad = o_grads[k].float()
            named[k].grad.copy_(v.grad.to(dev))     # same gradients on both sides: this checks the optimizer alone
    no_decay = ("bias", "LayerNorm.weight")
    groups_ref = [{"params": [p for n, p in ref_params.items() if not any(nd in n for nd in no_decay)], "weight_decay": 2.13e-5},
                  {"params": [p for n, p in ref_params.items() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    ropt = torch.optim.AdamW(groups_ref, lr=1e-3)
    torch.nn.utils.clip_grad_norm_([p for p in ref_params.values() if p.grad is not None], 43.0)
    ropt.step()
    opt = training.TcvnAdamW(training.reference_param_groups(net, 2.13e-5), lr=1e-3, max_grad_norm=43.0)
    opt.step()
    torch.cuda.synchronize()
    errs = [(rel(named[k].detach() - state[k].to(dev), ref_params[k].detach() - state[k]), k) for k in ref_params if k in o_grads]
    errs.sort(reverse=True)
    print("AdamW update (delta) worst rel errs:", [(f"{e:.2e}", k) for e, k in errs[:5]])
    unchanged = [k for k in ref_params if k not in o_grads and not torch.equal(named[k].detach().cpu(), state[k])]
    print("no-grad params changed by the optimizer:", unchanged)

    # ---- dropout: forward/backward mask consistency by a directional finite difference
    opts2 = PathOptions.tutorial()
    net2 = NeutrinoDenseNetwork(opts2, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=precision)
    net2.load_state_dict(synth.init_state(net2.specs, seed=2, perturb=True))
    net2 = net2.to(dev).train()
    eng = net2.train_engine

    def run(step_index):
        eng.step_index = step_index
        e, p = net2(db.features, db.extra, ev_px, db.event_mask, pr_px, db.prong_mask)
        return restate.training_loss(e, p, ev_t.to(dev), pr_t.to(dev), opts2)

    l0 = run(100)
    l0.backward()
    torch.cuda.synchronize()
    gflat = eng.arena.gflat.clone()
    flat0 = eng.arena.flat.clone()
    l0b = run(100)
    print("dropout: same seed -> same loss", float(l0), float(l0b), " other seed ->", float(run(101)))
    sel = torch.zeros_like(gflat)
    for n, v in eng.arena.gviews.items():
        o = eng.arena.offset[n]
        sel[o:o + v.numel()] = 1.0
    d = gflat * sel
    d = d / d.norm()
    for eps in (1e-4, 3e-5, 1e-5):
        eng.arena.flat.copy_(flat0 + eps * d)
        lp = float(run(100))
        eng.arena.flat.copy_(flat0 - eps * d)
        lm = float(run(100))
        print(f"dropout FD eps={eps}: (L+ - L-)/2eps = {(lp - lm) / (2 * eps):.6f}  g.d = {float((gflat * d).sum()):.6f}")
    eng.arena.flat.copy_(flat0)
    # timing of one fp32 training step
    for _ in range(2):
        l = run(200)
        l.backward()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        l = run(200)
        l.backward()
    torch.cuda.synchronize()
    print(f"{precision} train fwd+bwd, {len(prongs)} events / {sum(prongs)} prongs: {(time.time() - t0) / 3 * 1e3:.1f} ms/step")


if __name__ == "__main__":
    main()
