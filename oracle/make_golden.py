"""TEST INFRASTRUCTURE — generates tests/golden/* by running the UNMODIFIED reference on CPU.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

Fixtures (small on purpose; weights are NOT stored — they are regenerated from a seed by
``dune_transformercvn_b200.synth.init_state`` and fingerprinted):
  state_dict_keys.json   names/shapes/dtypes of the reference network's state_dict + param order
  densify.pt             COO hits and the reference ``sparse_to_dense`` output as a sparse checksum
  forward_eval.pt        eval-mode logits + intermediates of the reference for a 2-event batch
  forward_train.pt       train-mode (batch-stat BN, dropout 0) logits, updated running stats, and
                         gradients of the reference loss w.r.t. a sample of parameters
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_import                                     # noqa: E402
from dune_transformercvn_b200 import synth                              # noqa: E402
from dune_transformercvn_b200.params import network_specs               # noqa: E402
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
H, W = 400, 280


def sample_points(t: torch.Tensor, k: int = 64, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, t.numel(), (k,), generator=g)
    return idx, t.flatten()[idx].clone()


def summarize(t: torch.Tensor):
    idx, vals = sample_points(t)
    return {"shape": list(t.shape), "sum": float(t.double().sum()), "abs_sum": float(t.double().abs().sum()),
            "idx": idx, "vals": vals}


def build_reference(ref, options, state):
    net = ref.NeutrinoDenseNetwork(options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    missing = net.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    ref = reference_import.load()
    options = ref.tutorial_options()

    # ---- 1. state_dict inventory ------------------------------------------------------------
    net = ref.NeutrinoDenseNetwork(options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    sd = net.state_dict()
    inv = {"keys": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
           "params": [n for n, _ in net.named_parameters()],
           "num_params": sum(p.numel() for p in net.parameters())}
    with open(os.path.join(GOLDEN, "state_dict_keys.json"), "w") as f:
        json.dump(inv, f)
    print("state_dict:", len(inv["keys"]), "tensors,", inv["num_params"], "params")

    # ---- 2. densify --------------------------------------------------------------------------
    batch = synth.make_batch(3, seed=11, prongs_per_event=[2, 1, 3])
    cases = {}
    for tag, vals, coords in (("event", batch.event_values, batch.event_coords),
                              ("prong", batch.prong_values, batch.prong_coords)):
        dense = ref.sparse_to_dense(vals / 255.0, coords, (H, W))
        nz = dense.nonzero()
        cases[tag] = {"values": vals, "coords": coords, "shape": list(dense.shape),
                      "nz_index": nz.to(torch.int32), "nz_value": dense[tuple(nz.t())],
                      "sum": float(dense.double().sum())}
    torch.save(cases, os.path.join(GOLDEN, "densify.pt"))
    print("densify:", {k: v["shape"] for k, v in cases.items()})

    # ---- 3. eval forward ---------------------------------------------------------------------
    specs = network_specs(options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    out = {}
    for tag, perturb, seed in (("default", False, 0), ("perturbed", True, 1)):
        state = synth.init_state(specs, seed=seed, perturb=perturb)
        net = build_reference(ref, options, state).eval()
        batch = synth.make_batch(2, seed=21, prongs_per_event=[3, 1])
        ev = ref.sparse_to_dense(batch.event_values / 255.0, batch.event_coords, (H, W))
        pr = ref.sparse_to_dense(batch.prong_values / 255.0, batch.prong_coords, (H, W))
        taps = {}
        hooks = []
        cnn = net.prong_embedding.prong_pixel_embedding
        for name in ("pooling0", "dense1", "transition1", "dense2", "transition2", "dense3", "transition3",
                     "dense4", "transition4", "dense5"):
            mod = getattr(cnn.features, name)
            hooks.append(mod.register_forward_hook(lambda m, i, o, name=name: taps.__setitem__(name, o.detach())))
        with torch.no_grad():
            ev_emb = net.prong_embedding.event_pixel_embedding(ev)
            pr_emb = net.prong_embedding.prong_pixel_embedding(pr)
            tokens, mask = net.prong_embedding(batch.features, batch.extra, ev, batch.event_mask, pr, batch.prong_mask)
            hidden = net.encoder(tokens, mask)[0]
            ev_logits, pr_logits = net(batch.features, batch.extra, ev, batch.event_mask, pr, batch.prong_mask)
        for h in hooks:
            h.remove()
        out[tag] = {"seed": seed, "perturb": perturb, "state_checksum": synth.state_checksum(state),
                    "batch_seed": 21, "prongs": [3, 1],
                    "event_embedding": ev_emb, "prong_embedding": pr_emb, "tokens": tokens, "hidden": hidden,
                    "event_logits": ev_logits, "prong_logits": pr_logits.contiguous(),
                    "prong_cnn_stages": {k: summarize(v) for k, v in taps.items()}}
        print(tag, "event_logits", ev_logits[0].tolist())
    torch.save(out, os.path.join(GOLDEN, "forward_eval.pt"))

    # ---- 4. train-mode forward/backward (dropout 0, no pixel noise) ---------------------------
    options_t = ref.tutorial_options()
    options_t.dropout = 0.0
    specs_t = network_specs(options_t, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(specs_t, seed=2, perturb=True)
    # fp64 on purpose: train-mode BN over 2..5 images is badly conditioned, fp32-vs-fp32 gradient
    # noise reaches 2e-3; an fp64 run of the unmodified reference modules is the ground truth.
    net = build_reference(ref, options_t, state).double().train()
    batch = synth.make_batch(2, seed=31, prongs_per_event=[2, 3])
    ev = ref.sparse_to_dense(batch.event_values / 255.0, batch.event_coords, (H, W)).double()
    pr = ref.sparse_to_dense(batch.prong_values / 255.0, batch.prong_coords, (H, W)).double()
    ev_logits, pr_logits = net(batch.features.double(), batch.extra.double(), ev, batch.event_mask, pr,
                               batch.prong_mask)
    ev_t = torch.tensor([1, 3])
    pr_t = torch.tensor([[0, 5, -1], [7, 2, 4]])
    trainer = ref.dense_trainer.NeutrinoFullDenseTrainer
    fake = type("T", (), {"gamma": options_t.loss_gamma})()
    sel = pr_t >= 0
    le = trainer.loss(fake, ev_logits, ev_t)
    lp = trainer.loss(fake, pr_logits[sel], pr_t[sel])
    loss = options_t.event_prong_loss_proportion * le + (1 - options_t.event_prong_loss_proportion) * lp
    loss.backward()
    named = dict(net.named_parameters())
    grad_names = [
        "prong_embedding.prong_pixel_embedding.features.conv0.weight",
        "prong_embedding.prong_pixel_embedding.features.norm0.weight",
        "prong_embedding.prong_pixel_embedding.features.relu0.weight",
        "prong_embedding.prong_pixel_embedding.features.dense1.layers.0.bottleneck_block.conv1.weight",
        "prong_embedding.prong_pixel_embedding.features.dense1.layers.2.output_block.conv2.weight",
        "prong_embedding.prong_pixel_embedding.features.dense1.layers.1.output_block.norm2.bias",
        "prong_embedding.prong_pixel_embedding.features.dense3.layers.5.bottleneck_block.relu1.weight",
        "prong_embedding.prong_pixel_embedding.features.transition2.conv.weight",
        "prong_embedding.prong_pixel_embedding.output_block.linear.weight",
        "prong_embedding.event_pixel_embedding.features.dense5.layers.2.output_block.conv2.bias",
        "prong_embedding.event_pixel_embedding.features.final_norm.weight",
        "prong_embedding.event_position_embedding",
        "prong_embedding.combined_embedding.linear.weight",
        "prong_embedding.combined_embedding.norm.weight",
        "encoder.encoder.layers.0.self_attn.in_proj_weight",
        "encoder.encoder.layers.3.linear1.weight",
        "encoder.encoder.layers.5.norm2.bias",
        "event_decoder.hidden_layer.weight",
        "prong_decoder.hidden_layers.0.weight",
        "prong_decoder.hidden_layers.4.weight",
        "prong_decoder.output_layer.bias",
    ]
    grads = {}
    for n in grad_names:
        g = named[n].grad
        grads[n] = summarize(g) if g.numel() > 4096 else {"full": g.clone()}
    no_grad = [n for n, p in named.items() if p.grad is None]
    grad_norm = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in named.values() if p.grad is not None)))
    sd = net.state_dict()
    stats_names = [k for k in sd if k.endswith("running_var") or k.endswith("running_mean")]
    pick = stats_names[:4] + stats_names[100:104] + stats_names[-6:]
    train_out = {"seed": 2, "batch_seed": 31, "prongs": [2, 3], "dtype": "float64", "state_checksum": synth.state_checksum(state),
                 "event_targets": ev_t, "prong_targets": pr_t,
                 "event_logits": ev_logits.detach(), "prong_logits": pr_logits.detach().contiguous(),
                 "loss": float(loss), "grad_norm": grad_norm, "grads": grads, "no_grad": no_grad,
                 "running": {k: sd[k].clone() for k in pick}}
    torch.save(train_out, os.path.join(GOLDEN, "forward_train.pt"))
    print("train loss", float(loss), "grad_norm", grad_norm, "no-grad params:", len(no_grad))
    for fn in os.listdir(GOLDEN):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)), "bytes")


if __name__ == "__main__":
    main()
