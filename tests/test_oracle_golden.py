"""Pins the CPU oracle (oracle/restate.py) against outputs of the unmodified reference (CPU)."""
import os

import pytest
import torch

from conftest import rel_err
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.params import network_specs
from oracle import restate

H, W = 400, 280
FP32_TOL = 2e-5   # oracle-vs-reference reorder noise is 1e-7..5e-6


def test_densify_bit_exact(golden_dir):
    cases = torch.load(os.path.join(golden_dir, "densify.pt"))
    for tag, c in cases.items():
        dense = restate.densify(restate.preprocess_values(c["values"]), c["coords"], H, W)
        assert list(dense.shape) == c["shape"]
        nz = dense.nonzero()
        assert torch.equal(nz.to(torch.int32), c["nz_index"])
        assert torch.equal(dense[tuple(nz.t())], c["nz_value"])        # bit-exact
        assert float(dense.double().sum()) == c["sum"]


@pytest.mark.parametrize("tag", ["default", "perturbed"])
def test_eval_forward_matches_reference(golden_dir, tutorial_options, tag):
    g = torch.load(os.path.join(golden_dir, "forward_eval.pt"))[tag]
    specs = network_specs(tutorial_options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(specs, seed=g["seed"], perturb=g["perturb"])
    assert synth.state_checksum(state) == pytest.approx(g["state_checksum"], rel=1e-12)
    batch = synth.make_batch(2, seed=g["batch_seed"], prongs_per_event=g["prongs"])
    taps = {}
    with torch.no_grad():
        ev, pr = restate.sparse_forward(state, tutorial_options, batch, taps=taps)
    assert rel_err(taps["event_embedding"], g["event_embedding"]) < FP32_TOL
    assert rel_err(taps["prong_embedding"], g["prong_embedding"]) < FP32_TOL
    assert rel_err(taps["tokens"], g["tokens"]) < FP32_TOL
    assert rel_err(taps["hidden"], g["hidden"]) < FP32_TOL
    assert rel_err(ev, g["event_logits"]) < FP32_TOL
    assert rel_err(pr, g["prong_logits"]) < FP32_TOL
    names = {"pooling0": "stem_pool"}
    for k, summ in g["prong_cnn_stages"].items():
        t = taps["prong_cnn"][names.get(k, k)]
        assert list(t.shape) == summ["shape"]
        assert rel_err(t.flatten()[summ["idx"]], summ["vals"]) < FP32_TOL
        assert float(t.double().abs().sum()) == pytest.approx(summ["abs_sum"], rel=1e-5)


@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-9, 1e-8), (torch.float32, 1e-4, 1e-2)])
def test_train_forward_backward_matches_reference(golden_dir, dtype, tol, gtol):
    """Golden = the unmodified reference run in fp64 (train-mode BN on 2..5 images is badly
    conditioned: fp32 gradient noise reaches ~2e-3, hence the loose fp32 gate and the tight fp64 one)."""
    g = torch.load(os.path.join(golden_dir, "forward_train.pt"))
    opts = PathOptions.tutorial()
    opts.dropout = 0.0
    specs = network_specs(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(specs, seed=g["seed"], perturb=True)
    assert synth.state_checksum(state) == pytest.approx(g["state_checksum"], rel=1e-12)
    state = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in state.items()}
    for s in specs:
        if s.is_param:
            state[s.name].requires_grad_(True)
    batch = synth.make_batch(2, seed=g["batch_seed"], prongs_per_event=g["prongs"])
    stats = restate.Stats()
    ev, pr = restate.sparse_forward(state, opts, batch, train=True, stats=stats, dtype=dtype)
    assert rel_err(ev, g["event_logits"]) < tol
    assert rel_err(pr, g["prong_logits"]) < tol
    loss = restate.training_loss(ev, pr, g["event_targets"], g["prong_targets"], opts)
    assert float(loss) == pytest.approx(g["loss"], rel=10 * tol)
    loss.backward()
    for name, summ in g["grads"].items():
        grad = state[name].grad
        if "full" in summ and float(summ["full"].abs().max()) < 1e-12:
            # a conv bias in front of a batch-stat BN has an identically zero gradient
            assert float(grad.abs().max()) < (1e-12 if dtype == torch.float64 else 1e-5), name
        elif "full" in summ:
            assert rel_err(grad, summ["full"]) < gtol, name
        else:
            assert rel_err(grad.flatten()[summ["idx"]], summ["vals"]) < gtol, name
    assert sorted(n for n in (s.name for s in specs if s.is_param) if state[n].grad is None) == sorted(g["no_grad"])
    for k, v in g["running"].items():
        ours = stats.updated[k].detach() if k in stats.updated else state[k]   # dead modules keep their buffers
        assert rel_err(ours, v) < 10 * tol, k


def test_collate_matches_reference_golden(golden_dir):
    """oracle.restate.collate_sparse against the frozen outputs of the reference's own MinkowskiCollection.collate_sparse."""
    cases = torch.load(os.path.join(golden_dir, "collate.pt"))
    for name, c in cases.items():
        got_c, got_v = restate.collate_sparse(c["coords"], c["values"], c["masks"])
        assert torch.equal(got_c, c["out_coords"]) and got_c.dtype == c["out_coords"].dtype, name
        assert torch.equal(got_v, c["out_values"]), name


def test_loss_matches_reference_golden(golden_dir):
    """oracle.restate.focal_loss / training_loss against the frozen fp64 outputs of the reference's own
    NeutrinoFullBaseTrainer.loss + the training_step masking (oracle/make_golden_loss.py): gamma 0 / 0.5 / 1 / 2."""
    import types
    cases = torch.load(os.path.join(golden_dir, "loss.pt"))
    assert set(cases) == {"tutorial_gamma1", "cross_entropy", "gamma2", "gamma_half"}
    for name, c in cases.items():
        ev = c["event_logits"].double().requires_grad_(True)
        pr = c["prong_logits"].double().requires_grad_(True)
        opts = types.SimpleNamespace(loss_gamma=c["gamma"], event_prong_loss_proportion=c["event_scale"])
        total = restate.training_loss(ev, pr, c["event_targets"], c["prong_targets"], opts)
        total.backward()
        assert abs(float(total) - c["total"]) < 1e-9 * abs(c["total"]), name
        sel = c["prong_targets"] >= 0
        assert abs(float(restate.focal_loss(ev.detach(), c["event_targets"], c["gamma"])) - c["event_loss"]) < 1e-9, name
        assert abs(float(restate.focal_loss(pr.detach()[sel], c["prong_targets"][sel], c["gamma"])) - c["prong_loss"]) < 1e-9, name
        assert float((ev.grad - c["d_event_logits"]).abs().max()) < 1e-10, name
        assert float((pr.grad - c["d_prong_logits"]).abs().max()) < 1e-10, name


def test_export_surface_matches_reference_golden(golden_dir, tutorial_options):
    """oracle.restate.export_combined against the frozen outputs of the notebook wrapper's operation sequence run on the
    unmodified reference network (oracle/make_golden_export.py)."""
    from dune_transformercvn_b200 import synth
    from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES
    from dune_transformercvn_b200.params import network_specs
    g = torch.load(os.path.join(golden_dir, "export.pt"))
    specs = network_specs(tutorial_options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(specs, seed=g["state_seed"], perturb=True)
    assert synth.state_checksum(state) == pytest.approx(g["state_checksum"], rel=1e-12)
    c = g["events"]["one_prong"]        # the six-prong event is checked on the GPU (tests/test_gpu_export.py)
    batch = synth.make_batch(1, seed=c["batch_seed"], prongs_per_event=[c["n_prongs"]])
    px = torch.cat((restate.densify(batch.event_values.float(), batch.event_coords, 400, 280),
                    restate.densify(batch.prong_values.float(), batch.prong_coords, 400, 280))).to(torch.uint8)
    assert int(px.long().sum()) == c["pixel_sum"]
    with torch.no_grad():
        got = restate.export_combined(state, tutorial_options, px)
    for a, k in zip(got, ("event_prob", "prong_prob", "event_features", "prong_features")):
        assert rel_err(a, c[k]) < 2e-5, k
