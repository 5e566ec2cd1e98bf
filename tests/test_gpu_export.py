"""GPU parity of the single-event export surface (dune_transformercvn_b200/export.py) against the oracle's restatement
of the reference's CreateCompiled.ipynb wrappers (cells 6-8): pixels/255 -> network with all-true masks -> softmax of
the logits + the encoder's hidden vectors.  fp32 path: 1e-4 (north_star); the CUDA-graph replay must be bit-identical to
the un-captured call sequence."""
import pytest
import torch

from conftest import rel_err
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.export import EventClassifier, ExportableEventClassifier
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
from oracle import restate

pytestmark = pytest.mark.gpu
H, W = 400, 280


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _pixels(n_prongs, seed):
    """(1 + n, 3, H, W) uint8 maps: the dense form of one synthetic event, as LArSoft hands it over."""
    batch = synth.make_batch(1, seed=seed, prongs_per_event=[n_prongs])
    ev = restate.densify(batch.event_values.float(), batch.event_coords, H, W)
    pr = restate.densify(batch.prong_values.float(), batch.prong_coords, H, W)
    return torch.cat((ev, pr)).to(torch.uint8)


def _oracle(state, opts, pixels):
    with torch.no_grad():
        return restate.export_combined(state, opts, pixels)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_event_classifier_matches_oracle_and_graph_replay_is_exact(dev, precision, tol):
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=precision)
    state = synth.init_state(net.specs, seed=3, perturb=True)
    net.load_state_dict(state)
    net = net.to(dev).eval()
    graphed = EventClassifier(net, "combined", use_graph=True)
    plain = EventClassifier(net, "combined", use_graph=False)
    for n_prongs, seed in ((6, 11), (1, 12), (6, 13)):      # the third call replays the graph captured by the first
        px = _pixels(n_prongs, seed)
        want = _oracle(state, opts, px)
        got = graphed(px.to(dev))
        ref = plain(px.to(dev))
        assert [tuple(t.shape) for t in got] == [(4,), (n_prongs, 8), (128,), (n_prongs, 128)]
        for g, r, w in zip(got, ref, want):
            assert torch.equal(g, r), "graph replay differs from the un-captured call sequence"
            assert rel_err(g.cpu(), w) < tol
        assert abs(float(got[0].sum()) - 1) < 1e-5
    assert len(graphed._plans) == 2
    pid = EventClassifier(net, "pid", use_graph=False)(_pixels(2, 5).to(dev))
    emb = EventClassifier(net, "embeddings", use_graph=False)(_pixels(2, 5).to(dev))
    assert [tuple(t.shape) for t in pid] == [(4,), (2, 8)] and [tuple(t.shape) for t in emb] == [(128,), (2, 128)]


def test_graph_plans_survive_growing_workspaces(dev):
    """a captured graph bakes in workspace / packed-block pointers.  Prong counts 1, 6, 1 make the second plan
    grow the workspaces the first plan was captured over, and a batched forward on the same network grows them again;
    every replay must still equal the un-captured call sequence bit for bit (stale plans are re-captured)."""
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    net.load_state_dict(synth.init_state(net.specs, seed=3, perturb=True))
    net = net.to(dev).eval()
    graphed = EventClassifier(net, "combined", use_graph=True)
    plain = EventClassifier(net, "combined", use_graph=False)
    big = synth.make_batch(6, seed=5, max_prongs=10).to(dev)
    for i, (n_prongs, seed) in enumerate(((1, 12), (6, 11), (1, 14), (6, 15), (1, 16))):
        px = _pixels(n_prongs, seed).to(dev)
        got = graphed(px)
        junk = torch.full((1 << 24,), float("nan"), device=dev)     # whatever the allocator hands out next is poisoned
        del junk
        if i == 2:
            with torch.no_grad():
                net.forward_sparse(big)                              # a larger batched forward: workspaces grow again
        ref = plain(px)
        for g, r in zip(got, ref):
            assert torch.isfinite(g).all()
            assert torch.equal(g, r), (i, n_prongs)


def test_custom_op_runs_and_exports(dev):
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES).to(dev).eval()
    clf = EventClassifier(net, "pid", use_graph=False)
    mod = ExportableEventClassifier(clf)
    px = _pixels(3, 21).to(dev)
    a = mod(px)
    b = clf(px)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    ep = torch.export.export(mod, (px,))
    assert "tcvn.classify_event" in str(ep.graph)
    c = ep.module()(px)
    assert all(torch.equal(x, y) for x, y in zip(c, b))


def test_event_classifier_matches_the_reference_golden(dev, golden_dir):
    """EventClassifier (fp32 path, CUDA-graph plan) against tests/golden/export.pt: outputs of the notebook wrapper's
    operation sequence on the UNMODIFIED reference network (oracle/make_golden_export.py); 1e-4 (north_star, fp32)."""
    import os
    g = torch.load(os.path.join(golden_dir, "export.pt"))
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(net.specs, seed=g["state_seed"], perturb=True)
    assert synth.state_checksum(state) == pytest.approx(g["state_checksum"], rel=1e-12)
    net.load_state_dict(state)
    clf = EventClassifier(net.to(dev).eval(), "combined")
    for name, c in g["events"].items():
        px = _pixels(c["n_prongs"], c["batch_seed"])
        assert int(px.long().sum()) == c["pixel_sum"]
        got = clf(px.to(dev))
        for a, k in zip(got, ("event_prob", "prong_prob", "event_features", "prong_features")):
            assert rel_err(a.cpu(), c[k]) < 1e-4, (name, k)
