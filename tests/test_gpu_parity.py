"""GPU parity (run on the B200 box: pytest -m gpu): CUDA path through the C ABI vs the CPU oracle.

Tolerances: densify is bit-exact; the fp32 path must match the oracle's logits within 1e-4
relative (max|d|/max|ref|, BASELINE.json north_star); the bf16 path within 2e-2.
"""
import os

import pytest
import torch

from conftest import rel_err
from dune_transformercvn_b200 import lib as tl
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.ingest import densify, sparse_to_dense
from dune_transformercvn_b200.network import NeutrinoDenseNetwork
from oracle import restate

pytestmark = pytest.mark.gpu
H, W = 400, 280
FP32_TOL = 1e-4
BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _net(state_seed, perturb, dev, precision="fp32"):
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=precision)
    state = synth.init_state(net.specs, seed=state_seed, perturb=perturb)
    net.load_state_dict(state, strict=True)
    return net.to(dev).eval(), state, opts


# ------------------------------------------------------------------------------------------ ingest
def test_densify_matches_golden_bit_exact(golden_dir, dev):
    cases = torch.load(os.path.join(golden_dir, "densify.pt"))
    for tag, c in cases.items():
        dense = densify(c["values"].to(dev), c["coords"].to(dev), (H, W), divisor=255.0).cpu()
        assert list(dense.shape) == c["shape"]
        nz = dense.nonzero()
        assert torch.equal(nz.to(torch.int32), c["nz_index"])
        assert torch.equal(dense[tuple(nz.t())], c["nz_value"])
        # the drop-in signature (caller already divided, as preprocess_pixels does)
        d2 = sparse_to_dense((c["values"] / 255.0).to(dev), c["coords"].to(dev), (H, W)).cpu()
        assert torch.equal(d2, dense)


@pytest.mark.parametrize("dtype", ["f32", "u8"])
def test_densify_matches_oracle_bit_exact(dev, dtype):
    import numpy as np
    b = synth.make_batch(5, seed=77, value_dtype=np.uint8 if dtype == "u8" else np.float32)
    for vals, coords, n in ((b.event_values, b.event_coords, b.num_events), (b.prong_values, b.prong_coords, b.num_prongs)):
        want = restate.densify(restate.preprocess_values(vals.float()), coords, H, W)
        got = densify(vals.to(dev), coords.to(dev), (H, W), num_images=n, divisor=255.0).cpu()
        assert torch.equal(got, want)


def test_densify_edge_cases(dev):
    # empty hit list with a known image count -> all zeros; single hit in the last pixel; trailing empty images
    z = densify(torch.zeros(0, 3, device=dev), torch.zeros(0, 3, dtype=torch.int32, device=dev), (H, W), num_images=2)
    assert z.shape == (2, 3, H, W) and float(z.abs().sum()) == 0.0
    coords = torch.tensor([[0, 0, 0], [1, H - 1, W - 1]], dtype=torch.int32, device=dev)
    vals = torch.tensor([[255.0, 1.0, 2.0], [3.0, 4.0, 5.0]], device=dev)
    d = densify(vals, coords, (H, W), num_images=4, divisor=255.0).cpu()
    assert d.shape == (4, 3, H, W) and int((d != 0).sum()) == 6
    assert d[0, 0, 0, 0] == 1.0 and d[1, 2, H - 1, W - 1] == torch.tensor(5.0) / 255.0
    assert float(d[2:].abs().sum()) == 0.0
    # round trip at full size: sum of the dense map == sum of the scaled hit values, in fp64
    b = synth.make_batch(64, seed=5)
    dd = densify(b.prong_values.to(dev), b.prong_coords.to(dev), (H, W), num_images=b.num_prongs, divisor=255.0)
    assert dd.shape[0] == b.num_prongs
    assert float(dd.double().sum()) == pytest.approx(float((b.prong_values / 255.0).double().sum()), rel=1e-12)
    assert int((dd != 0).sum()) == b.prong_values.numel()


# ------------------------------------------------------------------------------------------ forward
@pytest.mark.parametrize("tag", ["default", "perturbed"])
def test_fp32_forward_matches_golden_and_oracle(golden_dir, dev, tag):
    g = torch.load(os.path.join(golden_dir, "forward_eval.pt"))[tag]
    net, state, opts = _net(g["seed"], g["perturb"], dev)
    assert synth.state_checksum(state) == pytest.approx(g["state_checksum"], rel=1e-12)
    batch = synth.make_batch(2, seed=g["batch_seed"], prongs_per_event=g["prongs"])
    gb = batch.to(dev)
    with torch.no_grad():
        ev_logits, pr_logits = net.forward_sparse(gb)
    assert rel_err(ev_logits.cpu(), g["event_logits"]) < FP32_TOL
    assert rel_err(pr_logits.cpu(), g["prong_logits"]) < FP32_TOL
    # per-stage: CNN feature maps, embeddings, tokens, encoder output
    ev = densify(gb.event_values, gb.event_coords, (H, W), batch.num_events, 255.0)
    pr = densify(gb.prong_values, gb.prong_coords, (H, W), batch.num_prongs, 255.0)
    eng = net.engine
    pr_emb = eng.cnn("prong", pr, tl.TCVN_FP32)
    taps = {}
    with torch.no_grad():
        restate.sparse_forward(state, opts, batch, taps=taps)
    names = ["stem_pool"]
    for b in range(5):
        names.append(f"dense{b + 1}")
        if b < 4:
            names.append(f"transition{b + 1}")
    for stage, name in enumerate(names):
        got = eng.read_stage("prong", batch.num_prongs, stage, tl.TCVN_FP32, dev).cpu()
        want = taps["prong_cnn"][name]
        assert got.shape == want.shape, name
        assert rel_err(got, want) < FP32_TOL, name
    assert rel_err(pr_emb.cpu(), g["prong_embedding"]) < FP32_TOL
    ev_emb = eng.cnn("event", ev, tl.TCVN_FP32)
    assert rel_err(ev_emb.cpu(), g["event_embedding"]) < FP32_TOL
    with torch.no_grad():
        tokens, mask = net.prong_embedding(gb.features, gb.extra, ev, gb.event_mask, pr, gb.prong_mask)
        hidden = net.encoder(tokens, mask)[0]
        ev2 = net.event_decoder(hidden[0])
        pr2 = net.prong_decoder(hidden[1:])
    assert rel_err(tokens.cpu(), g["tokens"]) < FP32_TOL
    assert torch.equal(mask.cpu(), torch.cat((batch.event_mask, batch.prong_mask), 1))
    assert rel_err(hidden.cpu(), g["hidden"]) < FP32_TOL
    assert rel_err(ev2.cpu(), g["event_logits"]) < FP32_TOL
    assert rel_err(pr2.transpose(0, 1).cpu(), g["prong_logits"]) < FP32_TOL   # (L,B,C) like the reference module


def test_fp32_forward_ragged_batch_vs_oracle(dev):
    """Ragged prong counts incl. L=1 events and a chunk boundary (> 32 prong images), padded slots constant."""
    net, state, opts = _net(3, True, dev)
    batch = synth.make_batch(6, seed=99, prongs_per_event=[10, 1, 7, 9, 3, 8])
    with torch.no_grad():
        ev, pr = net.forward_sparse(batch.to(dev))
        want_ev, want_pr = restate.sparse_forward(state, opts, batch)
    assert rel_err(ev.cpu(), want_ev) < FP32_TOL
    assert rel_err(pr.cpu(), want_pr) < FP32_TOL
    assert (ev.argmax(-1).cpu() == want_ev.argmax(-1)).all()
    # padded slots all carry the head's response to a zero vector (Evaluate.ipynb cell 19 behaviour)
    pad = pr.cpu()[~batch.prong_mask]
    assert pad.shape[0] > 1 and float((pad - pad[0]).abs().max()) == 0.0


def test_general_mask_and_single_event(dev):
    """Non-prefix masks pack in (event, slot) order like masked_pack_1d (packed_data.py:59-66)."""
    net, state, opts = _net(4, True, dev)
    batch = synth.make_batch(2, seed=5, prongs_per_event=[2, 2])
    mask = torch.tensor([[True, False, True], [False, True, True]])
    feats = torch.zeros(2, 3, 1)
    ev = restate.densify(restate.preprocess_values(batch.event_values), batch.event_coords, H, W)
    pr = restate.densify(restate.preprocess_values(batch.prong_values), batch.prong_coords, H, W)
    with torch.no_grad():
        want_ev, want_pr = restate.network_forward(state, opts, ev, batch.event_mask, pr, mask)
        got_ev, got_pr = net(feats.to(dev), batch.extra.to(dev), ev.to(dev), batch.event_mask.to(dev), pr.to(dev),
                             mask.to(dev))
    assert rel_err(got_ev.cpu(), want_ev) < FP32_TOL
    assert rel_err(got_pr.cpu(), want_pr) < FP32_TOL


# ------------------------------------------------------------------------------------------ bf16 tcgen05 path
def test_bf16_forward_vs_oracle_and_fp32_path(dev):
    """bf16 storage + tcgen05 MMA (fp32 accumulate): logits within 2e-2 of the fp32 oracle, same top-1."""
    net, state, opts = _net(1, True, dev, precision="bf16")
    batch = synth.make_batch(6, seed=42, prongs_per_event=[10, 1, 7, 9, 3, 8])
    gb = batch.to(dev)
    with torch.no_grad():
        ev, pr = net.forward_sparse(gb)
        want_ev, want_pr = restate.sparse_forward(state, opts, batch)
    assert not torch.isnan(ev).any() and not torch.isnan(pr).any()
    assert rel_err(ev.cpu(), want_ev) < BF16_TOL
    assert rel_err(pr.cpu(), want_pr) < BF16_TOL
    assert (ev.argmax(-1).cpu() == want_ev.argmax(-1)).all()
    valid = batch.prong_mask
    assert (pr.argmax(-1).cpu()[valid] == want_pr.argmax(-1)[valid]).float().mean() >= 0.999
    # run-to-run determinism of the tensor-core path
    with torch.no_grad():
        ev2, pr2 = net.forward_sparse(gb)
    assert torch.equal(ev, ev2) and torch.equal(pr, pr2)


def test_bf16_feature_maps_stage_by_stage(dev):
    """Every dense block / transition output of the tcgen05 path against the oracle (bf16 rounding level)."""
    net, state, opts = _net(1, True, dev, precision="bf16")
    batch = synth.make_batch(2, seed=21, prongs_per_event=[3, 1])
    gb = batch.to(dev)
    taps = {}
    with torch.no_grad():
        restate.sparse_forward(state, opts, batch, taps=taps)
    pr = densify(gb.prong_values, gb.prong_coords, (H, W), batch.num_prongs, 255.0)
    eng = net.engine
    eng.ensure_packed(tl.TCVN_BF16)
    emb = eng.cnn("prong", pr, tl.TCVN_BF16)
    names = ["stem_pool"]
    for b in range(5):
        names.append(f"dense{b + 1}")
        if b < 4:
            names.append(f"transition{b + 1}")
    for stage, name in enumerate(names):
        got = eng.read_stage("prong", batch.num_prongs, stage, tl.TCVN_BF16, dev).cpu()
        assert rel_err(got, taps["prong_cnn"][name]) < BF16_TOL, name
    assert rel_err(emb.cpu(), taps["prong_embedding"]) < BF16_TOL


def test_bf16_linearity_of_conv_tiles_at_full_size(dev):
    """Size-independent property at BASELINE size (256 events): permuting the events permutes the logits
    exactly (tiles, chunks and the haloed 3x3 loads never mix images)."""
    net, state, opts = _net(0, False, dev, precision="bf16")
    batch = synth.make_batch(256, seed=1234)
    gb = batch.to(dev)
    with torch.no_grad():
        ev, pr = net.forward_sparse(gb)
        # same events, reversed order
        ev_px = densify(gb.event_values, gb.event_coords, (H, W), batch.num_events, 255.0)
        pr_px = densify(gb.prong_values, gb.prong_coords, (H, W), batch.num_prongs, 255.0)
        counts = torch.tensor(batch.prongs_per_event)
        starts = torch.cumsum(counts, 0) - counts
        order = torch.arange(255, -1, -1)
        pr_idx = torch.cat([torch.arange(int(starts[i]), int(starts[i] + counts[i])) for i in order]).to(dev)
        ev2, pr2 = net(gb.features[order.to(dev)], gb.extra[order.to(dev)], ev_px[order.to(dev)],
                       gb.event_mask[order.to(dev)], pr_px[pr_idx], gb.prong_mask[order.to(dev)])
    assert torch.isfinite(ev).all() and torch.isfinite(pr).all()
    assert torch.equal(ev2.flip(0), ev)
    assert torch.equal(pr2.flip(0), pr)


# ------------------------------------------------------------------------------------------ COO-direct stem
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sparse_direct_stem_equals_densify_then_forward(dev, precision):
    """The stem fed from the hit list must give what densify -> dense forward gives (same arithmetic, same
    summation order for (y,x)-sorted hits: bit-identical), for f32 and u8 hit values, incl. an image whose
    only hit is the dummy (0,0) hit of an empty prong (fully_sparse_dataset.py:148-153)."""
    import numpy as np
    net, state, opts = _net(2, True, dev, precision=precision)
    for vdt in (np.float32, np.uint8):
        batch = synth.make_batch(5, seed=8, prongs_per_event=[4, 1, 6, 2, 3], value_dtype=vdt)
        # make prong image 2 "empty": a single zero-valued dummy hit at (0,0)
        pc, pv = batch.prong_coords.clone(), batch.prong_values.clone()
        keep = pc[:, 0] != 2
        dummy_c = torch.tensor([[2, 0, 0]], dtype=torch.int32)
        dummy_v = torch.zeros(1, 3, dtype=pv.dtype)
        first = int((pc[:, 0] < 2).sum())
        pc = torch.cat((pc[keep][:first], dummy_c, pc[keep][first:]))
        pv = torch.cat((pv[keep][:first], dummy_v, pv[keep][first:]))
        batch.prong_coords, batch.prong_values = pc, pv
        gb = batch.to(dev)
        with torch.no_grad():
            ev_a, pr_a = net.forward_sparse(gb)
            ev_b, pr_b = net.forward_sparse(gb, materialize=True)
        assert torch.equal(ev_a, ev_b) and torch.equal(pr_a, pr_b)
        if precision == "fp32" and vdt is np.float32:
            with torch.no_grad():
                want_ev, want_pr = restate.sparse_forward(state, opts, batch)
            assert rel_err(ev_a.cpu(), want_ev) < FP32_TOL and rel_err(pr_a.cpu(), want_pr) < FP32_TOL


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sparse_direct_stem_dense_windows_and_borders(dev, precision):
    """The warp-per-tile stem (csrc/stem_coo.cu) on inputs that leave its sparse fast paths: windows holding hundreds of
    hits (record lists longer than one 32-lane batch), a completely filled image, hits on all four borders and corners, and
    hits outside the map (dropped, as by the densify kernel).  Bit-identical to densify -> dense-window stem."""
    net, state, opts = _net(2, True, dev, precision=precision)
    batch = synth.make_batch(2, seed=21, prongs_per_event=[2, 2], event_occupancy=0.55, prong_occupancy=0.25)
    # prong image 1: every pixel is a hit; prong image 3: only border / corner hits plus two coordinates outside the map
    pc, pv = batch.prong_coords, batch.prong_values
    keep = (pc[:, 0] != 1) & (pc[:, 0] != 3)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.int32), torch.arange(W, dtype=torch.int32), indexing="ij")
    full_c = torch.stack((torch.full_like(yy, 1), yy, xx), -1).reshape(-1, 3)
    g = torch.Generator().manual_seed(3)
    full_v = torch.rand(full_c.shape[0], 3, generator=g).to(pv.dtype)
    edge = [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (0, 137), (H - 1, 5), (200, 0), (17, W - 1), (1, 1), (H, 3), (5, W)]
    edge_c = torch.tensor([[3, y, x] for (y, x) in sorted(edge)], dtype=torch.int32)
    edge_v = torch.rand(edge_c.shape[0], 3, generator=g).to(pv.dtype)
    parts_c = [pc[keep & (pc[:, 0] < 1)], full_c, pc[keep & (pc[:, 0] == 2)], edge_c]
    parts_v = [pv[keep & (pc[:, 0] < 1)], full_v, pv[keep & (pc[:, 0] == 2)], edge_v]
    batch.prong_coords, batch.prong_values = torch.cat(parts_c), torch.cat(parts_v)
    gb = batch.to(dev)
    with torch.no_grad():
        ev_a, pr_a = net.forward_sparse(gb)
        # the dense maps for the comparison are built without the two out-of-map hits (sparse_to_dense would raise on them)
        inside = (batch.prong_coords[:, 1] < H) & (batch.prong_coords[:, 2] < W)
        batch.prong_coords, batch.prong_values = batch.prong_coords[inside], batch.prong_values[inside]
        ev_b, pr_b = net.forward_sparse(batch.to(dev), materialize=True)
    assert torch.isfinite(ev_a).all() and torch.isfinite(pr_a).all()
    assert torch.equal(ev_a, ev_b) and torch.equal(pr_a, pr_b)


# ------------------------------------------------------------------------------------------ config 5 shape
@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_max_prong_count_ingest_and_forward(dev, precision, tol):
    """BASELINE configs[4]: every event carries the dataset's maximum of 20 prongs (21-token sequences), sparse
    Minkowski hits densified on the GPU and forwarded end to end; also the literal densify -> dense forward sequence."""
    net, state, opts = _net(6, True, dev, precision=precision)
    batch = synth.make_batch(2, seed=17, prongs_per_event=[20, 20])
    gb = batch.to(dev)
    with torch.no_grad():
        ev, pr = net.forward_sparse(gb)
        ev_m, pr_m = net.forward_sparse(gb, materialize=True)
        want_ev, want_pr = restate.sparse_forward(state, opts, batch)
    assert pr.shape == (2, 20, NUM_PRONG_CLASSES)
    assert rel_err(ev.cpu(), want_ev) < tol and rel_err(pr.cpu(), want_pr) < tol
    assert torch.equal(ev, ev_m) and torch.equal(pr, pr_m)          # COO-direct stem == densify kernel + dense stem
    assert (ev.argmax(-1).cpu() == want_ev.argmax(-1)).all()
    assert (pr.argmax(-1).cpu() == want_pr.argmax(-1)).all()


# ------------------------------------------------------------------------------------------ collate
def test_collate_bit_exact_vs_golden_and_oracle(golden_dir, dev):
    from dune_transformercvn_b200.ingest import collate_sparse
    cases = torch.load(os.path.join(golden_dir, "collate.pt"))
    for name, c in cases.items():
        got_c, got_v = collate_sparse([t.to(dev) for t in c["coords"]], [t.to(dev) for t in c["values"]],
                                      [t.to(dev) for t in c["masks"]])
        assert got_c.dtype == torch.int32
        assert torch.equal(got_c.cpu(), c["out_coords"]), name
        assert torch.equal(got_v.cpu(), c["out_values"]), name
    # a full-size batch: collate -> densify equals densify of the already-global list the generator makes
    b = synth.make_batch(64, seed=9)
    starts = torch.cumsum(torch.tensor([0] + b.prongs_per_event), 0)
    per_event, vals, masks = [], [], []
    img = b.prong_coords[:, 0].long()
    for e, p in enumerate(b.prongs_per_event):
        sel = (img >= starts[e]) & (img < starts[e + 1])
        ce = b.prong_coords[sel].clone()
        ce[:, 0] -= int(starts[e])
        per_event.append(ce.to(dev)); vals.append(b.prong_values[sel].to(dev)); masks.append(b.prong_mask[e].to(dev))
    got_c, got_v = collate_sparse(per_event, vals, masks)
    assert torch.equal(got_c.cpu(), b.prong_coords) and torch.equal(got_v.cpu(), b.prong_values)
    want_c, _ = restate.collate_sparse([t.cpu() for t in per_event], [t.cpu() for t in vals], [t.cpu() for t in masks])
    assert torch.equal(got_c.cpu(), want_c)


def test_bf16_top1_agreement_at_full_size(dev):
    """BASELINE configs[1] size (256 events, 1679 maps): bf16 tcgen05 path vs the fp32 path of this library: logits within
    2e-2, top-1 event / prong class agreement >= 99.9 % (north_star).  Measured: 1e-4 / 100 % (scripts/gpu_top1.py)."""
    opts = PathOptions.tutorial()
    batch = synth.make_batch(256, seed=1234, max_prongs=10).to(dev)
    outs = {}
    for prec in ("fp32", "bf16"):
        net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision=prec)
        net.load_state_dict(synth.init_state(net.specs, seed=1, perturb=True))
        net = net.to(dev).eval()
        with torch.no_grad():
            outs[prec] = net.forward_sparse(batch)
    (ev32, pr32), (ev16, pr16) = outs["fp32"], outs["bf16"]
    m = batch.prong_mask
    assert rel_err(ev16, ev32) < BF16_TOL and rel_err(pr16[m], pr32[m]) < BF16_TOL
    assert float((ev16.argmax(-1) == ev32.argmax(-1)).float().mean()) >= 0.999
    assert float((pr16.argmax(-1) == pr32.argmax(-1))[m].float().mean()) >= 0.999


def test_prefetcher_stages_batches_in_order(dev):
    """ingest.Prefetcher: batches copied on the copy stream come out in submission order and give bit-identical logits
    to batches copied on the compute stream."""
    from dune_transformercvn_b200.ingest import Prefetcher
    net, state, opts = _net(3, True, dev, precision="bf16")
    batches = [synth.make_batch(2, seed=40 + i, prongs_per_event=[1 + i, 2]).pin() for i in range(3)]
    with torch.no_grad():
        want = [tuple(t.clone() for t in net.forward_sparse(b.to(dev))) for b in batches]
        pf = Prefetcher(dev)
        pf.submit(batches[0])
        got = []
        for i in range(3):
            if i + 1 < 3:
                pf.submit(batches[i + 1])
            got.append(tuple(t.clone() for t in net.forward_sparse(pf.take())))
    for w, g in zip(want, got):
        assert torch.equal(w[0], g[0]) and torch.equal(w[1], g[1])
    with pytest.raises(tl.TcvnError):
        pf.take()


def test_sm_budgets_do_not_change_the_result(dev):
    """tcvn_set_sm_limit only sizes the persistent grids: logits with disjoint SM budgets for the two CNN streams are
    bit-identical to the default whole-chip launch (tile -> CTA assignment never enters the arithmetic)."""
    net, state, opts = _net(3, True, dev, precision="bf16")
    batch = synth.make_batch(6, seed=77, max_prongs=6).to(dev)
    with torch.no_grad():
        net.partition_sms = False
        a = tuple(t.clone() for t in net.forward_sparse(batch))
        net.partition_sms = True
        b = tuple(t.clone() for t in net.forward_sparse(batch))
        net.partition_sms = False
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    L = tl.load()
    assert L.tcvn_set_sm_limit(-1) != 0 and L.tcvn_set_sm_limit(0) == 0


def test_batched_inference_graph_replay_is_exact(dev):
    """forward_sparse(graph=True): one CUDA-graph replay per batch shape must give the bits of the kernel-by-kernel launch,
    for new data of the same shape, after a parameter change (re-pack -> stale plan re-captured), and for a second shape."""
    opts = PathOptions.tutorial()
    net = NeutrinoDenseNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    net.load_state_dict(synth.init_state(net.specs, seed=4, perturb=True))
    net = net.to(dev).eval()
    with torch.no_grad():
        for seed, events in ((1, 6), (2, 6), (3, 9), (4, 6)):
            b = synth.make_batch(events, seed=seed, prongs_per_event=[3, 1, 4, 2, 5, 2, 1, 1, 3][:events]).to(dev)
            want = net.forward_sparse(b)
            got = net.forward_sparse(b, graph=True)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]), (seed, events)
        assert len(net._plans) == 2 and net.graph_launches_per_replay > 50
        for p in net.parameters():                       # what an optimizer step or load_state_dict does
            p.mul_(1.01)
        want = net.forward_sparse(b)
        got = net.forward_sparse(b, graph=True)
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def test_training_pixel_noise_in_densify(dev):
    """preprocess_pixels' training noise (neutrino_full_dense_trainer.py:62-65: v *= 1 + randn * std) fused into densify:
    std = 0 is bit-identical to the plain kernel, the same seed repeats, and the multiplicative factors are standard
    normal draws (mean 0, variance 1 within sampling error, no value touched where there is no hit)."""
    from dune_transformercvn_b200.ingest import densify
    b = synth.make_batch(4, seed=9, max_prongs=3).to(dev)
    n = b.num_events
    base = densify(b.event_values, b.event_coords, (H, W), n, 255.0)
    assert torch.equal(densify(b.event_values, b.event_coords, (H, W), n, 255.0, noise_std=0.0, seed=5), base)
    std = 0.05
    a1 = densify(b.event_values, b.event_coords, (H, W), n, 255.0, noise_std=std, seed=5)
    a2 = densify(b.event_values, b.event_coords, (H, W), n, 255.0, noise_std=std, seed=5)
    a3 = densify(b.event_values, b.event_coords, (H, W), n, 255.0, noise_std=std, seed=6)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    hit = base != 0
    assert torch.equal(a1 == 0, ~hit)
    z = ((a1[hit] / base[hit]).double() - 1.0) / std
    m = z.numel()
    assert m > 10000
    assert abs(float(z.mean())) < 5.0 / m ** 0.5 and abs(float(z.var()) - 1.0) < 8.0 * (2.0 / m) ** 0.5
    assert 3.0 < float(z.abs().max()) < 6.5
