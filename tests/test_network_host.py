"""Host-side logic of the drop-in network (CPU): names, load_state_dict, loud failure without CUDA."""
import json
import os

import pytest
import torch

from dune_transformercvn_b200 import lib as tl
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES
from dune_transformercvn_b200.network import NeutrinoDenseNetwork


@pytest.fixture(scope="module")
def net(tutorial_options):
    return NeutrinoDenseNetwork(tutorial_options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)


def test_state_dict_is_the_references(net, golden_dir):
    inv = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    sd = net.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in inv["keys"]]
    assert [list(v.shape) for v in sd.values()] == [s for _, s, _ in inv["keys"]]
    assert [str(v.dtype) for v in sd.values()] == [d for _, _, d in inv["keys"]]
    assert [n for n, _ in net.named_parameters()] == inv["params"]


def test_weight_decay_grouping_matches_reference_rule(net):
    """trainers/neutrino_base.py:116-128 splits on 'bias' / 'LayerNorm.weight' substrings of the names."""
    no_decay = ["bias", "LayerNorm.weight"]
    b = [n for n, _ in net.named_parameters() if any(nd in n for nd in no_decay)]
    a = [n for n, _ in net.named_parameters() if not any(nd in n for nd in no_decay)]
    assert len(b) == 314 and len(a) == 464        # counts probed on the reference


def test_load_state_dict_strict_roundtrip(net):
    state = synth.init_state(net.specs, seed=5, perturb=True)
    res = net.load_state_dict(state, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    k = "prong_embedding.event_pixel_embedding.features.dense3.layers.7.output_block.conv2.weight"
    assert torch.equal(net.state_dict()[k], state[k])


def test_submodule_surface(net):
    for name in ("prong_embedding", "encoder", "event_decoder", "prong_decoder"):
        assert callable(getattr(net, name))


def test_cpu_call_fails_loudly(net):
    net.eval()
    b = synth.make_batch(1, seed=3, prongs_per_event=[1])
    ev = torch.zeros(1, 3, 400, 280)
    with pytest.raises(tl.TcvnError):
        net(b.features, b.extra, ev, b.event_mask, ev, b.prong_mask)
    net.train()
    with pytest.raises(tl.TcvnError):          # train mode: CUDA kernels only as well
        net(b.features, b.extra, ev, b.event_mask, ev, b.prong_mask)
    with pytest.raises(NotImplementedError):   # sub-modules on their own are eval-only
        net.prong_embedding(b.features, b.extra, ev, b.event_mask, ev, b.prong_mask)
    net.eval()


def test_export_surface_traces_without_a_gpu_and_rejects_cpu_tensors(net):
    """torch.export only runs the custom op's shape function, so the exported program can be built on a CPU-only host;
    actually calling it on CPU tensors fails loudly (no fallback)."""
    import torch
    from dune_transformercvn_b200.export import EventClassifier, ExportableEventClassifier
    from dune_transformercvn_b200.lib import TcvnError
    net.eval()
    clf = EventClassifier(net, "combined")
    px = torch.zeros(4, 3, 400, 280, dtype=torch.uint8)
    ep = torch.export.export(ExportableEventClassifier(clf), (px,))
    assert "tcvn.classify_event" in str(ep.graph)
    shapes = [tuple(n.meta["val"].shape) for n in ep.graph.nodes if n.op == "call_function" and "getitem" in str(n.target)]
    assert shapes == [(4,), (3, 8), (128,), (3, 128)]
    with pytest.raises(TcvnError):
        clf(px)


def test_sdxl_network_host_surface(tutorial_options):
    """--sdxl variant: the state_dict follows diffusers' published Encoder module tree (parity unpinned, see
    oracle/restate_sdxl.py), strict load round-trips, CPU calls and train mode fail loudly."""
    import torch
    from dune_transformercvn_b200.lib import TcvnError
    from dune_transformercvn_b200.sdxl import NeutrinoSDXLNetwork
    net = NeutrinoSDXLNetwork(tutorial_options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    sd = net.state_dict()
    p = "prong_embedding.event_pixel_embedding."
    assert tuple(sd[p + "encoder.conv_in.weight"].shape) == (64, 3, 3, 3)
    assert tuple(sd[p + "encoder.down_blocks.2.resnets.0.conv_shortcut.weight"].shape) == (128, 64, 1, 1)
    assert (p + "encoder.down_blocks.8.downsamplers.0.conv.weight") not in sd
    assert tuple(sd[p + "encoder.down_blocks.8.resnets.0.conv1.weight"].shape) == (288, 512, 3, 3)
    assert tuple(sd[p + "encoder.mid_block.attentions.0.to_out.0.weight"].shape) == (288, 288)
    assert tuple(sd[p + "output_layer.1.bias"].shape) == (288,)
    assert "encoder.encoder.layers.5.linear2.weight" in sd and "prong_decoder.output_layer.weight" in sd
    net.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    net.eval()
    z = torch.zeros(1, 3, 400, 280)
    with pytest.raises(TcvnError):
        net(torch.zeros(1, 1, 1), torch.zeros(1, 1), z, torch.ones(1, 1, dtype=torch.bool), z, torch.ones(1, 1, dtype=torch.bool))
    net.train()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 1, 1), torch.zeros(1, 1), z, torch.ones(1, 1, dtype=torch.bool), z, torch.ones(1, 1, dtype=torch.bool))


def test_loss_metrics_and_prefetcher_fail_loudly_without_cuda(tutorial_options):
    """The device-side loss / metrics / staging helpers have no CPU implementation either."""
    import torch
    from dune_transformercvn_b200 import loss as tloss
    from dune_transformercvn_b200.ingest import Prefetcher
    from dune_transformercvn_b200.lib import TcvnError
    ev, pr = torch.zeros(2, 4), torch.zeros(2, 3, 8)
    ev_t, pr_t = torch.zeros(2, dtype=torch.long), torch.zeros(2, 3, dtype=torch.long)
    with pytest.raises(TcvnError):
        tloss.training_loss(ev, pr, ev_t, pr_t, tutorial_options)
    with pytest.raises(TcvnError):
        tloss.DeviceMetrics().update(ev, pr, ev_t, pr_t)
    with pytest.raises(TcvnError):
        tloss.DeviceMetrics().compute()
    with pytest.raises(TcvnError):
        Prefetcher("cpu")
    assert tloss.LOSS_FIELDS[:3] == ("train_loss", "event_loss", "prong_loss")
