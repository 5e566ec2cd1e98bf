"""The parameter inventory must be the reference's state_dict, name for name (CPU)."""
import json
import os

from dune_transformercvn_b200.params import network_specs, arena_offsets
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES


def _specs(opts):
    return network_specs(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)


def test_inventory_matches_reference_state_dict(golden_dir, tutorial_options):
    inv = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    specs = _specs(tutorial_options)
    assert [s.name for s in specs] == [k for k, _, _ in inv["keys"]]
    assert [list(s.shape) for s in specs] == [shape for _, shape, _ in inv["keys"]]
    assert [s.name for s in specs if s.is_param] == inv["params"]
    assert sum(s.numel for s in specs if s.is_param) == inv["num_params"] == 5723512


def test_arena_is_dense_and_ordered(tutorial_options):
    specs = _specs(tutorial_options)
    off, total = arena_offsets(specs)
    cur = 0
    for s in specs:
        if s.in_arena:
            assert off[s.name] == cur
            cur += s.numel
    assert cur == total
    assert "prong_embedding.combined_embedding.norm.num_batches_tracked" not in off
