"""TEST INFRASTRUCTURE - golden vectors for the export surface: runs the forward of the reference's
`DynamicCombinedNetwork` wrapper (CreateCompiled.ipynb cell 8 - the notebook's operation sequence, on the UNMODIFIED
reference network modules) for single events given as (1 + Npng, 3, 400, 280) maps with 0..255 values, and freezes the
COO form of the inputs and the four outputs in tests/golden/export.pt.  Build container only (needs /root/reference)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import synth  # noqa: E402
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES  # noqa: E402
from dune_transformercvn_b200.params import network_specs  # noqa: E402
from oracle import reference_import  # noqa: E402
from oracle.make_golden import build_reference  # noqa: E402

H, W = 400, 280


def combined_forward(network, pixels, num_features=1, num_extra=1):
    """CreateCompiled.ipynb cell 8 (`DynamicCombinedNetwork.forward`, log_pixels False), operation by operation."""
    pixels = pixels.float() / 255
    pixels = pixels.reshape(-1, 3, H, W)
    num_images = pixels.shape[0]
    mask = torch.ones(num_images, dtype=torch.bool)
    features = torch.zeros(1, num_images - 1, num_features, dtype=pixels.dtype)
    extra = torch.zeros(1, num_extra, dtype=pixels.dtype)
    event_pixels, prong_pixels = pixels[:1], pixels[1:]
    event_mask, prong_mask = mask[:1], mask[1:]
    combined_embeddings, combined_mask = network.prong_embedding(features, extra, event_pixels, event_mask.unsqueeze(0),
                                                                 prong_pixels, prong_mask.unsqueeze(0))
    hidden_features, padding_mask, sequence_mask = network.encoder(combined_embeddings, combined_mask)
    event_features, prong_features = hidden_features[0], hidden_features[1:]
    event = network.event_decoder(event_features)
    prongs = network.prong_decoder(prong_features).transpose(0, 1)
    event_features, prong_features = event_features[0], prong_features[:, 0]
    event = torch.softmax(event[0], 0)
    prongs = torch.softmax(prongs[0], 1)
    return event, prongs, event_features, prong_features


def main():
    ref = reference_import.load()
    options = ref.tutorial_options()
    specs = network_specs(options, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(specs, seed=3, perturb=True)
    net = build_reference(ref, options, state).eval()
    cases = {"state_seed": 3, "state_checksum": synth.state_checksum(state), "events": {}}
    for name, n_prongs, seed in (("six_prongs", 6, 11), ("one_prong", 1, 12)):
        batch = synth.make_batch(1, seed=seed, prongs_per_event=[n_prongs])
        ev = ref.sparse_to_dense(batch.event_values, batch.event_coords, (H, W))     # raw 0..255 values, as LArSoft sends them
        pr = ref.sparse_to_dense(batch.prong_values, batch.prong_coords, (H, W))
        pixels = torch.cat((ev, pr)).to(torch.uint8)
        with torch.no_grad():
            out = combined_forward(net, pixels)
        cases["events"][name] = {"batch_seed": seed, "n_prongs": n_prongs, "event_prob": out[0], "prong_prob": out[1],
                                 "event_features": out[2], "prong_features": out[3].contiguous(),
                                 "pixel_sum": int(pixels.long().sum())}
        print(name, [tuple(t.shape) for t in out], out[0].tolist())
    torch.save(cases, os.path.join(ROOT, "tests", "golden", "export.pt"))


if __name__ == "__main__":
    main()
