"""TEST INFRASTRUCTURE — golden vectors for the collate row: runs the UNMODIFIED reference staticmethod
MinkowskiCollection.collate_sparse (transformercvn/dataset/minkowski_dataset.py:34-47) on seeded per-event hit lists and
freezes inputs + outputs in tests/golden/collate.pt.  Build container only (needs /root/reference)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_import  # noqa: E402


def make_events(seed, prongs_per_event, max_slots, hits_lo=0, hits_hi=40):
    g = torch.Generator().manual_seed(seed)
    coords, values, masks = [], [], []
    for p in prongs_per_event:
        per_image = torch.randint(hits_lo, hits_hi, (p,), generator=g)
        c = []
        for img, n in enumerate(per_image.tolist()):
            yx = torch.stack((torch.randint(0, 400, (n,), generator=g), torch.randint(0, 280, (n,), generator=g)), 1)
            c.append(torch.cat((torch.full((n, 1), img), yx), 1))
        c = torch.cat(c).to(torch.int32) if c else torch.zeros(0, 3, dtype=torch.int32)
        coords.append(c)
        values.append(torch.randint(1, 256, (c.shape[0], 3), generator=g).float())
        m = torch.zeros(max_slots, dtype=torch.bool)
        m[:p] = True
        masks.append(m)
    return coords, values, masks


def main():
    ref = reference_import.load()
    cases = {}
    for name, seed, prongs, slots, lo in (("ragged", 3, [3, 1, 10, 2, 7], 10, 1), ("with_empty_images", 4, [2, 5, 1], 6, 0),
                                          ("single_event", 5, [4], 4, 1), ("max_prongs", 6, [20, 20], 20, 1)):
        coords, values, masks = make_events(seed, prongs, slots, hits_lo=lo)
        out_c, out_v = ref.MinkowskiCollection.collate_sparse([c.clone() for c in coords], values, masks)
        cases[name] = {"coords": coords, "values": values, "masks": masks, "out_coords": out_c, "out_values": out_v}
        print(name, [int(c.shape[0]) for c in coords], "->", tuple(out_c.shape), "last image", int(out_c[-1, 0]))
    torch.save(cases, os.path.join(ROOT, "tests", "golden", "collate.pt"))


if __name__ == "__main__":
    main()
