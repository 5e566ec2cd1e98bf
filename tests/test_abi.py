"""The C-ABI library loads and exports every symbol include/tcvn.h declares (CPU; no compute calls)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from dune_transformercvn_b200 import lib as tl
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.params import arena_offsets, network_specs


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "tcvn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcvn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = tl.load()
    declared = _header_symbols()
    assert declared, "no declarations parsed from include/tcvn.h"
    assert sorted(tl.EXPORTS) == declared
    for name in declared:
        assert hasattr(L, name), name
    assert L.tcvn_abi_version() == 2


def test_host_side_size_queries_match_python_inventory():
    """tcvn_cnn_arena_floats walks the DenseNet in C++; it must agree with params.py (pinned to the reference)."""
    L = tl.load()
    opts = PathOptions.tutorial()
    specs = network_specs(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    for prefix, width in (("prong_embedding.prong_pixel_embedding.", 256), ("prong_embedding.event_pixel_embedding.", 288)):
        d = tl.CnnDesc()
        d.in_channels, d.init_features, d.growth, d.bn_size, d.num_blocks = 3, 64, 32, 4, 5
        for i, n in enumerate((3, 6, 12, 6, 3)):
            d.block_layers[i] = n
        d.out_features, d.height, d.width, d.bn_eps = width, 400, 280, 1e-5
        want = sum(s.numel for s in specs if s.name.startswith(prefix) and s.in_arena)
        assert L.tcvn_cnn_arena_floats(C.byref(d)) == want
        for prec in (tl.TCVN_FP32, tl.TCVN_BF16):
            assert L.tcvn_cnn_packed_bytes(C.byref(d), prec) > want
            assert L.tcvn_cnn_workspace_bytes(C.byref(d), prec, 4) > 4 * 3 * 400 * 280
    bad = tl.CnnDesc()
    assert L.tcvn_cnn_arena_floats(C.byref(bad)) == -1
    assert b"descriptor" in L.tcvn_last_error()


def test_eval_workspace_uses_line_aligned_block_rows():
    """The eval walk keeps every dense-block buffer on a row pitch of whole 128-byte lines (plan.h: BlockPlan::ld = channels
    rounded up to 64; csrc/umma.cu: a 64-channel TMA box row is then exactly one line).  Host-only size query: the bf16
    workspace for n images must hold, per image, the five block buffers at that pitch plus the 128-channel bottleneck map of
    the largest block - and not much more."""
    L = tl.load()
    d = tl.CnnDesc()
    d.in_channels, d.init_features, d.growth, d.bn_size, d.num_blocks = 3, 64, 32, 4, 5
    layers = (3, 6, 12, 6, 3)
    for i, n in enumerate(layers):
        d.block_layers[i] = n
    d.out_features, d.height, d.width, d.bn_eps = 256, 400, 280, 1e-5
    n = 8
    h, w = ((400 - 1) // 2 + 1 - 3) // 2 + 1, ((280 - 1) // 2 + 1 - 3) // 2 + 1   # stem conv s2, then AvgPool(3, 2): 99 x 69
    c, need, padded_only = 64, 0, 0
    rows0 = (h + 2) * (w + 2)
    for b, nl in enumerate(layers):
        c0p = (c + 7) // 8 * 8
        ctot = c0p + nl * 32
        ld = (ctot + 63) // 64 * 64
        assert ld % 64 == 0 and ld >= ctot
        rows = (h + 2) * (w + 2)
        need += n * rows * ld * 2
        padded_only += n * rows * (ld - ctot) * 2
        c = (c + nl * 32) // 2
        h, w = h // 2, w // 2
    need += n * rows0 * 128 * 2                     # bottleneck map of block 1
    got = L.tcvn_cnn_workspace_bytes(C.byref(d), tl.TCVN_BF16, n)
    assert got >= need, (got, need)
    assert got < need * 1.5, (got, need)             # pooled scratch, gap vector, hit offsets: small next to the maps
    assert padded_only > 0                           # the tutorial network does have unaligned channel totals (160, 272, ...)


def test_bad_arguments_are_reported_not_crashed():
    L = tl.load()
    rc = L.tcvn_densify(None, None, 0, 5, 3, 2, 400, 280, C.c_float(255.0), None, 0, None)
    assert rc == -1 and b"null" in L.tcvn_last_error()
    rc = L.tcvn_densify(None, None, 0, 0, 3, 0, 400, 280, C.c_float(255.0), None, 0, None)
    assert rc == 0   # empty batch is a no-op
