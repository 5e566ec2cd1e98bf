"""GPU parity of the --sdxl variant (BASELINE configs[3]; dune_transformercvn_b200/sdxl.py) against
oracle/restate_sdxl.py.  PARITY UNPINNED: the reference's arithmetic lives in un-vendored, unpinned `diffusers`
(layers/sdxl_net.py:4), absent here; the oracle restates its published layout, and this test holds the CUDA path to
that restatement - fp32, 1e-4 relative (north_star's fp32 tolerance) on both pixel embeddings and on the logits."""
import pytest
import torch

from conftest import rel_err
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.sdxl import NeutrinoSDXLNetwork
from oracle import restate, restate_sdxl

pytestmark = pytest.mark.gpu
H, W = 400, 280


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def test_sdxl_forward_matches_oracle(dev):
    opts = PathOptions.tutorial()
    net = NeutrinoSDXLNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)
    state = synth.init_state(net.specs, seed=4, perturb=True)
    net.load_state_dict(state, strict=True)
    net = net.to(dev).eval()
    batch = synth.make_batch(2, seed=8, prongs_per_event=[2, 1])
    ev = restate.densify(restate.preprocess_values(batch.event_values), batch.event_coords, H, W)
    pr = restate.densify(restate.preprocess_values(batch.prong_values), batch.prong_coords, H, W)
    taps = {}
    with torch.no_grad():
        want_ev, want_pr = restate_sdxl.network_forward(state, opts, ev, batch.event_mask, pr, batch.prong_mask, taps=taps)
        eng = net.engine
        eng.ensure_packed(0)
        got_pe = eng.cnn("prong", pr.to(dev), 0)
        got_ee = eng.cnn("event", ev.to(dev), 0)
        got_ev, got_pr = net(batch.features.to(dev), batch.extra.to(dev), ev.to(dev), batch.event_mask.to(dev), pr.to(dev),
                             batch.prong_mask.to(dev))
        sp_ev, sp_pr = net.forward_sparse(batch.to(dev))
    assert tuple(got_pe.shape) == (3, 256) and tuple(got_ee.shape) == (2, 288)
    assert rel_err(got_pe.cpu(), taps["prong_embedding"]) < 1e-4
    assert rel_err(got_ee.cpu(), taps["event_embedding"]) < 1e-4
    assert rel_err(got_ev.cpu(), want_ev) < 1e-4
    assert rel_err(got_pr.cpu(), want_pr) < 1e-4
    assert torch.equal(sp_ev, got_ev) and torch.equal(sp_pr, got_pr)     # densify kernel == the oracle's dense maps


def test_sdxl_kernels_against_torch(dev):
    """GroupNorm(1 group)+SiLU and the stride-2 patch GEMM on odd sizes (25x17 -> 12x8, as block 4 of the encoder sees)."""
    import ctypes as C
    import torch.nn.functional as F
    from dune_transformercvn_b200 import lib as tl
    L = tl.load()
    st = tl.stream_ptr(dev)
    g = torch.Generator().manual_seed(0)
    n, c, h, w = 3, 16, 25, 17
    x = torch.randn(n, c, h, w, generator=g)
    ring = torch.zeros(n, h + 2, w + 2, c)
    ring[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    d_ring = ring.to(dev).reshape(-1, c).contiguous()
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    out = torch.empty_like(d_ring)
    sums = torch.empty(2 * n, dtype=torch.float64, device=dev)
    d_gamma, d_beta = gamma.to(dev), beta.to(dev)      # keep device operands alive across the raw-pointer call
    tl.check(L.tcvn_sdxl_groupnorm(tl.ptr(d_ring), n, c, 1, h, w, tl.ptr(d_gamma), tl.ptr(d_beta), 1e-6, 1,
                                   tl.ptr(out), tl.ptr(sums), st), "groupnorm")
    want = F.silu(F.group_norm(x, 1, gamma, beta, 1e-6))
    got = out.view(n, h + 2, w + 2, c)
    assert rel_err(got[:, 1:-1, 1:-1].permute(0, 3, 1, 2).cpu(), want) < 1e-5
    assert float(got[:, 0].abs().max()) == 0.0 and float(got[:, :, -1].abs().max()) == 0.0
    # Downsample2D: pad (0,1,0,1) + conv 3x3 stride 2
    co = 8
    wt, b = torch.randn(co, c, 3, 3, generator=g) * 0.1, torch.randn(co, generator=g)
    ho, wo = h // 2, w // 2
    patches = torch.empty(n * (ho + 2) * (wo + 2), 9 * c, device=dev)
    tl.check(L.tcvn_sdxl_patch_s2(tl.ptr(d_ring), n, c, h, w, tl.ptr(patches), st), "patch_s2")
    wk = wt.permute(2, 3, 1, 0).reshape(9 * c, co).contiguous().to(dev)
    y = torch.empty(n * (ho + 2) * (wo + 2), co, device=dev)
    d_b = b.to(dev)
    tl.check(L.tcvn_t_gemm(tl.ptr(patches), 9 * c, patches.shape[0], 9 * c, 1, None, tl.ptr(wk), co, None, 0, 0,
                           tl.ptr(d_b), tl.ptr(y), co, 0, ho + 2, wo + 2, 0, st), "t_gemm")
    want = F.conv2d(F.pad(x, (0, 1, 0, 1)), wt, b, stride=2)
    got = y.view(n, ho + 2, wo + 2, co)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).cpu()
    assert tuple(want.shape[2:]) == (ho, wo)
    assert rel_err(got, want) < 1e-5


def _ring16(x, dev):
    """NCHW fp32 -> ringed channels-last bf16 [n*(H+2)*(W+2)][C] on the device"""
    n, c, h, w = x.shape
    ring = torch.zeros(n, h + 2, w + 2, c)
    ring[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    return ring.reshape(-1, c).to(dev).to(torch.bfloat16).contiguous()


def _unring(y, n, h, w):
    c = y.shape[1]
    return y.float().view(n, h + 2, w + 2, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).cpu()


def test_sdxl16_tensor_core_conv_against_torch(dev):
    """tcvn_sdxl16_conv (tcgen05 shifted GEMM): a 3x3 convolution with the residual input riding as a trailing K segment
    (identity weights, and a 1x1 shortcut with a channel change), on an odd-sized map that spans several 128-row tiles and
    images; tcvn_sdxl16_groupnorm and the stride-2 patch path.  bf16 operands, fp32 accumulation: 2e-2 of the output range."""
    import torch.nn.functional as F
    from dune_transformercvn_b200 import lib as tl
    from dune_transformercvn_b200.sdxl import _taps3
    L = tl.load()
    st = tl.stream_ptr(dev)
    g = torch.Generator().manual_seed(1)
    n, h, w = 3, 25, 17
    hp, wp = h + 2, w + 2
    rows = n * hp * wp
    ones = torch.ones(1024, device=dev)
    for cin, cout in ((64, 64), (64, 128), (128, 256)):
        a = torch.randn(n, cout, h, w, generator=g)           # activated map entering conv2
        x = torch.randn(n, cin, h, w, generator=g)            # the block input (residual)
        wt = torch.randn(cout, cout, 3, 3, generator=g) * 0.05
        b = torch.randn(cout, generator=g)
        if cin == cout:
            tail, want_res = torch.eye(cout), x
        else:
            ws = torch.randn(cout, cin, generator=g) * 0.1
            tail, want_res = ws, F.conv2d(x, ws.view(cout, cin, 1, 1))
        a16, x16 = _ring16(a, dev), _ring16(x, dev)
        want = F.conv2d(a16.float().cpu().view(n, hp, wp, cout)[:, 1:-1, 1:-1].permute(0, 3, 1, 2), wt.bfloat16().float(), b, padding=1)
        want = want + (want_res.bfloat16().float() if cin == cout else
                       F.conv2d(x.bfloat16().float(), tail.bfloat16().float().view(cout, cin, 1, 1)))
        wk = torch.cat((wt.permute(0, 2, 3, 1).reshape(cout, 9 * cout), tail), dim=1)
        n_tiles = (cout + 127) // 128
        wpad = torch.zeros(n_tiles * 128, wk.shape[1])
        wpad[:cout] = wk
        bpad = torch.zeros(n_tiles * 128)
        bpad[:cout] = b
        d_w, d_b = wpad.to(dev).to(torch.bfloat16).contiguous(), bpad.to(dev)
        out = torch.empty(rows, cout, dtype=torch.bfloat16, device=dev)
        tl.check(L.tcvn_sdxl16_conv(tl.ptr(a16), rows, cout, 9, _taps3(wp), tl.ptr(x16), cin, tl.ptr(d_w), n_tiles, tl.ptr(d_b),
                                    tl.ptr(ones), tl.ptr(out), cout, hp, wp, st), "sdxl16_conv")
        got = _unring(out, n, h, w)
        assert rel_err(got, want) < 2e-2, (cin, cout, rel_err(got, want))
        ring = out.float().view(n, hp, wp, cout)
        assert float(ring[:, 0].abs().max()) == 0.0 and float(ring[:, :, -1].abs().max()) == 0.0    # zero ring rows
    # GroupNorm(1) + SiLU, bf16
    c = 64
    x = torch.randn(n, c, h, w, generator=g) * 2 + 0.5
    x16 = _ring16(x, dev)
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(dev), torch.randn(c, generator=g).to(dev)
    ws = torch.empty(L.tcvn_sdxl16_groupnorm_workspace_bytes(n), dtype=torch.uint8, device=dev)
    out = torch.empty_like(x16)
    tl.check(L.tcvn_sdxl16_groupnorm(tl.ptr(x16), n, c, h, w, tl.ptr(gamma), tl.ptr(beta), 1e-6, 1, tl.ptr(out), tl.ptr(ws), ws.numel(),
                                     st), "sdxl16_groupnorm")
    xin = x16.float().cpu().view(n, hp, wp, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    want = F.silu(F.group_norm(xin, 1, gamma.cpu(), beta.cpu(), 1e-6))
    assert rel_err(_unring(out, n, h, w), want) < 1e-2
    # Downsample2D: pad (0,1,0,1) + conv 3x3 stride 2 as patches + plain GEMM
    wt, b = torch.randn(c, c, 3, 3, generator=g) * 0.05, torch.randn(c, generator=g)
    ho, wo = h // 2, w // 2
    orows = n * (ho + 2) * (wo + 2)
    pt = torch.empty(orows, 9 * c, dtype=torch.bfloat16, device=dev)
    tl.check(L.tcvn_sdxl16_patch_s2(tl.ptr(x16), n, c, h, w, tl.ptr(pt), st), "sdxl16_patch_s2")
    wpad = torch.zeros(128, 9 * c)
    wpad[:c] = wt.permute(0, 2, 3, 1).reshape(c, 9 * c)
    bpad = torch.zeros(128)
    bpad[:c] = b
    d_w, d_b = wpad.to(dev).to(torch.bfloat16).contiguous(), bpad.to(dev)
    y = torch.empty(orows, c, dtype=torch.bfloat16, device=dev)
    tl.check(L.tcvn_sdxl16_conv(tl.ptr(pt), orows, 9 * c, 1, None, None, 0, tl.ptr(d_w), 1, tl.ptr(d_b), tl.ptr(ones), tl.ptr(y), c,
                                ho + 2, wo + 2, st), "sdxl16_conv(patches)")
    want = F.conv2d(F.pad(xin, (0, 1, 0, 1)), wt.bfloat16().float(), b, stride=2)
    assert rel_err(_unring(y, n, ho, wo), want) < 2e-2
    # the same convolution without patches: space-to-depth matrix + nine shifted, column-grouped views (the product path);
    # the K chunks reach the MMA in the same order, so the two results are the same bits
    sd = torch.empty(orows, 4 * c, dtype=torch.bfloat16, device=dev)
    tl.check(L.tcvn_sdxl16_s2d(tl.ptr(x16), n, c, h, w, tl.ptr(sd), st), "sdxl16_s2d")
    y2 = torch.empty(orows, c, dtype=torch.bfloat16, device=dev)
    tl.check(L.tcvn_sdxl16_conv_s2(tl.ptr(sd), orows, c, tl.ptr(d_w), 1, tl.ptr(d_b), tl.ptr(ones), tl.ptr(y2), c, ho + 2, wo + 2, st),
             "sdxl16_conv_s2")
    assert torch.equal(y2, y)
    # odd map sizes: the last row / column of the map is reached through the ring positions of the space-to-depth matrix
    for (h2, w2) in ((7, 9), (35, 25)):
        xo = torch.randn(n, c, h2, w2, generator=g)
        xo16 = _ring16(xo, dev)
        ho2, wo2 = h2 // 2, w2 // 2
        r2 = n * (ho2 + 2) * (wo2 + 2)
        sd = torch.empty(r2, 4 * c, dtype=torch.bfloat16, device=dev)
        tl.check(L.tcvn_sdxl16_s2d(tl.ptr(xo16), n, c, h2, w2, tl.ptr(sd), st), "sdxl16_s2d")
        y3 = torch.empty(r2, c, dtype=torch.bfloat16, device=dev)
        tl.check(L.tcvn_sdxl16_conv_s2(tl.ptr(sd), r2, c, tl.ptr(d_w), 1, tl.ptr(d_b), tl.ptr(ones), tl.ptr(y3), c, ho2 + 2, wo2 + 2, st),
                 "sdxl16_conv_s2")
        xin2 = xo16.float().cpu().view(n, h2 + 2, w2 + 2, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        want2 = F.conv2d(F.pad(xin2, (0, 1, 0, 1)), wt.bfloat16().float(), b, stride=2)
        assert rel_err(_unring(y3, n, ho2, wo2), want2) < 2e-2, (h2, w2)


def test_sdxl_bf16_forward_tracks_oracle(dev):
    """The tcgen05 walk of the --sdxl network against the (unpinned) oracle: bf16 tolerance 2e-2 on the logits (north_star's
    bf16 bound), embeddings within 5e-2 of their range; top-1 classes agree."""
    opts = PathOptions.tutorial()
    net = NeutrinoSDXLNetwork(opts, 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16")
    state = synth.init_state(net.specs, seed=4, perturb=True)
    net.load_state_dict(state, strict=True)
    net = net.to(dev).eval()
    batch = synth.make_batch(2, seed=8, prongs_per_event=[2, 1])
    ev = restate.densify(restate.preprocess_values(batch.event_values), batch.event_coords, H, W)
    pr = restate.densify(restate.preprocess_values(batch.prong_values), batch.prong_coords, H, W)
    taps = {}
    with torch.no_grad():
        want_ev, want_pr = restate_sdxl.network_forward(state, opts, ev, batch.event_mask, pr, batch.prong_mask, taps=taps)
        eng = net.engine
        eng.ensure_packed(1)
        got_pe = eng.cnn("prong", pr.to(dev), 1)
        got_ev, got_pr = net.forward_sparse(batch.to(dev))
    e_emb = rel_err(got_pe.cpu(), taps["prong_embedding"])
    e_ev, e_pr = rel_err(got_ev.cpu(), want_ev), rel_err(got_pr.cpu(), want_pr)
    print("sdxl bf16: embedding", e_emb, "event logits", e_ev, "prong logits", e_pr)
    assert e_emb < 5e-2 and e_ev < 2e-2 and e_pr < 2e-2
    assert torch.equal(got_ev.argmax(1).cpu(), want_ev.argmax(1))


def test_sdxl16_conv2d_c64_fused_kernel_against_torch(dev):
    """tcvn_sdxl16_conv2d_c64 (2-D tiles, GroupNorm + SiLU on operand load, residual + next-GroupNorm statistics in the
    epilogue) against torch on odd-sized maps that need partial tiles in both directions and several images."""
    import torch.nn.functional as F
    from dune_transformercvn_b200 import lib as tl
    L = tl.load()
    st = tl.stream_ptr(dev)
    g = torch.Generator().manual_seed(3)
    c = 64
    for n, h, w in ((3, 25, 17), (2, 50, 35)):
        hp, wp = h + 2, w + 2
        x = torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3
        res = torch.randn(n, c, h, w, generator=g)
        wt = torch.randn(c, c, 3, 3, generator=g) * 0.05
        b = torch.randn(c, generator=g)
        gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
        x16, r16 = _ring16(x, dev), _ring16(res, dev)
        xin = x16.float().cpu().view(n, hp, wp, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        rin = r16.float().cpu().view(n, hp, wp, c)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        # statistics of the input, by the library's own kernel
        ws = torch.empty(L.tcvn_sdxl16_groupnorm_workspace_bytes(n), dtype=torch.uint8, device=dev)
        stat_in = torch.empty(n, 2, device=dev)
        tl.check(L.tcvn_sdxl16_gn_stats(tl.ptr(x16), n, c, h, w, 1e-6, tl.ptr(stat_in), tl.ptr(ws), ws.numel(), st), "gn_stats")
        mean = xin.mean(dim=(1, 2, 3))
        var = xin.var(dim=(1, 2, 3), unbiased=False)
        assert rel_err(stat_in[:, 0].cpu(), mean) < 1e-4 and rel_err(stat_in[:, 1].cpu(), torch.rsqrt(var + 1e-6)) < 1e-4
        wk = wt.permute(2, 3, 0, 1).reshape(9 * c, c).contiguous().to(dev).to(torch.bfloat16)
        d_b, d_g, d_be = b.to(dev), gamma.to(dev), beta.to(dev)
        parts = torch.empty(L.tcvn_sdxl16_conv2d_stat_bytes(n, h, w), dtype=torch.uint8, device=dev)
        for with_res in (False, True):
            out = torch.full((n * hp * wp, c), 7.0, dtype=torch.bfloat16, device=dev)      # poisoned: every row must be written
            stat_out = torch.empty(n, 2, device=dev)
            tl.check(L.tcvn_sdxl16_conv2d_c64(tl.ptr(x16), n, h, w, tl.ptr(stat_in), tl.ptr(d_g), tl.ptr(d_be), tl.ptr(wk), tl.ptr(d_b),
                                              tl.ptr(r16) if with_res else None, tl.ptr(out), tl.ptr(parts), tl.ptr(stat_out), 1e-6, st),
                     "conv2d_c64")
            torch.cuda.synchronize()
            a = F.silu(F.group_norm(xin, 1, gamma, beta, 1e-6)).bfloat16().float()
            want = F.conv2d(a, wt.bfloat16().float(), b, padding=1) + (rin if with_res else 0.0)
            got = _unring(out, n, h, w)
            assert rel_err(got, want) < 2e-2, (n, h, w, with_res, rel_err(got, want))
            ring = out.float().view(n, hp, wp, c)
            assert float(ring[:, 0].abs().max()) == 0.0 and float(ring[:, -1].abs().max()) == 0.0
            assert float(ring[:, :, 0].abs().max()) == 0.0 and float(ring[:, :, -1].abs().max()) == 0.0
            o = got.bfloat16().float()
            assert rel_err(stat_out[:, 0].cpu(), o.mean(dim=(1, 2, 3))) < 2e-3
            assert rel_err(stat_out[:, 1].cpu(), torch.rsqrt(o.var(dim=(1, 2, 3), unbiased=False) + 1e-6)) < 2e-3
