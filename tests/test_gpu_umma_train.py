"""tcgen05 kernels of the bf16 training path against plain fp32 matrix products of the same bf16-rounded inputs
(pytest -m gpu).  Inputs are bf16-exact, accumulation is fp32 in TMEM: tolerance 2e-3 of the result scale covers the
summation order; the dgrad output is rounded to bf16 (2^-8)."""
import ctypes as C

import pytest
import torch

from dune_transformercvn_b200 import lib as tl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _i32(vals):
    return (C.c_int32 * len(vals))(*vals)


def _wgrad(A, G, items, fold=None, g_col0=0):
    L = tl.load()
    rows = A.shape[0]
    n = len(items)
    dw = torch.zeros(n, 128, 128, dtype=torch.float32, device=A.device)
    cols, shifts, valid = zip(*items)
    fold_cols = 0 if fold is None else fold.shape[1]
    ws = torch.empty(L.tcvn_t_umma_wgrad_workspace_bytes(n), dtype=torch.uint8, device=A.device)
    tl.check(L.tcvn_t_umma_wgrad(tl.ptr(A), rows, A.shape[1], A.stride(0), n, _i32(cols), _i32(shifts), _i32(valid),
                                 tl.ptr(fold), fold_cols, tl.ptr(G), G.shape[1], G.stride(0), g_col0, tl.ptr(dw),
                                 tl.ptr(ws), ws.numel(), tl.stream_ptr(A.device)), "tcvn_t_umma_wgrad")
    torch.cuda.synchronize()
    return dw


def _shifted(A, s):
    out = torch.zeros_like(A)
    rows = A.shape[0]
    if s >= 0:
        out[: rows - s] = A[s:]
    else:
        out[-s:] = A[: rows + s]
    return out


@pytest.mark.parametrize("rows,k", [(64, 128), (1000, 200), (20000, 488)])
def test_wgrad_column_blocks_mn_major(dev, rows, k):
    g = torch.Generator(device="cpu").manual_seed(rows + k)
    ctot = (k + 32 + 7) // 8 * 8
    A = torch.randn(rows, ctot, generator=g).to(dev).bfloat16()
    G = torch.randn(rows, 128, generator=g).to(dev).bfloat16()
    n_items = (k + 127) // 128
    items = [(128 * j, 0, min(128, k - 128 * j)) for j in range(n_items)]
    dw = _wgrad(A, G, items)
    want = A.float()[:, :k].t() @ G.float()
    for j, (c0, _, valid) in enumerate(items):
        got = dw[j, :valid]
        ref = want[c0:c0 + valid]
        assert float((got - ref).abs().max()) < 2e-3 * float(ref.abs().max()), (j, rows, k)
        assert float(dw[j, valid:].abs().max()) == 0.0 if valid < 128 else True


def test_wgrad_with_fused_activation(dev):
    rows, k = 5000, 136
    g = torch.Generator(device="cpu").manual_seed(7)
    A = torch.randn(rows, 168, generator=g).to(dev).bfloat16()
    G = torch.randn(rows, 128, generator=g).to(dev).bfloat16()
    fold = torch.stack([torch.rand(k, generator=g) + 0.5, torch.randn(k, generator=g) * 0.3, torch.rand(k, generator=g) * 0.5]).to(dev)
    items = [(0, 0, 128), (128, 0, 8)]
    dw = _wgrad(A, G, items, fold=fold.contiguous())
    y = A.float()[:, :k] * fold[0] + fold[1]
    y = y.bfloat16()                                             # the kernel rounds BN to bf16, then PReLU on bf16 pairs
    act = (torch.clamp(y, min=0) + fold[2].bfloat16() * torch.clamp(y, max=0)).float()
    want = act.bfloat16().float().t() @ G.float()
    got = torch.cat([dw[0], dw[1, :8]])
    assert float((got - want).abs().max()) < 1e-2 * float(want.abs().max())


def test_wgrad_row_shifted_items(dev):
    rows, wp = 7000, 71
    g = torch.Generator(device="cpu").manual_seed(11)
    A = torch.randn(rows, 128, generator=g).to(dev).bfloat16()
    G = torch.randn(rows, 128, generator=g).to(dev).bfloat16()
    items = [(0, -wp, 128), (0, 0, 128), (0, wp, 128)]
    dw = _wgrad(A, G, items)
    for j, (_, s, _) in enumerate(items):
        want = _shifted(A.float(), s).t() @ G.float()
        assert float((dw[j] - want).abs().max()) < 2e-3 * float(want.abs().max()), j


@pytest.mark.parametrize("n,h,w", [(3, 99, 69), (5, 24, 17), (7, 6, 4)])
def test_conv2_dgrad_matches_conv_transpose(dev, n, h, w):
    L = tl.load()
    hp, wp = h + 2, w + 2
    rows = n * hp * wp
    g = torch.Generator(device="cpu").manual_seed(h)
    grad = torch.randn(n, 32, h, w, generator=g).bfloat16().float()
    w2 = (torch.randn(32, 128, 3, 3, generator=g) * 0.1).bfloat16().float()
    # reference: input gradient of conv2d(mid, w2, padding=1)
    want = torch.nn.grad.conv2d_input((n, 128, h, w), w2, grad, padding=1)
    # ringed channels-last gradient and its three horizontal shifts
    G = torch.zeros(n, hp, wp, 32)
    G[:, 1:-1, 1:-1] = grad.permute(0, 2, 3, 1)
    G = G.reshape(rows, 32)
    g2x = torch.zeros(rows, 128)
    g2x[:-1, 0:32] = G[1:]
    g2x[:, 32:64] = G
    g2x[1:, 64:96] = G[:-1]
    wd = torch.zeros(3, 128, 128)
    for dy in range(3):
        for dx in range(3):
            wd[dy, :, dx * 32:(dx + 1) * 32] = w2[:, :, dy, dx].t()
    out = torch.empty(rows, 128, dtype=torch.bfloat16, device=dev)
    g2x_d, wd_d = g2x.to(dev).bfloat16(), wd.to(dev).bfloat16()   # keep alive: the C ABI only sees raw pointers
    tl.check(L.tcvn_t_umma_conv2_dgrad(tl.ptr(g2x_d), tl.ptr(wd_d), rows, hp, wp, tl.ptr(out), tl.stream_ptr(dev)),
             "tcvn_t_umma_conv2_dgrad")
    got = out.float().cpu().reshape(n, hp, wp, 128)
    ring = torch.cat([got[:, 0].reshape(-1), got[:, -1].reshape(-1), got[:, :, 0].reshape(-1), got[:, :, -1].reshape(-1)])
    assert float(ring.abs().max()) == 0.0
    inner = got[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    assert float((inner - want).abs().max()) < 1e-2 * float(want.abs().max())
