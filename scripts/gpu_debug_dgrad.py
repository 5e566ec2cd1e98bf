import sys, os, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_transformercvn_b200 import lib as tl
dev = torch.device("cuda:0")
L = tl.load()
for (n, h, w) in [(3, 99, 69), (5, 24, 17), (7, 6, 4), (4, 12, 8), (2, 49, 34), (40, 24, 17)]:
    hp, wp = h + 2, w + 2
    rows = n * hp * wp
    g = torch.Generator(device="cpu").manual_seed(h)
    grad = torch.randn(n, 32, h, w, generator=g).bfloat16().float()
    w2 = (torch.randn(32, 128, 3, 3, generator=g) * 0.1).bfloat16().float()
    want = torch.nn.grad.conv2d_input((n, 128, h, w), w2, grad, padding=1)
    G = torch.zeros(n, hp, wp, 32); G[:, 1:-1, 1:-1] = grad.permute(0, 2, 3, 1); G = G.reshape(rows, 32)
    g2x = torch.zeros(rows, 128); g2x[:-1, 0:32] = G[1:]; g2x[:, 32:64] = G; g2x[1:, 64:96] = G[:-1]
    wd = torch.zeros(3, 128, 128)
    for dy in range(3):
        for dx in range(3):
            wd[dy, :, dx * 32:(dx + 1) * 32] = w2[:, :, dy, dx].t()
    out = torch.empty(rows, 128, dtype=torch.bfloat16, device=dev)
    a, b = g2x.to(dev).bfloat16(), wd.to(dev).bfloat16()
    tl.check(L.tcvn_t_umma_conv2_dgrad(tl.ptr(a), tl.ptr(b), rows, hp, wp, tl.ptr(out), tl.stream_ptr(dev)), "dgrad")
    torch.cuda.synchronize()
    got = out.float().cpu()
    ref = torch.zeros(n, hp, wp, 128); ref[:, 1:-1, 1:-1] = want.permute(0, 2, 3, 1); ref = ref.reshape(rows, 128)
    err = (got - ref).abs().max(dim=1).values
    tiles = (rows + 127) // 128
    per_tile = [float(err[t * 128:(t + 1) * 128].max()) for t in range(tiles)]
    bad = [t for t, e in enumerate(per_tile) if e > 0.05]
    print(f"n={n} {h}x{w} rows={rows} tiles={tiles} bad tiles: {bad[:40]} (count {len(bad)}) max err {float(err.max()):.3f}")
    if bad:
        t = bad[0]
        e = err[t * 128:(t + 1) * 128]
        print("   first bad tile row errors:", [round(float(x), 2) for x in e[:128:4]])
        # which dy contribution is missing? compare against partial references
        for skip in range(3):
            w2s = w2.clone(); w2s[:, :, skip] = 0
            ws = torch.nn.grad.conv2d_input((n, 128, h, w), w2s, grad, padding=1)
            r2 = torch.zeros(n, hp, wp, 128); r2[:, 1:-1, 1:-1] = ws.permute(0, 2, 3, 1); r2 = r2.reshape(rows, 128)
            print(f"   without dy={skip}: err in that tile {float((got - r2)[t*128:(t+1)*128].abs().max()):.3f}")
