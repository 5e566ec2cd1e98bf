"""--sdxl variant of the drop-in network (BASELINE configs[3]): ``NeutrinoSDXLNetwork``.

Reference: transformercvn/network/networks/neutrino_full_sdxl_network.py:6-21 (the base network with both pixel
embeddings replaced by ``SDXLNet``), transformercvn/network/layers/sdxl_net.py:7-42 (``diffusers.models.vae.Encoder``
with 9 ``DownEncoderBlock2D`` of widths [64, 64, 128, 128, 256, 256, 512, 512, out], GroupNorm with ONE group, then
Flatten + Linear).  diffusers is un-vendored, unpinned third-party code that is absent here: the arithmetic follows its
published layout (see oracle/restate_sdxl.py) and **parity is unpinned** (SURVEY 8c).

Round-1 state: eval-mode forward, fp32, through the C ABI only - every convolution / linear layer is the shifted GEMM
of the fp32 parity path (``tcvn_t_gemm`` over ringed channels-last maps, the residual add is its accumulate mode),
GroupNorm+SiLU, the stride-2 patch gather and the NCHW->ringed conversion are ``csrc/sdxl.cu``.  The token assembly,
encoder and heads are the fused sequence kernel shared with the DenseNet network.  Not built: train mode, the bf16
tcgen05 path, attention over more than one spatial position (400x280 inputs reach the mid block at 1x1).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from . import lib as _lib
from .network import NeutrinoDenseNetwork, _Engine
from .params import embedding_dims, network_specs, sdxl_block_channels

GN_EPS = 1e-6
_IMAGE_CHUNK = 16   # images per walk: ~150 MB of fp32 feature maps per image at 400x280


def _taps3(wp: int):
    return (C.c_int32 * 9)(*[(dy - 1) * wp + (dx - 1) for dy in range(3) for dx in range(3)])


class _SdxlCnn:
    """One SDXLNet: packed (GEMM-layout) weights and the layer walk."""

    def __init__(self, prefix: str, in_ch: int, out_features: int, init_block_dim: int):
        self.prefix = prefix
        self.in_ch = in_ch
        self.out = out_features
        self.ch = sdxl_block_channels(init_block_dim, out_features)
        self.w: Dict[str, torch.Tensor] = {}

    # ---- packing: conv (Cout, Cin, kh, kw) -> [kh*kw][Cin][Cout]; linear (Cout, Cin) -> [Cin][Cout] -------------------
    def pack(self, tensors: Dict[str, torch.Tensor]) -> None:
        self.w.clear()
        for name, t in tensors.items():
            if not name.startswith(self.prefix):
                continue
            key = name[len(self.prefix):]
            t = t.detach().float()
            if t.dim() == 4:
                co, ci, kh, kw = t.shape
                t = t.permute(2, 3, 1, 0).reshape(kh * kw, ci, co)
            elif t.dim() == 2:
                t = t.t()
            self.w[key] = t.contiguous()

    # ---- primitives -------------------------------------------------------------------------------------------------
    def _gemm(self, L, st, a, lda, rows, k, taps, tap_off, w, n_out, bias, out, ldo, ring, accumulate=False, a_offset=0):
        a_ptr = C.c_void_p(a.data_ptr() + 4 * a_offset)
        _lib.check(L.tcvn_t_gemm(a_ptr, lda, rows, k, taps, tap_off, _lib.ptr(w), n_out, None, 0, 0, _lib.ptr(bias),
                                 _lib.ptr(out), ldo, 0, ring[0], ring[1], 1 if accumulate else 0, st), "tcvn_t_gemm")

    def _conv3(self, L, st, name, a, n, h, w_, cin, cout, out=None, accumulate=False):
        hp, wp = h + 2, w_ + 2
        rows = n * hp * wp
        if out is None:
            out = torch.empty((rows, cout), dtype=torch.float32, device=a.device)
        self._gemm(L, st, a, cin, rows, cin, 9, _taps3(wp), self.w[name + ".weight"], cout, self.w[name + ".bias"], out, cout,
                   (hp, wp), accumulate)
        return out

    def _norm(self, L, st, name, x, n, h, w_, c, silu, sums):
        out = torch.empty_like(x)
        _lib.check(L.tcvn_sdxl_groupnorm(_lib.ptr(x), n, c, 1, h, w_, _lib.ptr(self.w[name + ".weight"]),
                                         _lib.ptr(self.w[name + ".bias"]), GN_EPS, 1 if silu else 0, _lib.ptr(out),
                                         _lib.ptr(sums), st), "tcvn_sdxl_groupnorm")
        return out

    def _resnet(self, L, st, p, x, n, h, w_, cin, cout, sums):
        a = self._norm(L, st, p + "norm1", x, n, h, w_, cin, True, sums)
        t = self._conv3(L, st, p + "conv1", a, n, h, w_, cin, cout)
        a = self._norm(L, st, p + "norm2", t, n, h, w_, cout, True, sums)
        if cin != cout:   # 1x1 shortcut into a new map, then conv2 accumulates onto it
            rows = n * (h + 2) * (w_ + 2)
            y = torch.empty((rows, cout), dtype=torch.float32, device=x.device)
            self._gemm(L, st, x, cin, rows, cin, 1, None, self.w[p + "conv_shortcut.weight"], cout,
                       self.w[p + "conv_shortcut.bias"], y, cout, (h + 2, w_ + 2))
            x = y
        self._conv3(L, st, p + "conv2", a, n, h, w_, cout, cout, out=x, accumulate=True)   # x += conv2(...) + bias
        return x

    # ---- the walk ---------------------------------------------------------------------------------------------------
    def forward(self, pixels: torch.Tensor) -> torch.Tensor:
        outs = [self._forward_chunk(pixels[i:i + _IMAGE_CHUNK]) for i in range(0, pixels.shape[0], _IMAGE_CHUNK)]
        if not outs:
            return torch.empty((0, self.out), dtype=torch.float32, device=pixels.device)
        return torch.cat(outs) if len(outs) > 1 else outs[0]

    def _forward_chunk(self, pixels: torch.Tensor) -> torch.Tensor:
        L = _lib.load()
        dev = pixels.device
        st = _lib.stream_ptr(dev)
        n, cin0, h, w_ = pixels.shape
        if cin0 != self.in_ch:
            raise _lib.TcvnError(f"sdxl: pixels have {cin0} channels, expected {self.in_ch}")
        f32 = dict(dtype=torch.float32, device=dev)
        sums = torch.empty(2 * n, dtype=torch.float64, device=dev)
        pixels = pixels.contiguous().float()
        ring = torch.empty((n * (h + 2) * (w_ + 2), cin0), **f32)
        _lib.check(L.tcvn_sdxl_pixels_to_ring(_lib.ptr(pixels), n, cin0, h, w_, _lib.ptr(ring), st), "tcvn_sdxl_pixels_to_ring")
        e = "encoder."
        x = self._conv3(L, st, e + "conv_in", ring, n, h, w_, cin0, self.ch[0])
        cin = self.ch[0]
        for i, cout in enumerate(self.ch[:-1]):
            for j in range(2):
                x = self._resnet(L, st, f"{e}down_blocks.{i}.resnets.{j}.", x, n, h, w_, cin if j == 0 else cout, cout, sums)
            if True:
                if h < 2 or w_ < 2:
                    raise _lib.TcvnError(f"sdxl: a {h}x{w_} map cannot be down-sampled again (input too small)")
                ho, wo = h // 2, w_ // 2
                rows = n * (ho + 2) * (wo + 2)
                patches = torch.empty((rows, 9 * cout), **f32)
                _lib.check(L.tcvn_sdxl_patch_s2(_lib.ptr(x), n, cout, h, w_, _lib.ptr(patches), st), "tcvn_sdxl_patch_s2")
                d = f"{e}down_blocks.{i}.downsamplers.0.conv"
                x = torch.empty((rows, cout), **f32)
                self._gemm(L, st, patches, 9 * cout, rows, 9 * cout, 1, None, self.w[d + ".weight"], cout, self.w[d + ".bias"],
                           x, cout, (ho + 2, wo + 2))
                del patches
                h, w_ = ho, wo
            cin = cout
        if (h, w_) != (1, 1):
            raise _lib.TcvnError(f"sdxl: the mid block is reached at {h}x{w_}; only the 1x1 case (400x280 inputs) is built")
        return self._tail(L, st, x, n, cin)

    def _tail(self, L, st, x: torch.Tensor, n: int, cin: int) -> torch.Tensor:
        """The 1x1-spatial end of the encoder on a ringed fp32 map [n*9][cin]: last down block, mid block (ResNet, attention
        over one position, ResNet), GroupNorm + SiLU + conv_out, Flatten + Linear."""
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        sums = torch.empty(2 * n, dtype=torch.float64, device=dev)
        e = "encoder."
        h = w_ = 1
        i, cout = len(self.ch) - 1, self.ch[-1]
        for j in range(2):
            x = self._resnet(L, st, f"{e}down_blocks.{i}.resnets.{j}.", x, n, h, w_, cin if j == 0 else cout, cout, sums)
        c = self.ch[-1]
        m = e + "mid_block."
        x = self._resnet(L, st, m + "resnets.0.", x, n, h, w_, c, c, sums)
        # Attention over ONE position: softmax over a single key is 1, so out = to_out(to_v(group_norm(x))) + x
        a = self._norm(L, st, m + "attentions.0.group_norm", x, n, h, w_, c, False, sums)
        rows = n * 9
        v = torch.empty((rows, c), **f32)
        att = m + "attentions.0."
        self._gemm(L, st, a, c, rows, c, 1, None, self.w[att + "to_v.weight"], c, self.w[att + "to_v.bias"], v, c, (3, 3))
        self._gemm(L, st, v, c, rows, c, 1, None, self.w[att + "to_out.0.weight"], c, self.w[att + "to_out.0.bias"], x, c, (3, 3),
                   accumulate=True)
        x = self._resnet(L, st, m + "resnets.1.", x, n, h, w_, c, c, sums)
        a = self._norm(L, st, e + "conv_norm_out", x, n, h, w_, c, True, sums)
        z = self._conv3(L, st, e + "conv_out", a, n, h, w_, c, self.out)
        out = torch.empty((n, self.out), **f32)
        # Flatten + Linear on the interior row (index 4 of the 3x3 ringed map) of every image
        self._gemm(L, st, z, 9 * self.out, n, self.out, 1, None, self.w["output_layer.1.weight"], self.out,
                   self.w["output_layer.1.bias"], out, self.out, (0, 0), a_offset=4 * self.out)
        return out


def _pad_to(v: int, a: int) -> int:
    return (v + a - 1) // a * a


class _SdxlCnn16(_SdxlCnn):
    """bf16 / tcgen05 walk of one SDXLNet (round 2).  Down blocks 0..7 (400x280 ... 3x2 maps, 99.9 % of the FLOPs) run on
    the tensor cores: every 3x3 convolution is ONE shifted GEMM (``tcvn_sdxl16_conv``: nine row-shifted views of the
    activated map as the K loop), and a ResNet block's residual add / 1x1 shortcut rides in the same GEMM as a trailing K
    segment over the block input with identity / shortcut weights.  The 1x1-spatial tail (last down block, mid block,
    conv_out, Linear: 9 rows per image) stays on the fp32 kernels of the base class."""

    image_chunk = 48   # images per walk: ~60 MB of bf16 maps per image at 400x280

    def pack(self, tensors: Dict[str, torch.Tensor]) -> None:
        super().pack(tensors)     # fp32 GEMM layouts for the tail
        self.w16: Dict[str, Tuple[torch.Tensor, torch.Tensor, int, int]] = {}
        self.w2d: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}   # 64 -> 64 3x3 convs for the fused 2-D tile kernel
        dev = None
        raw = {n[len(self.prefix):]: t.detach().float() for n, t in tensors.items() if n.startswith(self.prefix)}
        dev = next(iter(raw.values())).device
        self.ones = torch.ones(1024, dtype=torch.float32, device=dev)

        def conv_k(wt: torch.Tensor, cin_pad: int) -> torch.Tensor:
            """(Cout, Cin, kh, kw) -> [Cout][kh*kw*cin_pad], K index = tap * cin_pad + c"""
            co, ci, kh, kw = wt.shape
            out = torch.zeros((co, kh * kw, cin_pad), dtype=torch.float32, device=wt.device)
            out[:, :, :ci] = wt.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
            return out.reshape(co, kh * kw * cin_pad)

        def finish(key: str, wk: torch.Tensor, bias: torch.Tensor) -> None:
            co, k = wk.shape
            n_tiles = _pad_to(co, 128) // 128
            w = torch.zeros((n_tiles * 128, _pad_to(k, 64)), dtype=torch.float32, device=wk.device)
            w[:co, :k] = wk
            b = torch.zeros(n_tiles * 128, dtype=torch.float32, device=wk.device)
            b[:co] = bias
            self.w16[key] = (w.to(torch.bfloat16).contiguous(), b, n_tiles, co)

        e = "encoder."
        # conv_in: one 64-wide K chunk, column (tap, channel) with stride 3 (tcvn_sdxl16_patch27)
        wt = raw[e + "conv_in.weight"]
        co, ci = wt.shape[0], wt.shape[1]
        wk = torch.zeros((co, 64), dtype=torch.float32, device=dev)
        for c in range(ci):
            wk[:, c:27:3] = wt[:, c].reshape(co, 9)
        finish("conv_in", wk, raw[e + "conv_in.bias"])
        cin = self.ch[0]
        for i, cout in enumerate(self.ch[:-1]):
            for j in range(2):
                p = f"{e}down_blocks.{i}.resnets.{j}."
                c_in = cin if j == 0 else cout
                if c_in == 64 and cout == 64:   # [9*64][64]: row t*64 + n, column k = weight[n][k][t/3][t%3]
                    for cv in ("conv1", "conv2"):
                        wt = raw[p + cv + ".weight"]
                        self.w2d[p + cv] = (wt.permute(2, 3, 0, 1).reshape(9 * 64, 64).to(torch.bfloat16).contiguous(),
                                            raw[p + cv + ".bias"].contiguous())
                    continue
                finish(p + "conv1", conv_k(raw[p + "conv1.weight"], c_in), raw[p + "conv1.bias"])
                w2 = conv_k(raw[p + "conv2.weight"], cout)
                if c_in != cout:   # 1x1 shortcut as the trailing K segment
                    tail = raw[p + "conv_shortcut.weight"].reshape(cout, c_in)
                    bias = raw[p + "conv2.bias"] + raw[p + "conv_shortcut.bias"]
                else:              # identity: out = conv2(a) + x
                    tail = torch.eye(cout, dtype=torch.float32, device=dev)
                    bias = raw[p + "conv2.bias"]
                finish(p + "conv2", torch.cat((w2, tail), dim=1), bias)
            d = f"{e}down_blocks.{i}.downsamplers.0.conv"
            finish(d, conv_k(raw[d + ".weight"], cout), raw[d + ".bias"])
            cin = cout

    def _conv16(self, L, st, key, a, rows, a_cols, taps, wp, x2, x2_cols, hp_wp):
        w, b, n_tiles, co = self.w16[key]
        out = torch.empty((rows, co), dtype=torch.bfloat16, device=a.device)
        _lib.check(L.tcvn_sdxl16_conv(_lib.ptr(a), rows, a_cols, taps, _taps3(wp) if taps == 9 else None, _lib.ptr(x2), x2_cols,
                                      _lib.ptr(w), n_tiles, _lib.ptr(b), _lib.ptr(self.ones), _lib.ptr(out), co, hp_wp[0], hp_wp[1],
                                      st), "tcvn_sdxl16_conv")
        return out

    def _norm16(self, L, st, name, x, n, h, w_, c, ws):
        out = torch.empty_like(x)
        _lib.check(L.tcvn_sdxl16_groupnorm(_lib.ptr(x), n, c, h, w_, _lib.ptr(self.w[name + ".weight"]), _lib.ptr(self.w[name + ".bias"]),
                                           GN_EPS, 1, _lib.ptr(out), _lib.ptr(ws), ws.numel(), st), "tcvn_sdxl16_groupnorm")
        return out

    def _resnets_c64(self, L, st, block: str, x, n, h, w_, ws):
        """The two ResNet blocks of a 64-channel stage on the fused 2-D tile kernel: GroupNorm + SiLU applied on operand load,
        residual added and the next GroupNorm's statistics taken in the epilogue - four tensor-core launches, no activated
        map, no separate statistics pass except at the stage's entry."""
        dev = x.device
        rows = n * (h + 2) * (w_ + 2)
        stat = torch.empty((n, 2), dtype=torch.float32, device=dev)
        _lib.check(L.tcvn_sdxl16_gn_stats(_lib.ptr(x), n, 64, h, w_, GN_EPS, _lib.ptr(stat), _lib.ptr(ws), ws.numel(), st),
                   "tcvn_sdxl16_gn_stats")
        parts = torch.empty(L.tcvn_sdxl16_conv2d_stat_bytes(n, h, w_), dtype=torch.uint8, device=dev)
        for j in range(2):
            p = f"{block}resnets.{j}."
            for cv, nm, res in (("conv1", "norm1", None), ("conv2", "norm2", x)):
                wk, bias = self.w2d[p + cv]
                out = torch.empty((rows, 64), dtype=torch.bfloat16, device=dev)
                stat_out = torch.empty((n, 2), dtype=torch.float32, device=dev)
                _lib.check(L.tcvn_sdxl16_conv2d_c64(_lib.ptr(x if cv == "conv1" else t), n, h, w_, _lib.ptr(stat),
                                                    _lib.ptr(self.w[p + nm + ".weight"]), _lib.ptr(self.w[p + nm + ".bias"]),
                                                    _lib.ptr(wk), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), _lib.ptr(parts),
                                                    _lib.ptr(stat_out), GN_EPS, st), "tcvn_sdxl16_conv2d_c64")
                if cv == "conv1":
                    t = out
                else:
                    x = out
                stat = stat_out
        return x

    def forward(self, pixels: torch.Tensor) -> torch.Tensor:
        n = pixels.shape[0]
        if n == 0:
            return torch.empty((0, self.out), dtype=torch.float32, device=pixels.device)
        # the big maps are walked in chunks of images; the 1x1-spatial tail (26 small fp32 GEMMs whose cost is reading
        # their weights: 5 ms per call whatever the image count) runs ONCE over all images
        maps = [self._forward_chunk(pixels[i:i + self.image_chunk]) for i in range(0, n, self.image_chunk)]
        x32 = torch.cat(maps) if len(maps) > 1 else maps[0]
        return self._tail(_lib.load(), _lib.stream_ptr(pixels.device), x32, n, self.ch[-2])

    def _forward_chunk(self, pixels: torch.Tensor) -> torch.Tensor:
        """Down blocks 0..7 of a chunk of images on the tensor cores -> its ringed fp32 1x1 maps [n*9][512]."""
        L = _lib.load()
        dev = pixels.device
        st = _lib.stream_ptr(dev)
        n, cin0, h, w_ = pixels.shape
        if cin0 != self.in_ch or cin0 > 3:
            raise _lib.TcvnError(f"sdxl bf16: pixels have {cin0} channels, expected {self.in_ch} (<= 3)")
        pixels = pixels.contiguous().float()
        ws = torch.empty(L.tcvn_sdxl16_groupnorm_workspace_bytes(n), dtype=torch.uint8, device=dev)
        e = "encoder."
        hp, wp = h + 2, w_ + 2
        rows = n * hp * wp
        patches = torch.empty((rows, 64), dtype=torch.bfloat16, device=dev)
        _lib.check(L.tcvn_sdxl16_patch27(_lib.ptr(pixels), n, cin0, h, w_, 0.0, _lib.ptr(patches), st), "tcvn_sdxl16_patch27")
        x = self._conv16(L, st, "conv_in", patches, rows, 64, 1, wp, None, 0, (hp, wp))
        del patches
        cin = self.ch[0]
        for i, cout in enumerate(self.ch[:-1]):
            hp, wp = h + 2, w_ + 2
            rows = n * hp * wp
            if cin == 64 and cout == 64:
                x = self._resnets_c64(L, st, f"{e}down_blocks.{i}.", x, n, h, w_, ws)
            for j in range(2 if not (cin == 64 and cout == 64) else 0):
                p = f"{e}down_blocks.{i}.resnets.{j}."
                c_in = cin if j == 0 else cout
                a = self._norm16(L, st, p + "norm1", x, n, h, w_, c_in, ws)
                t = self._conv16(L, st, p + "conv1", a, rows, c_in, 9, wp, None, 0, (hp, wp))
                a = self._norm16(L, st, p + "norm2", t, n, h, w_, cout, ws)
                x = self._conv16(L, st, p + "conv2", a, rows, cout, 9, wp, x, c_in, (hp, wp))   # conv2(a) + shortcut(x) + biases
            if h < 2 or w_ < 2:
                raise _lib.TcvnError(f"sdxl: a {h}x{w_} map cannot be down-sampled again (input too small)")
            ho, wo = h // 2, w_ // 2
            orows = n * (ho + 2) * (wo + 2)
            # stride-2 convolution: nine shifted views of the space-to-depth matrix (512 B per output pixel instead of 1152 B
            # of materialised patches)
            pt = torch.empty((orows, 4 * cout), dtype=torch.bfloat16, device=dev)
            _lib.check(L.tcvn_sdxl16_s2d(_lib.ptr(x), n, cout, h, w_, _lib.ptr(pt), st), "tcvn_sdxl16_s2d")
            wd, bd, n_tiles, co = self.w16[f"{e}down_blocks.{i}.downsamplers.0.conv"]
            x = torch.empty((orows, co), dtype=torch.bfloat16, device=dev)
            _lib.check(L.tcvn_sdxl16_conv_s2(_lib.ptr(pt), orows, cout, _lib.ptr(wd), n_tiles, _lib.ptr(bd), _lib.ptr(self.ones),
                                             _lib.ptr(x), co, ho + 2, wo + 2, st), "tcvn_sdxl16_conv_s2")
            del pt
            h, w_ = ho, wo
            cin = cout
        if (h, w_) != (1, 1):
            raise _lib.TcvnError(f"sdxl: the mid block is reached at {h}x{w_}; only the 1x1 case (400x280 inputs) is built")
        x32 = torch.empty(tuple(x.shape), dtype=torch.float32, device=dev)
        _lib.check(L.tcvn_sdxl16_to_f32(_lib.ptr(x), x.numel(), _lib.ptr(x32), st), "tcvn_sdxl16_to_f32")
        return x32



class _SdxlEngine(_Engine):
    """The dense engine with the two pixel-map CNNs swapped; the sequence stage (tokens, encoder, heads) is shared."""

    def __init__(self, owner):
        super().__init__(owner)
        o = owner.options
        pix, feat, _ = embedding_dims(o)
        pe = "prong_embedding."
        cnn_in = owner.pixel_dim
        cls = _SdxlCnn16 if owner.precision == "bf16" else _SdxlCnn
        self.cnns = {"prong": cls(pe + "prong_pixel_embedding.", cnn_in, pix, o.initial_pixel_dim),
                     "event": cls(pe + "event_pixel_embedding.", cnn_in, pix + feat, o.initial_pixel_dim)}

    def ensure_packed(self, prec: int) -> None:
        net = self.owner[0]
        if self.frozen and self.key is not None and self.key[1] == prec:
            return
        key = self._state_key(prec)
        if key == self.key:
            return
        tensors = dict(net.named_parameters())
        tensors.update(dict(net.named_buffers()))
        self.generation += 1
        L = _lib.load()
        dev = next(iter(tensors.values())).device
        if dev.type != "cuda":
            raise _lib.TcvnError("NeutrinoSDXLNetwork: parameters are on the CPU; move the module to a CUDA device "
                                 "(this path has no CPU implementation)")
        for cnn in self.cnns.values():
            cnn.pack(tensors)
        st = _lib.stream_ptr(dev)
        pe = "prong_embedding."
        sd = self.seq_desc()
        nbytes = L.tcvn_seq_packed_bytes(C.byref(sd))
        if nbytes == 0:
            raise _lib.TcvnError("sequence descriptor rejected: " + L.tcvn_last_error().decode())
        buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        position = tensors[pe + "event_position_embedding"].detach().reshape(-1).float().contiguous()
        arenas = [self._arena(tensors, p) for p in (pe + "combined_embedding.", "encoder.", "event_decoder.", "prong_decoder.")]
        _lib.check(L.tcvn_seq_pack(C.byref(sd), _lib.ptr(position), *[_lib.ptr(a) for a in arenas], _lib.ptr(buf), nbytes, st),
                   "tcvn_seq_pack")
        self.packed["seq"] = buf
        self.key = key

    def cnn(self, tag: str, pixels: torch.Tensor, prec: int, ws_kind: str = "cnn") -> torch.Tensor:
        _lib.require_cuda(pixels, f"{tag} pixels")
        net = self.owner[0]
        if tuple(pixels.shape[2:]) != tuple(net.image_size):
            raise _lib.TcvnError(f"{tag} pixels have shape {tuple(pixels.shape)}, expected (N,{net.pixel_dim},{net.image_size[0]},{net.image_size[1]})")
        return self.cnns[tag].forward(pixels)

    def cnn_sparse(self, tag, values, coords, n, prec, divisor: float = 255.0, ws_kind: str = "cnn"):
        from .ingest import densify
        return self.cnn(tag, densify(values, coords, self.owner[0].image_size, n, divisor), prec)


class NeutrinoSDXLNetwork(NeutrinoDenseNetwork):
    """Same constructor and forward as the reference class of this name
    (networks/neutrino_full_sdxl_network.py:19-21); eval mode, fp32."""

    cnn_kind = "sdxl"

    def _make_engine(self):
        return _SdxlEngine(self)

    def __init__(self, options, features_dim: int, extra_dim: int, pixel_dim: int, num_prong_classes: int,
                 num_event_classes: int, image_size=None, precision: str = "fp32", seed: int = 0):
        if precision not in ("fp32", "bf16"):
            raise _lib.TcvnError("NeutrinoSDXLNetwork: precision must be fp32 (CUDA-core parity path) or bf16 (tcgen05 path)")
        kw = {} if image_size is None else {"image_size": image_size}
        super().__init__(options, features_dim, extra_dim, pixel_dim, num_prong_classes, num_event_classes,
                         precision=precision, seed=seed, **kw)

    def forward(self, features, extra, event_pixels, event_mask, prong_pixels, prong_mask):
        if self.training:
            raise NotImplementedError("NeutrinoSDXLNetwork: the train-mode forward / backward of the --sdxl variant is not "
                                      "built (round 1: eval-mode inference, BASELINE configs[3]); call .eval(). There is "
                                      "deliberately no PyTorch fallback")
        return super().forward(features, extra, event_pixels, event_mask, prong_pixels, prong_mask)

    def forward_sparse(self, batch, materialize: bool = True):
        """/255 + sparse_to_dense (densify kernel) + network, as the --sdxl trainer does
        (trainers/neutrino_full_sdxl_trainer.py:8 inherits neutrino_full_dense_trainer.py:46-66)."""
        return super().forward_sparse(batch, materialize=True)
