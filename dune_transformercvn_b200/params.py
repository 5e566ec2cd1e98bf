"""Canonical parameter/buffer inventory of the hot path.

One ordered list of ``TensorSpec`` describes every tensor the reference network
keeps in its ``state_dict`` (names, shapes, order — reference module tree:
transformercvn/network/networks/neutrino_full_base_network.py:38-85,128-164,
transformercvn/network/layers/dense_net.py:8-162,
transformercvn/network/layers/prong_target_decoder.py:19-36,
transformercvn/network/layers/prong_feature_embedding.py:7-71).

The same walk is mirrored in C (csrc/arena.cpp) so that a flat fp32 "arena"
holding the float tensors back to back can be handed to the CUDA side as one
pointer.  ``tests/test_params.py`` pins the list against the key/shape dump of
the real reference (tests/golden/state_dict_keys.json).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Sequence, Tuple

from .config import round_channels


@dataclass(frozen=True)
class TensorSpec:
    name: str
    shape: Tuple[int, ...]
    role: str          # conv_w conv_b bn_w bn_b bn_rm bn_rv bn_nbt prelu lin_w lin_b ln_w ln_b attn_w attn_b pos
    is_param: bool     # False => buffer

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def in_arena(self) -> bool:
        """int64 ``num_batches_tracked`` counters live outside the fp32 arena."""
        return self.role != "bn_nbt"


def _bn(prefix: str, c: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (c,), "bn_w", True)
    yield TensorSpec(prefix + ".bias", (c,), "bn_b", True)
    yield TensorSpec(prefix + ".running_mean", (c,), "bn_rm", False)
    yield TensorSpec(prefix + ".running_var", (c,), "bn_rv", False)
    yield TensorSpec(prefix + ".num_batches_tracked", (), "bn_nbt", False)


def _prelu(prefix: str, c: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (c,), "prelu", True)


def _conv(prefix: str, cout: int, cin: int, k: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (cout, cin, k, k), "conv_w", True)
    yield TensorSpec(prefix + ".bias", (cout,), "conv_b", True)


def _linear(prefix: str, cout: int, cin: int, bias: bool) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (cout, cin), "lin_w", True)
    if bias:
        yield TensorSpec(prefix + ".bias", (cout,), "lin_b", True)


def densenet_specs(prefix: str, in_ch: int, out_features: int, init_features: int, growth: int,
                   bn_size: int, blocks: Sequence[int]) -> Iterator[TensorSpec]:
    """Tensors of one pixel-map CNN in state_dict order (dense_net.py:97-162)."""
    f = prefix + "features."
    yield from _conv(f + "conv0", init_features, in_ch, 7)
    yield from _bn(f + "norm0", init_features)
    yield from _prelu(f + "relu0", init_features)
    c = init_features
    mid = bn_size * growth
    for bi, nl in enumerate(blocks):
        for li in range(nl):
            p = f"{f}dense{bi + 1}.layers.{li}."
            cin = c + li * growth
            yield from _bn(p + "bottleneck_block.norm1", cin)
            yield from _prelu(p + "bottleneck_block.relu1", cin)
            yield from _conv(p + "bottleneck_block.conv1", mid, cin, 1)
            yield from _bn(p + "output_block.norm2", mid)
            yield from _prelu(p + "output_block.relu2", mid)
            yield from _conv(p + "output_block.conv2", growth, mid, 3)
        c += nl * growth
        if bi != len(blocks) - 1:
            t = f"{f}transition{bi + 1}."
            yield from _bn(t + "norm", c)
            yield from _prelu(t + "relu", c)
            yield from _conv(t + "conv", c // 2, c, 1)
            c //= 2
    yield from _bn(f + "final_norm", c)
    yield from _prelu(f + "final_relu", c)
    o = prefix + "output_block."
    yield from _linear(o + "linear", out_features, c, bias=False)
    yield from _bn(o + "norm", out_features)
    yield from _prelu(o + "relu", out_features)


def sdxl_block_channels(init_block_dim: int, out_features: int, repeat: int = 2, num_blocks: int = 4) -> List[int]:
    """``block_out_channels`` of the --sdxl encoder (layers/sdxl_net.py:20-26): [64, 64, 128, 128, 256, 256, 512, 512, out]."""
    ch: List[int] = []
    dim = init_block_dim
    for _ in range(num_blocks):
        ch.extend([dim] * repeat)
        dim *= 2
    ch.append(out_features)
    return ch


def _gn(prefix: str, c: int) -> Iterator[TensorSpec]:
    yield TensorSpec(prefix + ".weight", (c,), "ln_w", True)
    yield TensorSpec(prefix + ".bias", (c,), "ln_b", True)


def _resnet(prefix: str, cin: int, cout: int) -> Iterator[TensorSpec]:
    yield from _gn(prefix + "norm1", cin)
    yield from _conv(prefix + "conv1", cout, cin, 3)
    yield from _gn(prefix + "norm2", cout)
    yield from _conv(prefix + "conv2", cout, cout, 3)
    if cin != cout:
        yield from _conv(prefix + "conv_shortcut", cout, cin, 1)


def sdxl_specs(prefix: str, in_ch: int, out_features: int, init_block_dim: int) -> Iterator[TensorSpec]:
    """Tensors of ``SDXLNet`` (layers/sdxl_net.py:7-42) in module-registration order.  The encoder is
    ``diffusers.models.vae.Encoder(down_block_types=("DownEncoderBlock2D",)*9, layers_per_block=2, norm_num_groups=1,
    double_z=False)``; diffusers is not vendored by the reference (and absent here), so these names restate its
    published module tree (conv_in / down_blocks.i.resnets.j / downsamplers.0.conv / mid_block.attentions.0 +
    resnets / conv_norm_out / conv_out) - parity unpinned, SURVEY 8c."""
    e = prefix + "encoder."
    ch = sdxl_block_channels(init_block_dim, out_features)
    yield from _conv(e + "conv_in", ch[0], in_ch, 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(2):
            yield from _resnet(f"{e}down_blocks.{i}.resnets.{j}.", cin if j == 0 else cout, cout)
        if i != len(ch) - 1:
            yield from _conv(f"{e}down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        cin = cout
    c = ch[-1]
    a = e + "mid_block.attentions.0."
    yield from _gn(a + "group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        yield from _linear(a + n, c, c, True)
    for j in range(2):
        yield from _resnet(f"{e}mid_block.resnets.{j}.", c, c)
    yield from _gn(e + "conv_norm_out", c)
    yield from _conv(e + "conv_out", out_features, c, 3)
    yield from _linear(prefix + "output_layer.1", out_features, out_features, True)


def densenet_final_channels(init_features: int, growth: int, blocks: Sequence[int]) -> int:
    c = init_features
    for bi, nl in enumerate(blocks):
        c += nl * growth
        if bi != len(blocks) - 1:
            c //= 2
    return c


def _linear_block(prefix: str, cin: int, cout: int, batch_norm: bool, prelu: bool) -> Iterator[TensorSpec]:
    """LinearBlock (prong_feature_embedding.py:7-33): linear / norm / activation attribute names."""
    yield from _linear(prefix + ".linear", cout, cin, bias=not batch_norm)
    if batch_norm:
        yield from _bn(prefix + ".norm", cout)
    if prelu:
        yield from _prelu(prefix + ".activation", cout)


def embedding_dims(options) -> Tuple[int, int, int]:
    """(pixel, feature, position) embedding widths, each rounded to a multiple of 8."""
    return (round_channels(options.pixel_embedding_dim, 8),
            round_channels(options.feature_embedding_dim, 8),
            round_channels(options.position_embedding_dim, 8))


def feature_embedding_widths(options, out_dim: int) -> List[int]:
    """Widths after each LinearBlock of the (dead) smart-feature embedding."""
    widths = [options.initial_feature_dim]
    cur = options.initial_feature_dim
    for _ in range(options.num_embedding_layers):
        nxt = 2 * cur
        if nxt >= out_dim:
            break
        widths.append(nxt)
        cur = nxt
    widths.append(out_dim)
    return widths


def prong_decoder_widths(options) -> List[int]:
    """Hidden widths of the prong-class MLP (prong_target_decoder.py:19-34)."""
    widths = []
    cur = options.hidden_dim
    for _ in range(options.num_prong_decoder_layers):
        nxt = cur // 2
        if nxt < 8:
            break
        widths.append(nxt)
        cur = nxt
    return widths


def network_specs(options, features_dim: int, extra_dim: int, pixel_dim: int,
                  num_prong_classes: int, num_event_classes: int, cnn: str = "dense") -> List[TensorSpec]:
    """All tensors of ``NeutrinoDenseNetwork`` (``cnn="sdxl"``: ``NeutrinoSDXLNetwork``) in state_dict order."""
    pix, feat, pos = embedding_dims(options)
    hidden = options.hidden_dim
    bn1d = bool(options.linear_batch_norm)
    prelu = bool(options.linear_prelu_activation)
    blocks = tuple(options.densenet_structure)
    out: List[TensorSpec] = []
    pe = "prong_embedding."
    out.append(TensorSpec(pe + "event_position_embedding", (1, pos), "pos", True))
    out.append(TensorSpec(pe + "prong_position_embedding", (1, pos), "pos", True))
    # smart-feature embedding: present in the state_dict, never used when disable_smart_features
    cin = features_dim + extra_dim
    for i, w in enumerate(feature_embedding_widths(options, feat)):
        out.extend(_linear_block(f"{pe}feature_embedding.embedding.{i}", cin, w, bn1d, prelu))
        cin = w
    cnn_in = pixel_dim * 256 if options.one_hot_pixels else pixel_dim
    for name, width in (("prong_pixel_embedding.", pix), ("event_pixel_embedding.", pix + feat)):
        if cnn == "sdxl":
            out.extend(sdxl_specs(pe + name, cnn_in, width, options.initial_pixel_dim))
        else:
            out.extend(densenet_specs(pe + name, cnn_in, width, options.initial_pixel_dim,
                                      options.densenet_growth_rate, options.densenet_batch_norm_size, blocks))
    out.extend(_linear_block(pe + "combined_embedding", feat + pix + pos, hidden, bn1d, prelu))
    for li in range(options.num_encoder_layers):
        p = f"encoder.encoder.layers.{li}."
        out.append(TensorSpec(p + "self_attn.in_proj_weight", (3 * hidden, hidden), "attn_w", True))
        out.append(TensorSpec(p + "self_attn.in_proj_bias", (3 * hidden,), "attn_b", True))
        out.extend(_linear(p + "self_attn.out_proj", hidden, hidden, True))
        out.extend(_linear(p + "linear1", hidden, hidden, True))
        out.extend(_linear(p + "linear2", hidden, hidden, True))
        for n in ("norm1", "norm2"):
            out.append(TensorSpec(p + n + ".weight", (hidden,), "ln_w", True))
            out.append(TensorSpec(p + n + ".bias", (hidden,), "ln_b", True))
    out.extend(_linear("event_decoder.hidden_layer", num_event_classes, hidden, True))
    stride = 1 + int(bn1d) + 1 + int(options.dropout > 0.0)
    cin = hidden
    widths = prong_decoder_widths(options)
    for i, w in enumerate(widths):
        base = i * stride
        out.extend(_linear(f"prong_decoder.hidden_layers.{base}", w, cin, True))
        k = base + 1
        if bn1d:
            out.extend(_bn(f"prong_decoder.hidden_layers.{k}", w))
            k += 1
        if prelu:
            out.extend(_prelu(f"prong_decoder.hidden_layers.{k}", w))
        cin = w
    out.extend(_linear("prong_decoder.output_layer", num_prong_classes, cin, True))
    return out


def arena_offsets(specs: Sequence[TensorSpec]) -> Tuple[dict, int]:
    """Float offset of every fp32 tensor in the flat arena; returns (offsets, total)."""
    off = {}
    cur = 0
    for s in specs:
        if s.in_arena:
            off[s.name] = cur
            cur += s.numel
    return off, cur
