"""TEST INFRASTRUCTURE — CPU restatement (torch, fp32 or fp64) of the reference hot path.

This is the oracle the CUDA path is checked against.  It is a functional re-derivation over a
plain ``state_dict`` (no nn.Module of the reference is used), pinned against outputs of the
unmodified reference by ``oracle/make_golden.py`` -> ``tests/golden/*.pt`` and
``tests/test_oracle_golden.py``.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs may import it; the product never does.

Each function cites the reference lines it follows (paths relative to /root/reference).
Dropout is the identity here (eval, or train with p=0): stochastic masks cannot be pinned.
Device-agnostic (tensors stay on the device of the inputs): bench.py also times it with CUDA tensors as the
"PyTorch eager on the same B200" baseline BASELINE.md asks for.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LN_EPS = 1e-5
# bench.py's "PyTorch eager on the same B200" baseline sets this: BatchNorm / PReLU through the fused ATen ops the
# reference's nn.BatchNorm2d / nn.PReLU modules call (same arithmetic; the elementwise restatement below is the oracle)
FUSED_ATEN = False


# ----------------------------------------------------------------------------- ingest
def preprocess_values(values: torch.Tensor) -> torch.Tensor:
    """transformercvn/network/trainers/neutrino_full_dense_trainer.py:59-60 (the /255.0 branch)."""
    return values / 255.0


def densify(values: torch.Tensor, coords: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """COO -> dense NCHW.  neutrino_full_dense_trainer.py:15-24.

    N is taken from the last hit's image index (coords are sorted by image); hits are unique so
    the indexed ``+=`` of the reference is a plain scatter onto zeros.
    """
    c = coords.long()
    n = int(c[-1, 0]) + 1
    out = torch.zeros(n, values.shape[1], h, w, dtype=values.dtype, device=values.device)
    out[c[:, 0], :, c[:, 1], c[:, 2]] = values
    return out


def collate_sparse(coordinates, values, masks):
    """dataset/minkowski_dataset.py:34-47 (MinkowskiCollection.collate_sparse): column 0 of every event's coordinates
    is offset by the number of images (mask.sum()) of all earlier events; lists are concatenated in event order."""
    out, total = [], 0
    for coord, mask in zip(coordinates, masks):
        c = coord.clone()
        c[:, 0] += total
        out.append(c)
        total += int(mask.sum())
    return torch.cat(out), torch.cat(list(values))


# ----------------------------------------------------------------------------- DenseNet
class Stats:
    """Collects updated BN running statistics in train mode (momentum 0.1, unbiased var)."""

    def __init__(self):
        self.updated: Dict[str, torch.Tensor] = {}


def _bn_prelu(state, bn: str, act: str, x: torch.Tensor, train: bool, stats: Optional[Stats]) -> torch.Tensor:
    """BatchNorm (torch defaults) followed by per-channel PReLU.  dense_net.py:19-20,30-31,85-86."""
    w, b = state[bn + ".weight"], state[bn + ".bias"]
    rm, rv = state[bn + ".running_mean"], state[bn + ".running_var"]
    if FUSED_ATEN and stats is None:
        y = F.batch_norm(x, rm.detach().clone() if train else rm, rv.detach().clone() if train else rv, w, b, training=train,
                         momentum=BN_MOMENTUM, eps=BN_EPS)
        return F.prelu(y, state[act + ".weight"])
    shape = [1, -1] + [1] * (x.dim() - 2)
    if train:
        dims = [0] + list(range(2, x.dim()))
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        if stats is not None:
            n = x.numel() // x.shape[1]
            stats.updated[bn + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
            stats.updated[bn + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
    else:
        mean, var = rm, rv
    y = (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS) * w.view(shape) + b.view(shape)
    a = state[act + ".weight"].view(shape)
    return torch.where(y >= 0, y, a * y)


def _conv(state, name: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    return F.conv2d(x, state[name + ".weight"], state[name + ".bias"], stride=stride, padding=padding)


def _drop(x: torch.Tensor, p: float, train: bool) -> torch.Tensor:
    """nn.Dropout at the reference's sites (dense_net.py:39,161; prong_feature_embedding.py:23; encoder.py:21-22; the four
    sites of nn.TransformerEncoderLayer).  p_drop = 0 (every parity test) is the identity."""
    return F.dropout(x, p, training=True) if (train and p > 0.0) else x


def densenet_forward(state, prefix: str, x: torch.Tensor, blocks, growth: int = 32, train: bool = False,
                     stats: Optional[Stats] = None, taps: Optional[Dict[str, torch.Tensor]] = None,
                     p_drop: float = 0.0) -> torch.Tensor:
    """One pixel-map CNN.  dense_net.py:97-167 (stem :111-122, Bottleneck :8-45, Transition :78-94,
    tail :147-162).  ``taps`` receives named intermediates (NCHW) for per-stage parity checks."""
    f = prefix + "features."
    x = _conv(state, f + "conv0", x, stride=2, padding=3)
    x = _bn_prelu(state, f + "norm0", f + "relu0", x, train, stats)
    if taps is not None:
        taps["stem_act"] = x
    x = F.avg_pool2d(x, kernel_size=3, stride=2)
    if taps is not None:
        taps["stem_pool"] = x
    for bi, nl in enumerate(blocks):
        for li in range(nl):
            p = f"{f}dense{bi + 1}.layers.{li}."
            y = _bn_prelu(state, p + "bottleneck_block.norm1", p + "bottleneck_block.relu1", x, train, stats)
            y = _conv(state, p + "bottleneck_block.conv1", y)
            y = _bn_prelu(state, p + "output_block.norm2", p + "output_block.relu2", y, train, stats)
            y = _drop(_conv(state, p + "output_block.conv2", y, padding=1), p_drop, train)
            x = torch.cat((x, y), dim=1)
        if taps is not None:
            taps[f"dense{bi + 1}"] = x
        if bi != len(blocks) - 1:
            t = f"{f}transition{bi + 1}."
            x = _bn_prelu(state, t + "norm", t + "relu", x, train, stats)
            x = _conv(state, t + "conv", x)
            x = F.avg_pool2d(x, kernel_size=2, stride=2)
            if taps is not None:
                taps[f"transition{bi + 1}"] = x
    x = _bn_prelu(state, f + "final_norm", f + "final_relu", x, train, stats)
    x = x.mean(dim=(2, 3))
    o = prefix + "output_block."
    x = x @ state[o + "linear.weight"].t()
    x = _drop(_bn_prelu(state, o + "norm", o + "relu", x, train, stats), p_drop, train)
    if taps is not None:
        taps["embedding"] = x
    return x


# ----------------------------------------------------------------------------- token assembly
def pack_indices(mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(I1, I2) = event index and slot index of every valid prong.  layers/packed_data.py:59-66."""
    nz = mask.nonzero()
    return nz[:, 0], nz[:, 1]


def tokens_forward(state, event_emb: torch.Tensor, prong_emb: torch.Tensor, prong_mask: torch.Tensor,
                   event_mask: torch.Tensor, feature_dim: int, train: bool = False,
                   stats: Optional[Stats] = None, p_drop: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """networks/neutrino_full_base_network.py:99-125 with disable_smart_features=True
    (prong_feature_embedding.py:75-76 returns zeros).  NOTE :107 — prong rows get the *event*
    position embedding; ``prong_position_embedding`` is never read."""
    pe = "prong_embedding."
    b, l = prong_mask.shape
    pos = state[pe + "event_position_embedding"]
    ev = torch.cat((event_emb, pos.expand(b, -1)), dim=1)
    t = prong_emb.shape[0]
    pr = torch.cat((torch.zeros(t, feature_dim, dtype=prong_emb.dtype, device=prong_emb.device), prong_emb, pos.expand(t, -1)), dim=1)
    rows = torch.cat((ev, pr), dim=0) @ state[pe + "combined_embedding.linear.weight"].t()
    rows = _drop(_bn_prelu(state, pe + "combined_embedding.norm", pe + "combined_embedding.activation", rows, train, stats),
                 p_drop, train)
    i1, i2 = pack_indices(prong_mask)
    padded = torch.zeros(b, l, rows.shape[1], dtype=rows.dtype, device=rows.device)
    padded[i1, i2] = rows[b:]
    tokens = torch.cat((rows[:b].unsqueeze(1), padded), dim=1)
    return tokens, torch.cat((event_mask, prong_mask), dim=1)


# ----------------------------------------------------------------------------- encoder
def _layer_norm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


def encoder_forward(state, tokens: torch.Tensor, mask: torch.Tensor, num_layers: int, num_heads: int, train: bool = False,
                    p_drop: float = 0.0) -> torch.Tensor:
    """layers/prong_custom_bert_encoder.py:57-75 around torch's post-norm TransformerEncoderLayer
    (:45-54): x=LN1(x+Wo.MHA(x)); x=LN2(x+W2.gelu_erf(W1 x)).  Keys of padded slots get -inf; the
    input and the output are multiplied by the mask.  Returns (S,B,D) like the reference."""
    b, s, d = tokens.shape
    dh = d // num_heads
    m = mask.unsqueeze(-1).to(tokens.dtype)
    x = tokens * m
    neg = torch.zeros(b, 1, 1, s, dtype=tokens.dtype, device=tokens.device).masked_fill(~mask.view(b, 1, 1, s), float("-inf"))
    for li in range(num_layers):
        p = f"encoder.encoder.layers.{li}."
        qkv = x @ state[p + "self_attn.in_proj_weight"].t() + state[p + "self_attn.in_proj_bias"]
        q, k, v = (t.view(b, s, num_heads, dh).transpose(1, 2) for t in qkv.split(d, dim=-1))
        att = _drop(torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh) + neg, dim=-1), p_drop, train)
        ctx = (att @ v).transpose(1, 2).reshape(b, s, d)
        ctx = ctx @ state[p + "self_attn.out_proj.weight"].t() + state[p + "self_attn.out_proj.bias"]
        x = _layer_norm(x + _drop(ctx, p_drop, train), state[p + "norm1.weight"], state[p + "norm1.bias"])
        hmid = x @ state[p + "linear1.weight"].t() + state[p + "linear1.bias"]
        hmid = _drop(0.5 * hmid * (1.0 + torch.erf(hmid / math.sqrt(2.0))), p_drop, train)
        ff = hmid @ state[p + "linear2.weight"].t() + state[p + "linear2.bias"]
        x = _layer_norm(x + _drop(ff, p_drop, train), state[p + "norm2.weight"], state[p + "norm2.bias"])
    return (x * m).transpose(0, 1).contiguous()


# ----------------------------------------------------------------------------- heads
def heads_forward(state, hidden: torch.Tensor, widths: List[int], stride: int, train: bool = False,
                  stats: Optional[Stats] = None, p_drop: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Event head: layers/prong_decoder.py:13-16 on token 0.  Prong head: prong_target_decoder.py:35-41
    over ALL L*B rows (padded rows included — they enter the BN1d batch statistics in train mode);
    caller-visible layout (B,L,C) per neutrino_full_base_network.py:186-188."""
    ev = hidden[0] @ state["event_decoder.hidden_layer.weight"].t() + state["event_decoder.hidden_layer.bias"]
    l, b, d = hidden[1:].shape
    x = hidden[1:].reshape(l * b, d)
    for i, _ in enumerate(widths):
        base = f"prong_decoder.hidden_layers.{i * stride}"
        x = x @ state[base + ".weight"].t() + state[base + ".bias"]
        x = _drop(_bn_prelu(state, f"prong_decoder.hidden_layers.{i * stride + 1}",
                            f"prong_decoder.hidden_layers.{i * stride + 2}", x, train, stats), p_drop, train)
    x = x @ state["prong_decoder.output_layer.weight"].t() + state["prong_decoder.output_layer.bias"]
    return ev, x.reshape(l, b, -1).transpose(0, 1)


# ----------------------------------------------------------------------------- whole path
def network_forward(state, options, event_pixels: torch.Tensor, event_mask: torch.Tensor,
                    prong_pixels: torch.Tensor, prong_mask: torch.Tensor, train: bool = False,
                    stats: Optional[Stats] = None, taps: Optional[dict] = None, p_drop: float = 0.0):
    """networks/neutrino_full_base_network.py:166-188."""
    from dune_transformercvn_b200.params import embedding_dims, prong_decoder_widths
    blocks = tuple(options.densenet_structure)
    g = options.densenet_growth_rate
    _, feat, _ = embedding_dims(options)
    et = {} if taps is not None else None
    pt = {} if taps is not None else None
    ev = densenet_forward(state, "prong_embedding.event_pixel_embedding.", event_pixels, blocks, g, train, stats, et, p_drop)
    pr = densenet_forward(state, "prong_embedding.prong_pixel_embedding.", prong_pixels, blocks, g, train, stats, pt, p_drop)
    tokens, mask = tokens_forward(state, ev, pr, prong_mask, event_mask, feat, train, stats, p_drop)
    hidden = encoder_forward(state, tokens, mask, options.num_encoder_layers, options.num_attention_heads, train, p_drop)
    stride = 3 + int(options.dropout > 0.0)
    ev_logits, pr_logits = heads_forward(state, hidden, prong_decoder_widths(options), stride, train, stats, p_drop)
    if taps is not None:
        taps.update(event_cnn=et, prong_cnn=pt, event_embedding=ev, prong_embedding=pr, tokens=tokens,
                    hidden=hidden)
    return ev_logits, pr_logits


def sparse_forward(state, options, batch, h: int = 400, w: int = 280, train: bool = False, taps=None, stats=None,
                   dtype=torch.float32, p_drop: float = 0.0):
    """Trainer-level path: preprocess -> densify -> network (neutrino_full_base_trainer.py:113-116).
    ``dtype``: the /255 is always done in fp32 like the reference; float64 only widens what follows."""
    ev = densify(preprocess_values(batch.event_values.float()), batch.event_coords, h, w).to(dtype)
    pr = densify(preprocess_values(batch.prong_values.float()), batch.prong_coords, h, w).to(dtype)
    return network_forward(state, options, ev, batch.event_mask, pr, batch.prong_mask, train, stats, taps, p_drop)


def focal_loss(logits: torch.Tensor, targets: torch.Tensor, gamma: float) -> torch.Tensor:
    """neutrino_full_base_trainer.py:148-160."""
    logp = torch.log_softmax(logits, dim=-1).gather(1, targets.view(-1, 1)).squeeze(1)
    p = logp.exp()
    return (-logp * (1 - p) ** gamma).mean()


def training_loss(ev_logits, pr_logits, ev_targets, pr_targets, options) -> torch.Tensor:
    """neutrino_full_base_trainer.py:162-177: 0.9*event + 0.1*prong over slots with target >= 0."""
    sel = pr_targets >= 0
    le = focal_loss(ev_logits, ev_targets, options.loss_gamma)
    lp = focal_loss(pr_logits[sel], pr_targets[sel], options.loss_gamma)
    a = options.event_prong_loss_proportion
    return a * le + (1 - a) * lp


def export_combined(state, options, pixels: torch.Tensor):
    """``DynamicCombinedNetwork.forward`` of CreateCompiled.ipynb cell 8 (log_pixels False): one event given as
    (1 + Npng, 3, H, W) maps with 0..255 values -> (event probabilities, prong probabilities, event hidden vector,
    prong hidden vectors).  Pinned on tests/golden/export.pt (oracle/make_golden_export.py)."""
    px = pixels.float() / 255
    n = px.shape[0]
    taps: dict = {}
    ev, pr = network_forward(state, options, px[:1], torch.ones(1, 1, dtype=torch.bool), px[1:],
                             torch.ones(1, n - 1, dtype=torch.bool), taps=taps)
    hidden = taps["hidden"]
    return torch.softmax(ev[0], 0), torch.softmax(pr[0], 1), hidden[0, 0], hidden[1:, 0]
