"""Hyper-parameters of the TransformerCVN hot path.

The reference keeps these in an ``Options`` namespace (reference:
transformercvn/options.py:21-162) filled from a JSON file
(option_files/fdhd_beam_2018prod_aiml_tutorial_2025_04_21.json).  The drop-in
network accepts that object unchanged (duck-typed attribute access); this
module only provides a stand-alone equivalent for benches/tests that run where
the reference tree is absent (the GPU box).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field, asdict
from typing import List


def round_channels(v: float, divisor: int = 8) -> int:
    """Round a channel count to a multiple of ``divisor`` (never by more than -10 %).

    Same arithmetic as the reference's ``make_divisible_channel_count``
    (transformercvn/network/layers/prong_masked_mobilenet_embedding.py:10-23).
    """
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


@dataclass
class PathOptions:
    """Attribute names follow the reference's ``Options`` so either can be passed."""
    hidden_dim: int = 128
    initial_feature_dim: int = 8
    initial_pixel_dim: int = 64
    final_decoder_dim: int = 16
    feature_embedding_dim: int = 32
    pixel_embedding_dim: int = 256
    position_embedding_dim: int = 32
    num_embedding_layers: int = 100
    num_encoder_layers: int = 6
    num_prong_decoder_layers: int = 4
    num_attention_heads: int = 8
    transformer_activation: str = "gelu"
    transformer_norm_first: bool = False
    linear_prelu_activation: bool = True
    linear_batch_norm: bool = True
    disable_smart_features: bool = True
    normalize_features: bool = True
    one_hot_pixels: bool = False
    log_pixels: bool = False
    densenet_structure: List[int] = field(default_factory=lambda: [3, 6, 12, 6, 3])
    densenet_growth_rate: int = 32
    densenet_batch_norm_size: int = 4
    pixel_noise_std: float = 0.001
    dropout: float = 0.1
    optimizer: str = "AdamW"
    event_prong_loss_proportion: float = 0.9
    learning_rate: float = 0.0000075665331738495
    l2_penalty: float = 2.1300208077466406e-05
    gradient_clip: float = 43.0
    loss_gamma: float = 1.0

    @classmethod
    def tutorial(cls) -> "PathOptions":
        """The DenseNet TransformerCVN tutorial configuration (BASELINE.json configs[0..2])."""
        return cls()

    @classmethod
    def load(cls, path: str) -> "PathOptions":
        with open(path) as f:
            raw = json.load(f)
        known = {k: v for k, v in raw.items() if k in cls.__dataclass_fields__}
        return cls(**known)

    def to_json(self) -> str:
        return json.dumps(asdict(self), indent=2)


# Fixed geometry of the pixel maps (reference README.md:82-95; full_pixels_shape = (3, 400, 280)).
PIXEL_CHANNELS = 3
PIXEL_H = 400
PIXEL_W = 280
NUM_EVENT_CLASSES = 4
NUM_PRONG_CLASSES = 8
