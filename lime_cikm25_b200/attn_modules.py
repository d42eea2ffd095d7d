"""Parameter containers for the two ``layers.py`` classes of the reference that sit on the hot path.

They keep the reference's constructor signatures, attribute names and ``state_dict`` keys
(checkpoints load strictly).  Their arithmetic lives in liblime_b200.so and is
driven by ``engine.py``; calling them directly is not part of the reference's public surface.
"""
from __future__ import annotations

import torch.nn as nn


def xavier_(linear, gain=1.0):
    """xavier-uniform weight + zero bias, the reference's initialisation for every Linear it owns."""
    nn.init.xavier_uniform_(linear.weight, gain=gain)
    if linear.bias is not None:
        nn.init.zeros_(linear.bias)


class Attention(nn.Module):
    """Additive attention over the k intents, reference layers.py:269-300.
    Evaluated by lime_linear (affine1) + lime_intent_pool (tanh, affine2, softmax, weighted sum)."""

    def __init__(self, feature_dim, attention_dim):
        super().__init__()
        self.affine1 = nn.Linear(feature_dim, attention_dim, bias=True)
        self.affine2 = nn.Linear(attention_dim, 1, bias=False)

    def initialize(self):   # layers.py:275-278
        xavier_(self.affine1, nn.init.calculate_gain("tanh"))
        xavier_(self.affine2)


class CandidateAware_ClickedNewsAttention(nn.Module):
    """Candidate-aware lifetime attention, reference layers.py:15-93.  Fused into csrc/score.cu
    (phase 1: topic attention weights, phase 2: gated residual + LayerNorm).  ``value_proj`` exists
    in checkpoints but the reference discards its output (:68,76-77): dead parameter, kept."""

    num_heads = 10      # layers.py:22

    def __init__(self, config, news_encoder):
        super().__init__()
        D = news_encoder.news_embedding_dim
        if D % self.num_heads:
            raise AssertionError("embedding_dim must be divisible by num_heads")
        self.news_embedding_dim = D
        self.topic_embedding_dim = config.category_embedding_dim
        self.use_residual_connection = config.use_residual_connection
        self.head_dim = D // self.num_heads
        self.scale = D ** 0.5                                   # layers.py:35 (sqrt(D), not sqrt(head_dim))
        self.query_proj = nn.Linear(self.topic_embedding_dim, D)
        self.key_proj = nn.Linear(self.topic_embedding_dim, D)
        self.value_proj = nn.Linear(D, D)
        self.dropout = nn.Dropout(p=0.2)                        # layers.py:36, fixed p
        self.gate_proj = nn.Linear(D, D)
        self.layernorm = nn.LayerNorm(D)

    def initialize(self):   # layers.py:42-50
        for lin in (self.query_proj, self.key_proj, self.value_proj, self.gate_proj):
            xavier_(lin)
