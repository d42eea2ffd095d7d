"""Drop-in user-encoder plugin class ``CROWN`` (``--user_encoder=CROWN``, reference model.py:67-68).

Keeps the reference's constructor signature, attributes, ``initialize()`` and ``state_dict`` keys
(reference userEncoders.py:16-89; PyG key names ``graph_sage.convs.0.lin_l.{weight,bias}``,
``graph_sage.convs.0.lin_r.weight``, ``lightgcn.embedding.weight``).  The arithmetic of
``forward`` (candidate-aware attention -> GraphSAGE mean aggregation -> candidate-query pooling)
returns the [B, N, D] user representation through the kernels of csrc/train_user.cu; the eval hot
path (``Model.forward`` / ``util.score_impressions``) fuses the same arithmetic with the click score
in csrc/score_tc.cu / score.cu and never materialises it.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from .attn_modules import CandidateAware_ClickedNewsAttention, xavier_


class _PygLinear(nn.Module):
    """Parameter holder with torch_geometric.nn.dense.linear.Linear's names and default init
    (Kaiming-uniform(a=sqrt(5)) weight, uniform(+-1/sqrt(fan_in)) bias)."""

    def __init__(self, in_channels, out_channels, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
            nn.init.uniform_(self.bias, -1.0 / math.sqrt(in_channels), 1.0 / math.sqrt(in_channels))
        else:
            self.register_parameter("bias", None)


class _SAGEConv(nn.Module):
    """SAGEConv(mean): out_i = lin_l(mean_j x_j) + lin_r(x_i); bias on lin_l only."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_l = _PygLinear(in_channels, out_channels, bias=True)
        self.lin_r = _PygLinear(in_channels, out_channels, bias=False)


class GraphSAGE(nn.Module):
    """torch_geometric.nn.GraphSAGE(num_layers=1) as the reference builds it (userEncoders.py:54-58):
    a single in->out SAGEConv, no activation/dropout after it.  PyG is a third-party dependency that
    the reference does not pin and that is not vendored; only the parameter layout is mirrored here,
    the algorithm (published SAGEConv semantics) is in csrc/score.cu."""

    def __init__(self, in_channels, hidden_channels, num_layers, out_channels=None, dropout=0.0):
        super().__init__()
        if num_layers != 1:
            raise NotImplementedError("the reference uses GraphSAGE(num_layers=1)")
        self.convs = nn.ModuleList([_SAGEConv(in_channels, out_channels or hidden_channels)])


class LightGCN(nn.Module):
    """Constructed by the reference (userEncoders.py:59-61) and never called; checkpoints hold its
    embedding table [batch_size*max_history_num, D]."""

    def __init__(self, num_nodes, embedding_dim, num_layers):
        super().__init__()
        self.embedding = nn.Embedding(num_nodes, embedding_dim)


class UserEncoder(nn.Module):
    def __init__(self, news_encoder, config):   # reference userEncoders.py:16-24
        super().__init__()
        self.news_embedding_dim = news_encoder.news_embedding_dim
        self.news_encoder = news_encoder
        self.device = torch.device("cuda")
        self.auxiliary_loss = None
        self.word_embedding_dim = config.word_embedding_dim
        self.batch_size = config.batch_size

    def forward(self, *args, **kwargs):
        raise Exception("Function forward must be implemented at sub-class")


class CROWN(UserEncoder):
    """reference userEncoders.py:49-175."""

    def __init__(self, news_encoder, config):
        super().__init__(news_encoder, config)
        D = self.news_embedding_dim
        self.config = config
        self.attention_dim = config.attention_dim
        self.graph_sage = GraphSAGE(in_channels=D, hidden_channels=D, num_layers=1, out_channels=D,
                                    dropout=config.dropout_rate)
        self.lightgcn = LightGCN(num_nodes=config.batch_size * config.max_history_num, embedding_dim=D,
                                 num_layers=1)
        self.user_node_embedding = nn.Parameter(torch.zeros([config.batch_size, D]))
        self.K = nn.Linear(D, self.attention_dim, bias=False)
        self.Q = nn.Linear(D, self.attention_dim, bias=True)
        self.max_history_num = config.max_history_num
        self.attention_scalar = math.sqrt(float(self.attention_dim))
        self.affine = nn.Linear(D, D, bias=True)                      # dead (userEncoders.py:171)
        self.dropout = nn.Dropout(p=config.dropout_rate, inplace=True)
        self.dropout_ = nn.Dropout(p=config.dropout_rate, inplace=False)
        self.use_candidate_aware_attn = config.use_candidate_ware_clicked_news_attention
        if self.use_candidate_aware_attn:
            self.candidate_aware_attn = CandidateAware_ClickedNewsAttention(config, news_encoder)

    def initialize(self):   # reference userEncoders.py:80-89
        nn.init.zeros_(self.user_node_embedding)
        xavier_(self.K)
        xavier_(self.Q)
        xavier_(self.affine, nn.init.calculate_gain("relu"))
        if self.use_candidate_aware_attn:
            self.candidate_aware_attn.initialize()

    def forward(self, user_title_text, user_title_mask, user_title_entity, user_content_text,
                user_content_mask, user_content_entity, category, subCategory, user_category,
                user_subCategory, user_history_mask, user_history_graph, user_history_category_mask,
                user_history_category_indices, user_embedding, candidate_news_representation,
                user_freshness, user_user_topic_lifetime):
        """-> user representation [B, N, D] (reference userEncoders.py:101-175), every step a liblime_b200 kernel
        (training.py composes them; each has its backward, so this is also the training path of a reference
        ``Model`` built over these plugin classes).  ``Model.forward`` of this package does not come through
        here in eval mode: it scores with the fused kernel, which never materialises the user vector."""
        from . import training
        B, H = user_category.shape[0], user_category.shape[1]
        N = candidate_news_representation.shape[1]
        i32 = torch.int32
        calls = getattr(self, "_fwd_calls", 0) + 1
        self._fwd_calls = calls
        seed = int(getattr(self.config, "seed", 0)) * 1000003 + calls * 211 + 7
        if self.training:
            flat = lambda t: t.reshape(B * H, -1).to(i32).contiguous()
            hist = training.encode_news(self.news_encoder, flat(user_title_text), flat(user_content_text),
                                        user_category.reshape(-1).to(i32).contiguous(),
                                        user_subCategory.reshape(-1).to(i32).contiguous(),
                                        user_freshness.reshape(-1).float().contiguous(),
                                        user_user_topic_lifetime.reshape(-1).float().contiguous(), seed)
        else:
            hist = self.news_encoder(user_title_text, user_title_mask, user_title_entity, user_content_text,
                                     user_content_mask, user_content_entity, user_category, user_subCategory,
                                     user_embedding, user_freshness, user_user_topic_lifetime)
        D = self.news_embedding_dim
        u = training.user_representation(
            self, hist.reshape(B * H, D).contiguous(), candidate_news_representation.reshape(B * N, D).contiguous(),
            user_category.reshape(-1).to(i32).contiguous(), user_subCategory.reshape(-1).to(i32).contiguous(),
            category.reshape(-1).to(i32).contiguous(), subCategory.reshape(-1).to(i32).contiguous(),
            user_history_mask.to(torch.uint8).contiguous(), B, H, N, seed)
        return u.view(B, N, D)
