"""Drop-in news-encoder plugin classes: ``LIME`` and ``CROWN`` (+ the sub-modules that own their
parameters), selected by the reference's unchanged flags
``--news_encoder=LIME --content_encoder=CROWN`` (reference model.py:15-17,39).

Contract kept from the reference: constructor signatures, the attributes other
modules read (``news_embedding_dim``, ``auxiliary_loss``, ``category_embedding``,
``subCategory_embedding``, ``category_affine``), ``initialize()``, the 11-argument ``forward`` and
the exact ``state_dict`` keys/shapes (Appendix A), including parameters that exist in checkpoints
but never take part in ``forward`` (ISAB, ``affine``, ``category_predictor`` at alpha = 0).
``forward`` runs on the B200 path (engine.NewsEncoderEngine -> liblime_b200.so); there is no
PyTorch implementation behind it and no CPU fallback.
"""
from __future__ import annotations

import math
import os
import pickle

import torch
import torch.nn as nn
from torch.nn import TransformerEncoder, TransformerEncoderLayer

from . import _lib
from .attn_modules import Attention, xavier_


def _require_eval(module, what):
    if module.training:
        raise NotImplementedError(
            what + ": the B200 path implements the eval forward; the training forward/backward "
            "(dropout + autograd) is not built yet — call model.eval() (no PyTorch fallback exists)")


class PositionalEncoding(nn.Module):
    """Sinusoidal table registered as buffer ``pe`` [1, max_len, d] (reference newsEncoders.py:806-828);
    the add is fused into lime_embed_pe."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        table = torch.zeros(max_len, d_model)
        table[:, 0::2] = torch.sin(pos * div)
        table[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", table.unsqueeze(0))


class CategoryPredictor(nn.Module):
    """Auxiliary category classifier (reference newsEncoders.py:375-393).  Its loss is multiplied by
    config.alpha = 0.0 (:259,364; config.py:81), so only the parameters are kept."""

    def __init__(self, title_embedding, category_num):
        super().__init__()
        self.fc = nn.Linear(title_embedding, category_num)


class MAB(nn.Module):
    """Parameters of a set-attention block (reference newsEncoders.py:395-406); never called."""

    def __init__(self, dim_Q, dim_K, dim_V, num_heads, ln=False):
        super().__init__()
        self.dim_V, self.num_heads = dim_V, num_heads
        self.fc_q = nn.Linear(dim_Q, dim_V)
        self.fc_k = nn.Linear(dim_K, dim_V)
        self.fc_v = nn.Linear(dim_K, dim_V)
        if ln:
            self.ln0 = nn.LayerNorm(dim_V)
            self.ln1 = nn.LayerNorm(dim_V)
        self.fc_o = nn.Linear(dim_V, dim_V)


class ISAB(nn.Module):
    """Constructed by CROWN (reference newsEncoders.py:250-254, 424-430) but never called: dead
    parameters that checkpoints contain."""

    def __init__(self, dim_in, dim_out, num_heads, num_inds, ln=False):
        super().__init__()
        self.I = nn.Parameter(torch.empty(1, num_inds, dim_out))
        nn.init.xavier_uniform_(self.I)
        self.mab0 = MAB(dim_out, dim_in, dim_out, num_heads, ln=ln)
        self.mab1 = MAB(dim_in, dim_out, dim_out, num_heads, ln=ln)


class NewsEncoder(nn.Module):
    """Base class, reference newsEncoders.py:167-225: trainable word table (no padding_idx), frozen
    category / sub-category tables, and the unused ``affine``."""

    def __init__(self, config):
        super().__init__()
        self.word_embedding_dim = config.word_embedding_dim
        self.category_num = config.category_num
        self.word_embedding = nn.Embedding(config.vocabulary_size, config.word_embedding_dim)
        self._load_pretrained_words(config)
        self.category_embedding = nn.Embedding(config.category_num, config.category_embedding_dim)
        self.category_embedding.weight.requires_grad = False
        self.subCategory_embedding = nn.Embedding(config.subCategory_num, config.subCategory_embedding_dim)
        self.subCategory_embedding.weight.requires_grad = False
        self.dropout = nn.Dropout(p=config.dropout_rate, inplace=True)
        self.dropout_ = nn.Dropout(p=config.dropout_rate, inplace=False)
        self.auxiliary_loss = None
        self.affine = nn.Linear(config.word_embedding_dim, config.word_embedding_dim, bias=True)

    def _load_pretrained_words(self, config):
        """The reference unpickles ``word_embedding-<thr>-<dim>-<tok>-<T>-<L>-<dataset>.pkl`` from the
        working directory (newsEncoders.py:173-174) and fails without it.  Same here, except that a
        harness may hand the tensor over as ``config.word_embedding_init`` (synthetic runs)."""
        init = getattr(config, "word_embedding_init", None)
        if init is None:
            name = "word_embedding-%s-%s-%s-%s-%s-%s.pkl" % (
                config.word_threshold, config.word_embedding_dim, config.tokenizer,
                config.max_title_length, config.max_abstract_length, config.dataset)
            with open(name, "rb") as f:                      # FileNotFoundError like the reference
                init = pickle.load(f)
        if not isinstance(init, str):
            self.word_embedding.weight.data.copy_(torch.as_tensor(init))

    def initialize(self):   # reference newsEncoders.py:193-198
        nn.init.uniform_(self.category_embedding.weight, -0.1, 0.1)
        nn.init.uniform_(self.subCategory_embedding.weight, -0.1, 0.1)
        nn.init.zeros_(self.subCategory_embedding.weight[0])
        xavier_(self.affine)

    def forward(self, title_text, title_mask, title_entity, content_text, content_mask, content_entity,
                category, subCategory, user_embedding, news_freshness, news_user_topic_lifetime):
        raise Exception("Function forward must be implemented at sub-class")


class CROWN(NewsEncoder):
    """Content encoder, reference newsEncoders.py:228-373: one post-LN transformer layer over the 32
    title and 128 body tokens (no padding mask), unmasked mean pool, k category-aware intent FCs +
    additive attention, cosine title/body gate, category features -> 900-d."""

    def __init__(self, config):
        super().__init__(config)
        self.config = config
        self.max_title_length = config.max_title_length
        self.max_body_length = config.max_abstract_length
        self.max_history_num = config.max_history_num
        self.category_embedding_dim = config.category_embedding_dim
        self.intent_embedding_dim = config.intent_embedding_dim
        self.category_embedding = nn.Embedding(config.category_num, config.category_embedding_dim)  # trainable copy (:237)
        self.news_embedding_dim = (config.intent_embedding_dim * 2 + config.category_embedding_dim
                                   + config.subCategory_embedding_dim)
        d, p = config.word_embedding_dim, config.dropout_rate
        self.title_pos_encoder = PositionalEncoding(d, p, config.max_title_length)
        self.body_pos_encoder = PositionalEncoding(d, p, config.max_abstract_length)
        self.title_transformer = TransformerEncoder(
            TransformerEncoderLayer(d, config.head_num, config.feedforward_dim, p, batch_first=True),
            config.num_layers)
        self.body_transformer = TransformerEncoder(
            TransformerEncoderLayer(d, config.head_num, config.feedforward_dim, p, batch_first=True),
            config.num_layers)
        self.ISAB = ISAB(dim_in=d, dim_out=d, num_heads=config.isab_num_heads,
                         num_inds=config.isab_num_inds, ln=True)
        self.category_affine = nn.Linear(config.category_embedding_dim + config.subCategory_embedding_dim,
                                         config.category_embedding_dim)
        self.intent_num = config.intent_num
        self.alpha = config.alpha
        self.title_intent_attention = Attention(config.intent_embedding_dim, config.attention_dim)
        self.body_intent_attention = Attention(config.intent_embedding_dim, config.attention_dim)
        self.intent_layers = nn.ModuleList(
            [nn.Linear(d + config.category_embedding_dim, config.intent_embedding_dim, bias=True)
             for _ in range(self.intent_num)])
        self.category_predictor = CategoryPredictor(config.intent_embedding_dim, config.category_num)
        self._standalone_engine = None

    def initialize(self):   # reference newsEncoders.py:271-281
        super().initialize()
        self.title_intent_attention.initialize()
        self.body_intent_attention.initialize()
        xavier_(self.category_affine)
        for lin in self.intent_layers:
            xavier_(lin)
        nn.init.uniform_(self.category_embedding.weight, -0.1, 0.1)

    def forward(self, title_text, title_mask, title_entity, content_text, content_mask, content_entity,
                category, subCategory, user_embedding, news_freshness, news_user_topic_lifetime):
        """[B,n,T] / [B,n,L] int32 ids, [B,n] int32 categories -> [B,n,900].  Masks / entities /
        user_embedding / lifetimes are accepted and ignored exactly as the reference does (:307-308)."""
        _require_eval(self, "CROWN.forward")
        if self.alpha != 0.0:
            raise NotImplementedError("category-predictor auxiliary loss (alpha != 0) is not on the B200 path")
        engine = getattr(self, "_engine_ref", None)
        if engine is None:
            raise _lib.LimeError("CROWN news encoder must be wrapped by LIME(config, base_news_encoder) "
                                 "(--news_encoder=LIME --content_encoder=CROWN)")
        B, n = title_text.shape[0], title_text.shape[1]
        content = engine().encode_content(
            title_text.reshape(B * n, -1).contiguous(), content_text.reshape(B * n, -1).contiguous(),
            category.reshape(-1).contiguous(), subCategory.reshape(-1).contiguous())
        self.auxiliary_loss = torch.zeros((), dtype=torch.float32, device=content.device)   # loss * alpha(=0)
        return content.view(B, n, self.news_embedding_dim)


class FreshnessEncoder(nn.Module):
    """Lifetime-aware freshness encoder, reference newsEncoders.py:38-83.  ``hidden_dim`` follows the
    reference's always-true ``fusion_method == 'add' or 'gated'`` test (:42): it is the content
    encoder's width (900), not ``lime_hidden_dim``."""

    def __init__(self, config, base_news_encoder):
        super().__init__()
        hidden_dim = base_news_encoder.news_embedding_dim
        self.num_buckets = config.num_buckets
        self.freshness_embedding = nn.Embedding(self.num_buckets, config.freshness_embedding_dim)
        self.lifetime_embedding = nn.Embedding(self.num_buckets, config.freshness_embedding_dim)
        self.dense = nn.Linear(config.freshness_embedding_dim * 2, hidden_dim)
        self.activation = nn.Tanh()

    def bucketize(self, x):
        """fp32 seconds -> int64 bucket ids, bit-exact with reference newsEncoders.py:53-58."""
        from . import ops
        return ops.bucketize(x.float(), self.num_buckets).long()


class LIME(nn.Module):
    """LIME news encoder (fusion_method='concat'), reference newsEncoders.py:87-161: content vector
    (900) || freshness vector (900) -> Linear(1800, 400).  Also owns the frozen topic tables and
    ``category_affine`` that the user encoder reads (userEncoders.py:103-105,115-117)."""

    def __init__(self, config, base_news_encoder):
        super().__init__()
        if config.fusion_method != "concat":
            raise NotImplementedError("only fusion_method='concat' (the reference default) is on the B200 path")
        self.config = config
        self.final_dim = config.lime_output_dim
        self.category_embedding = nn.Embedding(config.category_num, config.category_embedding_dim)
        self.category_embedding.weight.requires_grad = False
        self.subCategory_embedding = nn.Embedding(config.subCategory_num, config.subCategory_embedding_dim)
        self.subCategory_embedding.weight.requires_grad = False
        self.category_affine = nn.Linear(config.category_embedding_dim + config.subCategory_embedding_dim,
                                         config.category_embedding_dim)
        self.base_news_encoder = base_news_encoder
        self.freshness_encoder = FreshnessEncoder(config, base_news_encoder)
        self.fusion_method = config.fusion_method
        self.auxiliary_loss = getattr(base_news_encoder, "auxiliary_loss", None)
        content_dim = base_news_encoder.news_embedding_dim
        self.output_dim = content_dim + content_dim          # freshness_dim == content_dim (:106-107)
        if self.final_dim:
            self.project = nn.Linear(self.output_dim, self.final_dim)
            self.output_dim = self.final_dim
        else:
            raise NotImplementedError("lime_output_dim=0 (identity projection) is not on the B200 path")
        self.news_embedding_dim = self.output_dim
        self._engine = None
        self._scoring = None
        import weakref
        ref = weakref.ref(self)
        base_news_encoder._engine_ref = lambda: ref().engine

    @property
    def engine(self):
        if self._engine is None:
            from .engine import NewsEncoderEngine
            self._engine = NewsEncoderEngine(self, self.config)
        return self._engine

    def initialize(self):   # reference newsEncoders.py:130-138
        if hasattr(self.base_news_encoder, "initialize"):
            self.base_news_encoder.initialize()
        xavier_(self.freshness_encoder.dense)
        nn.init.uniform_(self.category_embedding.weight, -0.1, 0.1)
        nn.init.uniform_(self.subCategory_embedding.weight, -0.1, 0.1)
        xavier_(self.category_affine)

    def forward(self, title_text, title_mask, title_entity, content_text, content_mask, content_entity,
                category, subCategory, user_embedding, news_freshness, news_user_topic_lifetime):
        """-> [B, n, 400] = project(content || freshness) (reference :140-153)."""
        from . import ops
        B, n = title_text.shape[0], title_text.shape[1]
        if self.training:
            # differentiable path (training.py): the same kernels with their backward, dropout applied
            from . import training
            self._fwd_calls = getattr(self, "_fwd_calls", 0) + 1
            i32 = torch.int32
            fr = news_freshness.reshape(B, -1).float()
            lf = news_user_topic_lifetime.reshape(B, -1).float().expand_as(fr)
            vec = training.encode_news(self, title_text.reshape(B * n, -1).to(i32).contiguous(),
                                       content_text.reshape(B * n, -1).to(i32).contiguous(),
                                       category.reshape(-1).to(i32).contiguous(), subCategory.reshape(-1).to(i32).contiguous(),
                                       fr.reshape(-1).contiguous(), lf.reshape(-1).contiguous(),
                                       seed=int(getattr(self.config, "seed", 0)) * 1000003 + self._fwd_calls * 307 + 3)
            self.auxiliary_loss = torch.zeros((), dtype=torch.float32, device=vec.device)      # category loss * alpha (= 0)
            self.base_news_encoder.auxiliary_loss = self.auxiliary_loss
            return vec.view(B, n, self.news_embedding_dim)
        content = self.base_news_encoder(title_text, title_mask, title_entity, content_text, content_mask,
                                         content_entity, category, subCategory, user_embedding,
                                         news_freshness, news_user_topic_lifetime).view(B * n, -1)
        self.auxiliary_loss = self.base_news_encoder.auxiliary_loss
        if news_freshness.dim() == 1:                        # FreshnessEncoder.forward :67-73
            news_freshness = news_freshness.unsqueeze(1)
        if news_user_topic_lifetime.dim() == 1:
            news_user_topic_lifetime = news_user_topic_lifetime.unsqueeze(1)
        if news_freshness.shape != news_user_topic_lifetime.shape:
            news_user_topic_lifetime = news_user_topic_lifetime.expand_as(news_freshness)
        P = self.engine.prepare()
        nb = self.freshness_encoder.num_buckets
        T = self.engine.freshness_table()
        idx = (ops.bucketize(news_freshness.reshape(-1).float().contiguous(), nb).long() * nb
               + ops.bucketize(news_user_topic_lifetime.reshape(-1).float().contiguous(), nb).long())
        out = ops.linear(content, P["Wc"], residual=T.index_select(0, idx))
        return out.view(B, n, self.news_embedding_dim)
