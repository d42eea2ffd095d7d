"""Drop-in ``Model`` (reference model.py:11-187) for the LIME-CROWN-CROWN configuration.

Same constructor (``Model(config)``), attributes (``news_encoder``, ``user_encoder``,
``model_name``, ``news_embedding_dim``, ``config``), ``initialize()`` and 26-tensor ``forward``
returning logits ``[B, N]``.  The plugin classes are looked up by the reference's unchanged flag
values; any other encoder name is rejected (those encoders are out of the hot path's scope).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from . import news_modules as newsEncoders
from . import user_modules as userEncoders
from .engine import DeviceImpressions, ScoringEngine
from .util import RemainingLifetimeWeighting


def _rank():
    """Data-parallel rank (0 without a process group): mixed into the dropout seed so the ranks draw different masks."""
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class Model(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        if config.news_encoder == "LIME":                                    # model.py:15-39
            if config.content_encoder == "CROWN":
                base_encoder = newsEncoders.CROWN(config)
            else:
                raise ValueError("Unknown content encoder: %s (only CROWN is on the B200 path)"
                                 % config.content_encoder)
            self.news_encoder = newsEncoders.LIME(config=config, base_news_encoder=base_encoder)
        else:
            raise Exception(config.news_encoder + "is not implemented")      # model.py:64
        if config.user_encoder == "CROWN":                                   # model.py:67-68
            self.user_encoder = userEncoders.CROWN(self.news_encoder, config)
        else:
            raise Exception(config.user_encoder + "is not implemented")      # model.py:90
        self.model_name = "%s-%s-%s" % (config.news_encoder, config.content_encoder, config.user_encoder)
        self.news_embedding_dim = self.news_encoder.news_embedding_dim
        self.dropout = nn.Dropout(p=config.dropout_rate)
        self.use_user_embedding = False                                      # model.py:105-106
        if config.click_predictor != "dot_product":
            raise NotImplementedError("only click_predictor='dot_product' is on the B200 path")
        self.click_predictor = config.click_predictor
        self.remaining_lifetime_weighting = RemainingLifetimeWeighting(config)
        self._scoring = None

    @property
    def scoring(self):
        if self._scoring is None:
            self._scoring = ScoringEngine(self)
        return self._scoring

    def initialize(self):                                                    # model.py:133-145
        self.news_encoder.initialize()
        self.user_encoder.initialize()
        self.remaining_lifetime_weighting.initialize()

    def forward(self, user_ID, user_category, user_subCategory, user_title_text, user_title_mask,
                user_title_entity, user_content_text, user_content_mask, user_content_entity,
                user_freshness, user_user_topic_lifetime, user_history_mask, user_history_graph,
                user_history_category_mask, user_history_category_indices, news_category,
                news_subCategory, news_title_text, news_title_mask, news_title_entity,
                news_content_text, news_content_mask, news_content_entity, news_freshness,
                news_user_topic_lifetime, remaining_lifetime):
        """Training mode: differentiable path (training.py), candidates [B, N, ...] -> logits [B, N].
        Eval layout (model.py:158-169): candidate tensors carry no news dim; returns [B, 1].

        The B*(H+1) news of the batch are encoded once (Stage A kernels) into a batch-local vector
        cache, then one fused kernel (Stage B) scores the B pairs.  The GraphSAGE prefix is the
        runtime batch size B, exactly as in the reference (userEncoders.py:91-98,153)."""
        if self.training:
            # training layout (model.py:171-181): N = 1 + M candidates per sample, differentiable path
            if news_category.dim() != 2:
                raise _lib.LimeError("training-mode Model.forward expects candidate tensors [B, N, ...] (dataset.py:105-141)")
            from . import training
            self._train_calls = getattr(self, "_train_calls", 0) + 1
            logits = training.model_forward(
                self, user_category, user_subCategory, user_title_text, user_content_text, user_freshness,
                user_user_topic_lifetime, user_history_mask, news_category, news_subCategory, news_title_text,
                news_content_text, news_freshness, news_user_topic_lifetime, remaining_lifetime,
                seed=int(getattr(self.config, "seed", 0)) * 1000003 + self._train_calls * 101 + _rank() * 7919)
            self.news_encoder.auxiliary_loss = torch.zeros((), device=logits.device)   # category loss * alpha (= 0)
            self.news_encoder.base_news_encoder.auxiliary_loss = self.news_encoder.auxiliary_loss
            return logits
        if news_category.dim() != 1:
            raise _lib.LimeError("eval-mode Model.forward expects one candidate per sample (model.py:158-169)")
        B, H = user_category.shape
        se = self.scoring
        i32 = torch.int32
        title = torch.cat([user_title_text.reshape(B * H, -1), news_title_text.reshape(B, -1)]).to(i32).contiguous()
        body = torch.cat([user_content_text.reshape(B * H, -1), news_content_text.reshape(B, -1)]).to(i32).contiguous()
        cat = torch.cat([user_category.reshape(-1), news_category.reshape(-1)]).to(i32).contiguous()
        sub = torch.cat([user_subCategory.reshape(-1), news_subCategory.reshape(-1)]).to(i32).contiguous()
        hist_rows, cand_rows = se.build_rows(title, body, cat, sub)
        dimp = DeviceImpressions.from_pairs(user_history_mask, user_freshness, user_user_topic_lifetime,
                                            news_freshness, news_user_topic_lifetime, remaining_lifetime, 1)
        self.news_encoder.auxiliary_loss = torch.zeros((), device=hist_rows.device)   # category loss * alpha(=0)
        self.news_encoder.base_news_encoder.auxiliary_loss = self.news_encoder.auxiliary_loss
        scores = se.score(hist_rows, cand_rows, dimp, prefix_main=B)
        return scores.view(B, 1)
