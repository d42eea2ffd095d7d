"""Attribute bag with the reference's flag defaults (reference config.py:24-115).

The reference's ``Config()`` cannot be constructed without its dataset files and a GPU assert
(config.py:212,283-293); harnesses, tests and bench.py build this plain namespace instead.  When the
drop-in classes are used from the reference's own ``main.py``, the reference's ``Config`` object is
passed through unchanged — only attribute access is needed.
"""
from __future__ import annotations

import types


def default_config(**overrides):
    c = types.SimpleNamespace(
        mode="train", news_encoder="LIME", user_encoder="CROWN", content_encoder="CROWN",
        device_id=0, seed=0, dataset="mind", tokenizer="MIND", word_threshold=3,
        max_title_length=32, max_abstract_length=128,
        negative_sample_num=4, max_history_num=50, epoch=16, batch_size=32, lr=1e-4,
        weight_decay=0, gradient_clip_norm=4, world_size=1, dev_criterion="auc",
        early_stopping_epoch=5,
        fusion_method="concat", freshness_embedding_dim=500, lime_hidden_dim=200,
        lime_output_dim=400, num_buckets=10, use_candidate_ware_clicked_news_attention=True,
        use_residual_connection=True, lifetime_type="user_topic",
        use_remaining_lifetime_weighting=True, sigmoid_scaling_alpha=0.3,
        penalty_scaling_beta=0.3, use_expired_penalty=True, fixed_lifetime=36 * 3600,
        num_layers=1, feedforward_dim=512, head_num=10, head_dim=20, intent_embedding_dim=400,
        intent_num=3, dropout_rate=0.2, attention_dim=400, word_embedding_dim=300,
        isab_num_inds=4, isab_num_heads=4, alpha=0.0, beta=0.0,
        entity_embedding_dim=100, context_embedding_dim=100, cnn_method="naive",
        cnn_kernel_num=400, cnn_window_size=3, user_embedding_dim=50,
        category_embedding_dim=50, subCategory_embedding_dim=50, hidden_dim=400,
        click_predictor="dot_product",
        # data-derived in the reference (corpus.py:309-326): harness parameters here
        vocabulary_size=1000, category_num=18, subCategory_num=270, user_num=1000, entity_size=1,
        category_lifetime_map=None,
    )
    for k, v in overrides.items():
        setattr(c, k, v)
    return c
