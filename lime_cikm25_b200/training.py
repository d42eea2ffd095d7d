"""Training-mode forward of the drop-in model (reference model.py:151-187 with N = 1 + M candidates per
sample, trainer.py:131-146): the same math as the eval path, composed from differentiable ops whose
forward and backward are liblime_b200.so kernels (autograd.py).  No news-vector cache here: weights
change every step, so every news of the mini-batch is encoded (B * (H + N) encodes, as the reference).

Dropout: the reference drops activations at 7 sites with torch's RNG, which cannot be reproduced; this
path applies inverted dropout with its own stateless generator after the positional encoding, on the
category embeddings and on the user-node rows.  The dropouts inside nn.TransformerEncoderLayer and the
fixed p = 0.2 attention dropout of layers.py:36,74 are not applied (documented deviation; gradient
parity is tested with every p = 0).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import autograd as A
from . import ops

RELU, TANH = ops.ACT_RELU, ops.ACT_TANH


def _branch(base, ids, T, transformer, pos_encoder, heads, p, seed):
    """One transformer branch (title or body): int32 ids [n, T] -> mean-pooled features [n, 300]."""
    n = ids.shape[0]
    l = transformer.layers[0]
    d = base.word_embedding.weight.shape[1]
    pe = pos_encoder.pe.reshape(-1, d)
    x0 = A.EmbedPE.apply(base.word_embedding.weight, ids.reshape(-1).contiguous(), T, pe)
    x0 = A.dropout(x0, p, seed)
    qkv = A.linear(x0, l.self_attn.in_proj_weight, l.self_attn.in_proj_bias)
    ctx = A.MHA.apply(qkv, n, T, d, heads)
    y = A.linear(ctx, l.self_attn.out_proj.weight, l.self_attn.out_proj.bias, residual=x0)
    x1 = A.LayerNorm.apply(y, l.norm1.weight, l.norm1.bias, l.norm1.eps)
    hf = A.linear(x1, l.linear1.weight, l.linear1.bias, act=RELU)
    y2 = A.linear(hf, l.linear2.weight, l.linear2.bias, residual=x1)
    return A.LayerNormMeanPool.apply(y2, l.norm2.weight, l.norm2.bias, n, T, l.norm2.eps)


def topic_representation(owner, category, subCategory):
    """category_affine(category_embedding(c) || subCategory_embedding(s)) -> [n, 50]
    (newsEncoders.py:340-342 with CROWN's tables; userEncoders.py:103-105,115-117 with LIME's)."""
    ce = A.Gather.apply(owner.category_embedding.weight, category)
    se = A.Gather.apply(owner.subCategory_embedding.weight, subCategory)
    return A.linear(torch.cat([ce, se], dim=1), owner.category_affine.weight, owner.category_affine.bias)


def encode_news(lime, title_text, body_text, category, subCategory, freshness, lifetime, seed=0):
    """LIME(CROWN) news encoder, training mode, flat over news:
    int32 [n,32], [n,128], [n], [n], fp32 seconds [n], [n] -> fp32 [n, 400] with an autograd graph."""
    base, cfg = lime.base_news_encoder, lime.config
    p = float(cfg.dropout_rate) if lime.training else 0.0
    n = title_text.shape[0]
    heads = cfg.head_num
    ft = _branch(base, title_text, title_text.shape[1], base.title_transformer, base.title_pos_encoder, heads, p, seed + 11)
    fb = _branch(base, body_text, body_text.shape[1], base.body_transformer, base.body_pos_encoder, heads, p, seed + 23)
    t = topic_representation(base, category, subCategory)                                   # [n, 50]
    k = len(base.intent_layers)
    W = F.pad(torch.cat([lin.weight for lin in base.intent_layers], dim=0), (0, 2))         # [k*400, 352]
    b = torch.cat([lin.bias for lin in base.intent_layers], dim=0)
    pad = torch.zeros(n, 2, dtype=torch.float32, device=ft.device)
    pooled = []
    for feat, att in ((ft, base.title_intent_attention), (fb, base.body_intent_attention)):
        e = A.linear(torch.cat([feat, t, pad], dim=1), W, b, act=RELU).view(n * k, -1)      # [n*k, 400]
        pre = A.linear(e, att.affine1.weight, att.affine1.bias)
        pooled.append(A.IntentPool.apply(pre, e, att.affine2.weight.reshape(-1), n, k))
    content = A.ContentFuse.apply(pooled[0], pooled[1], base.category_embedding.weight,
                                  base.subCategory_embedding.weight, category, subCategory, p, seed + 37)
    fe = lime.freshness_encoder
    nb = fe.num_buckets
    pairs = A.BucketPairs.apply(fe.freshness_embedding.weight, fe.lifetime_embedding.weight)
    table = A.linear(pairs, fe.dense.weight, fe.dense.bias, act=TANH)                       # [nb*nb, 900]
    idx = (ops.bucketize(freshness.reshape(-1).float().contiguous(), nb) * nb
           + ops.bucketize(lifetime.reshape(-1).float().contiguous(), nb)).to(torch.int32)
    fresh = A.Gather.apply(table, idx)
    cd = content.shape[1]
    pw = lime.project.weight
    return A.linear(fresh, pw[:, cd:], None, residual=A.linear(content, pw[:, :cd], lime.project.bias))
