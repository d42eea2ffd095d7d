"""Synthetic MIND-/Adressa-shaped corpora and impressions.

The reference builds these arrays from real MIND/Adressa files in ``corpus.py:361-649`` (out of
scope: needs the datasets + GloVe).  The generators here produce arrays with the same dtypes and
layout the reference's datasets hand to the model (``corpus.py:361-368``, ``dataset.py:105-141,
192-227``): int32 ids, bool masks, fp32 seconds, history truncated to the last ``H`` items and
right-padded with news 0 / 0.0 seconds.

Everything is numpy on the host (seeded ``np.random.Generator``); nothing here touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Dataset statistics the reference states (README.md:14-15).
MIND = dict(news=65238, users=94057, title_mean=11.67, body_mean=41.01, body_full=False)
ADRESSA = dict(news=73844, users=601215, title_mean=6.63, body_mean=552.15, body_full=True)


@dataclass
class NewsTable:
    """Per-news id arrays, row 0 = the all-<PAD> news used for history padding (corpus.py:515,571)."""
    title_text: np.ndarray      # int32 [N, T]
    title_mask: np.ndarray      # bool  [N, T]
    body_text: np.ndarray       # int32 [N, L]
    body_mask: np.ndarray       # bool  [N, L]
    category: np.ndarray        # int32 [N]
    subCategory: np.ndarray     # int32 [N]
    vocabulary_size: int
    category_num: int
    subCategory_num: int

    @property
    def news_num(self):
        return int(self.title_text.shape[0])


@dataclass
class Impressions:
    """Impression-major eval set.  ``cand_*`` arrays are flat over all (impression, candidate)
    pairs in the reference's DevTest_Dataset order (``corpus.py:577-596``): pair p of impression i
    lives at ``cand_off[i] <= p < cand_off[i+1]``."""
    hist_news: np.ndarray       # int32 [I, H]   last-H history, right-padded with 0
    hist_mask: np.ndarray       # bool  [I, H]
    hist_fresh: np.ndarray      # fp32  [I, H]   seconds, 0.0 on padding (dataset.py:203-205)
    hist_life: np.ndarray       # fp32  [I, H]
    cand_off: np.ndarray        # int64 [I+1]
    cand_news: np.ndarray       # int32 [P]
    cand_fresh: np.ndarray      # fp32  [P]
    cand_life: np.ndarray       # fp32  [P]
    labels: np.ndarray          # uint8 [P]
    user_id: np.ndarray = field(default=None)   # int64 [I]

    @property
    def num_impressions(self):
        return int(self.hist_news.shape[0])

    @property
    def num_pairs(self):
        return int(self.cand_news.shape[0])

    def slice(self, lo, hi):
        """Impressions [lo, hi) as a new, re-based Impressions (used for rank sharding)."""
        p0, p1 = int(self.cand_off[lo]), int(self.cand_off[hi])
        return Impressions(self.hist_news[lo:hi], self.hist_mask[lo:hi], self.hist_fresh[lo:hi],
                           self.hist_life[lo:hi], self.cand_off[lo:hi + 1] - p0,
                           self.cand_news[p0:p1], self.cand_fresh[p0:p1], self.cand_life[p0:p1],
                           self.labels[p0:p1],
                           None if self.user_id is None else self.user_id[lo:hi])


def _lengths(rng, n, mean, lo, hi, kind):
    if kind == "poisson":
        x = rng.poisson(mean, size=n)
    else:  # long right tail (body text)
        x = np.rint(rng.gamma(shape=2.0, scale=mean / 2.0, size=n))
    return np.clip(x, lo, hi).astype(np.int64)


def make_news_table(num_news, *, vocabulary_size=40000, category_num=18, subCategory_num=270,
                    title_len=32, body_len=128, title_mean=MIND["title_mean"],
                    body_mean=MIND["body_mean"], body_full=False, seed=1) -> NewsTable:
    """``num_news`` real news + the pad row 0.  Word ids are Zipf-ish over [1, V)."""
    rng = np.random.default_rng(seed)
    n = num_news + 1
    tl = _lengths(rng, n, title_mean, 1, title_len, "poisson")
    bl = (np.full(n, body_len, np.int64) if body_full
          else _lengths(rng, n, body_mean, 1, body_len, "gamma"))
    tl[0] = 0
    bl[0] = 0

    def text(lengths, width):
        # id = floor(V^u) gives a log-uniform (Zipf-like) word distribution; 0 is reserved for <PAD>
        u = rng.random((n, width))
        ids = np.minimum((vocabulary_size ** u).astype(np.int64), vocabulary_size - 1)
        ids = np.maximum(ids, 1)
        mask = np.arange(width)[None, :] < lengths[:, None]
        return np.where(mask, ids, 0).astype(np.int32), mask

    title_text, title_mask = text(tl, title_len)
    body_text, body_mask = text(bl, body_len)
    # every sub-category belongs to exactly one category (as in MIND); Zipf-ish popularity
    sub_to_cat = rng.integers(0, category_num, size=subCategory_num)
    w = 1.0 / np.arange(1, subCategory_num + 1) ** 0.8
    sub = rng.choice(subCategory_num, size=n, p=w / w.sum())
    cat = sub_to_cat[sub]
    sub[0] = 0
    cat[0] = 0
    return NewsTable(title_text, title_mask, body_text, body_mask, cat.astype(np.int32),
                     sub.astype(np.int32), vocabulary_size, category_num, subCategory_num)


def _seconds(rng, n, lo, hi):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.float32)


def make_impressions(num_impressions, news_num, *, max_history=50, cand_mean=37.0, cand_fixed=None,
                     cand_max=300, num_users=MIND["users"], near_zero_frac=0.05, seed=1) -> Impressions:
    """History length uniform in [0, max_history] (empty included); candidates per impression from a
    long-tailed (log-normal) law with mean ``cand_mean`` clipped to [2, cand_max], or ``cand_fixed``;
    every impression has >=1 positive and >=1 negative (sklearn's AUC needs both, evaluate.py:77).
    Freshness is log-uniform on [1, 1e7] s and lifetime on [600, 6e5] s, plus a ``near_zero_frac``
    slice with |lifetime - freshness| < 60 s so the un-saturated sigmoid regime is exercised."""
    rng = np.random.default_rng(seed)
    I, H = num_impressions, max_history
    hl = rng.integers(0, H + 1, size=I)
    hist_mask = np.arange(H)[None, :] < hl[:, None]
    hist_news = np.where(hist_mask, rng.integers(1, news_num, size=(I, H)), 0).astype(np.int32)
    hist_fresh = np.where(hist_mask, _seconds(rng, I * H, 1.0, 1e7).reshape(I, H), 0).astype(np.float32)
    hist_life = np.where(hist_mask, _seconds(rng, I * H, 600.0, 6e5).reshape(I, H), 0).astype(np.float32)
    if cand_fixed is not None:
        C = np.full(I, int(cand_fixed), np.int64)
    else:
        sigma = 0.9
        mu = np.log(cand_mean) - 0.5 * sigma * sigma
        C = np.clip(np.rint(rng.lognormal(mu, sigma, size=I)), 2, cand_max).astype(np.int64)
    cand_off = np.zeros(I + 1, np.int64)
    np.cumsum(C, out=cand_off[1:])
    P = int(cand_off[-1])
    cand_news = rng.integers(1, news_num, size=P).astype(np.int32)
    cand_fresh = _seconds(rng, P, 1.0, 1e7)
    cand_life = _seconds(rng, P, 600.0, 6e5)
    near = rng.random(P) < near_zero_frac
    cand_life = np.where(near, cand_fresh + rng.uniform(-60, 60, size=P).astype(np.float32),
                         cand_life).astype(np.float32)
    # labels: ~10 % positives, then force one positive and one negative per impression
    labels = (rng.random(P) < 0.1).astype(np.uint8)
    first = cand_off[:-1]
    pos_slot = first + rng.integers(0, C)
    neg_slot = first + (pos_slot - first + 1 + rng.integers(0, np.maximum(C - 1, 1))) % C
    labels[neg_slot] = 0
    labels[pos_slot] = 1           # a single-candidate impression (C == 1) keeps its positive only
    user_id = rng.integers(0, num_users, size=I).astype(np.int64)
    return Impressions(hist_news, hist_mask, hist_fresh, hist_life, cand_off, cand_news,
                       cand_fresh, cand_life, labels, user_id)


def impressions_to_pair_batches(news: NewsTable, imp: Impressions, batch_size):
    """Yield the reference's eval mini-batches: one (user, candidate) pair per sample, batches of
    ``batch_size`` consecutive pairs, last batch short (DevTest_Dataset + DataLoader(shuffle=False),
    ``dataset.py:192-227``, ``util.py:81``).  Each batch is the 25-tuple of numpy arrays in the
    reference's order, so it can be fed to the reference ``Model.forward`` / the oracle / the
    drop-in module unchanged."""
    H = imp.hist_news.shape[1]
    pair_imp = np.repeat(np.arange(imp.num_impressions), np.diff(imp.cand_off))
    cat_num = news.category_num
    for lo in range(0, imp.num_pairs, batch_size):
        hi = min(lo + batch_size, imp.num_pairs)
        ii = pair_imp[lo:hi]
        hn = imp.hist_news[ii]
        cn = imp.cand_news[lo:hi]
        b = hi - lo
        yield (
            imp.user_id[ii],
            news.category[hn], news.subCategory[hn],
            news.title_text[hn], news.title_mask[hn], np.zeros_like(news.title_text[hn]),
            news.body_text[hn], news.body_mask[hn], np.zeros_like(news.body_text[hn]),
            imp.hist_fresh[ii], imp.hist_life[ii], imp.hist_mask[ii],
            np.zeros((b, H, H), np.float32), np.zeros((b, cat_num + 1), bool),
            np.zeros((b, H), np.int64),
            news.category[cn], news.subCategory[cn],
            news.title_text[cn], news.title_mask[cn], np.zeros_like(news.title_text[cn]),
            news.body_text[cn], news.body_mask[cn], np.zeros_like(news.body_text[cn]),
            imp.cand_fresh[lo:hi], imp.cand_life[lo:hi],
        )


def make_train_batch(news: NewsTable, batch_size, *, max_history=50, negatives=4, seed=1):
    """One training mini-batch in the reference's Train_Dataset layout (``dataset.py:105-141``):
    candidate tensors carry the news dimension N = 1 + negatives, positive first."""
    rng = np.random.default_rng(seed)
    B, H, N = batch_size, max_history, 1 + negatives
    imp = make_impressions(B, news.news_num, max_history=H, cand_fixed=N, seed=seed)
    hn = imp.hist_news
    cn = imp.cand_news.reshape(B, N)
    fresh = np.repeat(imp.cand_fresh.reshape(B, N)[:, :1], N, axis=1)   # dataset.py:53,73: one freshness per impression
    life = imp.cand_life.reshape(B, N)
    del rng
    return (
        imp.user_id,
        news.category[hn], news.subCategory[hn],
        news.title_text[hn], news.title_mask[hn], np.zeros_like(news.title_text[hn]),
        news.body_text[hn], news.body_mask[hn], np.zeros_like(news.body_text[hn]),
        imp.hist_fresh, imp.hist_life, imp.hist_mask,
        np.zeros((B, H, H), np.float32), np.zeros((B, news.category_num + 1), bool),
        np.zeros((B, H), np.int64),
        news.category[cn], news.subCategory[cn],
        news.title_text[cn], news.title_mask[cn], np.zeros_like(news.title_text[cn]),
        news.body_text[cn], news.body_mask[cn], np.zeros_like(news.body_text[cn]),
        fresh.astype(np.float32), life.astype(np.float32),
    )


# ---------------------------------------------------------------------------------------------
# synthetic weights (checkpoints cannot be downloaded; Model.initialize() leaves every bias, the
# LayerNorm affine and user_node_embedding at 0/1, which would hide errors in those terms)
# ---------------------------------------------------------------------------------------------
def synthetic_parameters(model, seed=0):
    """Overwrite every parameter of a (reference or drop-in) Model with seeded values drawn by numpy
    (platform-independent, unlike torch's CPU-capability-dispatched RNG kernels), at the scale the
    reference's initialisers use, but with non-trivial biases / LayerNorm affine / user-node rows.
    Buffers (the positional-encoding tables) are left alone.  Returns a float64 checksum."""
    import torch
    rng = np.random.default_rng(seed)
    checksum = 0.0
    with torch.no_grad():
        for name, p in model.named_parameters():
            shape = tuple(p.shape)
            leaf = name.split(".")[-1]
            if "norm" in name.split(".")[-2] or name.split(".")[-2] in ("ln0", "ln1", "layernorm"):
                v = (1.0 + 0.1 * rng.standard_normal(shape)) if leaf == "weight" else 0.05 * rng.standard_normal(shape)
            elif leaf.endswith("bias"):
                v = 0.05 * rng.standard_normal(shape)
            elif "word_embedding" in name:
                v = 0.1 * rng.standard_normal(shape)
                v[0] *= 0.1                      # the <PAD> row: trainable, zero-initialised, small after training
            elif "freshness_embedding" in name or "lifetime_embedding" in name or "lightgcn" in name:
                v = rng.standard_normal(shape)
            elif "category_embedding" in name or "subCategory_embedding" in name:
                v = rng.uniform(-0.1, 0.1, shape)
            elif "user_node_embedding" in name:
                v = 0.05 * rng.standard_normal(shape)
            elif len(shape) >= 2:
                fan_out, fan_in = shape[-2], shape[-1]
                bound = np.sqrt(6.0 / (fan_in + fan_out))
                v = rng.uniform(-bound, bound, shape)
            else:
                v = 0.05 * rng.standard_normal(shape)
            v = np.asarray(v, np.float32)
            p.copy_(torch.from_numpy(v).to(p.device))
            checksum += float(np.abs(v.astype(np.float64)).sum())
    return checksum
