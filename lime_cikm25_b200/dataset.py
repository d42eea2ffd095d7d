"""Dataset gathers of the hot path (reference dataset.py:41-76, 105-141, 192-227).

The reference builds every sample on CPU workers with numpy fancy indexing over the corpus id tables and ships
25 tensors per sample through the DataLoader.  Here the id tables live in HBM once (``DeviceNewsTables``) and a
mini-batch is assembled ON THE DEVICE from per-behaviour index lists (``DeviceTrainSet.batch`` /
``engine.DeviceImpressions`` for the eval side): the same gathers, the same truncate-to-last-H / right-zero-pad of
the freshness and lifetime lists, the same dummy graph tensors, bit for bit
(tests/test_dataset_parity.py pins them against the reference's own ``Train_Dataset`` / ``DevTest_Dataset``).

Index building and gathers are integer / byte plumbing: torch indexing on device tensors, no arithmetic.
"""
from __future__ import annotations

import numpy as np
import torch
from numpy.random import randint


def pad_history_seconds(values, max_history_num):
    """dataset.py:123-128 / 203-205: keep the LAST ``max_history_num`` entries, then zero-pad on the right by
    ``max_history_num - len(values)`` (not by the truncated length: a list longer than H gets no padding)."""
    values = list(values)
    pad = max_history_num - len(values)
    return values[-max_history_num:] + [0] * max(0, pad)


def negative_sampling(train_behaviors, negative_sample_num):
    """Train_Dataset.negative_sampling (dataset.py:41-76), same draws from numpy's GLOBAL generator
    (``numpy.random.randint(0, n - 1)``: the upper bound is exclusive, so the reference never draws the last
    negative when it samples without replacement -- kept).  Returns int64 samples [n, 1+M] and float64 freshness /
    lifetime [n, 1+M] (python floats in the reference)."""
    n, M = len(train_behaviors), negative_sample_num
    samples = np.zeros((n, 1 + M), np.int64)
    fresh = np.zeros((n, 1 + M), np.float64)
    life = np.zeros((n, 1 + M), np.float64)
    for i, b in enumerate(train_behaviors):
        pos_index, neg_indices, freshness, pos_lifetime, neg_lifetimes = b[3], b[4], b[6], b[7], b[8]
        samples[i, 0], fresh[i, 0], life[i, 0] = pos_index, freshness, pos_lifetime
        news_num = len(neg_indices)
        used = set()
        for j in range(M):
            if news_num <= M:
                k = j % news_num
            else:
                while True:
                    k = randint(0, news_num - 1)
                    if k not in used:
                        used.add(k)
                        break
            samples[i, j + 1], fresh[i, j + 1], life[i, j + 1] = neg_indices[k], freshness, neg_lifetimes[k]
    return samples, fresh, life


class DeviceNewsTables:
    """The corpus id tables (corpus.py:361-368) resident on the device: int32 text / entity ids, bool masks,
    int32 category ids.  ``news`` is a synth.NewsTable or any object with the reference corpus' attribute names."""

    def __init__(self, news, device):
        g = lambda *names: next(getattr(news, n) for n in names if hasattr(news, n))
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(device=device, dtype=dt)
        self.category = t(g("category", "news_category"), torch.int32)
        self.subCategory = t(g("subCategory", "news_subCategory"), torch.int32)
        self.title_text = t(g("title_text", "news_title_text"), torch.int32)
        self.title_mask = t(g("title_mask", "news_title_mask"), torch.bool)
        self.body_text = t(g("body_text", "news_abstract_text"), torch.int32)
        self.body_mask = t(g("body_mask", "news_abstract_mask"), torch.bool)
        te = getattr(news, "news_title_entity", None)
        be = getattr(news, "news_abstract_entity", None)
        self.title_entity = t(te, torch.int32) if te is not None else torch.zeros_like(self.title_text)
        self.body_entity = t(be, torch.int32) if be is not None else torch.zeros_like(self.body_text)
        self.category_num = int(getattr(news, "category_num", getattr(getattr(news, "config", None), "category_num", 0)))
        self.device = device

    def gather(self, idx):
        """The 8 per-news tensors of a (history or candidate) index tensor, in the reference's tuple order."""
        i = idx.long()
        return (self.category[i], self.subCategory[i], self.title_text[i], self.title_mask[i], self.title_entity[i],
                self.body_text[i], self.body_mask[i], self.body_entity[i])


class DeviceTrainSet:
    """Train_Dataset with device-resident gathers: per-behaviour index / seconds arrays are uploaded once per epoch
    (after ``negative_sampling``), ``batch(rows)`` returns the 25 tensors of Train_Dataset.__getitem__ stacked over
    ``rows`` exactly as the DataLoader's default collate does (dataset.py:105-141)."""

    def __init__(self, tables: DeviceNewsTables, train_behaviors, max_history_num, negative_sample_num):
        self.tables, self.H, self.M = tables, int(max_history_num), int(negative_sample_num)
        self.behaviors = train_behaviors
        dev = tables.device
        n = len(train_behaviors)
        self.user_id = torch.as_tensor(np.asarray([b[0] for b in train_behaviors], np.int64)).to(dev)
        self.hist_index = torch.as_tensor(np.asarray([b[1] for b in train_behaviors], np.int64).reshape(n, self.H)).to(dev)
        self.hist_mask = torch.as_tensor(np.asarray([b[2] for b in train_behaviors], bool).reshape(n, self.H)).to(dev)
        self.hist_fresh = torch.as_tensor(np.asarray([pad_history_seconds(b[9], self.H) for b in train_behaviors], np.float32)).to(dev)
        self.hist_life = torch.as_tensor(np.asarray([pad_history_seconds(b[10], self.H) for b in train_behaviors], np.float32)).to(dev)
        self.num = n
        self.samples = self.fresh = self.life = None

    def negative_sampling(self, rank=None):
        s, f, l = negative_sampling(self.behaviors, self.M)
        dev = self.tables.device
        self.samples = torch.as_tensor(s).to(dev)
        self.fresh = torch.as_tensor(f.astype(np.float32)).to(dev)
        self.life = torch.as_tensor(l.astype(np.float32)).to(dev)

    def __len__(self):
        return self.num

    def batch(self, rows):
        """rows: int64 device tensor [B] of behaviour indices -> the 25-tuple, every element [B, ...]."""
        if self.samples is None:
            raise RuntimeError("call negative_sampling() first (trainer.py:283-286)")
        T = self.tables
        B, H = rows.shape[0], self.H
        dev = T.device
        hist = self.hist_index[rows]
        samp = self.samples[rows]
        return (self.user_id[rows],) + T.gather(hist) + (
            self.hist_fresh[rows], self.hist_life[rows], self.hist_mask[rows],
            torch.zeros(B, H, H, dtype=torch.float32, device=dev),                 # dummy user_history_graph (:118)
            torch.zeros(B, T.category_num + 1, dtype=torch.bool, device=dev),      # dummy category mask (:119)
            torch.zeros(B, H, dtype=torch.int64, device=dev),                      # dummy category indices (:120)
        ) + T.gather(samp) + (self.fresh[rows], self.life[rows])


def epoch_order(num, seed, epoch, rank=0, world_size=1, shuffle=True):
    """DistributedSampler-equivalent index order (trainer.py:293-295): a seeded permutation per epoch, padded to a
    multiple of world_size by wrapping, every rank taking indices rank, rank + world_size, ..."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(int(seed) + int(epoch))
        idx = torch.randperm(num, generator=g)
    else:
        idx = torch.arange(num)
    total = ((num + world_size - 1) // world_size) * world_size
    if total > num:
        idx = torch.cat([idx, idx[:total - num]])
    return idx[rank:total:world_size]
