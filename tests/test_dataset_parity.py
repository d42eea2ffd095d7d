"""Dataset gathers: the impression-major / device-resident gathers of this package
against the reference's OWN ``DevTest_Dataset`` / ``Train_Dataset`` (dataset.py:41-76, 105-141, 192-227), bit for bit:
history truncated to the last 50, seconds lists right-zero-padded, dummy graph tensors, negative sampling.

Two pins: live against the unmodified reference module (build container, skipped where /root/reference is absent) and
against tests/golden/dataset_gathers.npz, which oracle/make_dataset_golden.py wrote from the same reference run."""
import os

import numpy as np
import pytest
import torch

from lime_cikm25_b200 import dataset as D
from lime_cikm25_b200 import engine, synth, util
from oracle import ref_import
from oracle.make_dataset_golden import H, M, fake_corpus

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_gathers.npz")


def _same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert a.dtype == b.dtype, (what, a.dtype, b.dtype)
    assert np.array_equal(a, b), what


def _dev_samples(corpus):
    """Every (user, candidate) sample of the dev set through corpus_to_tables -> impressions_to_pair_batches."""
    news, imp = util.corpus_to_tables(corpus, "dev")
    rows = []
    for b in synth.impressions_to_pair_batches(news, imp, 1):
        rows.append([np.asarray(x)[0] for x in b])
    return news, imp, rows


def _train_samples(corpus, device="cpu", np_seed=123):
    tables = D.DeviceNewsTables(corpus, device)
    ds = D.DeviceTrainSet(tables, corpus.train_behaviors, H, M)
    np.random.seed(np_seed)
    ds.negative_sampling()
    batch = ds.batch(torch.arange(len(ds), device=device))
    return ds, [[x[i].cpu().numpy() for x in batch] for i in range(len(ds))]


def _check_against(dev_ref, train_ref, samples_ref, corpus, device="cpu"):
    _, imp, dev = _dev_samples(corpus)
    assert len(dev) == len(dev_ref)
    for i, (got, want) in enumerate(zip(dev, dev_ref)):
        for f in range(25):
            _same(got[f], want[f], "dev sample %d field %d" % (i, f))
    ds, train = _train_samples(corpus, device)
    assert np.array_equal(ds.samples.cpu().numpy(), samples_ref)                   # negative sampling: same draws
    for i, (got, want) in enumerate(zip(train, train_ref)):
        for f in range(25):
            _same(got[f], want[f], "train sample %d field %d" % (i, f))
    # the device-resident eval layout holds the same history gathers (impression-major)
    first = imp.cand_off[:-1]
    for i in range(imp.num_impressions):
        want = dev_ref[int(first[i])]
        assert np.array_equal(imp.hist_fresh[i], want[9]) and np.array_equal(imp.hist_life[i], want[10])
        assert np.array_equal(imp.hist_mask[i], want[11])


@pytest.mark.skipif(not ref_import.reference_available(), reason="needs the reference tree (build container)")
def test_gathers_match_reference_datasets_live():
    from oracle.make_dataset_golden import reference_outputs
    _, corpus = fake_corpus()
    dev_ref, train_ref, samples_ref = reference_outputs(corpus)
    _check_against(dev_ref, train_ref, samples_ref, corpus)


def _golden():
    g = np.load(GOLD)
    n_dev, n_tr = g["dev_00"].shape[0], g["train_00"].shape[0]
    dev = [[g["dev_%02d" % f][i] for f in range(25)] for i in range(n_dev)]
    train = [[g["train_%02d" % f][i] for f in range(25)] for i in range(n_tr)]
    return dev, train, g["train_samples"]


def test_gathers_match_reference_goldens():
    _, corpus = fake_corpus()
    _check_against(*_golden(), corpus)


def test_seconds_padding_rule():
    """dataset.py:123-128: last H entries, padded by H - len(list) zeros (a longer list gets none)."""
    assert D.pad_history_seconds([], 3) == [0, 0, 0]
    assert D.pad_history_seconds([1.5], 3) == [1.5, 0, 0]
    assert D.pad_history_seconds([1, 2, 3, 4, 5], 3) == [3, 4, 5]


def test_epoch_order_is_a_distributed_partition():
    parts = [D.epoch_order(10, seed=3, epoch=2, rank=r, world_size=4) for r in range(4)]
    assert all(len(p) == 3 for p in parts)
    allidx = torch.cat(parts)
    assert set(allidx.tolist()) == set(range(10))                                  # padded by wrapping: every index at least once
    assert not torch.equal(D.epoch_order(10, 3, 2), D.epoch_order(10, 3, 3))       # reshuffled per epoch


@pytest.mark.gpu
def test_device_train_gathers_match_goldens_on_gpu():
    _, corpus = fake_corpus()
    dev_ref, train_ref, samples_ref = _golden()
    _, train = _train_samples(corpus, "cuda")
    for i, (got, want) in enumerate(zip(train, train_ref)):
        for f in range(25):
            _same(got[f], want[f], "train sample %d field %d" % (i, f))
