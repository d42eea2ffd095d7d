"""Golden vectors of the reference's dataset gathers (TEST INFRASTRUCTURE; run in the build container only):
a fake corpus -> the reference's own DevTest_Dataset / Train_Dataset (dataset.py, unmodified, imported from
/root/reference) -> tests/golden/dataset_gathers.npz.  The GPU box, which has no reference tree, replays the same
fake corpus through lime_cikm25_b200.dataset / util.corpus_to_tables and compares bit for bit.

    python -m oracle.make_dataset_golden
"""
from __future__ import annotations

import os
import types

import numpy as np

H, M = 50, 4


def fake_corpus(seed=5, n_news=60, n_imp=7):
    """News tables (synth) + dev / train behaviours in the reference's list layout (corpus.py:577-649), with the edge
    cases of dataset.py:119-128: empty history, history longer than H (seconds lists keep their raw length), short
    negative lists (<= M: cycled) and long ones (sampled without replacement)."""
    from lime_cikm25_b200 import synth
    rng = np.random.default_rng(seed)
    news = synth.make_news_table(n_news, vocabulary_size=500, seed=seed)
    lens = [0, 3, 50, 57, 12, 49, 51][:n_imp]
    dev, train = [], []
    for i, n in enumerate(lens):
        raw = rng.integers(1, n_news, size=n).tolist()
        kept = raw[-H:]
        hist_index = kept + [0] * (H - len(kept))
        hist_mask = np.asarray([True] * len(kept) + [False] * (H - len(kept)))
        fresh = rng.uniform(1.0, 1e6, size=n).astype(np.float32).astype(float).tolist()      # raw length (may exceed H)
        life = rng.uniform(600.0, 6e5, size=n).astype(np.float32).astype(float).tolist()
        ncand = int(rng.integers(1, 5))
        for _ in range(ncand):
            dev.append([100 + i, hist_index, hist_mask, int(rng.integers(1, n_news)), i,
                        float(np.float32(rng.uniform(1.0, 1e6))), float(np.float32(rng.uniform(600.0, 6e5))), fresh, life])
        nneg = [2, 4, 9, 1, 6, 5, 12][i]
        train.append([100 + i, hist_index, hist_mask, int(rng.integers(1, n_news)), rng.integers(1, n_news, size=nneg).tolist(), i,
                      float(np.float32(rng.uniform(1.0, 1e6))), float(np.float32(rng.uniform(600.0, 6e5))),
                      rng.uniform(600.0, 6e5, size=nneg).astype(np.float32).astype(float).tolist(), fresh, life])
    cfg = types.SimpleNamespace(category_num=news.category_num, user_encoder="CROWN", vocabulary_size=news.vocabulary_size,
                                subCategory_num=news.subCategory_num)
    corpus = types.SimpleNamespace(
        config=cfg, negative_sample_num=M, max_history_num=H,
        news_category=news.category, news_subCategory=news.subCategory,
        news_title_text=news.title_text, news_title_mask=news.title_mask, news_title_entity=np.zeros_like(news.title_text),
        news_abstract_text=news.body_text, news_abstract_mask=news.body_mask, news_abstract_entity=np.zeros_like(news.body_text),
        dev_behaviors=dev, test_behaviors=dev, train_behaviors=train,
        dev_indices=[b[4] for b in dev], test_indices=[b[4] for b in dev],
        dev_user_history_graph=None, dev_user_history_category_mask=None, dev_user_history_category_indices=None,
        test_user_history_graph=None, test_user_history_category_mask=None, test_user_history_category_indices=None)
    return news, corpus


def reference_outputs(corpus, np_seed=123):
    """Every DevTest sample and every Train sample (after negative_sampling under numpy seed ``np_seed``) of the
    reference's datasets as lists of 25 numpy arrays."""
    from oracle import ref_import
    import contextlib, io
    import torch
    ref = ref_import.load_reference()
    to_np = lambda x: x.numpy() if torch.is_tensor(x) else np.asarray(x)
    dev_ds = ref.dataset.DevTest_Dataset(corpus, "dev")
    dev = [[to_np(x) for x in dev_ds[i]] for i in range(len(dev_ds))]
    tr_ds = ref.dataset.Train_Dataset(corpus)
    np.random.seed(np_seed)
    with contextlib.redirect_stdout(io.StringIO()):
        tr_ds.negative_sampling()
    train = [[to_np(x) for x in tr_ds[i]] for i in range(len(tr_ds))]
    return dev, train, np.asarray(tr_ds.train_samples, np.int64)


def main():
    news, corpus = fake_corpus()
    dev, train, samples = reference_outputs(corpus)
    out = {"train_samples": samples}
    for name, rows in (("dev", dev), ("train", train)):
        for f in range(25):
            out["%s_%02d" % (name, f)] = np.stack([r[f] for r in rows])
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset_gathers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
