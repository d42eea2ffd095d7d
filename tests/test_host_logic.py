"""Host-side logic of the drop-in path: module surface, unit building, batching rules, adapters."""
import json
import os
import types

import numpy as np
import pytest
import torch

import lime_cikm25_b200 as L
from lime_cikm25_b200 import engine, parallel, synth, util
from oracle.ref_import import make_config


def small_model():
    cfg = make_config(vocabulary_size=500, word_embedding_init="skip")
    m = L.Model(cfg)
    m.initialize()
    return cfg, m


def test_state_dict_keys_match_reference_contract(golden_dir):
    """Appendix A of SURVEY.md: the key list captured from the reference's own Model."""
    cfg, m = small_model()
    want = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    got = {k: list(v.shape) for k, v in m.state_dict().items()}
    want = {k: [500 if (d == 2000 and "word_embedding" in k) else d for d in s] for k, s in want.items()}
    assert list(got) == list(want)
    assert got == want
    # every news-encoder key also appears under user_encoder.news_encoder.* (userEncoders.py:20)
    assert sum(k.startswith("user_encoder.news_encoder.") for k in got) == sum(k.startswith("news_encoder.") for k in got)
    assert m.user_encoder.news_encoder is m.news_encoder
    assert m.model_name == "LIME-CROWN-CROWN"
    assert m.news_embedding_dim == 400 and m.news_encoder.base_news_encoder.news_embedding_dim == 900


def test_plugin_dispatch_errors_like_reference():
    with pytest.raises(Exception, match="is not implemented"):
        L.Model(make_config(news_encoder="NAML", word_embedding_init="skip"))
    with pytest.raises(Exception, match="is not implemented"):
        L.Model(make_config(user_encoder="MINER", vocabulary_size=50, word_embedding_init="skip"))
    with pytest.raises(ValueError, match="Unknown content encoder"):
        L.Model(make_config(content_encoder="CNN", word_embedding_init="skip"))
    with pytest.raises(FileNotFoundError):            # newsEncoders.py:173: pickle must exist in CWD
        L.Model(make_config(vocabulary_size=50))


def test_frozen_and_dead_parameters():
    _, m = small_model()
    ne = m.news_encoder
    assert not ne.category_embedding.weight.requires_grad
    assert not ne.subCategory_embedding.weight.requires_grad
    assert ne.base_news_encoder.category_embedding.weight.requires_grad          # re-created trainable (:237)
    assert not ne.base_news_encoder.subCategory_embedding.weight.requires_grad
    assert float(m.user_encoder.user_node_embedding.abs().sum()) == 0.0          # zero-init (userEncoders.py:81)


def test_training_forward_needs_a_gpu():
    """Training-mode Model.forward is the differentiable B200 path (training.py); on a box without a
    CUDA device it must fail loudly instead of falling back to PyTorch."""
    from lime_cikm25_b200 import _lib
    cfg, m = small_model()
    m.train()
    news = synth.make_news_table(30, vocabulary_size=cfg.vocabulary_size, seed=1)
    tb = [torch.as_tensor(x) for x in synth.make_train_batch(news, 2, seed=2)]
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    with pytest.raises(_lib.LimeError):
        m(*tb, tb[24] - tb[23])


def test_build_units():
    off = np.array([0, 3, 3, 60, 61, 61 + 104])
    imp, p0, cnt = engine.build_units(off, 52)
    assert imp.tolist() == [0, 2, 2, 3, 4, 4]
    assert p0.tolist() == [0, 3, 55, 60, 61, 113]
    assert cnt.tolist() == [3, 52, 5, 1, 52, 52]
    assert cnt.sum() == off[-1]


def test_tail_rule_matches_dataloader():
    # DataLoader(shuffle=False, batch_size=bs): last batch has total % bs samples
    assert util._tail(100, 32) == (96, 4)
    assert util._tail(96, 32) == (96, 1)          # no tail: tail_start == total, prefix unused
    assert util._tail(5, 32) == (0, 5)


def test_synth_shapes_and_invariants():
    news = synth.make_news_table(300, vocabulary_size=1000, seed=2)
    assert news.title_text.dtype == np.int32 and news.title_text.shape == (301, 32)
    assert news.body_text.shape == (301, 128) and news.title_mask.dtype == bool
    assert news.title_text[0].sum() == 0 and news.body_text[0].sum() == 0      # pad news
    assert (news.title_text[news.title_mask] > 0).all() and (news.title_text[~news.title_mask] == 0).all()
    imp = synth.make_impressions(200, news.news_num, seed=3)
    assert imp.hist_news.dtype == np.int32 and imp.hist_fresh.dtype == np.float32
    assert (imp.hist_news[~imp.hist_mask] == 0).all() and (imp.hist_fresh[~imp.hist_mask] == 0).all()
    # mask is a prefix (corpus.py:516-517) and every impression has both classes
    assert (np.diff(imp.hist_mask.astype(int), axis=1) <= 0).all()
    for i in range(imp.num_impressions):
        y = imp.labels[imp.cand_off[i]:imp.cand_off[i + 1]]
        assert 0 < y.sum() < len(y)
    batches = list(synth.impressions_to_pair_batches(news, imp, 32))
    assert sum(len(b[0]) for b in batches) == imp.num_pairs and len(batches[0]) == 25
    assert batches[0][3].shape == (32, 50, 32) and batches[0][17].shape == (32, 32)
    assert batches[0][12].shape == (32, 50, 50) and batches[0][13].shape == (32, 19)


def test_shard_bounds_cover_and_balance():
    imp = synth.make_impressions(1000, 500, seed=5)
    b = parallel.shard_bounds(imp.cand_off, 50, 8)
    assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) > 0).all()
    work = np.diff(imp.cand_off) + 50
    per = [work[b[r]:b[r + 1]].sum() for r in range(8)]
    assert max(per) / (sum(per) / 8) < 1.05
    parts = [parallel.shard_impressions(imp, r, 8) for r in range(8)]
    assert sum(p[0].num_pairs for p in parts) == imp.num_pairs
    assert parts[3][1] == imp.cand_off[b[3]] and parts[0][2] == imp.num_pairs
    assert np.array_equal(np.concatenate([p[0].cand_news for p in parts]), imp.cand_news)


def test_corpus_adapter_roundtrip(tmp_path):
    """corpus_to_tables turns the reference's per-pair behaviour records (corpus.py:590-600) back
    into impression-major arrays, applying dataset.py:203-205's truncate-to-last-H + right-pad."""
    news = synth.make_news_table(50, vocabulary_size=200, seed=1)
    imp = synth.make_impressions(7, news.news_num, cand_mean=4, seed=2)
    beh, idx = [], []
    for i in range(imp.num_impressions):
        n = int(imp.hist_mask[i].sum())
        extra = [123.0] * 3 if i == 1 else []        # longer than H is impossible here; shorter lists get padded
        for p in range(imp.cand_off[i], imp.cand_off[i + 1]):
            beh.append([int(imp.user_id[i]), imp.hist_news[i].tolist(), imp.hist_mask[i].copy(),
                        int(imp.cand_news[p]), i, float(imp.cand_fresh[p]), float(imp.cand_life[p]),
                        imp.hist_fresh[i, :n].tolist(), imp.hist_life[i, :n].tolist()])
            idx.append(i)
        del extra
    cfg = types.SimpleNamespace(vocabulary_size=200, category_num=18, subCategory_num=270)
    corpus = types.SimpleNamespace(
        dev_behaviors=beh, dev_indices=idx, max_history_num=50, config=cfg,
        news_title_text=news.title_text, news_title_mask=news.title_mask,
        news_abstract_text=news.body_text, news_abstract_mask=news.body_mask,
        news_category=news.category, news_subCategory=news.subCategory)
    news2, imp2 = util.corpus_to_tables(corpus, "dev")
    for f in ("hist_news", "hist_mask", "hist_fresh", "hist_life", "cand_off", "cand_news", "cand_fresh", "cand_life"):
        assert np.array_equal(getattr(imp, f), getattr(imp2, f)), f
    assert news2.title_text is news.title_text


def test_rank_file_format(tmp_path):
    p = tmp_path / "res.txt"
    util.write_rank_file(str(p), torch.tensor([2, 1, 3, 1, 2]), np.array([0, 3, 5]))
    assert p.read_text() == "1 [2,1,3]\n2 [1,2]"       # util.py:123: no spaces, no trailing newline
