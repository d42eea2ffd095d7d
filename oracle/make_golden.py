"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_import.py) on seeded synthetic inputs.

Run in the build container only:   python -m oracle.make_golden
Inputs and weights are NOT stored: they are regenerated from the seeds recorded in each fixture
(lime_cikm25_b200.synth uses numpy's Generator, which is platform independent); a checksum of the
weights is stored so a mismatch in regeneration is told apart from a kernel error.
"""
from __future__ import annotations

import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lime_cikm25_b200 import synth          # noqa: E402
from oracle import ref_import as R          # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (config overrides, news, impressions kwargs, batch_size)
CASES = {
    # P = 8 < H: GraphSAGE mean over the first 8 history rows; short last batch
    "small_bs8": dict(cfg=dict(vocabulary_size=2000, batch_size=8), news=150, news_seed=3,
                      imp=dict(num_impressions=10, cand_mean=6.0, near_zero_frac=0.6, seed=4), weights_seed=11),
    # config.batch_size = 64 > H = 50 (what config.py:149,159 forces): user-node rows enter the mean
    "bs64": dict(cfg=dict(vocabulary_size=1500, batch_size=64), news=120, news_seed=5,
                 imp=dict(num_impressions=3, cand_fixed=30, near_zero_frac=0.7, seed=6), weights_seed=12),
    # B = 20 buckets, beta = 0.5, alpha = 0.1 : non-default LIME hyper-parameters
    "buckets20": dict(cfg=dict(vocabulary_size=1200, batch_size=16, num_buckets=20, sigmoid_scaling_alpha=0.1,
                               penalty_scaling_beta=0.5), news=100, news_seed=7,
                      imp=dict(num_impressions=6, cand_mean=5.0, near_zero_frac=0.5, seed=8), weights_seed=13),
}


def case_inputs(case):
    spec = CASES[case]
    cfg = R.make_config(**spec["cfg"])
    news = synth.make_news_table(spec["news"], vocabulary_size=cfg.vocabulary_size, seed=spec["news_seed"])
    imp = synth.make_impressions(news_num=news.news_num, **spec["imp"])
    # adversarial rows: an impression with an empty history and one with a full history
    imp.hist_mask[0, :] = False
    imp.hist_news[0, :] = 0
    imp.hist_fresh[0, :] = 0
    imp.hist_life[0, :] = 0
    return spec, cfg, news, imp


def run_case(case):
    spec, cfg, news, imp = case_inputs(case)
    model = R.build_reference_model(cfg, seed=0)
    checksum = synth.synthetic_parameters(model, spec["weights_seed"])
    model.eval()
    ref = R.load_reference()
    scores, base_scores, cand_vecs, hist_vecs = [], [], [], []
    with torch.no_grad():
        for batch in synth.impressions_to_pair_batches(news, imp, cfg.batch_size):
            tb = [torch.as_tensor(x) for x in batch]
            rem = tb[24] - tb[23]                                       # util.py:103-104
            scores.append(model(*tb, rem).squeeze(1))
            # base score (no lifetime weight) exposes the un-saturated value of every pair
            model.remaining_lifetime_weighting.use_remaining_lifetime_weighting = False
            base_scores.append(model(*tb, rem).squeeze(1))
            model.remaining_lifetime_weighting.use_remaining_lifetime_weighting = True
        # stage goldens: content vectors and LIME vectors of the first 24 news / first batch
        n0 = min(24, news.news_num)
        ids = lambda a: torch.as_tensor(a[:n0]).unsqueeze(0)
        content = model.news_encoder.base_news_encoder(
            ids(news.title_text), ids(news.title_mask), ids(news.title_text) * 0, ids(news.body_text),
            ids(news.body_mask), ids(news.body_text) * 0, ids(news.category), ids(news.subCategory),
            None, None, None)[0]
        fresh = torch.as_tensor(synth._seconds(np.random.default_rng(99), n0, 1.0, 1e7))
        life = torch.as_tensor(synth._seconds(np.random.default_rng(98), n0, 600.0, 6e5))
        lime_vec = model.news_encoder(
            ids(news.title_text), ids(news.title_mask), ids(news.title_text) * 0, ids(news.body_text),
            ids(news.body_mask), ids(news.body_text) * 0, ids(news.category), ids(news.subCategory),
            None, fresh.unsqueeze(0), life.unsqueeze(0))[0]
    scores = torch.cat(scores).numpy()
    base_scores = torch.cat(base_scores).numpy()
    # the reference's own ranking + metrics through its file round trip (util.py:113-127)
    indices = np.repeat(np.arange(imp.num_impressions), np.diff(imp.cand_off))
    sub_scores = [[] for _ in range(indices[-1] + 1)]
    sl = scores.tolist()
    for i, index in enumerate(indices):
        sub_scores[index].append([sl[i], len(sub_scores[index])])
    res, truth = io.StringIO(), io.StringIO()
    ranks = []
    for i, sub in enumerate(sub_scores):
        sub.sort(key=lambda x: x[0], reverse=True)
        result = [0 for _ in range(len(sub))]
        for j in range(len(sub)):
            result[sub[j][1]] = j + 1
        ranks += result
        res.write(("" if i == 0 else "\n") + str(i + 1) + " " + str(result).replace(" ", ""))
        lab = imp.labels[imp.cand_off[i]:imp.cand_off[i + 1]].tolist()
        truth.write(("" if i == 0 else "\n") + str(i + 1) + " " + str(lab).replace(" ", ""))
    res.seek(0)
    truth.seek(0)
    metrics = ref.evaluate.scoring(truth, res)
    np.savez_compressed(
        os.path.join(GOLDEN, case + ".npz"),
        weights_checksum=np.float64(checksum), scores=scores, base_scores=base_scores,
        ranks=np.asarray(ranks, np.int32), metrics=np.asarray(metrics, np.float64),
        content=content.numpy(), lime_vec=lime_vec.numpy(), stage_fresh=fresh.numpy(), stage_life=life.numpy())
    print(case, "pairs", len(scores), "metrics", metrics, "nonzero scores", int((scores != 0).sum()))


def bucket_goldens():
    """Bucket ids from the reference's FreshnessEncoder.bucketize on random and knife-edge inputs."""
    ref = R.load_reference()
    rng = np.random.default_rng(5)
    xs = [np.exp(rng.uniform(np.log(0.5), np.log(2e8), 20000)).astype(np.float32),
          np.asarray([0.0, -5.0, 0.5, 1.0, 1.0000001, 2.0, 59.9, 60.0, 3600.0, 86399.0, 86400.0, 86401.0,
                      6.04e5, 6.05e5, 1e7, 1e8, 3e38], np.float32)]
    out = {}
    for nb in (10, 20, 50):
        fe = ref.newsEncoders.FreshnessEncoder.__new__(ref.newsEncoders.FreshnessEncoder)
        fe.num_buckets = nb
        # knife edges: the fp32 neighbours of every bucket boundary 86400^(7 b / nb)
        edges = []
        for b in range(1, nb):
            e = np.float32(np.exp(np.log(86400.0) * 7.0 * b / nb))
            v = e
            for _ in range(6):
                v = np.nextafter(v, np.float32(0))
            for _ in range(13):
                edges.append(v)
                v = np.nextafter(v, np.float32(np.inf))
        x = np.concatenate(xs + [np.asarray(edges, np.float32)])
        out["x_%d" % nb] = x
        out["b_%d" % nb] = fe.bucketize(torch.as_tensor(x)).numpy().astype(np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "buckets.npz"), **out)
    print("buckets", {k: v.shape for k, v in out.items()})


def metric_goldens():
    """evaluate.scoring (sklearn AUC etc.) on random score lists with heavy ties."""
    ref = R.load_reference()
    rng = np.random.default_rng(17)
    scores, labels, off = [], [], [0]
    for i in range(200):
        c = int(rng.integers(2, 60)) if i % 10 else 300
        s = rng.standard_normal(c).astype(np.float32)
        s[rng.random(c) < 0.5] = 0.0                       # saturated lifetime weights -> exact ties
        s[rng.random(c) < 0.1] = -0.0
        y = (rng.random(c) < 0.15).astype(np.uint8)
        y[rng.integers(0, c)] = 1
        y[(np.flatnonzero(y)[0] + 1) % c] = 0
        if y.sum() == 0:
            y[0] = 1
        scores.append(s)
        labels.append(y)
        off.append(off[-1] + c)
    res, truth = io.StringIO(), io.StringIO()
    ranks = []
    for i, (s, y) in enumerate(zip(scores, labels)):
        sub = [[float(v), j] for j, v in enumerate(s)]
        sub.sort(key=lambda x: x[0], reverse=True)          # util.py:119
        result = [0] * len(sub)
        for j in range(len(sub)):
            result[sub[j][1]] = j + 1
        ranks += result
        res.write(("" if i == 0 else "\n") + str(i + 1) + " " + str(result).replace(" ", ""))
        truth.write(("" if i == 0 else "\n") + str(i + 1) + " " + str(y.tolist()).replace(" ", ""))
    res.seek(0)
    truth.seek(0)
    m = ref.evaluate.scoring(truth, res)
    np.savez_compressed(os.path.join(GOLDEN, "metrics.npz"), scores=np.concatenate(scores),
                        labels=np.concatenate(labels), cand_off=np.asarray(off, np.int64),
                        ranks=np.asarray(ranks, np.int32), metrics=np.asarray(m, np.float64))
    print("metrics", m)


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    bucket_goldens()
    metric_goldens()
    for c in CASES:
        run_case(c)
