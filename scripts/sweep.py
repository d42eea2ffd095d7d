#!/usr/bin/env python
"""Scaling sweep of the scoring stage (BASELINE.json configs[3], configs[4]): history 50 -> 200, candidates per
impression 5 -> 300, lifetime buckets 10 -> 50, and an Adressa-shaped run (body-heavy news only affect the
cache build).  Prints one JSON object per point: impressions/s, pairs/s, algorithmic GB/s and the fraction of the
measured HBM peak; writes gpurun_out/sweep.jsonl.  Runs on the GPU box:  python scripts/sweep.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import lime_cikm25_b200 as L
from lime_cikm25_b200 import engine, synth, util
from lime_cikm25_b200.config import default_config as make_config

PEAK = 6540.2
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(p):
    PEAK = float(json.load(open(p))["hbm_gbs"])


def point(name, H=50, cand_fixed=None, cand_mean=37.0, nb=10, news_num=20000, impressions=20000, body_full=False, steps=5):
    cfg = make_config(vocabulary_size=20000, batch_size=32, word_embedding_init="skip", max_history_num=H, num_buckets=nb)
    model = L.Model(cfg)
    model.initialize()
    synth.synthetic_parameters(model, seed=0)
    model = model.cuda().eval()
    news = synth.make_news_table(news_num, vocabulary_size=20000, seed=1, body_full=body_full,
                                 title_mean=6.63 if body_full else 11.67)
    imp = synth.make_impressions(impressions, news.news_num, max_history=H, cand_fixed=cand_fixed, cand_mean=cand_mean, seed=7)
    with torch.no_grad():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cache = util.build_news_cache(model, news)
        torch.cuda.synchronize()
        cache_s = time.perf_counter() - t0
        dimp = engine.DeviceImpressions(imp, "cuda")
        scores = torch.empty(dimp.num_pairs, dtype=torch.float32, device="cuda")
        for _ in range(3):
            util.evaluate_device(model, cache, dimp, 32, scores_out=scores, want_ranks=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            util.score_impressions(model, cache, dimp, 32, out=scores)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        fallback = int(dimp.work_counter[1])
    C = np.diff(imp.cand_off).astype(np.float64)
    alg = float(np.sum((H + C) * (400 * 4 + 20) + H + 4 * C))
    out = {"point": name, "history": H, "mean_candidates": float(C.mean()), "num_buckets": nb, "news": news_num,
           "impressions": impressions, "pairs": int(imp.num_pairs), "kernel": "score_tc" if H <= 56 else ("score_tc, %d chunks" % dimp.chunked(nb).chunks if dimp.chunked(nb) is not None else "score (exact)"),
           "ms_per_launch": ms, "impressions_per_sec": impressions / (ms * 1e-3), "pairs_per_sec": imp.num_pairs / (ms * 1e-3),
           "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / PEAK,
           "units_sent_to_exact_fallback": fallback, "cache_build_news_per_sec": news_num / cache_s}
    print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rows = []
    rows.append(point("MIND-shaped baseline"))
    for C in (5, 20, 100, 300):
        rows.append(point("candidates=%d" % C, cand_fixed=C, impressions=max(2000, 400000 // C)))
    for H in (100, 200):
        rows.append(point("history=%d" % H, H=H, impressions=10000))
    for nb in (20, 50):
        rows.append(point("buckets=%d" % nb, nb=nb))
    rows.append(point("Adressa-shaped (full 128-token bodies)", body_full=True, news_num=20000))
    with open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
