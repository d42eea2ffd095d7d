"""GPU box: the scoring call alone on a bench-shaped workload (for ncu): scripts/prof_score.py [news] [impressions] [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import engine, synth, util
from lime_cikm25_b200.config import default_config
n_news = int(sys.argv[1]) if len(sys.argv) > 1 else 65238
n_imp = int(sys.argv[2]) if len(sys.argv) > 2 else 73152
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = default_config(vocabulary_size=40000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(n_news, vocabulary_size=40000, seed=1)
imp = synth.make_impressions(n_imp, news.news_num, seed=100)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=0)
model = model.cuda().eval()
model.news_encoder.engine.bf16 = True       # cache build speed only; the scoring kernel is what is profiled
with torch.no_grad():
    cache = util.build_news_cache(model, news, "cuda")
    dimp = engine.DeviceImpressions(imp, "cuda")
    out = torch.empty(dimp.num_pairs, dtype=torch.float32, device="cuda")
    for _ in range(2):
        util.score_impressions(model, cache, dimp, 32, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        util.score_impressions(model, cache, dimp, 32, out=out)
    e1.record(); torch.cuda.synchronize()
    print("units %d pairs %d  %.3f ms per call  fallback %d" % (dimp.num_units, dimp.num_pairs, e0.elapsed_time(e1) / reps, int(dimp.work_counter[1])))
