"""GPU box: run the tensor-core scoring kernel on a mid-size impression set and compare with the exact kernel
(use LIME_B200_LIB=.../liblime_b200_dbg.so for the wait-timeout build)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import engine, ops, synth, util
from lime_cikm25_b200.config import default_config
n_news, n_imp = int(sys.argv[1]) if len(sys.argv) > 1 else 3000, int(sys.argv[2]) if len(sys.argv) > 2 else 4000
cfg = default_config(vocabulary_size=5000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(n_news, vocabulary_size=5000, seed=11)
imp = synth.make_impressions(n_imp, news.news_num, seed=12)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=7)
model = model.cuda().eval()
with torch.no_grad():
    cache = util.build_news_cache(model, news, "cuda")
    dimp = engine.DeviceImpressions(imp, "cuda")
    print("units", dimp.num_units, "pairs", dimp.num_pairs, flush=True)
    ops.score_configure(ops.SCORE_EXACT)
    ex = util.score_impressions(model, cache, dimp, 32).clone(); torch.cuda.synchronize()
    print("exact done", flush=True)
    ops.score_configure(ops.SCORE_AUTO, 1e-6)
    t0 = time.time()
    tc = util.score_impressions(model, cache, dimp, 32).clone(); torch.cuda.synchronize()
    print("tc done %.3fs fallback %d" % (time.time() - t0, int(dimp.work_counter[1])), flush=True)
    d = (tc - ex).abs().cpu().numpy(); b = ex.abs().cpu().numpy()
    floor = 0.1 * float(np.sqrt(np.mean(b * b)))
    print("max rel err %.3e" % float(np.max(d / np.maximum(b, floor))))
