"""GPU box: where the wall time of build_news_cache goes (bf16 mode), section by section (synchronised timers)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import synth, util, engine
from lime_cikm25_b200.config import default_config
cfg = default_config(vocabulary_size=40000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(65238, vocabulary_size=40000, seed=1)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=0)
model = model.cuda().eval()
model.news_encoder.engine.bf16 = True
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-28s %7.2f ms" % (label, 1e3 * (time.perf_counter() - t0))); return r
with torch.no_grad():
    model.scoring.fold()
    util.build_news_cache(model, news, "cuda")
    for rep in range(2):
        print("--- pass", rep)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.int32)).to("cuda")
        ins = T("H2D of the id tables", lambda: (t(news.title_text), t(news.body_text), t(news.category), t(news.subCategory)))
        hist, cand = T("build_rows", lambda: model.scoring.build_rows(*ins, chunk=8192))
        c16 = T("split_candidates", lambda: model.scoring.split_candidates(cand))
        meta = T("news_meta", lambda: model.scoring.news_meta(hist, cand))
        vg = T("interleave_vg", lambda: engine.interleave_vg(hist))
        T("whole build_news_cache", lambda: util.build_news_cache(model, news, "cuda"))
    # inside build_rows
    import types
    se = model.scoring
    tt, bt, ct, sb = ins
    T("encode_content (65k)", lambda: [se.news.encode_content(tt[lo:lo+8192].contiguous(), bt[lo:lo+8192].contiguous(), ct[lo:lo+8192].contiguous(), sb[lo:lo+8192].contiguous()) for lo in range(0, 65238, 8192)])
    T("zeros hist+cand", lambda: (torch.zeros(65238, 852, device="cuda"), torch.zeros(65238, 1720, device="cuda")))
    T("_register_topics", lambda: se._register_topics(ct, sb, hist, cand))
