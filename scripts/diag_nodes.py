"""Diagnostic (GPU box): how many work units of the bench workload take the 2-node / 4-node / exact path,
for several gate tolerances, with the kernel time of each setting."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import engine, ops, synth, util
from lime_cikm25_b200.config import default_config

n_news, n_imp = int(sys.argv[1]) if len(sys.argv) > 1 else 65238, int(sys.argv[2]) if len(sys.argv) > 2 else 73152
cfg = default_config(vocabulary_size=40000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(n_news, vocabulary_size=40000, seed=1)
imp = synth.make_impressions(n_imp, news.news_num, seed=100)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=0)
model = model.cuda().eval()
with torch.no_grad():
    cache = util.build_news_cache(model, news, "cuda")
    dimp = engine.DeviceImpressions(imp, "cuda")
    hist = cache.hist_rows
    print("gw absmax per news (natural units): median %.3f  p99 %.3f  max %.3f; table %.3f" % (
        float(hist[:, 851].median()) / 1.4427, float(hist[:, 851].quantile(0.99)) / 1.4427, float(hist[:, 851].max()) / 1.4427,
        model.scoring.fold()["tab_gw_absmax"] / 1.4427))
    out = torch.empty(dimp.num_pairs, dtype=torch.float32, device="cuda")
    for tol in (1e-6, 1e-13):
        ops.score_configure(ops.SCORE_AUTO, tol)
        for _ in range(2):
            util.score_impressions(model, cache, dimp, 32, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            util.score_impressions(model, cache, dimp, 32, out=out)
        e1.record(); torch.cuda.synchronize()
        import ctypes
        from lime_cikm25_b200 import _lib
        buf = (ctypes.c_uint64 * 16)()
        _lib.load().lime_score_phase_clocks(buf)
        names = ["I:wait_Wfull", "I:wait_Ofull", "attn", "centres", "produce", "mma_wait", "epilogue+pool", "merge", "tail", "units", "I:front_hist+dedup", "P:copy", "P:compute", "P:o_slot_wait", "P:store+arrive", "I:front_cand"]
        tot = sum(buf[i] for i in (2, 3, 4, 5, 6, 7, 8))
        print("  phase clocks per unit (thread 0): " + "  ".join("%s %.0f" % (n, buf[i] / max(buf[9], 1)) for i, n in enumerate(names) if i != 9 and n != "-") + "  | total %.0f" % (tot / max(buf[9], 1)))
        wc = dimp.work_counter[:4].tolist()
        print(json.dumps({"tol": tol, "units": dimp.num_units, "fallback": wc[1], "ms": e0.elapsed_time(e1) / 3}))
