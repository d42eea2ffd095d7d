"""GPU box: Stage A (news-vector cache build) timing per encoder mode: scripts/time_stage_a.py [news] [chunk] [mode]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import ops, synth, util
from lime_cikm25_b200.config import default_config
n_news = int(sys.argv[1]) if len(sys.argv) > 1 else 65238
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = default_config(vocabulary_size=40000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(n_news, vocabulary_size=40000, seed=1)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=0)
model = model.cuda().eval()
ref = None
with torch.no_grad():
    model.scoring.fold()
    modes = (False, "x3-3pass", "x3-3pass", "x3", "x3", "bf16-3pass", True, True)
    if len(sys.argv) > 3:          # one mode only, twice (launch lists): "x3" or "bf16"
        modes = (True, True) if sys.argv[3] == "bf16" else (sys.argv[3], sys.argv[3])
    for mode in modes:
        model.news_encoder.engine.bf16 = mode is True
        model.news_encoder.engine.x3 = str(mode).startswith("x3")
        model.news_encoder.engine.bf16 = mode is True or str(mode).startswith("bf16")
        model.news_encoder.engine.x3_small = not str(mode).endswith("ffma-small")
        ops.X3_FUSED = not str(mode).endswith("3pass")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cache = util.build_news_cache(model, news, "cuda", **({"chunk": chunk} if chunk else {}))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        if ref is None:
            ref = cache.hist_rows[:, :400].clone()
        err = float((cache.hist_rows[:, :400] - ref).abs().max() / ref.abs().max())
        print("mode=%s  %.1f ms  %.0f news/s  %.1f TFLOP/s  max |dv| / max |v| vs fp32 = %.2e" % (mode, dt * 1e3, n_news / dt, n_news * 241.3e6 / dt / 1e12, err))
        del cache
