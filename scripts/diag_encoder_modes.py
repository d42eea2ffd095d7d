"""GPU box: worst relative error of the 900-d content vectors against the reference's golden vectors, per encoder mode."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
gd = os.path.join(ROOT, "tests", "golden")
for case in ("small_bs8", "buckets20"):
    cfg, news, imp, g, sd, model = T.load_case(case, gd)
    n0 = g["content"].shape[0]
    t = lambda a: torch.as_tensor(a[:n0]).to("cuda").contiguous()
    eng = model.news_encoder.engine
    for mode in ("fp32", "fp32x3", "bf16"):
        eng.bf16, eng.x3 = mode == "bf16", mode == "fp32x3"
        with torch.no_grad():
            c = eng.encode_content(t(news.title_text), t(news.body_text), t(news.category), t(news.subCategory))
        a, b = c.cpu().numpy().astype(np.float64), g["content"].astype(np.float64)
        floor = 0.1 * np.sqrt(np.mean(b * b))
        e = np.abs(a - b) / np.maximum(np.abs(b), floor)
        print("%-10s %-7s max rel %.3e  p99.9 %.3e  median %.3e  (rms abs err / rms %.3e)" % (case, mode, e.max(), np.quantile(e, 0.999), np.median(e), np.sqrt(np.mean((a - b) ** 2)) / np.sqrt(np.mean(b * b))))
    eng.bf16 = eng.x3 = False
