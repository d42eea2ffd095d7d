"""Diagnostic (GPU box): error of the exact and tensor-core scoring kernels against the fp64 oracle on a
golden case's inputs, Stage B isolated (GPU-built cache for both)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import engine, ops, synth, util
from oracle import lime_oracle as O
from oracle.make_golden import case_inputs

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    floor = 0.1 * float(np.sqrt(np.mean(b * b))) + 1e-30
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))

for case in ("bs64", "small_bs8", "buckets20"):
    spec, cfg, news, imp = case_inputs(case)
    g = np.load(os.path.join(ROOT, "tests", "golden", case + ".npz"))
    cfg.word_embedding_init = "skip"
    cfg.use_remaining_lifetime_weighting = False
    model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, spec["weights_seed"])
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    with torch.no_grad():
        truth = O.score_pairs_reference_style(sd, news, imp, cfg, cfg.batch_size, dtype=torch.float64).numpy()
        cache = util.build_news_cache(model, news)
        dimp = engine.DeviceImpressions(imp, "cuda")
        out = {}
        for name, mode in (("exact", ops.SCORE_EXACT), ("tc", ops.SCORE_AUTO)):
            ops.score_configure(mode)
            out[name] = util.score_impressions(model, cache, dimp, cfg.batch_size).cpu().numpy()
        ops.score_configure(ops.SCORE_AUTO)
    print(case, "ref32-vs-fp64 %.2e  exact-vs-fp64 %.2e  tc-vs-fp64 %.2e  tc-vs-exact %.2e  exact-vs-ref32 %.2e tc-vs-ref32 %.2e  fallback_units %d"
          % (rel(g["base_scores"], truth), rel(out["exact"], truth), rel(out["tc"], truth), rel(out["tc"], out["exact"]),
             rel(out["exact"], g["base_scores"]), rel(out["tc"], g["base_scores"]), int(dimp.work_counter[1])))
