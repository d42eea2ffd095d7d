"""GPU box: where the fp32x3 encoder mode differs from the FFMA mode on the cache rows (signed error per column block vs fp64 folds)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import synth, util, engine
from lime_cikm25_b200.config import default_config as make_config
from oracle import lime_oracle as O
cfg = make_config(vocabulary_size=300, batch_size=8, word_embedding_init="skip", use_expired_penalty=False)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, 6)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
model = model.cuda().eval()
news = synth.make_news_table(40, vocabulary_size=300, seed=1)
imp = synth.make_impressions(3, news.news_num, cand_fixed=4, near_zero_frac=0.8, seed=2)
want32 = O.score_pairs_reference_style(sd, news, imp, cfg, 8).numpy().astype(np.float64)
want = O.score_pairs_reference_style(sd, news, imp, cfg, 8, dtype=torch.float64).double().numpy() if True else want32
print("fp32 oracle vs fp64 oracle: mean %+.2e max %.2e" % ((want32 - want).mean(), np.abs(want32 - want).max()))
from lime_cikm25_b200 import ops
rows = {}
with torch.no_grad():
    for mode in ("ffma", "x3"):
        for smode in (ops.SCORE_AUTO, ops.SCORE_EXACT):
            ops.score_configure(smode)
            e = model.news_encoder.engine
            e.x3 = mode.startswith("x3")
            cache = util.build_news_cache(model, news)
            got = util.score_impressions(model, cache, engine.DeviceImpressions(imp, "cuda"), 8).cpu().numpy().astype(np.float64)
            d = got - want
            print("%-5s scoring=%s  err vs fp64 oracle: mean %+.2e  max|.| %.2e   (rms want %.2f)" % (mode, "exact" if smode else "tc", d.mean(), np.abs(d).max(), np.sqrt((want ** 2).mean())))
            rows[mode] = (cache.hist_rows.double().cpu(), cache.cand_rows.double().cpu())
    ops.score_configure(ops.SCORE_AUTO)
    F = model.scoring.fold()
    G, Gg = F["G"].double().cpu(), F["Gg"].double().cpu()
    for mode in ("ffma", "x3"):
        h, c = rows[mode]
        vc = h[:, :400]
        gw = vc @ Gg.t()[:, :400] if Gg.shape[1] == 400 else None
        cf = vc @ G[:engine.CAND_NFOLD].t()
        blocks = [("w1", 0, 400), ("w2", 400, 800), ("w3", 800, 1200), ("scalars", 1200, engine.CAND_NFOLD)]
        for name, a, b in blocks:
            dd = c[:, a:b] - cf[:, a:b]
            print("%-5s fold %-8s given its own vc: mean %+.2e max %.2e  (rms %.2e)" % (mode, name, dd.mean(), dd.abs().max(), cf[:, a:b].pow(2).mean().sqrt()))
        if gw is not None:
            dd = h[:, 400:800] - gw
            print("%-5s fold gw: mean %+.2e max %.2e (rms %.2e)" % (mode, dd.mean(), dd.abs().max(), gw.pow(2).mean().sqrt()))
    dv = rows["x3"][0][:, :400] - rows["ffma"][0][:, :400]
    print("vc x3 - ffma: mean %+.2e max %.2e rms vc %.2e" % (dv.mean(), dv.abs().max(), rows["ffma"][0][:, :400].pow(2).mean().sqrt()))
    dc = rows["x3"][1][:, 1200:engine.CAND_NFOLD] - rows["ffma"][1][:, 1200:engine.CAND_NFOLD]
    print("scalars x3 - ffma per column: ", dc.mean(0).numpy(), " values ", rows["ffma"][1][:, 1200:engine.CAND_NFOLD].mean(0).numpy())
