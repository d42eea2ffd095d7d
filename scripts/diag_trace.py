"""Diagnostic (GPU box).  Needs the trace instrumentation, kept out of the product kernel because its extra arguments cost
4 % (registers): git apply scripts/score_tc_trace_and_stacking.patch; make -C lime_cikm25_b200/csrc VARIANT=trace
EXTRA=-DLIME_TC_TRACE; LIME_B200_LIB=$PWD/lime_cikm25_b200/liblime_b200_trace.so python scripts/diag_trace.py
prints the pipeline timeline (clock64, relative to the unit start) of CTA 0's units 2..5."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import _lib, engine, synth, util
from lime_cikm25_b200.config import default_config

cfg = default_config(vocabulary_size=40000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(65238, vocabulary_size=40000, seed=1)
imp = synth.make_impressions(73152, news.news_num, seed=100)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=0)
model = model.cuda().eval()
with torch.no_grad():
    cache = util.build_news_cache(model, news, "cuda")
    dimp = engine.DeviceImpressions(imp, "cuda")
    out = torch.empty(dimp.num_pairs, dtype=torch.float32, device="cuda")
    for _ in range(3):
        util.score_impressions(model, cache, dimp, 32, out=out)
    torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
n = 4 * 4 * 16 * 8
buf = (ctypes.c_int64 * n)()
lib.lime_score_trace.argtypes = [ctypes.c_void_p, ctypes.c_int64]
assert lib.lime_score_trace(buf, n) == 0
t = np.array(buf[:], dtype=np.int64).reshape(4, 4, 16, 8)
for u in range(4):
    t0 = t[u, 3, 0, 0]
    r = lambda x: int(x - t0) if x else -1
    print("unit slot %d: cnt %d U %d" % (u, t[u, 3, 8, 0], t[u, 3, 9, 0]))
    print("  start w0/w6/iss %d %d %d | attn end w0 %d w6 %d | after bar %d | nodes done %d | prod end w0 %d w6 %d iss %d | sync %d | accum %d | epi end w0 %d w6 %d | bar %d | pool end w0 %d w6 %d iss %d"
          % (r(t[u,3,0,0]), r(t[u,3,0,1]), r(t[u,3,0,2]), r(t[u,3,1,0]), r(t[u,3,1,1]), r(t[u,3,2,0]), r(t[u,3,3,0]), r(t[u,3,4,0]), r(t[u,3,4,1]), r(t[u,3,4,2]),
             r(t[u,3,5,0]), r(t[u,3,6,0]), r(t[u,3,7,0]), r(t[u,3,7,1]), r(t[u,3,10,0]), r(t[u,3,11,0]), r(t[u,3,11,1]), r(t[u,3,11,2])))
    print("  issuer front: hist start %d end %d | dedup+cand end %d" % (r(t[u,3,12,0]), r(t[u,3,12,1]), r(t[u,3,12,2])))
    print("  stage: w0[start computeEnd arrive freeOK copiesOut] | w6[...] | issuer[fullSeen mmaIssued]")
    for k in range(13):
        print("   %2d: w0 %6d %6d %6d %6d %6d | w6 %6d %6d %6d %6d %6d | iss %6d %6d" % (
            (k,) + tuple(r(t[u,0,k,e]) for e in range(5)) + tuple(r(t[u,1,k,e]) for e in range(5)) + (r(t[u,2,k,0]), r(t[u,2,k,1]))))
