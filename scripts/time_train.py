"""GPU box: one training step (B=32, bf16 GEMM mode) in a loop, for timing and ncu launch lists: scripts/time_train.py [steps]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lime_cikm25_b200 as L
from lime_cikm25_b200 import autograd, synth
from lime_cikm25_b200.config import default_config
from lime_cikm25_b200.trainer import Trainer
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cfg = default_config(vocabulary_size=40000, batch_size=32, word_embedding_init="skip")
news = synth.make_news_table(20000, vocabulary_size=40000, seed=1)
autograd.set_bf16(True)
autograd.set_tma(os.environ.get("LIME_TRAIN_TMA", "1") != "0")
torch.manual_seed(0)
model = L.Model(cfg); model.initialize(); synth.synthetic_parameters(model, seed=0)
model = model.cuda().train()
tr = Trainer(model, cfg)
batch = [torch.as_tensor(x).cuda() for x in synth.make_train_batch(news, 32, seed=500)]
for _ in range(2):
    tr.step(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.step(batch)
e1.record(); torch.cuda.synchronize()
print("train step %.2f ms  loss %.4f" % (e0.elapsed_time(e1) / steps, float(loss)))
