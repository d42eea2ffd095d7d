#!/usr/bin/env python
"""GEMM micro-benchmark at the training shapes of the body transformer (1760 news x 128 tokens): the tcgen05 bf16
kernels (forward NT, backward NN / TN with split-K) and the fp32 FFMA kernels, CUDA-event timed.  GPU box only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lime_cikm25_b200 import ops

M = 1760 * 128
def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

rows = []
for name, n, k in (("in_proj", 900, 300), ("out_proj", 300, 300), ("linear1", 512, 300), ("linear2", 300, 512)):
    x = torch.randn(M, k, device="cuda")
    w = torch.randn(n, k, device="cuda") * k ** -0.5
    dy = torch.randn(M, n, device="cuda")
    y = torch.empty(M, n, device="cuda")
    dx = torch.empty(M, k, device="cuda")
    dw = torch.empty(n, k, device="cuda")
    flop = 2.0 * M * n * k
    for bf16 in (True, False):
        t_f = timeit(lambda: ops.linear(x, w, out=y, bf16=bf16))
        t_x = timeit(lambda: ops.gemm(dy, True, w, False, M, k, n, out=dx, bf16=bf16))
        t_w = timeit(lambda: ops.gemm(dy, False, x, False, n, k, M, out=dw, bf16=bf16))
        r = {"layer": name, "m": M, "n": n, "k": k, "mode": "bf16 tcgen05" if bf16 else "fp32 FFMA",
             "fwd_ms": t_f, "dx_ms": t_x, "dw_ms": t_w, "fwd_tflops": flop / t_f / 1e9, "dx_tflops": flop / t_x / 1e9,
             "dw_tflops": flop / t_w / 1e9}
        rows.append(r)
        print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gemm_bench.jsonl"), "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
