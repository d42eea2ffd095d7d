"""GPU box: the four transformer GEMMs of Stage A through lime_linear_bf16_tma at cache-build shapes.
scripts/bench_gemm_tma.py [rows]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lime_cikm25_b200 import ops
m = int(sys.argv[1]) if len(sys.argv) > 1 else 2048 * 128
dev = "cuda"
def run(name, n, k, kp, out_bf16, residual, act, ld_out=None):
    a = torch.randn(m, kp, device=dev).bfloat16(); a[:, k:] = 0
    w = (torch.randn(n, kp, device=dev) * k ** -0.5).bfloat16(); w[:, k:] = 0
    b = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev) if residual else None
    out = torch.empty(m, ld_out or n, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
    for _ in range(3):
        ops.linear_tma(a, w, b, residual=r, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear_tma(a, w, b, residual=r, act=act, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byts = m * (kp * 2 + (ld_out or n) * (2 if out_bf16 else 4) + (n * 4 if residual else 0))
    print("%-8s m=%d n=%d k=%d: %.3f ms  %.0f TFLOP/s (real k)  %.0f GB/s of operand traffic" % (name, m, n, k, ms, 2 * m * n * k / ms / 1e9, byts / ms / 1e6))
    return ms
t = run("qkv", 960, 300, 320, True, False, 0) + run("out_proj", 300, 300, 320, False, True, 0) + run("ffn1", 512, 300, 320, True, False, 1) + run("ffn2", 300, 512, 512, False, True, 0)
print("sum %.3f ms for %d rows -> %.1f ms per 65,238 news (160 tokens each)" % (t, m, t * 65238 * 160 / m))
