"""Unit parity of every C-ABI kernel against a plain torch reference of the same op (fp64 where
rounding matters).  All calls go through ctypes into liblime_b200.so."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lime_cikm25_b200 import ops  # noqa: E402
from oracle import lime_oracle as O  # noqa: E402

DEV = "cuda"


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


def randn(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=gen(seed)) * scale).to(DEV)


def close(got, want, tol=2e-5):
    want = want.double()
    err = (got.double() - want).abs().max().item()
    scale = want.pow(2).mean().sqrt().item() + 1e-30
    assert err <= tol * scale, "max err %.3e vs rms %.3e" % (err, scale)


def test_bucketize_bit_exact(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "buckets.npz"))
    for nb in (10, 20, 50):
        x = torch.from_numpy(g["x_%d" % nb]).to(DEV)
        got = ops.bucketize(x, nb)
        # (1) the reference's own op sequence (newsEncoders.py:54-57) evaluated by torch on this GPU:
        #     bit-exact everywhere, knife edges included
        xc = torch.clamp(x.float(), min=1)
        want = torch.clamp((torch.log(xc) / torch.log(torch.tensor(60 * 60 * 24.0)) * (nb / 7)).long(), max=nb - 1)
        bad = (got.long() != want).nonzero().flatten()
        assert bad.numel() == 0, "differs from torch-CUDA at x=%s" % x[bad[:8]].tolist()
        # (2) the golden ids the reference produced on CPU.  The reference's two device paths disagree
        #     with each other on knife-edge inputs: ATen's CUDA kernel divides by the 0-dim tensor
        #     log(86400) through a reciprocal multiply while its CPU kernel truly divides, and the CPU
        #     (SLEEF) and CUDA logf differ by up to 1 ulp.  An input is a knife edge when any of those
        #     variations moves it across a bucket boundary; everywhere else the ids must be identical.
        xh = torch.clamp(torch.from_numpy(g["x_%d" % nb]), min=1)
        ld = torch.log(torch.tensor(60 * 60 * 24.0))
        lg = torch.log(xh)
        variants = []
        for l in (lg, torch.nextafter(lg, torch.full_like(lg, float("inf"))), torch.nextafter(lg, torch.full_like(lg, -float("inf")))):
            variants.append(torch.clamp((l / ld * (nb / 7)).long(), max=nb - 1))
            variants.append(torch.clamp((l * (torch.tensor(1.0) / ld) * (nb / 7)).long(), max=nb - 1))
        agree = torch.stack([v == variants[0] for v in variants]).all(0).numpy()
        assert agree.mean() > 0.9
        mism = (got.cpu().numpy() != g["b_%d" % nb]) & agree
        assert mism.sum() == 0, "differs from the CPU golden at x=%s" % g["x_%d" % nb][mism][:8].tolist()


@pytest.mark.parametrize("m,n,k", [(1, 7, 52), (37, 300, 300), (129, 900, 300), (256, 512, 300), (300, 300, 512),
                                   (77, 1207, 400), (100, 510, 52), (64, 400, 900), (100, 900, 1000), (5, 1200, 352)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear(lib, m, n, k, act):
    a, w, b = randn(m, k, seed=1), randn(n, k, seed=2, scale=k ** -0.5), randn(n, seed=3)
    r = randn(m, n, seed=4)
    got = ops.linear(a, w, b, residual=r, act=act)
    z = a.double() @ w.double().t() + b.double()
    z = [z, torch.relu(z), torch.tanh(z)][act] + r.double()
    close(got, z, 1e-5)
    close(ops.linear(a, w, act=act), [lambda t: t, torch.relu, torch.tanh][act](a.double() @ w.double().t()), 1e-5)


@pytest.mark.parametrize("m,n,k", [(128, 16, 64), (1, 7, 52), (300, 900, 300), (1000, 300, 512), (77, 512, 300),
                                   (4096, 300, 300), (513, 256, 64), (130, 1207, 400)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_bf16_tcgen05(lib, m, n, k, act):
    """bf16 mode: operands rounded to bf16, products exact, fp32 accumulation in TMEM -> equals an fp64
    matmul of the bf16-rounded operands up to fp32 accumulation error."""
    a, w, b = randn(m, k, seed=1), randn(n, k, seed=2, scale=k ** -0.5), randn(n, seed=3)
    r = randn(m, n, seed=4)
    got = ops.linear(a, w, b, residual=r, act=act, bf16=True)
    z = a.bfloat16().double() @ w.bfloat16().double().t() + b.double()
    z = [z, torch.relu(z)][act] + r.double()
    close(got, z, 1e-5)
    # and it is a bf16-accurate approximation of the fp32 layer
    z32 = a.double() @ w.double().t() + b.double()
    close(got, [z32, torch.relu(z32)][act] + r.double(), 2e-2)
    # strided views, no bias / residual
    big = torch.zeros(m, n + 12, device=DEV)
    ops.linear(a, w, out=big[:, 4:4 + n], bf16=True)
    close(big[:, 4:4 + n], a.bfloat16().double() @ w.bfloat16().double().t(), 1e-5)
    assert float(big[:, :4].abs().sum()) == 0 and float(big[:, 4 + n:].abs().sum()) == 0


@pytest.mark.parametrize("m,n,k", [(128, 32, 64), (1, 7, 52), (300, 900, 300), (1000, 300, 512), (77, 512, 300),
                                   (40000, 300, 300), (20000, 900, 300), (513, 256, 64)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_bf16_tma(lib, m, n, k, act):
    """Stage A dense layer on bf16 activations (TMA + resident W slice + two TMEM accumulators): equals an fp64 matmul
    of the bf16 operands up to fp32 accumulation error; fp32 output with residual, bf16 output with zeroed padding."""
    kp = (k + 63) // 64 * 64
    a, w, b = randn(m, k, seed=1), randn(n, k, seed=2, scale=k ** -0.5), randn(n, seed=3)
    r = randn(m, n, seed=4)
    a16 = torch.zeros(m, kp, dtype=torch.bfloat16, device=DEV)
    w16 = torch.zeros(n, kp, dtype=torch.bfloat16, device=DEV)
    a16[:, :k], w16[:, :k] = a, w
    z = a16.double() @ w16.double().t() + b.double()
    z = [z, torch.relu(z)][act]
    got = ops.linear_tma(a16, w16, b, residual=r, act=act, out_bf16=False)
    close(got, z + r.double(), 1e-5)
    ld = (n + 7) // 8 * 8 + 8
    got16 = ops.linear_tma(a16, w16, b, act=act, ld_out=ld)
    assert got16.dtype == torch.bfloat16 and got16.shape == (m, ld)
    assert torch.equal(got16[:, :n], z.float().bfloat16()) or float((got16[:, :n].double() - z).abs().max()) <= 2.0 ** -7 * float(z.abs().max())
    assert float(got16[:, n:].float().abs().sum()) == 0.0


@pytest.mark.parametrize("m,n,k", [(300, 900, 300), (1000, 300, 512), (77, 512, 300), (5, 7, 52), (333, 400, 900)])
@pytest.mark.parametrize("fp16", [True, False])
def test_linear_x3_fp32_accurate(lib, m, n, k, fp16):
    """Three accumulating tensor-core passes on hi / lo operand pairs reproduce the fp32 layer: fp16 pairs (pre-scaled by powers
    of two, undone by alpha) to ~1e-6 of the row scale, bf16 pairs to ~3e-5 (one bf16 product: 1e-2); the residual after (no
    activation) or the partial sums before the activation (ReLU)."""
    a, w, b, r = randn(m, k, seed=1), randn(n, k, seed=2, scale=k ** -0.5), randn(n, seed=3), randn(m, n, seed=4)
    sa, sw = (ops.X3_ACT_SCALE, ops.X3_W_SCALE) if fp16 else (1.0, 1.0)
    ah, al = ops.split16(a, scale=sa, fp16=fp16)
    assert float((ah.float()[:, :k] + al.float()[:, :k] - sa * a).abs().max()) <= (2.0 ** -21 if fp16 else 2.0 ** -16) * sa * float(a.abs().max())
    assert float(ah[:, k:].float().abs().sum()) == 0 and float(al[:, k:].float().abs().sum()) == 0
    wh, wl = ops.split16(w, scale=sw, fp16=fp16)
    z = a.double() @ w.double().t() + b.double()
    tol = 6e-6 if fp16 else 5e-5          # (fp16 pairs: the remaining error is the tensor core's own fp32 accumulation)
    close(ops.linear_x3(ah, al, wh, wl, b, residual=r, alpha=1.0 / (sa * sw)), z + r.double(), tol)
    if k <= ops.X3_MAX_K:                 # longer contractions accumulate column slices in place: no activation there
        close(ops.linear_x3(ah, al, wh, wl, b, act=ops.ACT_RELU, alpha=1.0 / (sa * sw)), torch.relu(z), tol)


def test_stage_a_bf16_glue(lib):
    """bf16 images written beside the fp32 rows: embed_pe_bf16, layernorm_bf16, mha_bf16 against their fp32 twins."""
    n, T, d, heads = 5, 32, 300, 10
    E, pe = randn(50, d, seed=1), randn(T, d, seed=2)
    ids = torch.arange(n * T, device=DEV, dtype=torch.int32) % 50
    x, x16 = torch.empty(n * T, d, device=DEV), torch.full((n * T, 320), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.embed_pe_bf16(E, ids, T, pe, x, x16)
    want = torch.empty(n * T, d, device=DEV)
    ops.embed_pe(E, ids, T, pe, want)
    assert torch.equal(x, want) and torch.equal(x16[:, :d], want.bfloat16()) and float(x16[:, d:].float().abs().sum()) == 0
    g, bta = randn(d, seed=3), randn(d, seed=4)
    y, y16 = torch.empty_like(x), torch.full((n * T, 320), 7.0, dtype=torch.bfloat16, device=DEV)
    ops.layernorm_bf16(x, g, bta, y, y16)
    wy = ops.layernorm(x, g, bta, torch.empty_like(x))
    assert torch.equal(y, wy) and torch.equal(y16[:, :d], wy.bfloat16()) and float(y16[:, d:].float().abs().sum()) == 0
    for TT in (32, 128):
        nn_ = 3
        plain = randn(nn_ * TT, 900, seed=5).bfloat16()
        hp = torch.zeros(nn_ * TT, 3, heads, 32, dtype=torch.bfloat16, device=DEV)       # head-padded q | k | v
        hp[..., :30] = plain.view(nn_ * TT, 3, heads, 30)
        ctx16 = torch.full((nn_ * TT, 320), 7.0, dtype=torch.bfloat16, device=DEV)
        ops.mha_bf16(hp.view(nn_ * TT, 960), ctx16, nn_, TT, d, heads)
        ctx = torch.empty(nn_ * TT, d, device=DEV)
        ops.mha(plain.float().contiguous(), ctx, nn_, TT, d, heads)
        assert float((ctx16[:, :d].float() - ctx).abs().max()) <= 2.0 ** -6 * float(ctx.abs().max())
        assert float(ctx16[:, d:].float().abs().sum()) == 0


def test_linear_strided_views(lib):
    big_a = randn(50, 852, seed=5)
    big_w = randn(400, 1800, seed=6, scale=0.03)
    out = torch.zeros(50, 1720, device=DEV)
    ops.linear(big_a[:, 400:800], big_w[:, 900:1300], out=out[:, 1208:1608])
    close(out[:, 1208:1608], big_a[:, 400:800].double() @ big_w[:, 900:1300].double().t(), 1e-5)
    assert float(out[:, :1208].abs().sum()) == 0 and float(out[:, 1608:].abs().sum()) == 0


def test_gemm_strided(lib):
    a, b = randn(400, 40, seed=7), randn(40, 50, seed=8)
    close(ops.gemm_strided(a, b, alpha=0.5), 0.5 * a.double() @ b.double())
    close(ops.gemm_strided(a.t(), a), a.double().t() @ a.double())           # transposed view
    out = torch.zeros(50, 10, 52, device=DEV)
    ops.gemm_strided(b.t(), b, out=out[:, 3, :50])
    close(out[:, 3, :50], b.double().t() @ b.double())


@pytest.mark.parametrize("T", [32, 128])
def test_embed_pe_and_mha(lib, T):
    n, d, heads, V = 9, 300, 10, 100
    E = randn(V, d, seed=9, scale=0.1)
    ids = torch.randint(0, V, (n * T,), generator=gen(10)).to(torch.int32).to(DEV)
    pe = O.positional_encoding(T, d, torch.float32).to(DEV)
    x = torch.empty(n * T, d, device=DEV)
    ops.embed_pe(E, ids, T, pe, x)
    assert torch.equal(x.view(n, T, d), E[ids.long()].view(n, T, d) + pe)
    qkv = randn(n * T, 3 * d, seed=11)
    ctx = torch.empty(n * T, d, device=DEV)
    ops.mha(qkv, ctx, n, T, d, heads)
    q, k, v = qkv.double().view(n, T, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d // heads), dim=-1)
    close(ctx.view(n, T, d), (att @ v).transpose(1, 2).reshape(n, T, d), 1e-5)
    # fp32x3 mode of the same core: fp16 hi / lo pairs on the tensor cores, same fp32-level bar against fp64
    ctx3 = torch.full((n * T, d), float("nan"), device=DEV)
    ops.mha(qkv, ctx3, n, T, d, heads, x3=True)
    close(ctx3.view(n, T, d), (att @ v).transpose(1, 2).reshape(n, T, d), 1e-5)
    ph, pl = ops.mha_x3_pairs(qkv, n, T, d, heads, 16.0)      # the same context as the next layer's fp16 operand pair
    assert ph.shape == (n * T, 320) and float(ph[:, d:].float().abs().sum()) == 0 and float(pl[:, d:].float().abs().sum()) == 0
    close((ph[:, :d].double() + pl[:, :d].double()).view(n, T, d) / 16.0, (att @ v).transpose(1, 2).reshape(n, T, d), 1e-5)
    big = qkv * 2.0                                          # sharper softmax (logits up to +-15), larger operands
    ops.mha(big, ctx3, n, T, d, heads, x3=True)
    q, k, v = big.double().view(n, T, 3, heads, d // heads).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d // heads), dim=-1)
    close(ctx3.view(n, T, d), (att @ v).transpose(1, 2).reshape(n, T, d), 1e-5)


@pytest.mark.parametrize("m,n,k", [(1000, 512, 300), (333, 200, 320), (128, 64, 512)])
def test_linear_x3_pairs_match_split16(lib, m, n, k):
    """lime_linear_x3_pairs_tma: the operand pair written by the GEMM's epilogue is, bit for bit, lime_split_bf16_pairs of the fp32
    result of lime_linear_x3_tma (same accumulators, same scale-then-split arithmetic); padding columns zero."""
    a, w, b = randn(m, k, seed=31), randn(n, k, seed=32, scale=k ** -0.5), randn(n, seed=33)
    sa, sw = ops.X3_ACT_SCALE, ops.X3_W_SCALE
    ah, al = ops.split16(a, scale=sa)
    wh, wl = ops.split16(w, scale=sw)
    for act in (ops.ACT_RELU, ops.ACT_NONE):
        if act == ops.ACT_NONE and k > ops.X3_FUSED_MAX_K:
            continue          # linear_x3 accumulates such a contraction in 320-column slices (another summation order)
        z = ops.linear_x3(ah, al, wh, wl, b, act=act, alpha=1.0 / (sa * sw))
        want_hi, want_lo = ops.split16(z, scale=sa)
        hi, lo = ops.linear_x3_pairs(ah, al, wh, wl, b, act=act, alpha=1.0 / (sa * sw), out_scale=sa)
        assert hi.shape == want_hi.shape and torch.equal(hi, want_hi) and torch.equal(lo, want_lo)
        assert float(hi[:, n:].float().abs().sum()) == 0 and float(lo[:, n:].float().abs().sum()) == 0


def test_pair_producers_match_split16(lib):
    """fp32x3 mode: embed_pe_pairs / layernorm_pairs write the same fp32 rows as their plain twins and the same fp16 operand pair
    (bit for bit) as lime_split_bf16_pairs of those rows."""
    n, T, d, V = 5, 32, 300, 70
    E = randn(V, d, seed=21, scale=0.3)
    ids = torch.randint(0, V, (n, T), generator=gen(22)).to(torch.int32).to(DEV)
    pe = O.positional_encoding(T, d, torch.float32).to(DEV)
    x, x2 = torch.empty(n * T, d, device=DEV), torch.empty(n * T, d, device=DEV)
    ops.embed_pe(E, ids, T, pe, x)
    hi, lo = ops.embed_pe_pairs(E, ids.reshape(-1), T, pe, x2, ops.X3_ACT_SCALE)
    wh, wl = ops.split16(x, scale=ops.X3_ACT_SCALE)
    assert torch.equal(x, x2) and torch.equal(hi, wh) and torch.equal(lo, wl) and hi.shape == (n * T, 320)
    g, b = randn(d, seed=23) * 0.1 + 1, randn(d, seed=24) * 0.1
    y, y2 = torch.empty_like(x), torch.empty_like(x)
    ops.layernorm(x, g, b, y)
    hi, lo = ops.layernorm_pairs(x, g, b, y2, ops.X3_ACT_SCALE)
    wh, wl = ops.split16(y, scale=ops.X3_ACT_SCALE)
    assert torch.equal(y, y2) and torch.equal(hi, wh) and torch.equal(lo, wl)


def test_layernorm_and_meanpool(lib):
    rows, d, T = 7 * 32, 300, 32
    x, g, b = randn(rows, d, seed=12) + 0.3, randn(d, seed=13) * 0.1 + 1, randn(d, seed=14) * 0.1
    y = torch.empty_like(x)
    ops.layernorm(x, g, b, y)
    want = torch.nn.functional.layer_norm(x.double(), (d,), g.double(), b.double(), 1e-5)
    close(y, want, 1e-5)
    out = torch.zeros(7, 352, device=DEV)
    ops.layernorm_meanpool(x, g, b, out, 7, T)
    close(out[:, :300], want.view(7, T, d).mean(1), 1e-5)
    assert float(out[:, 300:].abs().sum()) == 0
    y400 = torch.empty(5, 400, device=DEV)
    x400 = randn(5, 400, seed=15)
    ops.layernorm(x400, randn(400, seed=16), randn(400, seed=17), y400)
    close(y400, torch.nn.functional.layer_norm(x400.double(), (400,), randn(400, seed=16).double(), randn(400, seed=17).double()), 1e-5)


def test_topic_intent_content(lib):
    n = 33
    ce, se = randn(18, 50, seed=18, scale=0.1), randn(270, 50, seed=19, scale=0.1)
    W, b = randn(50, 100, seed=20, scale=0.1), randn(50, seed=21, scale=0.1)
    cat = torch.randint(0, 18, (n,), generator=gen(22)).to(torch.int32).to(DEV)
    sub = torch.randint(0, 270, (n,), generator=gen(23)).to(torch.int32).to(DEV)
    out = torch.full((n, 352), 7.0, device=DEV)
    ops.topic_rep(ce, se, W, b, cat, sub, out[:, 300:], 52)
    want = torch.cat([ce[cat.long()], se[sub.long()]], 1).double() @ W.double().t() + b.double()
    close(out[:, 300:350], want, 1e-5)
    assert float(out[:, 350:].abs().sum()) == 0 and float((out[:, :300] - 7).abs().sum()) == 0
    # intent attention pooling
    e, pre, w2 = randn(n, 3, 400, seed=24).relu(), randn(n, 3, 400, seed=25), randn(400, seed=26, scale=0.07)
    pooled = torch.empty(n, 400, device=DEV)
    ops.intent_pool(pre.view(n * 3, 400), e.view(n, 1200), w2, pooled, n, 3, 400)
    alpha = torch.softmax(torch.tanh(pre.double()) @ w2.double(), dim=1)
    close(pooled, (alpha.unsqueeze(-1) * e.double()).sum(1), 1e-5)
    # cosine gate + concat
    t, bd = randn(n, 400, seed=27).relu(), randn(n, 400, seed=28).relu()
    content = torch.empty(n, 900, device=DEV)
    ops.content_fuse(t, bd, ce, se, cat, sub, content)
    sim = (torch.nn.functional.cosine_similarity(t.double(), bd.double(), dim=1) + 1) / 2
    close(content, torch.cat([t.double(), sim.unsqueeze(1) * bd.double(), ce[cat.long()].double(), se[sub.long()].double()], 1), 1e-5)


def test_small_helpers(lib):
    Ef, El = randn(10, 500, seed=29), randn(10, 500, seed=30)
    out = torch.empty(100, 1000, device=DEV)
    ops.bucket_pairs(Ef, El, out)
    assert torch.equal(out.view(10, 10, 1000)[3, 7], torch.cat([Ef[3], El[7]]))
    M = randn(6, 40, seed=31)
    want = M.double() * (randn(6, seed=32).double() * 0.25).unsqueeze(1)
    ops.scale_rows(M, randn(6, seed=32), 0.25)
    close(M, want, 1e-6)
    Pm = randn(9, 33, seed=33)
    wantp = Pm.double().cumsum(0)
    ops.prefix_rows(Pm)
    close(Pm, wantp, 1e-6)


def test_rank_metrics_match_reference_golden(lib, golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    scores = torch.from_numpy(g["scores"]).to(DEV)
    labels = torch.from_numpy(g["labels"]).to(DEV)
    off = torch.from_numpy(g["cand_off"]).to(DEV)
    ranks, per = ops.rank_metrics(scores, labels, off)
    assert np.array_equal(ranks.cpu().numpy(), g["ranks"])          # stable ties, -0.0 == 0.0
    sums = ops.metrics_reduce(per).cpu().numpy()
    assert sums[4] == len(g["cand_off"]) - 1
    assert np.allclose(sums[:4] / sums[4], g["metrics"], rtol=0, atol=1e-12)
    # per-impression values against the oracle's restatement of evaluate.py
    per = per.cpu().numpy()
    o = g["cand_off"]
    for i in (0, 10, 57, 199):
        rk = O.rank_impression(g["scores"][o[i]:o[i + 1]])
        assert np.allclose(per[i], O.impression_metrics(rk, g["labels"][o[i]:o[i + 1]]), atol=1e-12)


def test_rank_metrics_edge_cases(lib):
    # empty impression is skipped (evaluate.py:44-45); a 1-candidate impression has no AUC
    scores = torch.tensor([0.5, 0.5, 0.5, -1.0, 2.0], device=DEV)
    labels = torch.tensor([0, 1, 0, 1, 0], dtype=torch.uint8, device=DEV)
    off = torch.tensor([0, 3, 3, 5], device=DEV)
    ranks, per = ops.rank_metrics(scores, labels, off)
    assert ranks.tolist() == [1, 2, 3, 2, 1]
    per = per.cpu().numpy()
    assert np.isnan(per[1]).all()
    assert per[0, 0] == 0.5 and per[0, 1] == 0.5 and per[2, 0] == 0.0
    sums = ops.metrics_reduce(torch.from_numpy(per).to(DEV)).cpu().numpy()
    assert sums[4] == 2
