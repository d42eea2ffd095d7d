"""torch.autograd.Function wrappers: forward AND backward of every op are calls into liblime_b200.so
(ops.py); torch only records the graph, owns the buffers and accumulates parameter gradients.

These make ``Model.forward`` differentiable in training mode, so that the reference's own training
step (trainer.py:131-148: loss, ``backward()``, ``clip_grad_norm_``, ``Adam.step``) runs unchanged on
top of the B200 kernels.  All tensors are fp32 CUDA, 2-D row-major views unless stated otherwise.
"""
from __future__ import annotations

import torch

from . import ops


# "bf16 mode" of the training path (BASELINE.json configs[2]): every nn.Linear forward and backward GEMM runs
# on the tcgen05 tensor cores with bf16 operands and fp32 accumulation (lime_linear_bf16 / lime_gemm_bf16);
# parameters, activations, gradients and the optimizer stay fp32.  Off by default: fp32 is the parity mode.
BF16 = False


def set_bf16(enabled):
    global BF16
    BF16 = bool(enabled)


# In bf16 mode the forward GEMM and dX = dY . W of every nn.Linear run on the TMA-fed tcgen05 kernel of the Stage A cache
# build (lime_linear_bf16_tma, csrc/gemm_tma.cu: resident W slice, two TMEM accumulators) over bf16 casts of the operands;
# dW = dY^T . X (contraction over the rows) stays on lime_gemm_bf16.  False: the round-1 kernels (lime_linear_bf16 /
# lime_gemm_bf16 convert fp32 tiles in their producers).
TMA = True
_TMA_K = 512          # contraction length of one lime_linear_bf16_tma pass; longer ones accumulate in place


def set_tma(enabled):
    global TMA
    TMA = bool(enabled)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _rows_ok(t):
    return t.dim() == 2 and t.stride(1) == 1 and t.dtype == torch.float32


def _linear_tma(a16, w16, n, bias, residual, act, out=None):
    """act(a @ w.T + bias) + residual -> fp32 [m, n] through lime_linear_bf16_tma on bf16 images a16 [m, kp], w16 [n, kp]
    (ops.cast_bf16).  Contractions longer than 512 run as accumulating passes (then without activation)."""
    m, kp = a16.shape
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a16.device)
    if kp <= _TMA_K:
        return ops.linear_tma(a16, w16, bias, residual=residual, act=act, out=out, n=n, out_bf16=False)
    assert act == 0
    for c0 in range(0, kp, _TMA_K):
        c1 = min(kp, c0 + _TMA_K)
        ops.linear_tma(a16[:, c0:c1], w16[:, c0:c1], bias if c0 == 0 else None, residual=residual if c0 == 0 else out,
                       out=out, n=n, out_bf16=False)
    return out


class Linear(torch.autograd.Function):
    """y = act(x @ w.T + b) + residual  (lime_linear); backward with lime_gemm / lime_col_sum."""

    @staticmethod
    def forward(ctx, x, w, b, act, residual):
        if act and residual is not None:
            raise ValueError("activation and residual cannot be combined (the activation output is needed)")
        tma = BF16 and TMA and x.shape[0] > 0 and _rows_ok(x) and _rows_ok(w) and (x.shape[1] <= _TMA_K or act == 0)
        if tma:
            x16 = ops.cast_bf16(x)                       # kept for dW = dZ^T X (the fp32 x is not needed again)
            y = _linear_tma(x16, ops.cast_bf16(w), w.shape[0], b, residual, act)
        else:
            y = ops.linear(x, w, b, residual=residual, act=act, bf16=BF16)
        ctx.act = act
        ctx.bf16 = BF16
        ctx.tma = tma
        ctx.xshape = tuple(x.shape)
        ctx.save_for_backward(x16 if tma else x, w, y if act else None)
        ctx.has_b, ctx.has_res = b is not None, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _c(dy)
        dz = ops.act_bwd(dy, y, ctx.act) if ctx.act else dy
        m, k = ctx.xshape
        n = w.shape[0]
        db = None
        if ctx.tma:
            # one bf16 image of dZ for both gradients; the bias gradient comes out of the same pass over dZ
            if ctx.has_b and ctx.needs_input_grad[2]:
                dz16, db = ops.cast_bf16_colsum(dz)
            else:
                dz16 = ops.cast_bf16(dz)
            # dX = dZ . W (contraction over n) and dW = dZ^T . X (contraction over the rows, x = the saved bf16 image)
            dx = _linear_tma(dz16, ops.cast_bf16(w.t().contiguous()), k, None, None, 0) if ctx.needs_input_grad[0] else None
            dw = ops.gemm_tn_tma(dz16, x, n, k) if ctx.needs_input_grad[1] else None
        else:
            dx = ops.gemm(dz, True, w, False, m, k, n, bf16=ctx.bf16) if ctx.needs_input_grad[0] else None
            dw = ops.gemm(dz, False, x, False, n, k, m, bf16=ctx.bf16) if ctx.needs_input_grad[1] else None
        if db is None and ctx.has_b and ctx.needs_input_grad[2]:
            db = ops.col_sum(dz)
        dres = dy if ctx.has_res and ctx.needs_input_grad[4] else None
        return dx, dw, db, None, dres


def linear(x, w, b=None, act=0, residual=None):
    return Linear.apply(x, w, b, act, residual)


class LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y = torch.empty_like(x)
        ops.layernorm(x, gamma, beta, out=y, eps=eps)
        ctx.eps = eps
        ctx.save_for_backward(x, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(x, gamma, _c(dy), ctx.eps)
        return dx, dg, db, None


class LayerNormMeanPool(torch.autograd.Function):
    """LayerNorm of every token + unmasked mean over the T tokens of a news -> [n, d]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, n, T, eps):
        out = torch.empty((n, x.shape[1]), dtype=torch.float32, device=x.device)
        ops.layernorm_meanpool(x, gamma, beta, out, n, T, eps=eps)
        ctx.T, ctx.eps = T, eps
        ctx.save_for_backward(x, gamma)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(x, gamma, _c(dout), ctx.eps, bcast_T=ctx.T)
        return dx, dg, db, None, None, None


class EmbedPE(torch.autograd.Function):
    """word_embedding(ids) + positional encoding; dense [V, d] gradient like nn.Embedding."""

    @staticmethod
    def forward(ctx, E, ids, T, pe):
        out = torch.empty((ids.numel(), E.shape[1]), dtype=torch.float32, device=E.device)
        ops.embed_pe(E, ids, T, pe, out)
        ctx.save_for_backward(ids)
        ctx.shape = E.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        dE = torch.zeros(ctx.shape, dtype=torch.float32, device=dout.device)
        ops.scatter_add_rows(_c(dout), ids.reshape(-1), dE)
        return dE, None, None, None


class Gather(torch.autograd.Function):
    """table[ids] -> [n, d]  (category / bucket-table lookups)."""

    @staticmethod
    def forward(ctx, table, ids):
        ctx.save_for_backward(ids)
        ctx.shape = table.shape
        return ops.gather_rows(table, ids)

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        dT = torch.zeros(ctx.shape, dtype=torch.float32, device=dout.device)
        ops.scatter_add_rows(_c(dout), ids, dT)
        return dT, None


class MHA(torch.autograd.Function):
    """nn.MultiheadAttention core; p_drop / seed: dropout on the attention weights (same stateless mask both ways)."""

    @staticmethod
    def forward(ctx, qkv, n, T, d, heads, p_drop=0.0, seed=0):
        out = torch.empty((qkv.shape[0], d), dtype=torch.float32, device=qkv.device)
        ctx.bf16 = BF16 and TMA                  # bf16 mode: forward and backward on the tensor cores (lime_mha_{fwd,bwd}_bf16)
        ops.mha(qkv, out, n, T, d, heads, p_drop, seed, bf16=ctx.bf16)
        ctx.dims = (n, T, d, heads, float(p_drop), int(seed))
        ctx.save_for_backward(qkv)
        return out

    @staticmethod
    def backward(ctx, dctx):
        (qkv,) = ctx.saved_tensors
        n, T, d, heads, p_drop, seed = ctx.dims
        return ops.mha_bwd(qkv, _c(dctx), n, T, d, heads, p_drop, seed, bf16=ctx.bf16), None, None, None, None, None, None


class IntentPool(torch.autograd.Function):
    """layers.Attention over the k intents: pre [n*k, D] (affine1 output), e [n*k, D], w2 [D] -> [n, D]."""

    @staticmethod
    def forward(ctx, pre, e, w2, n, k):
        D = e.shape[1]
        out = torch.empty((n, D), dtype=torch.float32, device=e.device)
        ops.intent_pool(pre, e, w2, out, n, k, D)
        ctx.dims = (n, k, D)
        ctx.save_for_backward(pre, e, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        pre, e, w2 = ctx.saved_tensors
        n, k, D = ctx.dims
        dpre, de, dw2 = ops.intent_pool_bwd(pre, e, w2, _c(dout), n, k, D)
        return dpre, de, dw2, None, None


class ContentFuse(torch.autograd.Function):
    """[title | sim * body | dropout(cat_emb[c]) | dropout(sub_emb[s])] -> [n, 900]."""

    @staticmethod
    def forward(ctx, title, body, cat_table, sub_table, cat, sub, p, seed):
        n, D = title.shape
        cd, sd = cat_table.shape[1], sub_table.shape[1]
        out = torch.empty((n, 2 * D + cd + sd), dtype=torch.float32, device=title.device)
        ops.content_fuse(title, body, cat_table, sub_table, cat, sub, out)
        if p > 0:                                   # feature_fusion drops the two category embeddings (:224)
            ops.dropout(out[:, 2 * D:2 * D + cd], p, seed, out=out[:, 2 * D:2 * D + cd])
            ops.dropout(out[:, 2 * D + cd:], p, seed + 1, out=out[:, 2 * D + cd:])
        ctx.meta = (D, cd, sd, p, seed, cat_table.shape)
        ctx.save_for_backward(title, body, cat)
        return out

    @staticmethod
    def backward(ctx, dout):
        title, body, cat = ctx.saved_tensors
        D, cd, sd, p, seed, tshape = ctx.meta
        dout = _c(dout)
        dt, db = ops.content_fuse_bwd(title, body, dout)
        dcat_table = None
        if ctx.needs_input_grad[2]:
            g = dout[:, 2 * D:2 * D + cd]
            if p > 0:
                g = ops.dropout(g, p, seed)
            dcat_table = torch.zeros(tshape, dtype=torch.float32, device=dout.device)
            ops.scatter_add_rows(g, cat, dcat_table)
        return dt, db, dcat_table, None, None, None, None, None


class Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout(x, p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(_c(dy), ctx.p, ctx.seed), None, None


def dropout(x, p, seed):
    return Dropout.apply(x, p, seed) if p > 0 else x


def _quad_ok(*ts):
    return all(t.dim() == 2 and t.stride(1) == 1 and t.shape[1] % 4 == 0 and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0
               for t in ts)


class DropoutAdd(torch.autograd.Function):
    """dropout(x) + res in one pass (the post-LN residual branches of nn.TransformerEncoderLayer)."""

    @staticmethod
    def forward(ctx, x, res, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout_fused(x, p, seed, res=res)

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        return ops.dropout(dy, ctx.p, ctx.seed), dy, None, None


def dropout_add(x, res, p, seed):
    if p <= 0:
        return x + res
    if _quad_ok(x, res):
        return DropoutAdd.apply(x, res, p, seed)
    return Dropout.apply(x, p, seed) + res


class EmbedPEDropout(torch.autograd.Function):
    """dropout(dropout(word_embedding(ids)) + pe): one forward pass, one backward pass + the embedding scatter."""

    @staticmethod
    def forward(ctx, E, ids, T, pe, p, seed_w, seed_x):
        ctx.save_for_backward(ids)
        ctx.meta = (E.shape, p, seed_w, seed_x)
        return ops.embed_pe_dropout(E, ids, T, pe, p, seed_w, seed_x)

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        shape, p, seed_w, seed_x = ctx.meta
        g = ops.dropout_fused(_c(dout), p, seed_x, seed2=seed_w)
        dE = torch.zeros(shape, dtype=torch.float32, device=dout.device)
        ops.scatter_add_rows(g, ids.reshape(-1), dE)
        return dE, None, None, None, None, None, None


class BucketPairs(torch.autograd.Function):
    """[Ef[bf] | El[bl]] for every (bf, bl) -> [nb*nb, 2*dim]."""

    @staticmethod
    def forward(ctx, Ef, El):
        nb, dim = Ef.shape
        out = torch.empty((nb * nb, 2 * dim), dtype=torch.float32, device=Ef.device)
        ops.bucket_pairs(Ef, El, out)
        ctx.meta = (nb, dim)
        return out

    @staticmethod
    def backward(ctx, dout):
        nb, dim = ctx.meta
        dout = _c(dout)
        idx = torch.arange(nb * nb, dtype=torch.int32, device=dout.device)
        dEf = torch.zeros((nb, dim), dtype=torch.float32, device=dout.device)
        dEl = torch.zeros((nb, dim), dtype=torch.float32, device=dout.device)
        ops.scatter_add_rows(dout[:, :dim], torch.div(idx, nb, rounding_mode="floor").to(torch.int32), dEf)
        ops.scatter_add_rows(dout[:, dim:], torch.remainder(idx, nb).to(torch.int32), dEl)
        return dEf, dEl


# ---- user encoder + click score (csrc/train_user.cu) -------------------------------------------------
class CAAttention(torch.autograd.Function):
    """(Qp [B*N,400], Kp [B*H,400], mask uint8 [B,H]) -> a [B*H] (layers.py:66-81); p_drop / seed: the dropout on
    the per-head attention weights (layers.py:36,74), same stateless mask in forward and backward."""

    @staticmethod
    def forward(ctx, Qp, Kp, mask, B, N, H, p_drop=0.0, seed=0):
        ctx.dims = (B, N, H, float(p_drop), int(seed))
        ctx.save_for_backward(Qp, Kp, mask)
        return ops.ca_attention_fwd(Qp, Kp, mask, B, N, H, p_drop, seed)

    @staticmethod
    def backward(ctx, da):
        Qp, Kp, mask = ctx.saved_tensors
        B, N, H, p_drop, seed = ctx.dims
        dQ, dK = ops.ca_attention_bwd(Qp, Kp, mask, B, N, H, _c(da), p_drop, seed)
        return dQ, dK, None, None, None, None, None, None


class RowScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, a):
        ctx.save_for_backward(v, a)
        return ops.row_scale_fwd(v, a)

    @staticmethod
    def backward(ctx, dwc):
        v, a = ctx.saved_tensors
        return ops.row_scale_bwd(v, a, _c(dwc))


class GateMix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, wc, v):
        ctx.save_for_backward(z, wc, v)
        return ops.gate_mix_fwd(z, wc, v)

    @staticmethod
    def backward(ctx, dout):
        z, wc, v = ctx.saved_tensors
        return ops.gate_mix_bwd(z, wc, v, _c(dout))


class SageMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, un, B, H, P):
        ctx.dims = (B, H, P, un.shape[0])
        return ops.sage_mean_fwd(x, un, B, H, P)

    @staticmethod
    def backward(ctx, dm):
        B, H, P, un_rows = ctx.dims
        dx, dun = ops.sage_mean_bwd(_c(dm), B, H, P, un_rows)
        return dx, dun, None, None, None


class AddRowBroadcast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r, l, H):
        ctx.dims = (l.shape[0], H)
        return ops.add_row_bcast(r, l, H)

    @staticmethod
    def backward(ctx, dg):
        B, H = ctx.dims
        dg = _c(dg)
        return dg, ops.sum_over_h(dg, B, H), None


class Pool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Kg, q, g, B, N, H):
        u, alpha = ops.pool_fwd(Kg, q, g, B, N, H)
        ctx.dims = (B, N, H)
        ctx.save_for_backward(Kg, q, g, alpha)
        return u

    @staticmethod
    def backward(ctx, du):
        Kg, q, g, alpha = ctx.saved_tensors
        B, N, H = ctx.dims
        dKg, dq, dg = ops.pool_bwd(Kg, q, g, alpha, _c(du), B, N, H)
        return dKg, dq, dg, None, None, None


class ClickScore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, c, remaining, alpha, beta, use_weighting, use_penalty):
        s, w = ops.click_score_fwd(u, c, remaining, alpha, beta, use_weighting, use_penalty)
        ctx.save_for_backward(u, c, w)
        return s

    @staticmethod
    def backward(ctx, ds):
        u, c, w = ctx.saved_tensors
        du, dc = ops.click_score_bwd(u, c, w, _c(ds))
        return du, dc, None, None, None, None, None
