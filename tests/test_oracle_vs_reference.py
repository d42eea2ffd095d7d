"""Live cross-check of the oracle against the reference's own modules.  Only runs where
/root/reference is mounted (the build container); skipped on the GPU box."""
import os

import numpy as np
import pytest
import torch

from lime_cikm25_b200 import synth
from oracle import lime_oracle as O
from oracle import ref_import as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference tree not mounted")


def test_drop_in_state_dict_loads_strictly_into_reference():
    import lime_cikm25_b200 as L
    cfg = R.make_config(vocabulary_size=300)
    ref = R.build_reference_model(cfg)
    mine = L.Model(R.make_config(vocabulary_size=300, word_embedding_init="skip"))
    ref.load_state_dict(mine.state_dict(), strict=True)
    mine.load_state_dict(ref.state_dict(), strict=True)


def test_oracle_forward_equals_reference_forward():
    torch.set_num_threads(os.cpu_count())
    cfg = R.make_config(vocabulary_size=700, batch_size=64)       # prefix 64 > H: user-node rows used
    ref = R.build_reference_model(cfg)
    synth.synthetic_parameters(ref, 21)
    ref.eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    news = synth.make_news_table(80, vocabulary_size=700, seed=9)
    imp = synth.make_impressions(2, news.news_num, cand_fixed=33, near_zero_frac=1.0, seed=10)
    for batch in synth.impressions_to_pair_batches(news, imp, 64):
        tb = [torch.as_tensor(x) for x in batch]
        with torch.no_grad():
            want = ref(*tb, tb[24] - tb[23]).squeeze(1)
            got = O.model_forward(sd, batch, cfg).squeeze(1)
        assert float(((got - want).abs() / want.abs().clamp_min(1e-6)).max()) < 2e-5


def test_oracle_candidate_attention_equals_reference_layer():
    cfg = R.make_config(vocabulary_size=50)
    ref = R.build_reference_model(cfg)
    synth.synthetic_parameters(ref, 22)
    ref.eval()
    sd = ref.state_dict()
    g = torch.Generator().manual_seed(0)
    B, H, N = 3, 50, 5
    hist = torch.randn(B, H, 400, generator=g)
    ht, ct = torch.randn(B, H, 50, generator=g), torch.randn(B, N, 50, generator=g)
    mask = torch.rand(B, H, generator=g) < 0.6
    mask[1] = False                                               # all-padding history
    with torch.no_grad():
        want, wa = ref.user_encoder.candidate_aware_attn(hist, ht, ct, mask=mask)
        got, ga = O.candidate_aware_attention(sd, hist, ht, ct, mask)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)
    assert torch.allclose(ga, wa, atol=1e-7)


def test_oracle_training_layout_and_gradients_equal_reference():
    """Training layout (N = 1 + M candidates, model.train(), every dropout p = 0): logits, the
    reference trainer's loss (trainer.py:71-73) and its parameter gradients from the reference's own
    autograd vs autograd through the oracle restatement.  This pins the gradient oracle that the GPU
    backward kernels are tested against (tests/test_gpu_training.py)."""
    torch.set_num_threads(os.cpu_count())
    cfg = R.make_config(vocabulary_size=400, batch_size=4, dropout_rate=0.0)
    ref = R.build_reference_model(cfg)
    synth.synthetic_parameters(ref, 23)
    ref.train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                             # incl. the hard-coded 0.2 of layers.py:36
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    news = synth.make_news_table(60, vocabulary_size=400, seed=11)
    batch = synth.make_train_batch(news, 4, seed=12)
    tb = [torch.as_tensor(x) for x in batch]
    logits = ref(*tb, tb[24] - tb[23])
    loss = (-torch.log_softmax(logits, dim=1)[:, 0]).mean()
    loss.backward()
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in ref.state_dict().items()}
    got = O.model_forward(sd, batch, cfg)
    assert got.shape == logits.shape == (4, 5)
    assert float(((got - logits).abs() / logits.abs().clamp_min(1e-4)).max()) < 5e-5
    (-torch.log_softmax(got, dim=1)[:, 0]).mean().backward()
    checked = 0
    for name, p in ref.named_parameters():
        if p.grad is None or float(p.grad.abs().max()) == 0.0:
            continue
        g = sd[name].grad
        scale = float(p.grad.abs().max())
        assert float((g - p.grad).abs().max()) < 2e-4 * scale + 1e-9, name
        checked += 1
    assert checked >= 50
