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
