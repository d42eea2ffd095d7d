"""Import the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  Used in the build container to (a) pin the oracle restatement in
``lime_oracle.py`` against the reference's own code and (b) generate the golden fixtures under
``tests/golden/`` (see ``make_golden.py``).  Nothing under ``lime_cikm25_b200/`` imports this, and
nothing that runs on the GPU box may: ``/root/reference`` does not exist there.

The reference cannot be imported as-is:
  * ``userEncoders.py:6``  imports torch_geometric  (absent)  -> stub with PyG-semantics GraphSAGE
  * ``userEncoders.py:11`` imports torch_scatter    (absent)  -> stub (unused by CROWN)
  * ``newsEncoders.py:6``  imports matplotlib       (absent)  -> stub (unused import)
  * ``corpus.py:8,11``     import nltk, torchtext   (absent)  -> stub (corpus is never constructed)
  * ``config.Config()`` needs dataset files + CUDA (``config.py:212,283-293``) -> attribute bag
  * ``NewsEncoder.__init__`` unpickles ``word_embedding-...pkl`` from CWD (``newsEncoders.py:173``)
    -> we write a synthetic pickle into a temp dir and chdir there while constructing the model.

The only third-party arithmetic on the hot path is ``torch_geometric.nn.GraphSAGE``
(``userEncoders.py:54-58,153``); its version is unpinned by the reference and PyG is not installed
here, so the stub below restates the published SAGEConv algorithm (mean aggregation over node dim
-2, ``lin_l(mean_j x_j) + lin_r(x_i)``, bias on ``lin_l`` only, no activation after the single
layer).  PARITY UNPINNED at exactly that boundary; everything else is the reference's own code.
"""
from __future__ import annotations

import contextlib
import io
import math
import os
import pickle
import sys
import tempfile
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("LIME_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


# --------------------------------------------------------------------------------------------
# torch_geometric stub: SAGEConv / GraphSAGE with PyG semantics (documented algorithm)
# --------------------------------------------------------------------------------------------
class _PygLinear(nn.Module):
    """torch_geometric.nn.dense.linear.Linear: weight [out,in], Kaiming-uniform(a=sqrt(5)) init."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(in_channels)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        return nn.functional.linear(x, self.weight, self.bias)


class _SAGEConv(nn.Module):
    """out_i = lin_l(mean_{j in N(i)} x_j) + lin_r(x_i); node dimension is -2 (MessagePassing
    default), flow source_to_target: edge_index[0]=source j, edge_index[1]=target i."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_l = _PygLinear(in_channels, out_channels, bias=True)
        self.lin_r = _PygLinear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        n = x.size(-2)
        msg = x.index_select(-2, src)                                  # x_j per edge
        agg = torch.zeros_like(x)
        agg.index_add_(-2, dst, msg)
        deg = torch.zeros(n, dtype=x.dtype, device=x.device)
        deg.index_add_(0, dst, torch.ones_like(dst, dtype=x.dtype))
        agg = agg / deg.clamp(min=1).view(*([1] * (x.dim() - 2)), n, 1)  # scatter-mean
        return self.lin_l(agg) + self.lin_r(x)


class _GraphSAGE(nn.Module):
    """BasicGNN with num_layers convs; num_layers == 1 -> a single in->out conv, no act/dropout."""

    def __init__(self, in_channels, hidden_channels, num_layers, out_channels=None, dropout=0.0, **kw):
        super().__init__()
        assert num_layers == 1, "stub restates only the configuration the reference uses"
        out_channels = hidden_channels if out_channels is None else out_channels
        self.convs = nn.ModuleList([_SAGEConv(in_channels, out_channels)])

    def forward(self, x, edge_index):
        return self.convs[0](x, edge_index)


class _LightGCN(nn.Module):
    """Constructed by userEncoders.CROWN (:59-61), never called; holds one Embedding."""

    def __init__(self, num_nodes, embedding_dim, num_layers, **kw):
        super().__init__()
        self.embedding = nn.Embedding(num_nodes, embedding_dim)


class _NoParam(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    mod("matplotlib")
    mod("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mod("torch_scatter", scatter_sum=None, scatter_softmax=None)
    mod("nltk")
    mod("nltk.tokenize", word_tokenize=None)
    mod("torchtext")
    mod("torchtext.vocab", GloVe=None)
    pyg_nn = mod("torch_geometric.nn", SAGEConv=_SAGEConv, GraphSAGE=_GraphSAGE, GCN=_NoParam,
                 LightGCN=_LightGCN, LGConv=_NoParam)
    pyg = mod("torch_geometric")
    pyg.nn = pyg_nn


# --------------------------------------------------------------------------------------------
# config attribute bag (config.py:24-115 defaults; dataset-derived sizes are harness parameters)
# --------------------------------------------------------------------------------------------
def make_config(**over):
    from lime_cikm25_b200.config import default_config
    return default_config(**over)


def _word_embedding_name(cfg):
    # newsEncoders.py:173
    return ("word_embedding-" + str(cfg.word_threshold) + "-" + str(cfg.word_embedding_dim) + "-" +
            cfg.tokenizer + "-" + str(cfg.max_title_length) + "-" + str(cfg.max_abstract_length) +
            "-" + cfg.dataset + ".pkl")


def load_reference():
    """Returns the reference's own modules (model, newsEncoders, userEncoders, layers, util, evaluate)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at " + REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # config.py / corpus.py are imported for their class names only; argparse is never run.
    with contextlib.redirect_stdout(io.StringIO()):      # newsEncoders.py:17 prints versions
        import model as ref_model                        # noqa
        import newsEncoders as ref_news                  # noqa
        import userEncoders as ref_user                  # noqa
        import layers as ref_layers                      # noqa
        import util as ref_util                          # noqa
        import evaluate as ref_eval                      # noqa
        import dataset as ref_dataset                    # noqa
    return types.SimpleNamespace(model=ref_model, newsEncoders=ref_news, userEncoders=ref_user,
                                 layers=ref_layers, util=ref_util, evaluate=ref_eval, dataset=ref_dataset)


def build_reference_model(cfg, seed=0, word_embedding=None):
    """Construct + initialize() the reference's Model on CPU with seeded weights.

    ``word_embedding``: optional [V,300] tensor; default N(0,0.1) with row 0 zero (the <PAD> row is
    zero-initialised by corpus.py:193-196)."""
    ref = load_reference()
    g = torch.Generator().manual_seed(seed + 12345)
    if word_embedding is None:
        word_embedding = torch.randn(cfg.vocabulary_size, cfg.word_embedding_dim, generator=g) * 0.1
        word_embedding[0].zero_()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        with open(os.path.join(td, _word_embedding_name(cfg)), "wb") as f:
            pickle.dump(word_embedding, f)
        os.chdir(td)
        try:
            torch.manual_seed(seed)
            m = ref.model.Model(cfg)
            m.initialize()
        finally:
            os.chdir(cwd)
    return m


def load_reference_model_over_plugins():
    """The reference's OWN ``model.py`` (unmodified source, loaded from REFERENCE_ROOT) with its three plugin imports --
    ``newsEncoders``, ``userEncoders`` and ``util.RemainingLifetimeWeighting`` -- resolved to lime_cikm25_b200's drop-in
    modules (variant 1).  Returns the module; ``.Model(config)`` is the reference's class."""
    import importlib.util
    if not reference_available():
        raise RuntimeError("reference tree not found at " + REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import lime_cikm25_b200.news_modules as b_news
    import lime_cikm25_b200.user_modules as b_user
    import lime_cikm25_b200.util as b_util
    saved = {k: sys.modules.get(k) for k in ("newsEncoders", "userEncoders", "util")}
    sys.modules["newsEncoders"], sys.modules["userEncoders"], sys.modules["util"] = b_news, b_user, b_util
    try:
        spec = importlib.util.spec_from_file_location("ref_model_over_b200", os.path.join(REFERENCE_ROOT, "model.py"))
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
