"""The C-ABI library loads and exports exactly what include/lime_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from lime_cikm25_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lime_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lime_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "header declares %s but the library does not export it" % n
    assert sorted(_lib.PROTOTYPES) == names, "ctypes binding and header disagree"


def test_abi_version(lib):
    assert lib.lime_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header(lib):
    # the ctypes mirrors against sizeof() as compiled from include/lime_b200.h
    assert ctypes.sizeof(_lib.LimeNewsCache) == lib.lime_sizeof_news_cache() == 12 * 8 + 12 * 4
    assert ctypes.sizeof(_lib.LimeImpressions) == lib.lime_sizeof_impressions() == 11 * 8 + 3 * 4 + 4


def test_smem_budget_query(lib):
    from lime_cikm25_b200.engine import choose_tile_c
    assert lib.lime_score_smem_bytes(50, 48) < 232448
    assert choose_tile_c(50) == 39 and choose_tile_c(56) == 39          # tensor-core path: 3 * (candidates + bucket pairs) <= 120 M rows
    assert choose_tile_c(64) in (8, 16, 24, 32, 40, 48)                 # beyond 56 slots: exact kernel
    assert lib.lime_score_smem_bytes(56, 39) <= 232448                  # the exact kernel doubles as the fallback
    assert choose_tile_c(200) in (8, 16, 24, 32, 40, 48)
    assert lib.lime_score_scratch_ints(10) == 14
    assert lib.lime_score_smem_bytes(200, choose_tile_c(200)) <= 232448


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lime_cikm25_b200 import ops
    with pytest.raises(_lib.LimeError):
        ops.bucketize(torch.ones(4), 10)
    assert lib.lime_device_count() == 0


def test_argument_errors_are_reported(lib):
    rc = lib.lime_bucketize(None, 4, 10, None, None)
    assert rc != 0
    assert b"lime_bucketize" in lib.lime_last_error()
