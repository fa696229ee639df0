"""CPU: the C-ABI library loads and exports every symbol include/sf_b200.h declares; record layouts agree;
the product fails loudly without a GPU (no fallback); host-only helpers work."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle.oracle import Record
from spacefortress_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sf_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_record_layout_matches_oracle_record():
    assert C.sizeof(_lib.StateRecord) == C.sizeof(Record) == 1816
    for (n1, t1), (n2, t2) in zip(_lib.StateRecord._fields_[:-1], Record._fields_[:-1]):
        assert n1 == n2 and C.sizeof(t1) == C.sizeof(t2)
        assert getattr(_lib.StateRecord, n1).offset == getattr(Record, n2).offset


def test_synthetic_action_stream_is_deterministic_and_uniform():
    L = _lib.lib()
    a = np.array([[L.sf_synthetic_action(7, e, t, 5) for e in range(64)] for t in range(200)])
    assert a.min() == 0 and a.max() == 4
    assert np.array_equal(a, np.array([[L.sf_synthetic_action(7, e, t, 5) for e in range(64)] for t in range(200)]))
    counts = np.bincount(a.reshape(-1), minlength=5) / a.size
    assert np.all(np.abs(counts - 0.2) < 0.02)
    assert not np.array_equal(a[0], a[1])


def test_invalid_arguments_return_status_codes():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.sf_create(b"nonsense", 1, 4, 0, C.byref(h)) == _lib.SF_ERR_INVALID
    assert b"Unknown config value" in L.sf_last_error()
    assert L.sf_create(b"youturn", 5, 4, 0, C.byref(h)) != 0
    assert L.sf_create(b"youturn", 1, 0, 0, C.byref(h)) != 0
    assert L.sf_num_envs(None) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (on a GPU box this test just creates an env)."""
    import torch
    from spacefortress_b200 import SFVecEnv
    if torch.cuda.is_available():
        env = SFVecEnv("youturn", num_envs=2)
        env.close()
    else:
        with pytest.raises(_lib.SFError):
            SFVecEnv("youturn", num_envs=2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spacefortress_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle's", "") or f in ("sf_tables.cpp",), (dirpath, f)
