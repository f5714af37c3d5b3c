"""CPU-side checks of the C ABI: the library builds, loads, and exports every symbol that
include/scqed.h declares; the product path refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT, load_golden


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from pycqed_b200 import engine
    return engine.load_library()


def test_header_symbols_are_exported(lib):
    from pycqed_b200 import engine
    header = open(os.path.join(ROOT, "include", "scqed.h")).read()
    declared = set(re.findall(r"\b(scqed_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(engine.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None


def test_version_and_counter(lib):
    assert b"sm_100a" in lib.scqed_version()
    assert lib.scqed_launch_count() >= 0


def test_no_cpu_fallback():
    import torch
    from pycqed_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _, plan = load_golden("c4_averin")
    with pytest.raises(engine.EngineError):
        engine.Engine(plan)
    with pytest.raises(engine.EngineError):
        engine.probe_fp64()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pycqed_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pycqed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert not re.search(r"^\s*(from|import)\s+scipy", src, re.M), f
                assert "eigsh(" not in src and "linalg.eigh(" not in src.replace("np.linalg.eigh(", ""), f
