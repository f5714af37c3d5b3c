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


def _ctype_name(t):
    """Canonical spelling of a ctypes type (c_int32 is an alias of c_int on this platform)."""
    import ctypes as C
    if hasattr(t, "_length_") and hasattr(t, "_type_") and not isinstance(t._type_, str):
        return "%s*%d" % (_ctype_name(t._type_), t._length_)
    if hasattr(t, "contents"):
        return "P(%s)" % _ctype_name(t._type_)
    return "%s%d" % ("f" if t in (C.c_double, C.c_float) else "i", C.sizeof(t))


def test_struct_layouts_header_engine_integration_agree():
    """include/scqed.h is the source of truth: pycqed_b200/engine.py's ctypes structures and the reference-side
    stub printed in INTEGRATION.md must list exactly its fields, in its order, with its types."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_stub
    from pycqed_b200 import engine
    header = open(os.path.join(ROOT, "include", "scqed.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# pyscqed/_scqed\.py.*?)```", doc, re.S).group(1)
    classes = "\n".join(m.group(0) for m in re.finditer(r"^class \w+\(C\.Structure\):.*?\n(?=\n|\Z)", block, re.S | re.M))
    ns = {"C": C}
    exec(classes, ns)
    for cname, ours, theirs in (("scqed_plan_desc", engine._PlanDesc, ns["PlanDesc"]),
                                ("scqed_sweep_opts", engine._SweepOpts, ns["SweepOpts"])):
        want = [(f, _ctype_name(eval(t, {"C": C}))) for f, t in gen_stub.parse_struct(header, cname)]
        assert len(want) > 10
        for cls in (ours, theirs):
            got = [(f, _ctype_name(t)) for f, t in cls._fields_]
            assert got == want, (cname, cls)
        assert C.sizeof(ours) == C.sizeof(theirs)
