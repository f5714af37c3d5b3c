#!/usr/bin/env python
"""Generate the ctypes field lists of INTEGRATION.md section 2 from include/scqed.h (the single source of truth
of the C ABI).  `python tools/gen_stub.py` prints the two class bodies; tests/test_abi.py checks that the
header, pycqed_b200/engine.py and INTEGRATION.md agree field by field."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTYPE = {"int32_t": "C.c_int32", "int64_t": "C.c_int64", "double": "C.c_double", "float": "C.c_float"}


def parse_struct(header_text, name):
    """[(field, ctypes expression)] of ``typedef struct { ... } name;`` in declaration order."""
    end = re.search(r"\}\s*%s\s*;" % re.escape(name), header_text)
    if not end:
        raise KeyError(name)
    start = header_text.rfind("typedef struct", 0, end.start())
    body = header_text[header_text.index("{", start) + 1:end.start()]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        mm = re.match(r"(?:const\s+)?(\w+)\s*(\*?)\s*(\w+)(?:\[(\d+)\])?$", decl)
        if not mm:
            raise ValueError("cannot parse %r" % decl)
        base, ptr, field, arr = mm.groups()
        t = CTYPE[base]
        if ptr:
            t = "C.POINTER(%s)" % t
        if arr:
            t = "%s * %s" % (t, arr)
        out.append((field, t))
    return out


def render(fields, indent="                "):
    lines, cur = [], indent
    for f, t in fields:
        item = '("%s", %s), ' % (f, t)
        if len(cur) + len(item) > 118:
            lines.append(cur.rstrip())
            cur = indent
        cur += item
    lines.append(cur.rstrip().rstrip(","))
    return "\n".join(lines)


def stub():
    h = open(os.path.join(ROOT, "include", "scqed.h")).read()
    return ("class PlanDesc(C.Structure):          # mirrors scqed_plan_desc (generated: tools/gen_stub.py)\n"
            "    _fields_ = [\n%s]\n\n"
            "class SweepOpts(C.Structure):         # mirrors scqed_sweep_opts (generated: tools/gen_stub.py)\n"
            "    _fields_ = [\n%s]\n" % (render(parse_struct(h, "scqed_plan_desc")), render(parse_struct(h, "scqed_sweep_opts"))))


if __name__ == "__main__":
    print(stub())
