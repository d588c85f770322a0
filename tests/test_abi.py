"""The C-ABI library: builds, loads, exports what include/mmalign.h declares, and
refuses to run without an sm_100 device (no CPU fallback).  No compute calls here."""
import ctypes as C
import re
import subprocess

import pytest

from conftest import ROOT


def test_header_symbols_are_exported(pkg):
    pkg._native.build()
    L = pkg._native.load()
    hdr = (ROOT / "include" / "mmalign.h").read_text()
    declared = set(re.findall(r"\b(mmalign_[a-z_]+)\s*\(", hdr))
    assert declared == set(pkg._native.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.mmalign_abi_version() == 3


def test_struct_layout_matches_header(pkg, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mmalign.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(mmalign_params),sizeof(mmalign_out),offsetof(mmalign_params,lam_lex),offsetof(mmalign_params,path),'
                   'offsetof(mmalign_params,pipeline_rows));return 0;}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), "-o", str(exe), str(src)], check=True)
    a, b, c, d, e = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    P, O = pkg._native.Params, pkg._native.Out
    assert (a, b, c, d, e) == (C.sizeof(P), C.sizeof(O), P.lam_lex.offset, P.path.offset, P.pipeline_rows.offset)


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.MMAlignError) as e:
        pkg.AlignmentEngine(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    pkg_dir = ROOT / "multimodal-alignment-of-noisy-image-text-pairs-using-weak-supervision_b200"
    for f in list(pkg_dir.rglob("*.py")) + list(pkg_dir.rglob("*.cu")) + list(pkg_dir.rglob("*.cuh")):
        text = f.read_text()
        assert "oracle" not in text.replace("oracle/mmalign_oracle.c: orc_dot", ""), f


def test_integration_md_stub_matches_the_abi(pkg):
    """The ctypes structs shown to maintainers in INTEGRATION.md are the real ones (same fields, same size)."""
    text = (ROOT / "INTEGRATION.md").read_text()
    ns = {"C": C}
    exec(text[text.index("class Params(C.Structure):"):text.index("def check(ctx, rc):")], ns)
    for name in ("Params", "Out"):
        doc, real = ns[name], getattr(pkg._native, name)
        assert [f[0] for f in doc._fields_] == [f[0] for f in real._fields_], name
        assert C.sizeof(doc) == C.sizeof(real), name
