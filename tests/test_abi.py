"""The drop-in boundary on a box without a GPU: libta_b200.so loads, exports every symbol include/ta_b200.h declares, the
ctypes table agrees with the header (names and argument counts), no torch / C++ types leak into the C-ABI, and the
product fails loudly (no fallback) when the library or a device is missing. No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "ta_b200.h")


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(ta_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = (n, args)
    return decls


def test_header_is_plain_c_abi():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    for banned in ("torch", "at::", "std::", "Tensor", "template"):
        assert banned not in re.sub(r"/\*.*?\*/", "", src, flags=re.S), banned
    # compiles as C, not only as C++
    subprocess.check_call(["gcc", "-std=c11", "-fsyntax-only", "-x", "c", HEADER])


def test_library_exports_every_declared_symbol():
    from transferattack_b200 import _lib
    decls = _declarations()
    assert len(decls) >= 30
    lib = ctypes.CDLL(_lib.SO_PATH)
    for name in decls:
        assert hasattr(lib, name), "libta_b200.so does not export %s" % name
    assert lib.ta_version() == 1


def test_ctypes_table_matches_header():
    from transferattack_b200 import _lib
    decls = _declarations()
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, (n, args) in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == n, (name, n, len(_lib.SIGNATURES[name][1]), args)


def test_every_declaration_cites_the_reference():
    src = open(HEADER).read()
    for cite in ("attack.py:124-128", "attack.py:145-153", "utils.py:68-69", "dim.py:42-68", "tim.py:68-73", "sim.py:36-46",
                 "emifgsm.py", "vmifgsm.py:42-58", "utils.py:72-79", "utils.py:63-66", "admix.py:40-51", "nifgsm.py:35-39"):
        assert cite in src, cite


def test_no_silent_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from transferattack_b200 import ops
    ops._install_backend_for_tests(None)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ops.backend()


def test_missing_library_raises(monkeypatch, tmp_path):
    from transferattack_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.KernelLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "transferattack_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "ta_oracle" not in txt, os.path.join(dp, f)
