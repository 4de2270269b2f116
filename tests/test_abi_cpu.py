"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol that
include/modcr_b200.h declares, the ctypes table mirrors the header, and -- without a GPU -- every compute
entry fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

import modcr_b200  # noqa: F401
from modcr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "modcr_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(modcr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header but not exported" % n
    assert lib.modcr_version() == 100


def test_ctypes_table_mirrors_header():
    declared = set(header_functions())
    bound = set(_lib.SIGNATURES) - _lib.OPTIONAL
    assert bound <= declared, bound - declared
    assert declared <= set(_lib.SIGNATURES), declared - set(_lib.SIGNATURES)


def test_struct_layouts_match_header_field_order():
    src = open(HEADER).read()
    for cname, cls in (("modcr_gemm_desc", _lib.GemmDesc), ("modcr_attn_desc", _lib.AttnDesc),
                       ("modcr_scorer_desc", _lib.ScorerDesc), ("modcr_rowmap", _lib.RowMap)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.replace("*", " ").split()
            # "const void* A" -> A ; "int32_t rows, K, C, G" -> rows K C G
            first = decl.split(",")
            fields.append(first[0].replace("*", " ").split()[-1])
            fields += [f.strip().lstrip("*") for f in first[1:]]
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_no_gpu_means_loud_failure_not_fallback():
    lib = _lib.load()
    buf = (ctypes.c_float * 8)()
    rc = lib.modcr_key_bias(ctypes.addressof(buf), ctypes.addressof(buf), 8, None)
    assert rc == -3                                          # MODCR_ERR_ARCH
    assert "CUDA" in _lib.last_error() or "sm_100" in _lib.last_error()
    from modcr_b200 import ops
    with pytest.raises(RuntimeError):
        ops.key_bias(torch.zeros(8), torch.zeros(8))         # CPU tensors are rejected, nothing is computed on the host


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()
