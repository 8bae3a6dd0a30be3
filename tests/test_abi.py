"""The C-ABI library must load on a CPU-only box and export exactly what include/stocs_b200.h
declares; creating a context without a B200 must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import model_matching_b200 as mm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "stocs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stocs_b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = mm.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(mm.SYMBOLS) == names
    assert L.stocs_b200_abi_version() == 2


def test_no_cpu_fallback():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(mm.StocsError):
        mm.Context(0)
    # the raw ABI reports the failure through its status code and error text
    h = ctypes.c_void_p()
    rc = mm.lib().stocs_b200_create(ctypes.byref(h), 0)
    assert rc < 0 and not h.value
    assert mm.lib().stocs_b200_last_error(None)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "model_matching_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in src.splitlines():
                    code = line.split("//")[0].split("#", 1)[0] if not line.lstrip().startswith("#include") else line
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", line), (f, line)
                    assert not (line.lstrip().startswith("#include") and "oracle" in line), (f, line)
                    # the only dlopen in the product binds NCCL at run time (csrc/comm.cu)
                    assert "liboracle" not in code and ("dlopen" not in code or f == "comm.cu"), (f, line)
                    assert not (f == "comm.cu" and "oracle" in line.lower()), (f, line)
