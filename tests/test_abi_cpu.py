"""The C-ABI library loads without a GPU and exports every symbol include/sgb200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from imagegenerator_b200 import build, ops
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in sgb200.h but not exported"
    for n in ops.EXPORTS:
        assert n in names, f"{n} bound in ops.py but not declared in sgb200.h"
    lib.sg_version.restype = ctypes.c_int
    assert lib.sg_version() == 1


def test_no_cpu_fallback_without_device():
    import pytest
    import torch
    from imagegenerator_b200.ops import CudaOps
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        CudaOps("bf16")


def test_product_package_never_imports_the_oracle_or_the_tests():
    """oracle/ and tests/emu_ops.py are the checker, never the thing shipped: no module of the package may import them
    (bench.py may, for its cpu_baseline / --impl reference leg only; __graft_entry__.smoke() for its check)."""
    import re
    pkg = os.path.join(ROOT, "imagegenerator_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                for m in re.finditer(r"^\s*(from|import)\s+(oracle|emu_ops|tests)\b", src, re.M):
                    bad.append((f, m.group(0).strip()))
    assert not bad, bad
    # and the ctypes binding has no CPU fallback: constructing the backend without the library / a device raises
    src = open(os.path.join(pkg, "ops.py")).read()
    assert "raise" in src and "libsgb200" in src
