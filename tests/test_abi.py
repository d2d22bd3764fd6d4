"""CPU tests of the drop-in boundary: libppp_gpu.so loads, exports every symbol include/ppp_gpu.h
declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ppp_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from polishpathplanning_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libppp_gpu.so does not export %s" % n
        assert n in _lib.SIGNATURES, "python binding lacks %s" % n
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().ppp_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from polishpathplanning_b200 import api
    with pytest.raises(api.PPPError) as e:
        api.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The package must not reference oracle/ (only tests, smoke and bench's CPU legs may)."""
    pkg = os.path.join(ROOT, "polishpathplanning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "ppp_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_argument_validation_without_device():
    from polishpathplanning_b200 import _lib
    L = _lib.load()
    assert L.ppp_create(0, None) == _lib.PPP_ERR_INVALID
    assert b"NULL" in L.ppp_last_error()
    assert L.ppp_cloud_size(None) == -1
    assert L.ppp_cloud_free(None) == _lib.PPP_OK
    assert L.ppp_sync(None) == _lib.PPP_ERR_INVALID


def test_header_is_plain_c(tmp_path):
    """include/ppp_gpu.h must be consumable from C (and C++) with nothing but the standard headers."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "ppp_gpu.h"\nint main(void) { ppp_ctx* c = 0; (void)c; return PPP_OK + PPP_PAIR_SECT - 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only", str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-fsyntax-only", str(src)])
