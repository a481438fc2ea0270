"""The C-ABI shared library: loads, exports every symbol include/is3d_b200.h declares, and refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from is3d_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "is3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(is3d_b200_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(api.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "libis3d_b200.so does not export %s" % n


def test_struct_sizes_match_header():
    # int32 x9 + 4 pad + 2 doubles; int64 + 37 pointers; ...
    assert ctypes.sizeof(api.Flags) == 56
    assert ctypes.sizeof(api.Surface) == 8 + 8 * 39
    assert ctypes.sizeof(api.Species) == 8 + 8 * 4
    assert ctypes.sizeof(api.Grid) == 16 + 8 * 5
    assert ctypes.sizeof(api.Options) == 4 + 4 + 8 + 4 + 4 + 16
    assert ctypes.sizeof(api.Stats) == 3 * 8 + 6 * 8 + 4 * 4 + 8 + 8  # ... + n_gpus, allreduce_ms, n_chunks_wanted, reserved


def test_error_strings():
    lib = api.lib()
    for code in range(8):
        assert lib.is3d_b200_strerror(code)
    assert lib.is3d_b200_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (IS3D_ERR_NO_DEVICE), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.Is3dError) as e:
        api.init()
    assert e.value.code == 5


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under is3d_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "is3d_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "cf_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
