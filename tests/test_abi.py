"""CPU-only checks of the C-ABI boundary: libvsm.so builds in-tree, loads without a GPU, exports exactly
what include/vsm.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vsm.h")
LIB = os.path.join(ROOT, "vggt-slam_b200", "csrc", "libvsm.so")


@pytest.fixture(scope="module")
def native():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.dirname(LIB), "-j", "8"], check=True)
    from vsm import _native

    return _native


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^VSM_API\s+[\w\s\*]+?\b(vsm_\w+)\s*\(", text, flags=re.M)))


def test_header_symbols_are_exported(native):
    want = declared_symbols()
    assert len(want) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], check=True, capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(vsm_\w+)", out)))
    assert exported == want
    assert sorted(native.SIGNATURES) == want  # the ctypes binding covers the whole header


def test_struct_layouts_match_header(native):
    # sizes the C compiler gives the structs in include/vsm.h
    src = '#include "vsm.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(vsm_config), sizeof(vsm_fuse_params), sizeof(vsm_fuse_stats));return 0;}'
    exe = "/tmp/vsm_sizes"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.dirname(HEADER), "-o", exe], input=src, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(native.Config), C.sizeof(native.FuseParams), C.sizeof(native.FuseStats)]


def test_no_cpu_fallback(native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert native.lib.vsm_abi_version() == 3
    cfg = native.Config(0.05, 16, native.F32, 1024, -1, 0)
    h = C.c_void_p()
    rc = native.lib.vsm_map_create(C.byref(cfg), C.byref(h))
    assert rc == native.E_CUDA and "no CPU fallback" in native.last_error()
    with pytest.raises(RuntimeError):
        native.check(rc)
    import vsm

    sm = vsm.Submap(0)
    import numpy as np

    with pytest.raises(RuntimeError):  # even the confidence threshold runs on the device
        sm.add_all_points(np.zeros((1, 2, 2, 3), np.float32), None, np.ones((1, 2, 2), np.float32), 25.0, None)
    with pytest.raises(RuntimeError):
        vsm.GraphMap().build_semantic_voxel_map(0.05)


def test_argument_validation_without_gpu(native):
    cfg = native.Config(0.0, 16, native.F32, 1024, -1, 0)
    h = C.c_void_p()
    assert native.lib.vsm_map_create(C.byref(cfg), C.byref(h)) == native.E_INVALID
    cfg = native.Config(0.05, 12, native.F32, 1024, -1, 0)
    assert native.lib.vsm_map_create(C.byref(cfg), C.byref(h)) == native.E_INVALID
    with pytest.raises(ValueError):
        native.check(native.E_INVALID)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "vggt-slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "voxel_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
