"""The C-ABI shared library loads and exports every symbol include/lisec_b200.h declares; the ctypes binding names
exactly that set; without a GPU the library refuses to work instead of falling back. No compute calls here."""
import ctypes as C
import os
import re

import pytest

from lisec_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    with open(os.path.join(ROOT, "include", "lisec_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return set(re.findall(r"\b(lisec_[a-z0-9_]+)\s*\(", src))


def test_library_is_built():
    assert os.path.exists(N.LIB_PATH), "run `python -m lisec_b200.build` (or __graft_entry__.build())"


def test_every_declared_symbol_is_exported_and_bound():
    lib = N.load()
    declared = header_symbols()
    assert declared == set(N.SIGNATURES), "ctypes binding and header disagree: %s" % (declared ^ set(N.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lisec_abi_version() == N.ABI_VERSION


def test_config_struct_layout_matches_header():
    # 3 doubles, 9 int32, (pad) int64, 2 int32 -> 80 bytes with natural alignment
    assert C.sizeof(N.lisec_config) == 80
    assert N.lisec_config.max_points.offset == 64
    assert N.lisec_config.fcn_post_dense.offset == 76
    assert C.sizeof(N.lisec_vfe_weights) == 15 * 8 + 8 + 3 * 8
    assert N.lisec_vfe_weights.post_dense_kernel.offset == 128


def test_bad_config_is_rejected_with_a_message():
    lib = N.load()
    cfg = N.lisec_config(voxel_x=0.5, voxel_y=0.25, voxel_z=0.25, sample_size=35, max_voxel_x=100, max_voxel_y=200,
                         max_voxel_z=8, c1=16, c2=48, c3=96, grid_dtype=0, max_sweeps=1, max_points=1000, device=0)
    h = C.c_void_p()
    st = lib.lisec_create(C.byref(cfg), C.byref(h))  # widths of neither graph the reference has had
    assert st == -6 and b"16,32,64" in lib.lisec_last_error(h) and b"16,64,128" in lib.lisec_last_error(h)
    lib.lisec_destroy(h)
    cfg.c2, cfg.c3, cfg.fcn_post_dense = 64, 128, 2
    st = lib.lisec_create(C.byref(cfg), C.byref(h))
    assert st == -2 and b"fcn_post_dense" in lib.lisec_last_error(h)
    lib.lisec_destroy(h)
    cfg.c2, cfg.c3, cfg.fcn_post_dense, cfg.sample_size = 32, 64, 0, 1
    st = lib.lisec_create(C.byref(cfg), C.byref(h))
    assert st == -2
    lib.lisec_destroy(h)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = N.load()
    cfg = N.lisec_config(voxel_x=0.5, voxel_y=0.25, voxel_z=0.25, sample_size=35, max_voxel_x=100, max_voxel_y=200,
                         max_voxel_z=8, c1=16, c2=32, c3=64, grid_dtype=0, max_sweeps=1, max_points=1000, device=0)
    h = C.c_void_p()
    assert lib.lisec_create(C.byref(cfg), C.byref(h)) == -4  # LISEC_ERR_CUDA
    assert lib.lisec_last_error(h)
    lib.lisec_destroy(h)
    from lisec_b200 import Frontend

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Frontend()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lisec_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
