"""CPU tests of the drop-in boundary: libb2gpu.so loads, exports every symbol include/b2gpu.h declares,
its record layouts match the ctypes/numpy mirror, and without a GPU it fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2gpu_[a-z0-9_]+)\s*\(", text)))


def abi_version_of_header():
    return int(re.search(r"#define B2GPU_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "b2gpu.h")).read()).group(1))


def test_library_exports_every_declared_symbol(built):
    from box2d_rs_b200 import lib
    L = C.CDLL(lib.DEFAULT_SO)
    names = _declared_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(L, n)]
    assert missing == []
    bound = lib.load()
    assert bound._b2gpu_missing == []
    assert sorted(set(names) - set(bound._b2gpu_declared)) == [], "lib.py does not bind every header symbol"
    assert bound.b2gpu_abi_version() == abi_version_of_header() == 2


def test_record_sizes_match_header(built):
    """sizeof of every record as the C compiler sees it == the numpy/ctypes mirror (abi.py)."""
    import subprocess
    import tempfile
    from box2d_rs_b200 import abi
    src = r'''
#include <stdio.h>
#include "b2gpu.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(b2gpu_joint_rec), sizeof(b2gpu_joint_def),
  sizeof(b2gpu_snapshot), sizeof(b2gpu_body_rec), sizeof(b2gpu_fixture_rec),
  sizeof(b2gpu_shape_rec), sizeof(b2gpu_proxy_rec), sizeof(b2gpu_tree_node_rec), sizeof(b2gpu_manifold), sizeof(b2gpu_contact_rec),
  sizeof(b2gpu_step_stats), sizeof(b2gpu_world_rec), sizeof(b2gpu_snapshot_sizes), sizeof(b2gpu_body_def), sizeof(b2gpu_fixture_def),
  sizeof(b2gpu_shape_def)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    mirror = [abi.JOINT_DTYPE.itemsize, C.sizeof(abi.JointDef), C.sizeof(abi.SnapshotC), abi.BODY_DTYPE.itemsize, abi.FIXTURE_DTYPE.itemsize, abi.SHAPE_DTYPE.itemsize, abi.PROXY_DTYPE.itemsize,
              abi.NODE_DTYPE.itemsize, abi.MANIFOLD_DTYPE.itemsize, abi.CONTACT_DTYPE.itemsize, abi.STATS_DTYPE.itemsize,
              C.sizeof(abi.WorldRec), C.sizeof(abi.SnapshotSizes), C.sizeof(abi.BodyDef), C.sizeof(abi.FixtureDef),
              C.sizeof(abi.ShapeDef)]
    assert sizes == mirror


def test_no_cpu_fallback_without_device(built):
    """On a machine without a CUDA device every stepping entry point reports B2GPU_E_NO_DEVICE."""
    from box2d_rs_b200 import abi, lib
    L = lib.load()
    n = L.b2gpu_device_count()
    if n > 0:
        pytest.skip("a CUDA device is present")
    assert n == abi.E_NO_DEVICE or n == 0
    h = C.c_void_p()
    assert L.b2gpu_init(0, None, C.byref(h)) == abi.E_NO_DEVICE
    assert not h
    with pytest.raises(lib.B2gpuError) as e:
        from box2d_rs_b200 import batch
        batch.Context(0)
    assert e.value.code == abi.E_NO_DEVICE


def test_argument_errors_are_codes_not_crashes(built):
    from box2d_rs_b200 import abi, lib
    L = lib.load()
    assert L.b2gpu_init(0, None, None) == abi.E_INVALID
    assert b"out is NULL" in L.b2gpu_last_error()
    assert L.b2gpu_batch_step(None, 0.016, 8, 3, 1) == abi.E_INVALID
    assert L.b2gpu_world_step(None, 0.016, 8, 3) == abi.E_INVALID
    s = abi.ShapeDef()
    tri = (C.c_float * 4)(0.0, 0.0, 1.0, 0.0)
    assert L.b2gpu_polygon_set(C.byref(s), tri, 2) == abi.E_INVALID  # the reference asserts 3 <= count <= 8
    assert L.b2gpu_world_set_continuous_physics(None, 1) == abi.E_INVALID


def test_rust_ffi_lists_every_symbol():
    """rust/ffi.rs (the extern "C" block a box2d-rs maintainer adds; not buildable here: no Rust toolchain)
    declares exactly the functions of include/b2gpu.h."""
    text = open(os.path.join(ROOT, "rust", "ffi.rs")).read()
    block = text[text.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    rust = sorted(set(re.findall(r"pub fn (b2gpu_[a-z0-9_]+)\s*\(", block)))
    assert rust == _declared_symbols()
