"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/aither_gpu.h declares; with no device, compute entry points fail loudly (no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

import aither_b200
from aither_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aither_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aither_gpu_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(aither_b200.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = aither_b200.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.aither_gpu_version()


def test_struct_sizes_match_header():
    """ctypes mirrors vs the C compiler's view of include/aither_gpu.h."""
    import subprocess
    import tempfile
    from aither_b200 import ctypes_abi as abi
    src = ('#include <stdio.h>\n#include "aither_gpu.h"\nint main(void){printf("%zu %zu %zu %zu %zu '
           '%zu\\n", sizeof(aither_cfg), sizeof(aither_bc_state), sizeof(aither_block_desc), '
           'sizeof(aither_conn), sizeof(aither_surface), sizeof(aither_linf));return 0;}\n')
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o",
                               os.path.join(tmp, "t"), os.path.join(tmp, "t.c")])
        out = subprocess.check_output([os.path.join(tmp, "t")]).split()
    mine = [C.sizeof(t) for t in (abi.Cfg, abi.BCState, abi.BlockDesc, abi.Conn, abi.Surface,
                                  abi.Linf)]
    assert [int(v) for v in out] == mine


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    prob = synthetic.box_problem(8, 6, 4)
    with pytest.raises(aither_b200.AitherGpuError):
        aither_b200.GridLevel(prob)


@pytest.mark.parametrize("field, value, needle", [
    ("numBCStates", 33, "numBCStates"),       # bcStates[] holds AITHER_MAX_BC_STATES = 32 entries
    ("numBCStates", -1, "numBCStates"),
    ("numGhosts", 1, "stencil"),              # MUSCL needs two ghost layers (ref src/input.cpp:1127-1144)
    ("numSpecies", 4, "numSpecies"),
])
def test_create_rejects_unsupported_configurations(field, value, needle):
    """aither_gpu_create validates the configuration before it touches a device, so this runs on a
    machine without a GPU: nothing unsupported is silently approximated or read out of bounds."""
    prob = synthetic.box_problem(8, 6, 4)
    setattr(prob.cfg, field, value)
    with pytest.raises(aither_b200.AitherGpuError, match=needle):
        aither_b200.GridLevel(prob)


def test_entry_points_refuse_a_null_handle():
    """Every entry point that takes a handle reports an error for a null one (and says so in
    aither_gpu_last_error) instead of touching a device: runs on a machine without a GPU."""
    lib = aither_b200.load_library()
    buf = (C.c_double * 16)()
    raw = C.create_string_buffer(64 * 2)
    calls = [
        lambda: lib.aither_gpu_compute_wall_distance(None, buf, 1),
        lambda: lib.aither_gpu_download_output(None, 0, 0, 0, 1.0, buf),
        lambda: lib.aither_gpu_download_wall_data(None, 0, 0, buf),
        lambda: lib.aither_gpu_upload_interior_async(None, 0, buf),
        lambda: lib.aither_gpu_upload_state_async(None, 0, buf),
        lambda: lib.aither_gpu_halo_p2p_export(None, raw),
        lambda: lib.aither_gpu_halo_p2p_import(None, raw),
        lambda: lib.aither_gpu_iterate(None, 1.0, 0, buf, None, None),
    ]
    for call in calls:
        assert call() != 0
        assert b"null" in lib.aither_gpu_last_error()
