"""The C-ABI library loads and exports every symbol include/smem_gpu.h declares (no compute here)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, pkg


def test_exports_match_header():
    sg = pkg("smem_gpu")
    sg.build()
    lib = sg.load_library()
    hdr = open(os.path.join(ROOT, "include", "smem_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(smem_gpu_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in smem_gpu.h but not exported"
    assert sorted(sg.EXPORTS) == declared


def test_struct_layouts_match_reference_types():
    sg = pkg("smem_gpu")
    # bwtintv_t is 4 x uint64 (bwt.h:60-62); smem_seed_opt_t is {int,double,int,int}
    assert C.sizeof(sg.SeedOpt) == 24
    assert C.sizeof(sg.IndexDesc) == 8 * 9
    assert C.sizeof(sg.Timing) == 48


def test_error_paths_without_gpu():
    import torch
    sg = pkg("smem_gpu")
    lib = sg.load_library()
    assert lib.smem_gpu_strerror(0) == b"ok"
    assert lib.smem_gpu_strerror(-5) == b"capacity exceeded"
    h = C.c_void_p()
    assert lib.smem_gpu_create(C.byref(h), 0, None, C.c_int64(10), 101) == -1      # bad n_devices
    if not torch.cuda.is_available():
        assert lib.smem_gpu_create(C.byref(h), 1, None, C.c_int64(10), 101) == -7  # no device, no fallback
        with pytest.raises(sg.SmemGpuError):
            sg.SmemGpu(10, 101)
