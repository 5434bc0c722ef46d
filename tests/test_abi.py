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


def test_mirror_structs_match_reference_headers(tmp_path):
    """include/bwa_abi.h against the reference's own headers (build container only)."""
    import subprocess
    ref = "/root/reference/software"
    if not os.path.isdir(ref):
        pytest.skip("reference sources absent")
    src = tmp_path / "abi.c"
    src.write_text(f"""
#include <stddef.h>
#include "{ref}/bwt.h"
#include "{ref}/bwamem.h"
#include "{ROOT}/include/bwa_abi.h"
#include "{ROOT}/include/smem_gpu.h"
#define SAME(T, U, f) _Static_assert(offsetof(T, f) == offsetof(U, f), #f)
_Static_assert(sizeof(bwtintv_t) == sizeof(harp_bwtintv_t) && sizeof(bwtintv_t) == sizeof(smem_intv_t), "bwtintv_t");
_Static_assert(sizeof(bwtintv_v) == sizeof(harp_bwtintv_v), "bwtintv_v");
_Static_assert(sizeof(bwt_t) == sizeof(harp_bwt_t), "bwt_t");
_Static_assert(sizeof(smem_i) == sizeof(harp_smem_i), "smem_i");
SAME(bwt_t, harp_bwt_t, primary); SAME(bwt_t, harp_bwt_t, L2); SAME(bwt_t, harp_bwt_t, seq_len); SAME(bwt_t, harp_bwt_t, bwt_size);
SAME(bwt_t, harp_bwt_t, bwt); SAME(bwt_t, harp_bwt_t, sa);
SAME(smem_i, harp_smem_i, bwt); SAME(smem_i, harp_smem_i, query); SAME(smem_i, harp_smem_i, start); SAME(smem_i, harp_smem_i, len);
SAME(smem_i, harp_smem_i, matches); SAME(smem_i, harp_smem_i, sub); SAME(smem_i, harp_smem_i, tmpvec);
_Static_assert(BWT_BATCHED_INIT == HARP_BWT_BATCHED_INIT && BWT_BATCHED_FREE == HARP_BWT_BATCHED_FREE && BWT_BATCHED_DO == HARP_BWT_BATCHED_DO, "status");
int main(void) {{ return 0; }}
""")
    subprocess.run(["gcc", "-w", "-c", str(src), "-o", str(tmp_path / "abi.o")], check=True)
