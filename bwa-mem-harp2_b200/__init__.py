"""B200-native SMEM seeding (BWA-MEM 0.7.8 hot path) -- host-side Python mirror.

The product is the C-ABI CUDA library (include/smem_gpu.h, csrc/); this package holds the
ctypes binding used by tests/bench plus workload tooling (synthetic data, FM-index build).
Import with ``importlib.import_module("bwa-mem-harp2_b200")`` (the name carries hyphens).
"""
