"""scgpu -- B200-native Scan Context loop-closure hot path (SC-LeGO-LOAM's SCManager).

Layout:
  csrc/scgpu*.cu(h)   hand-written sm_100a CUDA kernels + the C ABI declared in include/scgpu.h
  csrc/scangen.c      deterministic synthetic scans (host-only data generator)
  build.py            in-tree nvcc/gcc build of libscgpu.so / libscangen.so
  scgpu.py            ctypes binding of the C ABI + a Python mirror of the reference's SCManager surface
  sharded.py          database sharding over ranks (torch.distributed) -- the only place a collective is used
  synth.py            wrapper of the synthetic scan generator
The C++ drop-in for mapOptmization.cpp is include/Scancontext.h.
"""
PACKAGE_DIR = __path__[0] if "__path__" in globals() else None
