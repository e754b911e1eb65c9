"""simpleslam_b200 — B200-native (sm_100a) implementation of SimpleSLAM's PCR registration hot path.

Layout: csrc/ (CUDA kernels + the C ABI of include/pcr_cuda.h), capi.py (ctypes binding), registers.py (Python
mirror of the reference's PCR plugin interface), cpp/PCR (C++ adaptor with the reference's virtual interface),
synth/ (synthetic lidar world for tests and benchmarks).
"""
__all__ = ["capi", "registers", "synth", "build"]
