"""aligngraph2_b200 -- B200-native (sm_100a) implementation of AlignGraph2's compute hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of include/ag2_b200.h),
``lib`` (ctypes binding), ``mecat2ref`` (host-side mirror of the reference's extension interface),
``synth`` (the synthetic CLR generator of SURVEY.md 8d).  There is no CPU fallback.
"""
__version__ = "0.1.0"
