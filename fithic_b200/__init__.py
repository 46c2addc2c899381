"""fithic_b200 -- B200-native implementation of Fit-Hi-C's per-pair significance path.

Python host (same CLI and output format as ay-lab/fithic) over hand-written sm_100a CUDA kernels in
libfithic_b200.so, bound through ctypes (include/fithic_b200.h).  No CPU fallback: importing the engine without the
built library, or running it without a CUDA device, raises.
"""
__version__ = "0.1.0"
