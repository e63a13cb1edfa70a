"""B200-native signature-k-mer calling path of olsonanl/close_kmers.

``close_kmers_b200.api.KmerGuts`` mirrors the reference's KmerGuts interface over libckm.so (hand-written
sm_100a CUDA kernels behind the C ABI in include/ckm.h).  Importing the package does not load CUDA.
"""
__all__ = ["api", "synth", "build"]
