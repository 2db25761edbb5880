"""B200-native photon random walk behind monte_carloMPI's Python driver surface.

Python host code (this package) -> ctypes -> libmc3d.so (hand-written sm_100a CUDA + NCCL).  No PyTorch, no
Triton, no CPU fallback.  The drop-in import path ``from monte_carloMPI import monte_carlo3D`` is provided by the
top-level ``monte_carloMPI`` package, which re-exports this one.
"""
__version__ = '0.1.0'
