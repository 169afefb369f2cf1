"""nbots_b200 -- B200 (sm_100a) implementation of the NBOTS Jacobi-PCG / FEM-assembly hot path.

The product is the C-ABI shared library ``nbots_b200/lib/libnbgpu.so`` (sources
in ``nbots_b200/csrc``, interface in ``include/nbgpu.h``) plus the
reference-named C shims ``libnbots_b200.so``.  The Python modules here are
ctypes plumbing for tests and bench.py and the synthetic-input generators.
"""
__all__ = ["capi", "api", "meshgen"]
