"""cornell-moe_b200: B200-native GP-posterior + Monte-Carlo acquisition path of Cornell-MOE.

Layout:
  csrc/                     hand-written sm_100a CUDA kernels + the C ABI (include/cmoe_b200.h) + pybind11 `GPP` module
  libcornell_moe_b200.so    built in-tree by ``make -C cornell-moe_b200`` / ``__graft_entry__.build()``
  capi.py                   ctypes binding of the C ABI (what the parity tests and bench.py call)
  multigpu.py               one-process-per-GPU sharding of the multistart candidates (torch.distributed)

There is no CPU compute path: importing works anywhere, but every compute call needs the built .so and a CUDA device.
"""
__version__ = "0.1.0"
