"""two_pass_lanczos_b200 -- B200-native (sm_100a) two-pass Lanczos engine for f(A)b.

Python mirror of the reference crate's public surface (lukefleed/two-pass-lanczos, src/lib.rs:95-101):

    from two_pass_lanczos_b200 import lanczos, lanczos_two_pass          # solvers::*
    from two_pass_lanczos_b200 import algorithms, data_loader             # algorithms::*, utils::data_loader

All compute goes through the C ABI of libtplanczos.so (include/tplanczos.h); there is no CPU fallback.
"""
from . import algorithms, data_loader, datagen, error, operators, sharding, solvers  # noqa: F401
from .error import CudaError, DataLoaderError, LanczosError  # noqa: F401
from .operators import LinOp  # noqa: F401
from .solvers import lanczos, lanczos_sweep, lanczos_two_pass, lanczos_two_pass_sweep  # noqa: F401

__all__ = ["lanczos", "lanczos_two_pass", "algorithms", "solvers", "data_loader", "datagen", "operators", "LinOp",
           "LanczosError", "DataLoaderError", "CudaError"]
