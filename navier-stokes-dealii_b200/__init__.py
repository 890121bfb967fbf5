"""B200-native hot path of giuseppeegentile/Navier-Stokes-dealii: assemble_system + solve_system.

The directory name follows the reference repo (it contains '-'), so import it with
    importlib.import_module("navier-stokes-dealii_b200")
Layout: csrc/ (CUDA kernels + the two C-ABI libraries), host/ (C++ NavierStokesSolver shim that
keeps src/main.cpp compiling unchanged), and this Python mirror used by tests/ and bench.py.
"""
from .device import (PRECOND_BLOCK_DIAGONAL, PRECOND_BLOCK_TRIANGULAR, PRECOND_IDENTITY,  # noqa: F401
                     DeviceProblem)
from .solver import NavierStokesSolver, Parameters  # noqa: F401
from .topology import Dofs, Mesh, Part  # noqa: F401
