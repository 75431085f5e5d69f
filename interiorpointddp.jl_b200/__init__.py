"""B200-native batched IPDDP2: a drop-in for the hot path of mingu6/InteriorPointDDP.jl.

    from ipddp_b200 import Dynamics, Objective, Constraint, Bound, Options, Solver, solve, get_trajectory

mirrors the reference's exports (reference src/InteriorPointDDP.jl:29-45); `BatchSolver` is the thin
batched wrapper over the C ABI (include/ipddp_b200.h, libipddp_b200.so).  Import is lazy so that the
package can be inspected (and the CUDA library built) on a machine without a GPU.
"""
__all__ = ["Dynamics", "Objective", "Constraint", "Bound", "Options", "Solver", "solve", "get_trajectory",
           "BatchSolver"]


def __getattr__(name):
    if name in ("Dynamics", "Objective", "Constraint", "Bound", "Options", "Solver", "solve", "get_trajectory"):
        from . import api
        return getattr(api, name)
    if name == "BatchSolver":
        from .batch import BatchSolver
        return BatchSolver
    raise AttributeError(name)
