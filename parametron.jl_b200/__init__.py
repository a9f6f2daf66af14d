"""parametron-b200: B200-native `update!` + coefficient assembly for Parametron.jl models (C ABI + sm_100a kernels).

The directory name contains a dot, so import it through the root-level shim:  `import parametron_b200 as pm`.
"""
from . import _capi, macros, shard
from ._capi import (AlreadyInitialized, DimensionMismatch, PmbError, Unsupported, LIB_PATH)
from .model import (Expr, Maximize, Minimize, Model, Parameter, RecordParameter, Variable, adjoint, bilinear, constraint, convert, dot,
                    objective, transpose, vcat, vect)

__all__ = ["Model", "Parameter", "Variable", "Expr", "Minimize", "Maximize", "adjoint", "transpose", "dot", "bilinear", "vcat",
           "vect", "convert", "RecordParameter", "objective", "constraint", "shard", "macros", "PmbError", "DimensionMismatch", "AlreadyInitialized", "Unsupported", "LIB_PATH"]
