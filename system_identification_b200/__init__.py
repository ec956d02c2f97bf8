"""system_identification_b200 -- B200-native hot path of xiaohu97/system_identification.

regressor -> contact null-space projection -> Gram / normal equations -> LMI-constrained fit,
as hand-written sm_100a fp64 kernels behind a C-ABI (include/sysid_b200.h), with a Python host
layer that mirrors the reference's two classes (SystemIdentification, Solver).
"""
from .model import FlatModel  # noqa: F401

__all__ = ["FlatModel"]
