"""lnaphylodyn_b200 -- B200-native LNA + coalescent likelihood / elliptical-slice hot path.

Drop-in for the likelihood path of MingweiWilliamTang/LNAphyloDyn: a C-ABI CUDA library
(include/lnaphylodyn_b200.h, csrc/) plus this host-side mirror of the reference's Rcpp exports.
"""
from . import api  # noqa: F401
from .api import Problem, problem_for  # noqa: F401

__version__ = "0.1.0"
