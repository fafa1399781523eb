"""deepsee_b200: B200-native implementation of DeepSEE's data-parallel hot path.

Python host code (the reference's SRModel / BaseManager API) over a C-ABI CUDA library
(include/deepsee_b200.h) of hand-written sm_100a kernels. See DESIGN.md.
"""
__version__ = "0.1.0"
