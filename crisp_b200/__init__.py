"""crisp_b200 — B200-native (sm_100a) statistics engine for CRISP candidate sites.

Host-side mirror of the C ABI in include/crispgpu.h.  The compute path is the CUDA library
crisp_b200/csrc/libcrispgpu.so; there is no CPU fallback — importing `crisp_b200.engine.Engine`
without the built library raises.
"""
from .abi import Batch, EngineConfig, Results  # noqa: F401

__version__ = "0.1.0"
