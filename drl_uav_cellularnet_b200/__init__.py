"""B200-native batched MobiEnvironment step path.

Public API
    MobiEnvironment          single-env drop-in for the reference class (mobile_env.py:35-233)
    BatchedMobiEnvironment   E environments per call, torch CUDA tensors in / out
"""
from .env import BatchedMobiEnvironment, MobiEnvironment, StepInfo, shard_range  # noqa: F401

__all__ = ["BatchedMobiEnvironment", "MobiEnvironment", "StepInfo", "shard_range"]
