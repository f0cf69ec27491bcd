"""Batched LIPM / ZMP-preview MPC for bipedal walking (B200-native engine).

Host-side mirror of the reference package ``src/mpc_bipedal``: same class and
function names, same JSON schema; the per-tick QP, the state update and the
whole receding-horizon rollout run in the CUDA library behind
``mpc_bipedal.engine`` (C ABI in ``include/lipm_mpc.h``).
"""

__version__ = "0.1.0"

from .config import MPCConfig  # noqa: F401
