"""gym_po — B200-native build of gym-po-taxi's internally vectorized environment step.

Same import surface as the reference package (``from gym_po.envs import TaxiVecEnv, RoomsEnv, ...``,
reference gym_po/envs/__init__.py:1-4); the step itself runs in hand-written sm_100a CUDA kernels
behind the C ABI in ``include/gpt_b200.h``.
"""
from .envs import *  # noqa: F401,F403
