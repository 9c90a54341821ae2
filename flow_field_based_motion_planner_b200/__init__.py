"""B200-native gym_ffmp hot path: batched FFMP-v0 environments as hand-written sm_100a CUDA kernels
behind a C-ABI (include/ffmp_b200.h).  See SPEC.md / DESIGN.md."""
from .gym_compat import FFMP, make, register  # noqa: F401
from .robot import NUM_ACTIONS, RobotAction, RobotPose, RobotState, RobotVelocity  # noqa: F401
from .vector_env import FFMPConfig, FFMPVectorEnv, make_spaces, p_threshold  # noqa: F401
from . import native, ops, spaces  # noqa: F401
from .replay import ReplayRing  # noqa: F401
from .qnet import QNetwork  # noqa: F401
from .learner import DDQNLearner, TorchNetwork  # noqa: F401

__all__ = ["FFMP", "FFMPConfig", "FFMPVectorEnv", "RobotAction", "RobotPose", "RobotState", "RobotVelocity",
           "NUM_ACTIONS", "make", "register", "make_spaces", "ops", "spaces", "native", "ReplayRing", "QNetwork", "DDQNLearner", "TorchNetwork"]
