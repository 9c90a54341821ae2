"""`gym_ffmp` — drop-in import shim for the reference package of the same name.

The reference trainer does (/root/reference/src/train.py:35-37,456, /root/reference/src/gym_ffmp/__init__.py:1-6)

    import gym_ffmp
    from gym_ffmp.envs.robot.config import RobotPose, RobotVelocity, RobotState, RobotAction
    from gym_ffmp.envs.ffmp import FFMP
    env = gym.make('FFMP-v0')

With this directory on sys.path those lines resolve to the B200 implementation: `FFMP` is the single-env object of
flow_field_based_motion_planner_b200.gym_compat (same spaces, is_collision / is_goal / reward_calculator / rewarder /
rewarder2, plus the reset() / step() the reference left to ROS), evaluated by the CUDA library — there is no CPU fallback.
'FFMP-v0' is registered with `gym` (or `gymnasium`) when one of them is importable, and always with the package's own
registry (`gym_ffmp.make`), which is what runs on hosts without gym (this image has neither).
"""
from flow_field_based_motion_planner_b200.gym_compat import make, register as _register  # noqa: F401
from flow_field_based_motion_planner_b200.vector_env import FFMPVectorEnv  # noqa: F401

ENV_ID = "FFMP-v0"
REGISTERED_WITH = []


def _register_with(modname):
    try:
        registration = __import__(modname + ".envs.registration", fromlist=["register"])
    except ImportError:
        return
    try:
        registration.register(id=ENV_ID, entry_point="gym_ffmp.envs:FFMP")
        REGISTERED_WITH.append(modname)
    except Exception:           # already registered (re-import) or an incompatible registry: the own registry still works
        pass


for _m in ("gym", "gymnasium"):
    _register_with(_m)
