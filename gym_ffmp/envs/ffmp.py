"""gym_ffmp.envs.ffmp — the reference's module path of the env class (/root/reference/src/gym_ffmp/envs/ffmp.py:22)."""
from flow_field_based_motion_planner_b200.gym_compat import FFMP  # noqa: F401
from flow_field_based_motion_planner_b200.vector_env import (GOAL_THRESHOLD, MAP_GRID_NUM, MAP_RANGE,  # noqa: F401
                                                             MAP_RESOLUTION, ROBOT_RSIZE)
