"""gym_ffmp.envs.robot.config — value types and the 28-entry action table (/root/reference/src/gym_ffmp/envs/robot/config.py:6-58)."""
from flow_field_based_motion_planner_b200.robot import RobotAction, RobotPose, RobotState, RobotVelocity  # noqa: F401
