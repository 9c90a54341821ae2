"""Robot value types and the 28-way (v, w) action table.

Same names and meaning as /root/reference/src/gym_ffmp/envs/robot/config.py:6-58; the table is
generated from its rule (id = 7*iv + iw) and pinned against the reference by tests/golden."""
from dataclasses import dataclass

LINEAR_V = (0.0, 0.2, 0.4, 0.6)                        # m/s
ANGULAR_V = (-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6)     # rad/s
NUM_ACTIONS = len(LINEAR_V) * len(ANGULAR_V)


@dataclass
class RobotPose:
    x: float
    y: float
    yaw: float


@dataclass
class RobotVelocity:
    linear_v: float
    angular_v: float


class RobotState:
    def __init__(self, x, y, yaw, linear_v, angular_v):
        self.robot_position = RobotPose(x, y, yaw)
        self.robot_velocity = RobotVelocity(linear_v, angular_v)


class RobotAction:
    def __init__(self):
        self.cmd = [RobotVelocity(v, w) for v in LINEAR_V for w in ANGULAR_V]

    def commander(self, i):
        return self.cmd[i]
