"""Kinematic facts of the XBot-L asset that the hot path needs, as Isaac Gym reports them for
resources/robots/XBot/urdf/XBot-L.urdf loaded with collapse_fixed_joints=True
(reference legged_robot.py:588-681): 12 revolute DoF in URDF joint order and the 13 rigid bodies
that survive the fixed-joint collapse.  (lower, upper, velocity, effort) are the URDF <limit> values."""

_LEG = (
    # joint suffix         lower   upper   vel   effort   (left leg; the right leg mirrors these signs)
    ("leg_roll_joint",    -0.44,   1.57,  12.0, 100.0),
    ("leg_yaw_joint",     -1.05,   1.05,  12.0, 100.0),
    ("leg_pitch_joint",   -1.57,   1.31,  12.0, 250.0),
    ("knee_joint",        -1.05,   1.10,  12.0, 250.0),
    ("ankle_pitch_joint", -0.70,   0.87,  12.0, 100.0),
    ("ankle_roll_joint",  -0.44,   0.44,  12.0, 100.0),
)

DOF_NAMES, DOF_LOWER, DOF_UPPER, DOF_VELOCITY, DOF_EFFORT = [], [], [], [], []
for _side in ("left_", "right_"):
    for _name, _lo, _hi, _vel, _eff in _LEG:
        if _side == "right_":
            _lo, _hi = -_hi, -_lo
        DOF_NAMES.append(_side + _name)
        DOF_LOWER.append(_lo)
        DOF_UPPER.append(_hi)
        DOF_VELOCITY.append(_vel)
        DOF_EFFORT.append(_eff)

BODY_NAMES = ["base_link"] + [
    side + link for side in ("left_", "right_")
    for link in ("leg_roll_link", "leg_yaw_link", "leg_pitch_link", "knee_link", "ankle_pitch_link",
                 "ankle_roll_link")]
BASE_LINK_MASS = 5.0      # nominal; only body_mass/30 enters the privileged observation
NUM_DOF = len(DOF_NAMES)
NUM_BODIES = len(BODY_NAMES)
