"""B200-native drop-in for the `humanoid` package of roboterax/humanoid-gym
(humanoid_ppo hot path only)."""
import os

LEGGED_GYM_ROOT_DIR = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
LEGGED_GYM_ENVS_DIR = os.path.join(LEGGED_GYM_ROOT_DIR, "humanoid", "envs")
