from .vec_env import VecEnv  # noqa: F401
from .ppo import *  # noqa: F401,F403
