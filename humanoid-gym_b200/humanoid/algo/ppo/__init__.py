from .ppo import PPO  # noqa: F401
from .on_policy_runner import OnPolicyRunner  # noqa: F401
from .actor_critic import ActorCritic  # noqa: F401
from .rollout_storage import RolloutStorage  # noqa: F401
