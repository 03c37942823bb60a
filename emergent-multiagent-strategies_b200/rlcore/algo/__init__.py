from .ppo import JointPPO, PPO, magent_feed_forward_generator  # noqa: F401
