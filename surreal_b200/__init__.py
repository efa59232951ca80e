"""surreal_b200 -- B200-native actor -> HBM replay -> PPO / DDPG learner hot path behind
SurrealAI/surreal's Agent / Replay / Learner plugin surface."""
__version__ = '0.1.0'
