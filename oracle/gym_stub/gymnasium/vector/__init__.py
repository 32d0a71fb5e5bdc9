from gymnasium.vector.vector_env import AutoresetMode, SyncVectorEnv, VectorEnv  # noqa: F401
