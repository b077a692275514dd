class MultiAgentEnv:
    """Empty base class, as far as the reference uses it (core.py:73,87)."""
