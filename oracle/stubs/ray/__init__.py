"""Stand-in for `ray` — TEST INFRASTRUCTURE ONLY (ray is not installed in this image).
Only `ray.rllib.env.multi_agent_env.MultiAgentEnv` is needed to import the reference
(`/root/reference/warehouse/core.py:6`)."""
