"""Live differential test of the oracles against the UNMODIFIED reference, on fresh seeds.
Only runs where /root/reference exists (the build container); the GPU box relies on the committed
fixtures instead."""
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("WH_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "warehouse")), reason="reference not mounted")

import golden_util as gu  # noqa: E402


@pytest.fixture(scope="module")
def mg():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    import make_golden
    return make_golden


@pytest.mark.parametrize("size,A", [("small", 3), ("medium", 9), ("large", 11)])
def test_fresh_seeds(mg, size, A):
    from oracle import ref_port as rp
    from oracle import wh_oracle as wo
    for seed in (424242, 7):
        for policy in ("greedy", "random"):
            ep = mg.run_episode(size, A, seed, policy, T=205)
            d = {"x_" + k: v for k, v in ep.items()}
            gu.check_episode(lambda kw, n, a: wo.OracleEnv(wo.make_config(**kw), n, num_agents=a), d, "x_")
            if seed == 7:
                gu.check_episode(lambda kw, n, a: rp.PortEnv(kw, n, a), d, "x_")


def test_fresh_single_steps(mg):
    from oracle import wh_oracle as wo
    d = mg.run_single_steps("large", 300, 987)
    gu.check_single_steps(lambda kw, n, a: wo.OracleEnv(wo.make_config(**kw), n, num_agents=a), d)


def test_native_rng_return_statistics():
    """On-device-RNG mode is distribution-equivalent, not stream-equivalent: the oracle's native
    Philox mode must reproduce the reference's episode-return statistics (SURVEY.md §6 [probe]:
    greedy p=0 total return 84.3+-48.7 / 102.0+-57.3 / 127.3+-47.9 for S/M/L over 200 episodes;
    uniform-random policy 5.72 / 11.22 / 18.64)."""
    from oracle import wh_oracle as wo
    ref_greedy = {"small": (84.3, 48.7), "medium": (102.0, 57.3), "large": (127.3, 47.9)}
    ref_random = {"small": 5.72, "medium": 11.22, "large": 18.64}
    n = 4000
    for size in gu.SIZES:
        env = wo.OracleEnv(wo.variant_config(size), n, seed=99)
        env.reset()
        env.rollout(200, 8, policy="greedy", auto_reset=False)
        ret = (env.state["acc"][:, 0] + env.state["acc"][:, 1]).astype(np.float64)
        mean, std = ref_greedy[size]
        # 200-episode reference sample: standard error of its mean = std/sqrt(200)
        assert abs(ret.mean() - mean) < 4 * std / np.sqrt(200), (size, ret.mean(), mean)
        assert abs(ret.std() - std) < 0.2 * std, (size, ret.std(), std)
        env = wo.OracleEnv(wo.variant_config(size), n, seed=100)
        env.reset()
        rng = np.random.Generator(np.random.PCG64(1))
        for _ in range(200):
            env.step(rng.integers(0, 9, size=(n, env.R)), with_obs=False)
        ret = (env.state["acc"][:, 0] + env.state["acc"][:, 1]).astype(np.float64)
        assert abs(ret.mean() - ref_random[size]) < 0.5, (size, ret.mean())
