"""CPU-only checks of the host-side mirror of the reference's plugin interface."""
import numpy as np
import pytest

from blackbox_mpc_b200 import _lib, sharding
from blackbox_mpc_b200.dynamics_functions.deterministic_mlp import activation_id
from blackbox_mpc_b200.optimizers import BY_NAME
from blackbox_mpc_b200.spaces import Box
from blackbox_mpc_b200.utils import workloads


def test_optimizer_names_match_reference():
    """policies/mpc_policy.py:81-116 dispatches on exactly these strings."""
    assert sorted(BY_NAME) == sorted(["CEM", "CMA-ES", "PI2", "PSO", "SPSA", "RandomSearch"])


def test_reference_default_hyperparameters():
    a, o = Box(-np.ones(2, np.float32), np.ones(2, np.float32)), Box(-np.ones(3, np.float32), np.ones(3, np.float32))
    cem = BY_NAME["CEM"](a, o)        # cem.py:8-10
    assert (cem._planning_horizon, cem._max_iterations, cem._population_size, cem._num_elite, cem._num_agents,
            cem._epsilon, cem._alpha) == (50, 5, 500, 50, 5, 0.001, 0.25)
    pi2 = BY_NAME["PI2"](a, o)        # pi2.py:10-11
    assert (pi2._population_size, pi2._lamda) == (500, 1.0)
    rs = BY_NAME["RandomSearch"](a, o)  # random_search.py:8
    assert (rs._population_size, rs._max_iterations) == (1024, None)
    pso = BY_NAME["PSO"](a, o)        # pso.py:8-11
    assert (pso._c1, pso._c2, pso._w, pso._initial_velocity_fraction) == (0.3, 0.5, 0.2, 0.01)
    spsa = BY_NAME["SPSA"](a, o)      # spsa.py:8-12
    assert (spsa._alpha, spsa._gamma, spsa._a_par, spsa._noise_parameter) == (0.602, 0.101, 0.01, 0.3)
    cma = BY_NAME["CMA-ES"](a, o)     # cma_es.py:8-10
    assert (cma._num_elite, cma._h_sigma, cma._alpha_cov) == (50, 1.0, 2.0)
    # optimizer_base.py:46-50
    np.testing.assert_allclose(cem._exploration_variance, (np.square(a.low - a.high) / 16) * 0.05)
    np.testing.assert_allclose(cem._exploration_mean, (a.high + a.low) / 2)


def test_unimplemented_base_raises_like_reference():
    """optimizer_base.py:53,103 / evaluator_base.py:44,63,85 raise a plain Exception."""
    from blackbox_mpc_b200.optimizers.optimizer_base import OptimizerBase
    from blackbox_mpc_b200.trajectory_evaluators.evaluator_base import EvaluatorBase
    a, o = Box(-np.ones(2, np.float32), np.ones(2, np.float32)), Box(-np.ones(3, np.float32), np.ones(3, np.float32))
    base = OptimizerBase(None, 5, 1, 1, a, o)
    with pytest.raises(Exception):
        base._optimize(None, 0)
    with pytest.raises(Exception):
        base.reset()
    ev = EvaluatorBase(reward_function=None, system_dynamics_handler=None)
    for call in (lambda: ev(None, None, 0), lambda: ev.predict_next_state(None, None),
                 lambda: ev.evaluate_next_reward(None, None, None)):
        with pytest.raises(Exception):
            call()


def test_activation_names_and_callables():
    import torch
    assert activation_id(None) == _lib.ACT_NONE and activation_id("tanh") == _lib.ACT_TANH
    assert activation_id(torch.tanh) == _lib.ACT_TANH and activation_id(torch.relu) == _lib.ACT_RELU

    def tanh(x):  # stands for tf.math.tanh (tutorials/mujoco/tutorial_two.py:28-31): matched by __name__
        return x
    assert activation_id(tanh) == _lib.ACT_TANH
    with pytest.raises(ValueError):
        activation_id("gelu")


def test_reward_must_be_builtin():
    from blackbox_mpc_b200.trajectory_evaluators.deterministic import reward_id_of
    from blackbox_mpc_b200.utils import halfcheetah, pendulum
    assert reward_id_of(pendulum.pendulum_reward_function) == _lib.REWARD_PENDULUM
    assert reward_id_of(halfcheetah.reward_function) == _lib.REWARD_HALFCHEETAH
    with pytest.raises(TypeError):
        reward_id_of(lambda s, a, s2: 0)


def test_workload_flop_counts_match_survey():
    """F_step = 178 400 (C3, C5), 892 000 (C4), 9 088 (C2); per-iteration totals."""
    assert workloads.make("C2").flops_per_row_step() == 9088
    assert workloads.make("C3").flops_per_row_step() == 178400
    c4 = workloads.make("C4")
    assert c4.flops_per_row_step() == 892000
    assert c4.flops_per_iteration() == 267_600_000_000
    assert workloads.make("C5").flops_per_iteration() == 446_000_000_000
    assert workloads.make("C1").flops_per_row_step() == 0
    assert c4.state.shape == (1, 20) and len(c4.weights) == 5 and c4.weights[0][0].shape == (26, 200)
    # deterministic: same seed, same data on every rank
    np.testing.assert_array_equal(c4.weights[3][2], workloads.make("C4").weights[3][2])


@pytest.mark.parametrize("P,world", [(10000, 1), (10000, 8), (301, 3), (7, 8), (50000, 8)])
def test_shard_ranges_partition_the_population(P, world):
    edges = [sharding.shard_range(P, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == P
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(P, world, world)


def test_partial_message_sizes():
    assert sharding.partial_floats("CEM", 1, 30, 6, num_elite=50) == 50 * 182      # 36.4 KB / rank
    assert sharding.partial_floats("PI2", 1, 30, 6) == 182
    assert sharding.partial_floats("SPSA", 2, 30, 6) == 360
    assert sharding.partial_floats("CMA-ES", 1, 50, 6, num_elite=50) == 50 * 302
