"""GPU test of the handler's training half: learn the pendulum from trajectories, check that the
next act() uses the re-staged weights / statistics, and that `saved_model_<k>/` round-trips."""
import os

import numpy as np
import pytest
import torch

from blackbox_mpc_b200.dynamics_functions.deterministic_mlp import DeterministicMLP
from blackbox_mpc_b200.dynamics_handlers.system_dynamics_handler import SystemDynamicsHandler
from blackbox_mpc_b200.policies.mpc_policy import MPCPolicy
from blackbox_mpc_b200.spaces import Box
from blackbox_mpc_b200.utils import pendulum

pytestmark = pytest.mark.gpu


def _pendulum_step(s, u):
    """gym Pendulum-v0 (the closed form utils/pendulum.py:78-92 restates), numpy float64."""
    th, thdot = np.arctan2(s[..., 1], s[..., 0]), s[..., 2]
    thdot = thdot + (-3 * 10.0 / 2 * np.sin(th + np.pi) + 3.0 * u[..., 0]) * 0.05
    th = th + thdot * 0.05
    thdot = np.clip(thdot, -8, 8)
    return np.stack([np.cos(th), np.sin(th), thdot], -1)


def _episodes(n_ep, T, n_agents, rng):
    obs_l, act_l = [], []
    for _ in range(n_ep):
        th, om = rng.uniform(-np.pi, np.pi, n_agents), rng.uniform(-1, 1, n_agents)
        obs = np.zeros((T + 1, n_agents, 3)); obs[0] = np.stack([np.cos(th), np.sin(th), om], -1)
        acts = rng.uniform(-2, 2, (T, n_agents, 1))
        for t in range(T):
            obs[t + 1] = _pendulum_step(obs[t], acts[t])
        obs_l.append(obs.astype(np.float32)); act_l.append(acts.astype(np.float32))
    return obs_l, act_l


def test_train_then_act_and_reload(cuda_device, tmp_path):
    rng = np.random.default_rng(0)
    act_space = Box(np.array([-2.0], np.float32), np.array([2.0], np.float32))
    obs_space = Box(-np.ones(3, np.float32) * 8, np.ones(3, np.float32) * 8)
    mlp = DeterministicMLP([4, 64, 64, 3], ["tanh", "tanh", None], seed=1)
    handler = SystemDynamicsHandler(act_space, obs_space, dynamics_function=mlp, is_normalized=True, log_dir=str(tmp_path), seed=0)
    policy = MPCPolicy(reward_function=pendulum.pendulum_reward_function, env_action_space=act_space,
                       env_observation_space=obs_space, dynamics_handler=handler, optimizer_name="CEM", num_agents=1,
                       planning_horizon=10, population_size=256, max_iterations=3, num_elite=16)
    obs, acts = _episodes(12, 100, 4, rng)
    handler.train(obs, acts, None, epochs=25, learning_rate=2e-3, batch_size=128)
    tr, va = handler.last_training_loss, handler.last_validation_loss
    assert tr[-1] < 0.05 * tr[0] and np.isfinite(va).all() and va[-1] < 0.1
    # the trained model drives act(): its one-step prediction must be close to the true pendulum step
    s0 = obs[0][5, 0]
    action, next_obs, _ = policy.act(s0, 0)
    true_next = _pendulum_step(s0.astype(np.float64), action.astype(np.float64))
    np.testing.assert_allclose(next_obs, true_next, atol=0.08)
    # a second round of training changes the weights and the next act() sees them (version bump -> re-stage)
    w_before = mlp.weights[0].clone()
    handler.train(*_episodes(2, 100, 4, rng), None, epochs=2)
    assert not torch.equal(mlp.weights[0], w_before)
    a2, n2, _ = policy.act(s0, 1)
    assert np.isfinite(n2).all()
    # on-disk contract and reload through saved_model_dir
    d = os.path.join(str(tmp_path), "saved_model_2")
    assert os.path.exists(os.path.join(d, "mean_states.npy")) and os.path.exists(os.path.join(d, "weights.npz"))
    mlp2 = DeterministicMLP([4, 64, 64, 3], ["tanh", "tanh", None], seed=7)
    handler2 = SystemDynamicsHandler(act_space, obs_space, dynamics_function=mlp2, is_normalized=True, saved_model_dir=d, seed=0)
    assert torch.equal(mlp2.weights[1].cpu(), mlp.weights[1].cpu())
    x = torch.from_numpy(np.concatenate([s0, action])[None]).float()
    ev1, ev2 = policy._trajectory_evaluator, None
    from blackbox_mpc_b200.trajectory_evaluators.deterministic import DeterministicTrajectoryEvaluator
    ev2 = DeterministicTrajectoryEvaluator(reward_function=pendulum.pendulum_reward_function, system_dynamics_handler=handler2)
    st, ac = torch.from_numpy(s0[None]), torch.from_numpy(action[None])
    np.testing.assert_array_equal(ev1.predict_next_state(st, ac).cpu().numpy(), ev2.predict_next_state(st, ac).cpu().numpy())
    # the loaded statistics survive a later train() (reference :80-83: _first_time = False after loading) ...
    stats_loaded = [t.clone() for t in handler2._stats]
    handler2.train(*_episodes(2, 50, 4, rng), None, epochs=1)
    assert all(torch.equal(a, b) for a, b in zip(stats_loaded, handler2._stats))
    # ... and a directory with trained weights is never silently dropped
    with pytest.raises(ValueError):
        SystemDynamicsHandler(act_space, obs_space, dynamics_function=None, is_normalized=True, saved_model_dir=d, seed=0)


def test_iterative_mpc_driver_on_pendulum_env(cuda_device):
    """learn_dynamics_iteratively_w_mpc (utils/iterative_mpc.py) end to end on the dependency-free pendulum env:
    random-policy episodes -> train -> MPC episodes (shared handler, re-staged weights) -> retrain."""
    from blackbox_mpc_b200.environment_utils import PendulumVecEnv
    from blackbox_mpc_b200.policies.random_policy import RandomPolicy
    from blackbox_mpc_b200.utils.iterative_mpc import learn_dynamics_iteratively_w_mpc
    from blackbox_mpc_b200.utils.rollouts import perform_rollouts

    class Writer:
        def __init__(self):
            self.rows = {}

        def add_scalar(self, tag, value, step):
            self.rows.setdefault(tag, []).append(value)

    n = 4
    env = PendulumVecEnv(num_of_agents=n, seed=0)
    mlp = DeterministicMLP([4, 64, 64, 3], ["tanh", "tanh", None], seed=1)
    writer = Writer()
    handler, policy = learn_dynamics_iteratively_w_mpc(
        env, number_of_initial_rollouts=8, number_of_rollouts_for_refinement=1, number_of_refinement_steps=1, task_horizon=60,
        env_action_space=env.action_space, env_observation_space=env.observation_space,
        initial_policy=RandomPolicy(n, env.action_space, seed=0), planning_horizon=12,
        reward_function=pendulum.pendulum_reward_function, optimizer_name="CEM", num_agents=n, dynamics_function=mlp,
        tf_writer=writer, epochs=20, learning_rate=2e-3, population_size=256, max_iterations=3, num_elite=16)
    assert handler._training_iter == 2 and handler._model_training_in.shape[0] > 8 * 60 * n * 0.6
    # the MPC episode was logged with the reference's tags; the learned model predicts the next observation well
    err = writer.rows["states/predicted_observations_abs_error"]
    assert len(err) == 60 and np.mean(err) < 0.25, np.mean(err)
    assert "rewards/actual_episode_reward" in writer.rows and "system_model_val/loss" in writer.rows
    obs, acts, rews = perform_rollouts(env, 1, 40, policy)
    assert obs[0].shape == (41, n, 3) and acts[0].shape == (40, n, 1) and np.isfinite(rews[0]).all()
    assert (acts[0] >= -2 - 1e-6).all() and (acts[0] <= 2 + 1e-6).all()
