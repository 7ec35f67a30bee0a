"""MPCPolicy — the drop-in boundary (blackbox_mpc/policies/mpc_policy.py:10-245).

Same constructor keywords, optimizer_name dispatch (:81-116), `act` marshalling (:149-172: 1-D
observations are tiled to num_agents rows and un-batched on return), `reset`, `switch_optimizer`.
`act` is one call into libbbmpc (bbmpc_opt_call_host): H2D of the observation, the whole optimizer
loop on the GPU, D2H of (action, next observation, reward), one synchronisation."""
from __future__ import annotations

import numpy as np

from ..dynamics_handlers.system_dynamics_handler import SystemDynamicsHandler
from ..trajectory_evaluators.deterministic import DeterministicTrajectoryEvaluator
from .model_based_base_policy import ModelBasedBasePolicy


class MPCPolicy(ModelBasedBasePolicy):
    def __init__(self, trajectory_evaluator=None, optimizer=None, tf_writer=None, log_dir=None,
                 reward_function=None, env_action_space=None, env_observation_space=None,
                 dynamics_function=None, dynamics_handler=None, true_model=False, optimizer_name=None,
                 num_agents=None, save_model_frequency=1, saved_model_dir=None, **optimizer_args):
        if trajectory_evaluator is None:
            if dynamics_handler is None:
                dynamics_handler = SystemDynamicsHandler(
                    env_action_space=env_action_space, env_observation_space=env_observation_space,
                    true_model=true_model, dynamics_function=dynamics_function, log_dir=log_dir,
                    tf_writer=tf_writer, save_model_frequency=save_model_frequency,
                    saved_model_dir=saved_model_dir)
            trajectory_evaluator = DeterministicTrajectoryEvaluator(
                reward_function=reward_function, system_dynamics_handler=dynamics_handler)
        super().__init__(trajectory_evaluator=trajectory_evaluator)
        if optimizer is None:
            if num_agents is None:
                raise Exception("Please Specify Num Of Agents in the MPC")
            optimizer = self._make_optimizer(optimizer_name, env_action_space, env_observation_space,
                                             num_agents, optimizer_args)
        self._optimizer = optimizer
        self._tf_writer = tf_writer
        self._trajectory_evaluator = trajectory_evaluator
        self._optimizer.set_trajectory_evaluator(trajectory_evaluator)  # AttributeError if the name was unknown (:117-120)
        self._act_call_counter = 0

    @staticmethod
    def _make_optimizer(optimizer_name, action_space, observation_space, num_agents, optimizer_args):
        from ..optimizers import BY_NAME
        cls = BY_NAME.get(optimizer_name)
        if cls is None:
            return None  # the reference silently leaves optimizer=None for unknown names
        return cls(env_action_space=action_space, env_observation_space=observation_space,
                   num_agents=num_agents, **optimizer_args)

    def act(self, observations, t, exploration_noise=False):
        observations = np.asarray(observations)
        batched = observations
        if observations.ndim == 1:
            batched = np.tile(observations[None, :], (self._optimizer._num_agents, 1))
        if hasattr(self._optimizer, "call_host") and self._optimizer.KIND is not None:
            action, next_obs, reward = self._optimizer.call_host(batched, int(t), bool(exploration_noise))
        else:  # user-defined Python optimizer
            import torch
            a, n, r = self._optimizer(torch.as_tensor(batched, dtype=torch.float32), int(t), bool(exploration_noise))
            action, next_obs, reward = a.cpu().numpy(), n.cpu().numpy(), r.cpu().numpy()
        self._act_call_counter += 1
        if observations.ndim == 1:
            action, next_obs, reward = action[0], next_obs[0], reward[0]
        return action, next_obs, reward

    def reset(self):
        self._optimizer.reset()

    def switch_optimizer(self, optimizer=None, optimizer_name='', **optimizer_args):
        if optimizer is None:
            old = self._optimizer
            new = self._make_optimizer(optimizer_name, old._env_action_space, old._env_observation_space,
                                       old._num_agents, optimizer_args)
            if new is not None:
                self._optimizer = new
        else:
            self._optimizer = optimizer
        self._optimizer.set_trajectory_evaluator(self._trajectory_evaluator)
