"""Twin of the reference's tutorials/true_model_mpc/tutorial_one.py: MPC on Pendulum with the analytical
model and RandomSearch.  Only the imports and the environment line differ (gym is not in this image, so the
dependency-free PendulumVecEnv stands in for EnvironmentWrapper.make_standard_gym_env("Pendulum-v0", ...)).

    python tutorials/true_model_mpc_tutorial_one.py        # needs one B200
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from blackbox_mpc_b200.environment_utils import PendulumVecEnv
from blackbox_mpc_b200.policies.mpc_policy import MPCPolicy
from blackbox_mpc_b200.utils.pendulum import PendulumTrueModel, pendulum_reward_function
from blackbox_mpc_b200.utils.rollouts import perform_rollouts

logging.basicConfig(level=logging.INFO)

number_of_agents = 1
env = PendulumVecEnv(num_of_agents=number_of_agents, seed=0)
my_policy = MPCPolicy(reward_function=pendulum_reward_function,
                      env_action_space=env.action_space,
                      env_observation_space=env.observation_space,
                      true_model=True,
                      dynamics_function=PendulumTrueModel(),
                      optimizer_name='RandomSearch',
                      num_agents=number_of_agents,
                      planning_horizon=30, population_size=500)

observations, actions, rewards = perform_rollouts(env, number_of_rollouts=2, task_horizon=200, policy=my_policy)
print("episode rewards:", [float(r.sum(0).mean()) for r in rewards])
