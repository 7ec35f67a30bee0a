"""Twin of the reference's tutorials/model_based_RL/tutorial_two.py (iterative MPC with a learned model): collect
with a RandomPolicy, train a 2x64 tanh MLP, then alternate CEM-MPC episodes and retraining on the shared handler.

    python tutorials/model_based_rl_tutorial_two.py        # needs one B200
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from blackbox_mpc_b200.dynamics_functions.deterministic_mlp import DeterministicMLP
from blackbox_mpc_b200.environment_utils import PendulumVecEnv
from blackbox_mpc_b200.policies.random_policy import RandomPolicy
from blackbox_mpc_b200.utils.iterative_mpc import learn_dynamics_iteratively_w_mpc
from blackbox_mpc_b200.utils.pendulum import pendulum_reward_function

logging.basicConfig(level=logging.INFO)

number_of_agents = 5
env = PendulumVecEnv(num_of_agents=number_of_agents, seed=0)
dynamics_function = DeterministicMLP(layers=[env.action_space.shape[0] + env.observation_space.shape[0], 64, 64,
                                             env.observation_space.shape[0]],
                                     activation_functions=["tanh", "tanh", None])
initial_policy = RandomPolicy(number_of_agents=number_of_agents, env_action_space=env.action_space)
system_dynamics_handler, mpc_policy = learn_dynamics_iteratively_w_mpc(
    env=env, env_action_space=env.action_space, env_observation_space=env.observation_space,
    number_of_initial_rollouts=5, number_of_rollouts_for_refinement=2, number_of_refinement_steps=3, task_horizon=200,
    planning_horizon=30, initial_policy=initial_policy, optimizer_name='CEM', num_agents=number_of_agents,
    reward_function=pendulum_reward_function, dynamics_function=dynamics_function, epochs=30,
    population_size=500, max_iterations=5, num_elite=50)
print("validation loss of the last fit:", system_dynamics_handler.last_validation_loss[-1])
