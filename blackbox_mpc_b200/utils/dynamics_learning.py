"""collect -> train (blackbox_mpc/utils/dynamics_learning.py:7-90)."""
from ..dynamics_handlers.system_dynamics_handler import SystemDynamicsHandler
from .rollouts import perform_rollouts


def learn_dynamics_from_policy(env, policy, number_of_rollouts, task_horizon, dynamics_function=None,
                               system_dynamics_handler=None, epochs=30, learning_rate=1e-3, validation_split=0.2,
                               batch_size=128, is_normalized=True, nn_optimizer=None, tf_writer=None,
                               exploration_noise=False, log_dir=None, save_model_frequency=1, saved_model_dir=None,
                               start_episode=0):
    """Rolls `policy` out in `env`, trains the handler's dynamics function on the episodes, returns the handler
    (a new one around `dynamics_function` unless `system_dynamics_handler` is given)."""
    handler = system_dynamics_handler
    if handler is None:
        handler = SystemDynamicsHandler(env_action_space=env.action_space, env_observation_space=env.observation_space,
                                        true_model=False, dynamics_function=dynamics_function, tf_writer=tf_writer,
                                        is_normalized=is_normalized, log_dir=log_dir,
                                        save_model_frequency=save_model_frequency, saved_model_dir=saved_model_dir)
    observations, actions, rewards = perform_rollouts(env, number_of_rollouts, task_horizon, policy,
                                                      exploration_noise=exploration_noise, tf_writer=tf_writer,
                                                      start_episode=start_episode)
    handler.train(observations, actions, rewards, validation_split=validation_split, batch_size=batch_size,
                  learning_rate=learning_rate, epochs=epochs, nn_optimizer=nn_optimizer)
    return handler
