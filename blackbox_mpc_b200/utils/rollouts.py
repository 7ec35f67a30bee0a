"""Episode collection (blackbox_mpc/utils/rollouts.py:10-139): host-side orchestration around
`policy.act` and `env.step`.  Model-based policies return (action, predicted next observation, predicted
reward) and get their prediction errors logged; model-free ones return the action only.  The mean wall time
of `act` per step — the reference's own "Average action selection time" (:92-101,133) — is the MPC
steps/sec metric of bench.py seen from the driver.

`tf_writer` may be any object with `add_scalar(tag, value, step)` (e.g. torch.utils.tensorboard.SummaryWriter);
the tags are the reference's (:103-131)."""
import logging
import time

import numpy as np

from ..policies.model_free_base_policy import ModelFreeBasePolicy
from ..policies.random_policy import RandomPolicy

logger = logging.getLogger(__name__)


def _scalar(writer, tag, value, step):
    if writer is not None and hasattr(writer, "add_scalar"):
        writer.add_scalar(tag, float(value), int(step))


def _to_numpy(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _sample(env, horizon, policy, episode_step, exploration_noise=False, tf_writer=None):
    """One episode of `horizon` steps for all agents of the vectorised env ->
    {observations [T+1,n,dS], actions [T,n,dU], rewards [T,n], reward_sum [n]}."""
    model_based = not isinstance(policy, ModelFreeBasePolicy)
    logs_rewards = not isinstance(policy, RandomPolicy)
    policy.reset()
    observations, actions, rewards, act_seconds = [env.reset()], [], [], []
    reward_sum, predicted_sum = 0, 0
    for t in range(horizon):
        t0 = time.time()
        if model_based:
            action, predicted_obs, predicted_reward = policy.act(observations[t], t, exploration_noise)
            predicted_obs, predicted_reward = _to_numpy(predicted_obs), _to_numpy(predicted_reward)
            predicted_sum = predicted_sum + predicted_reward
        else:
            action = policy.act(observations[t], t)
        action = _to_numpy(action)
        act_seconds.append(time.time() - t0)
        obs, reward, _done, _info = env.step(action)
        step = episode_step * horizon + t
        if logs_rewards:
            _scalar(tf_writer, "rewards/actual_reward", np.mean(reward), step)
        if model_based:
            _scalar(tf_writer, "states/predicted_observations_abs_error", np.mean(np.sum(np.abs(predicted_obs - obs), axis=1)), step)
            _scalar(tf_writer, "rewards/predicted_reward_abs_error", np.mean(np.abs(predicted_reward - reward)), step)
        actions.append(action)
        observations.append(obs)
        rewards.append(reward)
        reward_sum = reward_sum + reward
    if logs_rewards:
        _scalar(tf_writer, "rewards/actual_episode_reward", np.mean(reward_sum), episode_step)
    if model_based:
        _scalar(tf_writer, "rewards/predicted_episode_reward", np.mean(predicted_sum), episode_step)
    logger.info("Average action selection time: %s", np.mean(act_seconds) if act_seconds else float("nan"))
    logger.info("Rollout length: %d", len(actions))
    return {"observations": np.array(observations), "actions": np.array(actions), "rewards": np.array(rewards),
            "reward_sum": reward_sum, "act_seconds": np.array(act_seconds)}


def perform_rollouts(env, number_of_rollouts, task_horizon, policy, exploration_noise=False, tf_writer=None, start_episode=0):
    """`number_of_rollouts` episodes -> (observations, actions, rewards): three lists of per-episode arrays."""
    logger.info("Started collecting samples for rollouts")
    episodes = [_sample(env, task_horizon, policy, start_episode + i, exploration_noise, tf_writer)
                for i in range(number_of_rollouts)]
    logger.info("Finished collecting samples for rollout")
    return ([e["observations"] for e in episodes], [e["actions"] for e in episodes], [e["rewards"] for e in episodes])
