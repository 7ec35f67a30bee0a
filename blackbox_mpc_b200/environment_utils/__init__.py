"""Vectorised environments.  The reference wraps gym environments in daemon subprocesses
(blackbox_mpc/environment_utils/subprocess_env.py); gym / MuJoCo are not part of this image, so this package
only fixes the protocol the drivers in utils/ rely on and ships one dependency-free environment.

Protocol (what SubprocVecEnv exposes): `action_space`, `observation_space` (objects with .shape/.low/.high),
`reset() -> obs [n_agents, dS]`, `step(actions [n_agents, dU]) -> (obs, rewards [n_agents], dones, infos)`."""
from .pendulum_env import PendulumVecEnv  # noqa: F401
