"""blackbox_mpc_b200 — B200-native sampling-MPC rollout engine behind the plugin surface of
ossamaAhmed/blackbox_mpc (MPCPolicy.act / OptimizerBase / EvaluatorBase / dynamics_function /
reward_function).  Python host code holds torch CUDA tensors and calls libbbmpc.so
(include/bbmpc.h) through ctypes; there is no CPU or eager-PyTorch fallback on the hot path."""
__version__ = "0.1.0"
