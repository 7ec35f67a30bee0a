"""SystemDynamicsHandler.

Mirrors blackbox_mpc/dynamics_handlers/system_dynamics_handler.py:7-161: constructor signature,
process_input (:97-126), process_output (:128-161), normalisation statistics (six fp32 vectors,
:84-95, :340-348).  The handler owns the Engine (one bbmpc_ctx = this model on this GPU) and
re-stages weights/statistics into libbbmpc whenever the dynamics function's version changes (the
handler is shared with the trainer in the reference, utils/iterative_mpc.py:147-157).

The training half (train, :163-243) lives in training.py (PyTorch stock ops: it is not
on the act() path); train() appends the trajectories, fits the model(s) in place and bumps the versions
so that the next act() re-stages weights and statistics."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..engine import Engine
from . import training

_STAT_NAMES = ("mean_states", "std_states", "mean_actions", "std_actions", "mean_targets", "std_targets")


def stage_model(engine: Engine, dynamics_function) -> None:
    """bbmpc_model_set_mlp / bbmpc_model_set_builtin for a dynamics_function object."""
    lib = engine.lib
    builtin = getattr(dynamics_function, "bbmpc_dynamics_id", None)
    if builtin is not None:
        engine.check(lib.bbmpc_model_set_builtin(engine.handle, builtin, dynamics_function.dim_S, dynamics_function.dim_U))
        return
    if not hasattr(dynamics_function, "members"):
        raise TypeError(
            "dynamics_function must be a DeterministicMLP / EnsembleMLP or a built-in analytical model "
            "(e.g. utils.pendulum.PendulumTrueModel): arbitrary Python callables cannot be fused into "
            "the sm_100a rollout kernel")
    members = dynamics_function.members()
    dims = dynamics_function.layer_sizes
    n_layers = len(dims) - 1
    n = len(members) * n_layers
    W = (C.c_void_p * n)(*[m.weights[l].data_ptr() for m in members for l in range(n_layers)])
    B = (C.c_void_p * n)(*[m.biases[l].data_ptr() for m in members for l in range(n_layers)])
    for m in members:
        for t in m.weights + m.biases:
            assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
    engine.check(lib.bbmpc_model_set_mlp(engine.handle, len(members), n_layers, (C.c_int * (n_layers + 1))(*dims), W, B,
                                         (C.c_int * n_layers)(*dynamics_function.activation_ids), engine.stream()))


class SystemDynamicsHandler:
    def __init__(self, env_action_space, env_observation_space, dynamics_function=None, true_model=False,
                 is_normalized=True, log_dir=None, tf_writer=None, save_model_frequency=1, saved_model_dir=None,
                 transform_targets_func=None, inverse_transform_targets_func=None,
                 device: Optional[int] = None, seed: int = 0, precision: str = "auto"):
        if transform_targets_func is not None or inverse_transform_targets_func is not None:
            raise NotImplementedError("only the default target transform (next - current) is fused")
        self._is_true_model = bool(true_model)
        self._dim_S = int(env_observation_space.shape[0])
        self._dim_U = int(env_action_space.shape[0])
        self._dynamics_function = dynamics_function
        self._is_normalized = bool(is_normalized)
        self._log_dir, self._tf_writer = log_dir, tf_writer
        self._save_model_frequency, self._saved_model_dir = save_model_frequency, saved_model_dir
        self.engine = Engine(device=device, seed=seed, precision=precision)
        self._stats = None          # six CUDA tensors once set
        self._stats_version = 0
        self._staged = (None, -1, -1)   # (id(dynamics_function), its version, stats version)
        # training state (:60-76)
        self._model_training_in = np.zeros((0, self._dim_S + self._dim_U), np.float32)
        self._model_training_out = np.zeros((0, self._dim_S), np.float32)
        self._model_validation_in = np.zeros((0, self._dim_S + self._dim_U), np.float32)
        self._model_validation_out = np.zeros((0, self._dim_S), np.float32)
        self._first_time, self._training_iter, self._refining_model_iter = True, 0, 0
        self._rng = np.random.default_rng(seed)
        self._torch_gen = torch.Generator().manual_seed(seed)
        self.last_training_loss = self.last_validation_loss = None
        if saved_model_dir is not None and not self._is_true_model:
            if self._is_normalized:   # same six file names as the reference (:84-95)
                self.set_normalization(*[np.load(os.path.join(saved_model_dir, n + ".npy")) for n in _STAT_NAMES])
            saved = training.load_weights(saved_model_dir)
            if saved is not None:
                if dynamics_function is None or not hasattr(dynamics_function, "members"):
                    # the reference restores the whole model from the directory (:78-83); weights.npz carries no
                    # activation list, so a model object of the right shape must be supplied instead of dropping them
                    raise ValueError("saved_model_dir holds trained weights: pass the DeterministicMLP / EnsembleMLP "
                                     "they belong to as dynamics_function")
                for member, (ws, bs) in zip(dynamics_function.members(), saved):
                    member.set_weights(ws, bs)
            # the loaded statistics are the ones the loaded weights were trained against: keep them when train() is
            # called later (reference :80-83 sets _first_time = False after loading)
            self._first_time = False

    # -- statistics ---------------------------------------------------------------------------
    def set_normalization(self, mean_states, std_states, mean_actions, std_actions, mean_targets, std_targets):
        dev = self.engine.device
        vals = [torch.as_tensor(np.asarray(v), dtype=torch.float32).to(dev).contiguous()
                for v in (mean_states, std_states, mean_actions, std_actions, mean_targets, std_targets)]
        for v, n in zip(vals, (self._dim_S, self._dim_S, self._dim_U, self._dim_U, self._dim_S, self._dim_S)):
            if v.numel() != n:
                raise ValueError("normalisation statistic of the wrong length")
        (self._mean_states, self._std_states, self._mean_actions, self._std_actions,
         self._mean_targets, self._std_targets) = vals
        self._stats = vals
        self._stats_version += 1

    # -- staging ------------------------------------------------------------------------------
    def ensure_staged(self) -> Engine:
        f = self._dynamics_function
        if f is None:
            raise RuntimeError("SystemDynamicsHandler has no dynamics_function")
        key = (id(f), getattr(f, "version", 0), self._stats_version)
        if key != self._staged:
            e, lib = self.engine, self.engine.lib
            use_norm = self._is_normalized and not self._is_true_model
            if use_norm and self._stats is None:
                raise RuntimeError("is_normalized=True but no statistics were set "
                                   "(set_normalization(...) or saved_model_dir)")
            p = [_lib.ptr(t) for t in self._stats] if use_norm else [None] * 6
            e.check(lib.bbmpc_model_set_norm(e.handle, self._dim_S, self._dim_U, p[0], p[1], p[2], p[3], p[4], p[5], e.stream()))
            stage_model(e, f)
            self._staged = key
        return self.engine

    # -- reference-API tensor functions (used outside the fused path, e.g. by user code) -------
    def process_input(self, states, actions):
        if self._is_true_model or not self._is_normalized:
            return torch.cat([states, actions], dim=-1)
        return torch.cat([(states - self._mean_states) / (self._std_states + 1e-7),
                          (actions - self._mean_actions) / (self._std_actions + 1e-7)], dim=-1)

    def process_output(self, inputs_states, raw_output):
        if self._is_true_model or not self._is_normalized:
            return raw_output + inputs_states
        return (self._mean_targets + raw_output * (self._std_targets + 1e-7)) + inputs_states

    def get_dynamics_function(self):
        return self._dynamics_function

    # -- training half (:163-243) ---------------------------------------------------------------
    def train(self, observations_trajectories, actions_trajectories, rewards_trajectories, validation_split=0.2,
              batch_size=128, learning_rate=1e-3, epochs=30, nn_optimizer=None):
        """Appends the trajectories to the dataset, (re)computes the statistics the first time, fits every
        member of the dynamics function (Adam + loss_fn, default MSE) and saves `saved_model_<k>/` every
        `save_model_frequency` calls.  `nn_optimizer` is accepted for signature compatibility (Adam is used)."""
        if self._is_true_model:
            raise Exception("a true model has nothing to train")
        f = self._dynamics_function
        if f is None or not hasattr(f, "members"):
            raise TypeError("train() needs a DeterministicMLP / EnsembleMLP dynamics_function")
        new_in, new_out = training.trajectories_to_samples(observations_trajectories, actions_trajectories, self._dim_S, self._dim_U)
        tr_in, tr_out, va_in, va_out = training.split_train_validation(new_in, new_out, validation_split, self._rng)
        self._model_training_in = np.concatenate([self._model_training_in, tr_in], 0)
        self._model_training_out = np.concatenate([self._model_training_out, tr_out], 0)
        self._model_validation_in = np.concatenate([self._model_validation_in, va_in], 0)
        self._model_validation_out = np.concatenate([self._model_validation_out, va_out], 0)
        if self._first_time:
            if self._is_normalized:
                self.set_normalization(*training.normalization_stats(self._model_training_in, self._model_training_out, self._dim_S))
            self._first_time = False
        if self._is_normalized:
            stats = [t.cpu().numpy() for t in self._stats]
            train_xy = training.normalize(self._model_training_in, self._model_training_out, stats, self._dim_S)
            val_xy = training.normalize(self._model_validation_in, self._model_validation_out, stats, self._dim_S)
        else:   # the reference normalises unconditionally and crashes here; raw data is the obvious intent
            stats = None
            train_xy = (self._model_training_in, self._model_training_out)
            val_xy = (self._model_validation_in, self._model_validation_out)

        def log_epoch(ep, _tr, va):
            if self._tf_writer is not None and hasattr(self._tf_writer, "add_scalar"):
                self._tf_writer.add_scalar("system_model_val/loss", va, self._refining_model_iter * epochs + ep)

        losses = []
        for member in f.members():
            losses.append(training.fit_mlp(member.weights, member.biases, member.activation_ids, train_xy, val_xy, epochs=epochs,
                                           learning_rate=learning_rate, batch_size=batch_size, loss_fn=member.loss_fn,
                                           generator=self._torch_gen, on_epoch=log_epoch))
            member.mark_dirty()
        self.last_training_loss, self.last_validation_loss = losses[0]
        self._refining_model_iter += 1
        self._training_iter += 1
        if self._training_iter % self._save_model_frequency == 0 and self._log_dir is not None:
            training.save_model(os.path.join(self._log_dir, f"saved_model_{self._refining_model_iter}"),
                                [m.weights for m in f.members()], [m.biases for m in f.members()], stats)
