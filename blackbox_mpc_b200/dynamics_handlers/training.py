"""Training half of the dynamics handler — the caller on the other side of the hot path: it
produces the weights and normalisation statistics that the rollout kernel stages.

Restates blackbox_mpc/dynamics_handlers/system_dynamics_handler.py:163-349 with PyTorch stock ops (a small
MLP, Adam, MSE: nothing here needs a hand-written kernel) and keeps its on-disk contract:
`<log_dir>/saved_model_<k>/` holding the six `mean_*/std_*.npy` files (:224-241) plus `weights.npz`
(the reference writes a TF SavedModel there, which has no meaning without TensorFlow).

Everything is device-agnostic (the tensors decide), so the `-m "not gpu"` tests run it on the CPU; the
handler calls it with the CUDA weight tensors of a DeterministicMLP and then bumps their version so the
engine re-stages them (utils/iterative_mpc.py:147-157: handler shared by trainer and evaluator)."""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

STAT_NAMES = ("mean_states", "std_states", "mean_actions", "std_actions", "mean_targets", "std_targets")
_ACTS = {0: None, 1: torch.tanh, 2: torch.relu, 3: torch.sigmoid}   # ids of include/bbmpc.h (BBMPC_ACT_*)


def trajectories_to_samples(observations_trajectories, actions_trajectories, dim_S: int, dim_U: int):
    """:300-318.  Per episode: observations [T+1, n_agents, dS], actions [T, n_agents, dU] ->
    inputs [N, dS+dU] = (s_t, a_t), targets [N, dS] = s_{t+1} - s_t (utils/transforms.py:4-17)."""
    ins, outs = [], []
    for obs, acs in zip(observations_trajectories, actions_trajectories):
        obs, acs = np.asarray(obs, np.float32), np.asarray(acs, np.float32)
        for agent in range(acs.shape[1]):
            states, nxt = obs[:-1, agent], obs[1:, agent]
            ins.append(np.concatenate([states, acs[:, agent]], axis=-1))
            outs.append(nxt - states)
    if not ins:
        return np.zeros((0, dim_S + dim_U), np.float32), np.zeros((0, dim_S), np.float32)
    return (np.concatenate(ins, 0).reshape(-1, dim_S + dim_U).astype(np.float32),
            np.concatenate(outs, 0).reshape(-1, dim_S).astype(np.float32))


def split_train_validation(data_in, data_out, validation_split: float, rng: np.random.Generator):
    """:319-327: every sample goes to the training set with probability 1 - validation_split."""
    train = rng.random(data_in.shape[0]) >= validation_split
    return data_in[train], data_out[train], data_in[~train], data_out[~train]


def normalization_stats(train_in, train_out, dim_S: int) -> List[np.ndarray]:
    """:340-348: mean / population std (ddof 0) of states, actions and targets of the training set."""
    s, a = train_in[:, :dim_S], train_in[:, dim_S:]
    return [np.mean(s, 0), np.std(s, 0), np.mean(a, 0), np.std(a, 0), np.mean(train_out, 0), np.std(train_out, 0)]


def normalize(data_in, data_out, stats, dim_S: int):
    """:334-338."""
    ms, ss, ma, sa, mt, st = stats
    x = np.concatenate([(data_in[:, :dim_S] - ms) / (ss + 1e-7), (data_in[:, dim_S:] - ma) / (sa + 1e-7)], axis=1)
    return x.astype(np.float32), ((data_out - mt) / (st + 1e-7)).astype(np.float32)


def mlp_forward(x, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor], activation_ids: Sequence[int]):
    """dynamics_functions/deterministic_mlp.py:49-51 with autograd-capable torch ops."""
    for w, b, a in zip(weights, biases, activation_ids):
        x = x @ w + b
        if _ACTS[a] is not None:
            x = _ACTS[a](x)
    return x


def fit_mlp(weights: List[torch.Tensor], biases: List[torch.Tensor], activation_ids: Sequence[int],
            train_xy: Tuple[np.ndarray, np.ndarray], val_xy: Tuple[np.ndarray, np.ndarray], epochs: int = 30,
            learning_rate: float = 1e-3, batch_size: int = 128, loss_fn: Optional[Callable] = None,
            generator: Optional[torch.Generator] = None, on_epoch: Optional[Callable[[int, float, float], None]] = None):
    """:245-298.  Adam (Keras defaults: beta 0.9/0.999, eps 1e-7) over shuffled batches with drop_remainder,
    loss_fn(expected, predicted) (default: mean squared error), per-epoch mean training / validation loss.
    Updates `weights` / `biases` IN PLACE and returns (training_loss[epochs], validation_loss[epochs])."""
    dev = weights[0].device
    loss_fn = loss_fn or (lambda expected, predicted: torch.mean((expected - predicted) ** 2))
    params = [t.detach().clone().requires_grad_(True) for t in list(weights) + list(biases)]
    nw = len(weights)
    opt = torch.optim.Adam(params, lr=learning_rate, betas=(0.9, 0.999), eps=1e-7)
    x, y = (torch.as_tensor(v, dtype=torch.float32, device=dev) for v in train_xy)
    xv, yv = (torch.as_tensor(v, dtype=torch.float32, device=dev) for v in val_xy)
    n_batches, n_val = x.shape[0] // batch_size, xv.shape[0] // batch_size
    tr_loss, va_loss = np.full(epochs, np.nan), np.full(epochs, np.nan)
    for ep in range(epochs):
        perm = torch.randperm(x.shape[0], generator=generator).to(dev)
        total = 0.0
        for b in range(n_batches):
            idx = perm[b * batch_size:(b + 1) * batch_size]
            loss = loss_fn(y[idx], mlp_forward(x[idx], params[:nw], params[nw:], activation_ids))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            total += float(loss.detach())
        if n_batches:
            tr_loss[ep] = total / n_batches
        with torch.no_grad():
            total = 0.0
            for b in range(n_val):
                sl = slice(b * batch_size, (b + 1) * batch_size)
                total += float(loss_fn(yv[sl], mlp_forward(xv[sl], params[:nw], params[nw:], activation_ids)))
            if n_val:
                va_loss[ep] = total / n_val
        if on_epoch is not None:
            on_epoch(ep, tr_loss[ep], va_loss[ep])
    with torch.no_grad():
        for dst, src in zip(list(weights) + list(biases), params):
            dst.copy_(src)
    return tr_loss, va_loss


def save_model(directory: str, members_weights, members_biases, stats) -> None:
    """`saved_model_<k>/`: weights.npz (member m, layer l -> W_m_l / b_m_l) + the six statistics files."""
    os.makedirs(directory, exist_ok=True)
    arrays = {}
    for m, (ws, bs) in enumerate(zip(members_weights, members_biases)):
        for l, (w, b) in enumerate(zip(ws, bs)):
            arrays[f"W_{m}_{l}"] = w.detach().cpu().numpy()
            arrays[f"b_{m}_{l}"] = b.detach().cpu().numpy()
    np.savez(os.path.join(directory, "weights.npz"), **arrays)
    if stats is not None:
        for name, v in zip(STAT_NAMES, stats):
            np.save(os.path.join(directory, name), np.asarray(v, np.float32))


def load_weights(directory: str):
    """-> [member][layer] lists of (W, b) numpy arrays, or None when the directory has no weights.npz."""
    path = os.path.join(directory, "weights.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    members = sorted({int(k.split("_")[1]) for k in z.files if k.startswith("W_")})
    out = []
    for m in members:
        layers = sorted({int(k.split("_")[2]) for k in z.files if k.startswith(f"W_{m}_")})
        out.append(([z[f"W_{m}_{l}"] for l in layers], [z[f"b_{m}_{l}"] for l in layers]))
    return out
