"""CPU tests of the dynamics-training half (dynamics_handlers/training.py): the reference's
train() (system_dynamics_handler.py:163-349) restated with PyTorch stock ops.  No GPU needed: the module is
device-agnostic; the handler wiring is covered by tests/test_gpu_training.py."""
import os

import numpy as np
import torch

from blackbox_mpc_b200.dynamics_handlers import training


def _episodes(n_ep=3, T=40, n_agents=2, dS=3, dU=1, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(dS, dS)).astype(np.float32) * 0.1
    B = rng.normal(size=(dU, dS)).astype(np.float32) * 0.3
    obs_l, act_l = [], []
    for _ in range(n_ep):
        obs = np.zeros((T + 1, n_agents, dS), np.float32)
        obs[0] = rng.normal(size=(n_agents, dS))
        acts = rng.uniform(-1, 1, size=(T, n_agents, dU)).astype(np.float32)
        for t in range(T):
            obs[t + 1] = obs[t] + obs[t] @ A + acts[t] @ B
        obs_l.append(obs); act_l.append(acts)
    return obs_l, act_l, A, B


def test_samples_targets_and_stats():
    obs, acts, A, B = _episodes()
    x, y = training.trajectories_to_samples(obs, acts, 3, 1)
    assert x.shape == (3 * 2 * 40, 4) and y.shape == (3 * 2 * 40, 3)
    # first sample of episode 0, agent 0: (s_0, a_0) -> s_1 - s_0 (default_transform_targets)
    np.testing.assert_allclose(x[0], np.concatenate([obs[0][0, 0], acts[0][0, 0]]))
    np.testing.assert_allclose(y[0], obs[0][1, 0] - obs[0][0, 0], rtol=1e-6)
    np.testing.assert_allclose(y, x[:, :3] @ A + x[:, 3:] @ B, atol=1e-5)
    stats = training.normalization_stats(x, y, 3)
    np.testing.assert_allclose(stats[0], x[:, :3].mean(0)); np.testing.assert_allclose(stats[1], x[:, :3].std(0))   # ddof 0
    np.testing.assert_allclose(stats[3], x[:, 3:].std(0)); np.testing.assert_allclose(stats[5], y.std(0))
    xn, yn = training.normalize(x, y, stats, 3)
    np.testing.assert_allclose(xn.mean(0), 0, atol=1e-5); np.testing.assert_allclose(yn.std(0), 1, atol=1e-4)


def test_split_is_bernoulli_and_disjoint():
    x = np.arange(4000, dtype=np.float32).reshape(1000, 4); y = x[:, :3]
    tr_x, tr_y, va_x, va_y = training.split_train_validation(x, y, 0.2, np.random.default_rng(1))
    assert tr_x.shape[0] + va_x.shape[0] == 1000 and 120 < va_x.shape[0] < 280
    assert not set(tr_x[:, 0].tolist()) & set(va_x[:, 0].tolist())
    np.testing.assert_array_equal(tr_y, tr_x[:, :3])


def test_fit_reduces_loss_and_updates_in_place(tmp_path):
    obs, acts, _, _ = _episodes(n_ep=6)
    x, y = training.trajectories_to_samples(obs, acts, 3, 1)
    stats = training.normalization_stats(x, y, 3)
    xn, yn = training.normalize(x, y, stats, 3)
    g = torch.Generator().manual_seed(0)
    W = [torch.randn(4, 32, generator=g) * 0.3, torch.randn(32, 3, generator=g) * 0.3]
    b = [torch.zeros(32), torch.zeros(3)]
    w0 = W[0].clone()
    tr, va = training.fit_mlp(W, b, [1, 0], (xn[:400], yn[:400]), (xn[400:], yn[400:]), epochs=40, learning_rate=3e-3,
                              batch_size=64, generator=g)
    assert tr.shape == (40,) and va.shape == (40,)
    assert tr[-1] < 0.2 * tr[0] and va[-1] < 0.3 * va[0]
    assert not torch.equal(W[0], w0)                       # trained in place
    pred = training.mlp_forward(torch.from_numpy(xn[400:]), W, b, [1, 0])
    assert float(torch.mean((pred - torch.from_numpy(yn[400:])) ** 2)) < 0.1
    # fewer samples than one batch: drop_remainder leaves no batch, losses stay NaN, weights unchanged
    w1 = W[0].clone()
    tr2, _ = training.fit_mlp(W, b, [1, 0], (xn[:10], yn[:10]), (xn[:5], yn[:5]), epochs=2, batch_size=64)
    assert np.isnan(tr2).all() and torch.equal(W[0], w1)
    # on-disk contract: saved_model_<k>/ with the reference's six statistics files + weights.npz
    d = os.path.join(tmp_path, "saved_model_1")
    training.save_model(d, [W], [b], stats)
    for n in training.STAT_NAMES:
        assert os.path.exists(os.path.join(d, n + ".npy"))
    (ws, bs), = training.load_weights(d)
    np.testing.assert_array_equal(ws[1], W[1].numpy()); np.testing.assert_array_equal(bs[0], b[0].numpy())
    assert training.load_weights(str(tmp_path)) is None
