"""GPU parity of the pipelined tensor-core rollout (csrc/rollout_pipe.cu) against the one-tile-per-CTA kernel
(csrc/rollout_tc.cu, BBMPC_TC_PIPE=0) and across its own scheduling modes.

Both kernels contract the same bf16 hi/lo products in the same order and form the member sum, the state update
and the reward with the same fp32 operations, so their returns must agree exactly; the number of member-tiles a
CTA interleaves (BBMPC_PIPE_MT = 1, 2, 3: single / half-offset / merged stage order) only changes WHEN a row
is processed, never its arithmetic."""
import numpy as np
import pytest
import torch

import helpers
from blackbox_mpc_b200.utils import workloads

pytestmark = pytest.mark.gpu
TOL = dict(atol=3e-3, rtol=3e-5)


@pytest.fixture(autouse=True)
def _force_pipe(monkeypatch):
    """The dispatcher picks the pipelined kernel only when a CTA gets two or more member-tiles; the tests force it for
    every shape (BBMPC_TC_PIPE=1) and switch to the single-tile kernel explicitly (BBMPC_TC_PIPE=0)."""
    monkeypatch.setenv("BBMPC_TC_PIPE", "1")


def _returns(w, P, precision, seed):
    policy = workloads.build_policy(w, precision=precision)
    ev = policy._trajectory_evaluator
    actions = helpers.random_actions(w, P, seed=seed)
    return ev(torch.from_numpy(w.state), actions, 0).cpu().numpy()


@pytest.mark.parametrize("name,P,A", [("C4", 700, 1), ("C4", 10000, 1), ("C3", 5000, 1), ("C2", 2000, 1), ("C4", 130, 3), ("C3", 40000, 1)])
def test_pipe_equals_single_tile_kernel(cuda_device, monkeypatch, name, P, A):
    w = workloads.make(name, population_size=P, num_agents=A, bias_scale=0.1)
    got = _returns(w, P, "bf16x3", 21)
    assert np.isfinite(got).all()
    again = _returns(w, P, "bf16x3", 21)
    assert np.array_equal(got, again)
    monkeypatch.setenv("BBMPC_TC_PIPE", "0")
    old = _returns(w, P, "bf16x3", 21)
    monkeypatch.setenv("BBMPC_TC_PIPE", "1")
    if not np.array_equal(got, old):
        helpers.compare_returns(got, old, max_jump_frac=0.02, **TOL)
        pytest.fail(f"pipelined and single-tile kernels agree only within tolerance (max |d| = {np.abs(got - old).max()})")


@pytest.mark.parametrize("name,P", [("C4", 10000), ("C3", 3000), ("C4", 3000)])
def test_pipe_modes_bit_identical(cuda_device, monkeypatch, name, P):
    """1, 2 or 3 member-tiles per CTA (several rounds when fewer are allowed) and a sharded population give the
    same bits per row."""
    w = workloads.make(name, population_size=P, bias_scale=0.1)
    ref = _returns(w, P, "bf16x3", 22)
    for mt in ("1", "2"):
        monkeypatch.setenv("BBMPC_PIPE_MT", mt)
        got = _returns(w, P, "bf16x3", 22)
        monkeypatch.delenv("BBMPC_PIPE_MT")
        assert np.array_equal(ref, got), f"BBMPC_PIPE_MT={mt}: max |d| = {np.abs(ref - got).max()}"
    # a shard of the population (rows 1000..) computes the same returns for its rows
    policy = workloads.build_policy(w, precision="bf16x3")
    actions = helpers.random_actions(w, P, seed=22)
    part = policy._trajectory_evaluator(torch.from_numpy(w.state), actions[1000:], 0).cpu().numpy()
    assert np.array_equal(ref[1000:], part)


@pytest.mark.parametrize("name,P", [("C4", 10000), ("C3", 5000)])
def test_pipe_full_size_vs_oracle(cuda_device, name, P):
    """BASELINE sizes against the float64 oracle (the CPU restatement finishes C4 at P = 10 000 in seconds)."""
    w = workloads.make(name, population_size=P, bias_scale=0.1)
    got = _returns(w, P, "bf16x3", 23)
    actions = helpers.random_actions(w, P, seed=23)
    state = torch.from_numpy(w.state)
    ref = helpers.oracle_evaluator(w, torch.float64)(state.double(), actions.double(), 0).numpy()
    helpers.compare_returns(got, ref, max_jump_frac=0.02, **TOL)


def test_pipe_single_pass_bf16(cuda_device, monkeypatch):
    """BBMPC_PREC_BF16 (one pass, 4 KB ring chunks): pipelined == single-tile kernel."""
    w = workloads.make("C4", population_size=1500, bias_scale=0.1)
    got = _returns(w, 1500, "bf16", 24)
    monkeypatch.setenv("BBMPC_TC_PIPE", "0")
    old = _returns(w, 1500, "bf16", 24)
    assert np.array_equal(got, old), f"max |d| = {np.abs(got - old).max()}"
