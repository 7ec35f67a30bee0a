"""GPU parity of CMA-ES (optimizers/cma_es.py of the reference): the CUDA path's own z draws and its own
eigendecompositions (eigenvector signs / degenerate subspaces are solver-specific) are injected into the
oracle, so that mean / step size / evolution paths / covariance are compared on identical inputs, one
optimizer iteration at a time (split API).  The eigensolver is checked separately against its definition."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from blackbox_mpc_b200 import _lib
from blackbox_mpc_b200.utils import workloads

pytestmark = pytest.mark.gpu


def _workload(P=96, A=2, H=8, E=16, iters=4):
    w = workloads.make("C2", population_size=P, planning_horizon=H, num_agents=A, bias_scale=0.1)
    w.optimizer_name, w.max_iterations = "CMA-ES", iters
    w.optimizer_args = dict(num_elite=E, alpha_cov=2.0, h_sigma=1.0)
    return w


def _drive(opts, w, world):
    """Runs one act() through begin / iter_local / all-gather (emulated) / iter_merge / finish on `world`
    handles living on one GPU; returns the per-iteration state of rank 0 and the final actions."""
    lib, dev = opts[0]._engine.lib, opts[0]._engine.device
    n = lib.bbmpc_opt_partial_floats(opts[0]._handle)
    gathered = torch.empty(world, n, device=dev)
    state = torch.from_numpy(w.state).to(dev)
    for o in opts:
        o._engine.check(lib.bbmpc_opt_begin(o._handle, _lib.ptr(state), 0, None))
    per_iter = []
    for it in range(lib.bbmpc_opt_num_iterations(opts[0]._handle)):
        for r, o in enumerate(opts):
            o._engine.check(lib.bbmpc_opt_iter_local(o._handle, it, gathered[r].data_ptr(), None))
        for o in opts:
            o._engine.check(lib.bbmpc_opt_iter_merge(o._handle, it, _lib.ptr(gathered), world, None))
        torch.cuda.synchronize()
        per_iter.append({k: opts[0].get_tensor(k).cpu() for k in ("m", "sigma", "C", "B", "D", "p_sigma", "p_C")})
    acts = []
    for o in opts:
        act = torch.empty(w.num_agents, w.dU, device=dev)
        o._engine.check(lib.bbmpc_opt_finish(o._handle, 0, _lib.ptr(act), None, None, None))
        torch.cuda.synchronize()
        acts.append(act.cpu())
    return per_iter, acts


def _workload_c5(P=2000, iters=2):
    """BASELINE C5 shape: HalfCheetah 3x200 MLP, H = 50, dU = 6 -> N = 300 search dimensions, 50 elites."""
    w = workloads.make("C5", population_size=P, bias_scale=0.1)
    w.max_iterations = iters
    return w


@pytest.mark.parametrize("which", ["small", "c5_n300"])
def test_cmaes_matches_oracle(cuda_device, which):
    w = _workload() if which == "small" else _workload_c5()
    N = w.num_agents * w.planning_horizon * w.dU
    assert which == "small" or N == 300
    policy = workloads.build_policy(w, precision="fp32")
    opt = policy._optimizer
    opt._ensure_handle()
    ztrace = opt.enable_sample_trace()            # CMA-ES records the raw z draws [iters, P*N]
    per_iter, acts = _drive([opt], w, 1)
    zs = [z.reshape(w.population_size, N).cpu() for z in ztrace]
    assert all(abs(float(z.mean())) < 0.2 and 0.8 < float(z.std()) < 1.2 for z in zs)

    # eigensolver: B orthogonal, D descending, B D^2 B^T = C
    for st in per_iter:
        B, d, Cm = st["B"].reshape(N, N).double(), st["D"].double(), st["C"].reshape(N, N).double()
        assert (d[:-1] >= d[1:] - 1e-6).all()
        np.testing.assert_allclose((B.T @ B).numpy(), np.eye(N), atol=5e-5)
        np.testing.assert_allclose((B @ torch.diag(d * d) @ B.T).numpy(), Cm.numpy(), atol=5e-5, rtol=1e-4)
        np.testing.assert_allclose(Cm.numpy(), Cm.T.numpy(), atol=0)

    eigs = [(st["D"].double() ** 2, st["B"].reshape(N, N).double()) for st in per_iter]
    o = helpers.oracle_optimizer(w, "CMA-ES", dtype=torch.float64, eig_fn=lambda C: eigs.pop(0))
    draws = oracle.InjectedDraws({"cmaes.z": zs}, dtype=torch.float64)
    ref_action = o._optimize(torch.from_numpy(w.state).double(), 0, draws)
    # N = 300: sums of 300 fp32 products per sample and 50 x 300 x 300 rank-mu terms; elite membership is exact unless two
    # returns tie within fp32 noise, the statistics agree to fp32 accumulation error
    rtol, atol = (2e-4, 2e-5) if which == "small" else (1e-3, 2e-4)
    for it, (st, ref) in enumerate(zip(per_iter, o.trace)):
        for key, rk in (("m", "m"), ("sigma", "sigma"), ("p_sigma", "p_sigma"), ("p_C", "p_C"), ("C", "C")):
            np.testing.assert_allclose(st[key].numpy().ravel(), ref[rk].numpy().ravel(), rtol=rtol, atol=atol,
                                       err_msg=f"iteration {it}: {key}")
    np.testing.assert_allclose(acts[0].numpy(), ref_action.numpy(), rtol=rtol, atol=atol)


def test_cmaes_shard_invariance(cuda_device):
    """2 and 3 emulated ranks reproduce the unsharded run bit for bit (draws keyed on the global row,
    exact global top-E, identical merge arithmetic)."""
    w = _workload(P=101)
    ref_policy = workloads.build_policy(w, precision="fp32")
    ref_policy._optimizer._ensure_handle()
    ref_iter, ref_acts = _drive([ref_policy._optimizer], w, 1)
    for world in (2, 3):
        pols = [workloads.build_policy(w, precision="fp32") for _ in range(world)]
        opts = [p._optimizer for p in pols]
        for r, o in enumerate(opts):
            o.shard(r, world)
            o._ensure_handle()
        per_iter, acts = _drive(opts, w, world)
        for k in ("m", "sigma", "C", "p_sigma", "p_C"):
            assert torch.equal(per_iter[-1][k], ref_iter[-1][k]), f"world={world}: {k}"
        for a in acts:
            assert torch.equal(a, ref_acts[0])


def test_cmaes_policy_act_and_reset(cuda_device):
    """MPCPolicy(optimizer_name='CMA-ES').act end to end; state persists across act() calls, reset()
    restores m and sigma only (cma_es.py:215-227)."""
    w = _workload(P=64, A=1, H=6, E=8, iters=2)
    policy = workloads.build_policy(w, precision="fp32")
    a1, nxt, rew = policy.act(w.state[0], 0)
    assert a1.shape == (w.dU,) and np.isfinite(a1).all() and (a1 >= w.lb - 1e-6).all() and (a1 <= w.ub + 1e-6).all()
    opt = policy._optimizer
    C1 = opt.get_tensor("C").cpu()
    policy.reset()
    m = opt.get_tensor("m").cpu().numpy()
    np.testing.assert_allclose(m, np.tile((w.lb + w.ub) / 2, m.size // w.dU))
    assert torch.equal(opt.get_tensor("C").cpu(), C1)       # covariance survives reset()
