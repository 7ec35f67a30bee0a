"""GPU parity of the optimizer loop: the CUDA path's own draws are recorded (sample trace) and
injected into the oracle, so sampling->rollout->refit is compared on identical inputs."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from blackbox_mpc_b200.utils import workloads

pytestmark = pytest.mark.gpu


def _run_with_trace(w, name, precision="fp32"):
    policy = workloads.build_policy(w, precision=precision, optimizer_name=name)
    opt = policy._optimizer
    trace = opt.enable_sample_trace()
    state = torch.from_numpy(w.state)
    action, nxt, rew = opt(state, 0, False)
    torch.cuda.synchronize()
    shape = (-1, w.num_agents, w.planning_horizon, w.dU)
    return policy, opt, [t.reshape(shape).cpu() for t in trace], action.cpu(), nxt.cpu(), rew.cpu()


@pytest.mark.parametrize("name,P,A", [("C2", 400, 1), ("C2", 300, 3), ("C4", 256, 1)])
def test_cem_matches_oracle(cuda_device, name, P, A):
    w = workloads.make(name, population_size=P, num_agents=A, bias_scale=0.1)
    w.optimizer_args = dict(num_elite=20, alpha=0.25)
    policy, opt, samples, action, nxt, rew = _run_with_trace(w, "CEM")
    o = helpers.oracle_optimizer(w, "CEM", dtype=torch.float64)
    draws = oracle.InjectedDraws({"cem.samples": samples}, dtype=torch.float64)
    ref_action, ref_next, ref_rew = o(torch.from_numpy(w.state).double(), 0, False, draws)
    # elite sets may differ only where returns tie within fp32 noise; compare the refit statistics
    np.testing.assert_allclose(opt.get_tensor("mean").cpu().numpy().reshape(A, -1),
                               o.trace[-1]["mean"].reshape(A, -1).numpy(), rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(opt.get_tensor("variance").cpu().numpy().reshape(A, -1),
                               o.trace[-1]["variance"].reshape(A, -1).numpy(), rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(action.numpy(), ref_action.numpy(), rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(nxt.numpy(), ref_next.numpy(), rtol=1e-4, atol=2e-4)
    # samples honour the +-2 sigma truncation and the constrained variance: always inside the box
    for s in samples:
        assert (s >= torch.from_numpy(w.lb) - 1e-6).all() and (s <= torch.from_numpy(w.ub) + 1e-6).all()


@pytest.mark.parametrize("opt_name", ["PI2", "RandomSearch", "SPSA"])
def test_other_optimizers_match_oracle(cuda_device, opt_name):
    w = workloads.make("C2", population_size=300, num_agents=2, bias_scale=0.1)
    policy, opt, samples, action, nxt, rew = _run_with_trace(w, opt_name)
    o = helpers.oracle_optimizer(w, opt_name, dtype=torch.float64)
    st = torch.from_numpy(w.state).double()
    if opt_name == "SPSA":
        # the trace holds [plus; minus] clipped parameters; recover delta = sign(plus - minus)
        P = w.population_size
        mids = []
        deltas = []
        for s in samples:
            plus, minus = s[:P], s[P:]
            deltas.append(torch.sign(plus - minus))
        assert all((d.abs() == 1).all() for d in deltas)
        draws = oracle.InjectedDraws({"spsa.delta": deltas}, dtype=torch.float64)
    else:
        tag = {"PI2": "pi2.samples", "RandomSearch": "rs.samples"}[opt_name]
        draws = oracle.InjectedDraws({tag: samples}, dtype=torch.float64)
    ref_action, ref_next, ref_rew = o(st, 0, False, draws)
    np.testing.assert_allclose(action.numpy(), ref_action.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(nxt.numpy(), ref_next.numpy(), rtol=2e-4, atol=2e-4)
    # second act(): warm start state (PI2/SPSA shift) must match too
    if opt_name in ("PI2", "SPSA"):
        key = "previous_solution"
        ref_prev = o._previous_solution if opt_name == "PI2" else o._current_parameters
        np.testing.assert_allclose(opt.get_tensor(key).cpu().numpy().reshape(ref_prev.shape), ref_prev.numpy(), rtol=2e-4, atol=2e-4)


def test_pso_matches_oracle(cuda_device):
    """optimizers/pso.py:79-138 step by step: the CUDA swarm is driven through begin / iter_local / iter_merge / finish
    and its x, v, personal and global bests are compared with the float64 oracle after EVERY iteration, with the same
    initial swarm, the same (r1, r2) scalars (read back through get_tensor("pso_r")) and, for the tail, the CUDA
    path's own re-seed draws injected into the oracle.  A second act() call then runs on the re-seeded swarm."""
    from blackbox_mpc_b200 import _lib
    P, A = 200, 2
    w = workloads.make("C2", population_size=P, num_agents=A, bias_scale=0.1)
    policy = workloads.build_policy(w, precision="fp32", optimizer_name="PSO")
    opt = policy._optimizer
    opt.reset()
    e, lib, h = opt._ensure_handle(), opt._engine.lib, opt._handle
    shape = (P, A, w.planning_horizon, w.dU)
    lb, ub = torch.from_numpy(w.lb), torch.from_numpy(w.ub)
    state = torch.from_numpy(w.state).to(e.device)
    n_iters = lib.bbmpc_opt_num_iterations(h)
    o = helpers.oracle_optimizer(w, "PSO", dtype=torch.float64)
    get = lambda name, shp: opt.get_tensor(name).cpu().reshape(shp).double()   # noqa: E731
    x0, v0 = get("x", shape), get("v", shape)
    assert (x0 >= lb).all() and (x0 <= ub).all() and (v0.abs() <= 0.01 * (ub - lb) + 1e-7).all()
    o._x, o._v, o._pbest_x = x0.clone(), v0.clone(), x0.clone()
    o._pbest_r = torch.full((P, A), -float("inf"), dtype=torch.float64)
    o._gbest_r = torch.full((A,), -float("inf"), dtype=torch.float64)

    for call in range(2):
        # ---- CUDA path, one iteration at a time
        e.check(lib.bbmpc_opt_begin(h, _lib.ptr(state), call, None))
        cuda_iters = []
        for it in range(n_iters):
            e.check(lib.bbmpc_opt_iter_local(h, it, None, None))
            e.check(lib.bbmpc_opt_iter_merge(h, it, None, 1, None))
            torch.cuda.synchronize()
            cuda_iters.append(dict(x=get("x", shape), v=get("v", shape), pbx=get("pbest_x", shape), pbr=get("pbest_r", (P, A)),
                                   gbx=get("gbest_x", shape[1:]), gbr=get("gbest_r", (A,))))
        r12 = opt.get_tensor("pso_r").cpu().double()[: 2 * n_iters].reshape(n_iters, 2)
        gb_final = cuda_iters[-1]["gbx"]
        act = torch.empty(A, w.dU, device=e.device)
        nxt = torch.empty(A, w.dS, device=e.device)
        e.check(lib.bbmpc_opt_finish(h, 0, _lib.ptr(act), _lib.ptr(nxt), None, None))
        torch.cuda.synchronize()
        x_seed, v_seed = get("x", shape), get("v", shape)
        # ---- oracle with the same scalars and the CUDA path's re-seed draws
        draws = oracle.InjectedDraws({"pso.r1": [r12[i, 0].reshape(()) for i in range(n_iters)],
                                      "pso.r2": [r12[i, 1].reshape(()) for i in range(n_iters)],
                                      "pso.reseed_x": [x_seed], "pso.reseed_v": [v_seed]}, dtype=torch.float64)
        ref_action, ref_next, _ = o(torch.from_numpy(w.state).double(), call, False, draws)
        assert len(o.trace) == n_iters
        for it, (c, r) in enumerate(zip(cuda_iters, o.trace)):
            msg = f"act() call {call}, iteration {it}"
            np.testing.assert_allclose(c["x"].numpy(), r["x"].numpy(), rtol=1e-4, atol=2e-5, err_msg="x, " + msg)
            np.testing.assert_allclose(c["v"].numpy(), r["v"].numpy(), rtol=1e-4, atol=2e-5, err_msg="v, " + msg)
            # the swarm contracts towards gbest, so late iterations have near-ties between neighbouring particles: the fp32 path
            # may pick the neighbour (positions then agree to ~1e-4, not to rounding)
            np.testing.assert_allclose(c["gbx"].numpy(), r["gbest_x"].numpy(), rtol=1e-3, atol=2e-4, err_msg="gbest_x, " + msg)
            for a in range(A):   # the global best IS the personal best of the arg-max particle (pso.py:97-103), lowest index on ties
                best_row = int(torch.argmax(c["pbr"][:, a]))
                np.testing.assert_allclose(c["pbx"][best_row, a].numpy(), c["gbx"][a].numpy(), rtol=0, atol=0)
                assert c["gbr"][a] == c["pbr"][:, a].max()
            # personal bests: monotone, and equal to the best clipped position seen so far
            if it > 0:
                assert (c["pbr"] >= cuda_iters[it - 1]["pbr"]).all()
        np.testing.assert_allclose(act.cpu().numpy(), ref_action.numpy(), rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(nxt.cpu().numpy(), ref_next.numpy(), rtol=2e-4, atol=2e-4)
        # ---- the tail (pso.py:114-138): re-seed support and bookkeeping
        H, dU = w.planning_horizon, w.dU
        g = gb_final                                           # [A, H, dU]
        shifted = torch.cat([g[:, 1:], g[:, -1:]], dim=1)
        var0 = ((ub - lb).double() / 4.0) ** 2                 # optimizer_base.py:44-46 solution variance
        cvar = torch.minimum(torch.minimum(((g - lb) / 2) ** 2, ((ub - g) / 2) ** 2), var0.expand_as(g))
        assert ((x_seed - shifted).abs() <= 2.0 * cvar.sqrt() + 1e-6).all(), "re-seeded positions outside +-2 sigma of shift(gbest)"
        assert (v_seed.abs() <= 0.01 * (ub - lb) + 1e-7).all()
        assert torch.equal(get("pbest_x", shape), x_seed)
        assert torch.isinf(get("pbest_r", (P, A))).all() and torch.isinf(get("gbest_r", (A,))).all()
        assert (x_seed[:, :, :, :].std(dim=0) > 0).any()


def test_mpc_policy_act_marshalling(cuda_device):
    """policies/mpc_policy.py:149-172: 1-D obs -> tiled to A rows -> un-batched outputs."""
    w = workloads.make("C2", population_size=200, num_agents=3, bias_scale=0.1)
    policy = workloads.build_policy(w)
    a, n, r = policy.act(w.state[0], 0)
    assert a.shape == (w.dU,) and n.shape == (w.dS,) and np.ndim(r) == 0
    a2, n2, r2 = policy.act(w.state, 1)
    assert a2.shape == (3, w.dU) and n2.shape == (3, w.dS) and r2.shape == (3,)
    assert (a2 >= w.lb).all() and (a2 <= w.ub).all()
    policy.reset()
    a3, _, _ = policy.act(w.state, 2, exploration_noise=True)
    assert (a3 >= w.lb).all() and (a3 <= w.ub).all()
    policy.switch_optimizer(optimizer_name="RandomSearch", planning_horizon=10, population_size=128)
    a4, _, _ = policy.act(w.state, 3)
    assert a4.shape == (3, w.dU)


def test_shard_invariance_single_gpu(cuda_device):
    """Sharded (2 and 3 ranks emulated on one GPU through begin/iter_local/iter_merge/finish) equals
    unsharded: samples are keyed on the global row, merge is exact."""
    import ctypes as C
    from blackbox_mpc_b200 import _lib
    w = workloads.make("C2", population_size=301, num_agents=2, bias_scale=0.1)
    w.optimizer_args = dict(num_elite=16, alpha=0.25)
    ref_policy = workloads.build_policy(w, precision="fp32")
    ref_action, _, _ = ref_policy._optimizer(torch.from_numpy(w.state), 0, False)
    ref_mean = ref_policy._optimizer.get_tensor("mean").cpu()
    for world in (2, 3):
        pols = [workloads.build_policy(w, precision="fp32") for _ in range(world)]
        opts = [p._optimizer for p in pols]
        for r, o in enumerate(opts):
            o.shard(r, world)
            o._ensure_handle()
        lib = opts[0]._engine.lib
        dev = opts[0]._engine.device
        n = lib.bbmpc_opt_partial_floats(opts[0]._handle)
        gathered = torch.empty(world, n, device=dev)
        state = torch.from_numpy(w.state).to(dev)
        for o in opts:
            o._engine.check(lib.bbmpc_opt_begin(o._handle, _lib.ptr(state), 0, None))
        for it in range(lib.bbmpc_opt_num_iterations(opts[0]._handle)):
            for r, o in enumerate(opts):
                o._engine.check(lib.bbmpc_opt_iter_local(o._handle, it, gathered[r].data_ptr(), None))
            for o in opts:
                o._engine.check(lib.bbmpc_opt_iter_merge(o._handle, it, _lib.ptr(gathered), world, None))
        for o in opts:
            act = torch.empty(2, w.dU, device=dev)
            o._engine.check(lib.bbmpc_opt_finish(o._handle, 0, _lib.ptr(act), None, None, None))
            torch.cuda.synchronize()
            assert torch.equal(o.get_tensor("mean").cpu(), ref_mean), f"world={world}"
            assert torch.equal(act.cpu(), ref_action.cpu())


@pytest.mark.parametrize("opt_name,opt_args,H,A", [
    ("CEM", dict(num_elite=16, alpha=0.25), 30, 2),       # 2 * 16 * 32 floats: 16-byte vector path
    ("CEM", dict(num_elite=5, alpha=0.25), 15, 1),        # 5 * 17 = 85 floats: not a multiple of 4
    ("PI2", dict(lamda=1.0), 15, 3),                      # 3 * 17 = 51 floats
    ("RandomSearch", dict(), 13, 1),                      # 15 floats
])
def test_peer_memory_exchange_single_rank(cuda_device, opt_name, opt_args, H, A):
    """bbmpc_opt_p2p_export / _connect with world = 1 (the rank pulls its own message through the exchange
    buffer): publish -> flag wait -> gather -> merge must reproduce the plain path bit for bit, for message
    lengths that are and are not multiples of 4 floats (the parity copies are 16-byte aligned either way)."""
    import ctypes as C
    w = workloads.make("C2", population_size=300, num_agents=A, planning_horizon=H, bias_scale=0.1)
    w.optimizer_name, w.optimizer_args = opt_name, dict(opt_args)
    ref_policy = workloads.build_policy(w, precision="fp32")
    ref_action, ref_next, _ = ref_policy._optimizer(torch.from_numpy(w.state), 0, False)
    policy = workloads.build_policy(w, precision="fp32")
    opt = policy._optimizer
    e = opt._ensure_handle()
    ptr = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    e.check(e.lib.bbmpc_opt_p2p_export(opt._handle, handle, C.byref(ptr)))
    assert ptr.value
    ptrs = (C.c_void_p * 1)(ptr.value)
    e.check(e.lib.bbmpc_opt_p2p_connect(opt._handle, None, ptrs))
    for t in range(3):      # sequence numbers / buffer parity advance across act() calls (odd and even parity)
        action, nxt, _ = opt(torch.from_numpy(w.state), t, False)
        torch.cuda.synchronize()
        ref_action_t, _, _ = ref_policy._optimizer(torch.from_numpy(w.state), t, False) if t else (ref_action, None, None)
        assert torch.equal(action.cpu(), ref_action_t.cpu()), f"act() call {t}"


def test_sampler_distributions(cuda_device):
    """The in-kernel Philox samplers follow the TF distributions they stand for [TF]: tf.random.truncated_normal
    (cem.py:90-94: N(0,1) re-drawn until |z| <= 2, i.e. variance 1 - 4 phi(2) / (Phi(2) - Phi(-2)) = 0.7737) and
    tf.random.uniform (random_search.py:40-41).  First-iteration draws of a fresh optimizer: mean = midpoint,
    constrained variance = ((ub - lb) / 4)^2."""
    w = workloads.make("C4", population_size=512, bias_scale=0.1)      # lb = -1, ub = +1: sigma = 0.5
    _, _, samples, _, _, _ = _run_with_trace(w, "CEM")
    x = samples[0].double().flatten()                                   # 512 * 30 * 6 = 92 160 draws
    assert float(x.abs().max()) <= 1.0 + 1e-6                            # +-2 sigma support, no clipping needed
    assert abs(float(x.mean())) < 0.006
    assert abs(float(x.std()) - 0.5 * 0.7737 ** 0.5) < 0.004
    z = x / 0.5
    assert abs(float((z ** 4).mean()) / float((z ** 2).mean()) ** 2 - 2.3655) < 0.05   # kurtosis of N(0,1) truncated at 2 sigma
    # independent across rows and time steps
    s0 = samples[0][:, 0].double()
    assert abs(float(torch.corrcoef(torch.stack([s0[:, 0, 0], s0[:, 1, 0]]))[0, 1])) < 0.15
    assert abs(float(torch.corrcoef(torch.stack([s0[:-1, 0, 0], s0[1:, 0, 0]]))[0, 1])) < 0.15
    _, _, samples_u, _, _, _ = _run_with_trace(w, "RandomSearch")
    u = samples_u[0].double().flatten()
    assert float(u.min()) >= -1.0 and float(u.max()) < 1.0
    assert abs(float(u.mean())) < 0.008 and abs(float(u.var()) - 1.0 / 3.0) < 0.005


@pytest.mark.parametrize("opt_name", ["CEM", "PI2", "RandomSearch", "SPSA", "PSO"])
def test_graph_replay_equals_eager(cuda_device, monkeypatch, opt_name):
    """bbmpc_opt_call captures the kernels of an act() into a CUDA graph after two eager calls and replays it
    (optimizer_base.py:55-56: one graph launch per act() in the reference too).  Replays must draw fresh samples (device-side
    Philox act-call counter) and reproduce the eager path bit for bit, call after call, including the warm start."""
    w = workloads.make("C2", population_size=256, num_agents=2, bias_scale=0.1)
    state = torch.from_numpy(w.state)

    def run(n_calls):
        policy = workloads.build_policy(w, precision="fp32", optimizer_name=opt_name)
        opt = policy._optimizer
        if opt_name == "PSO":
            opt.reset()
        outs = []
        for t in range(n_calls):
            a, n, r = opt(state, t, False)
            outs.append((a.cpu().clone(), n.cpu().clone(), r.cpu().clone()))
        return outs, opt._engine.launch_count

    monkeypatch.setenv("BBMPC_NO_GRAPH", "1")
    eager, launches_eager = run(6)
    monkeypatch.delenv("BBMPC_NO_GRAPH")
    graphed, launches_graph = run(6)
    assert launches_eager == launches_graph            # kernels are counted inside the graph
    for t, (e, g) in enumerate(zip(eager, graphed)):
        for x, y in zip(e, g):
            assert torch.equal(x, y), f"act() call {t} differs between the eager and the replayed path"
    # consecutive calls differ (fresh draws), except for optimizers whose answer does not depend on the draws' index
    assert not torch.equal(graphed[3][0], graphed[4][0]) or opt_name == "SPSA"


@pytest.mark.parametrize("name,P,A", [("C2", 300, 2), ("C4", 1000, 1)])
def test_cem_fused_topk_refit_equals_split_kernels(cuda_device, monkeypatch, name, P, A):
    """Inside bbmpc_opt_call an unsharded CEM iteration selects the elites and refits mean / variance in ONE kernel
    (cem.py:98-125); the split kernels (local top-E message + merge / refit, the form every sharded run uses) must give the
    same bits: identical elite order, identical summation order."""
    w = workloads.make(name, population_size=P, num_agents=A, bias_scale=0.1)
    state = torch.from_numpy(w.state)

    def run():
        policy = workloads.build_policy(w, precision="fp32")
        opt = policy._optimizer
        outs = []
        for t in range(3):
            a, n, r = opt(state, t, False)
            outs.append((a.cpu().clone(), opt.get_tensor("mean").cpu().clone(), opt.get_tensor("variance").cpu().clone()))
        return outs, opt._engine.launch_count

    monkeypatch.setenv("BBMPC_NO_CEM_FUSE", "1")
    split, launches_split = run()
    monkeypatch.delenv("BBMPC_NO_CEM_FUSE")
    fused, launches_fused = run()
    assert launches_fused < launches_split
    for t, (s, f) in enumerate(zip(split, fused)):
        for x, y in zip(s, f):
            assert torch.equal(x, y), f"act() call {t}: fused and split CEM refit differ"


def test_cem_elites_follow_top_k_tie_rule(cuda_device, monkeypatch):
    """tf.nn.top_k takes the LOWEST indices among equal values (cem.py:98-101).  A reward quantised to integers makes most
    returns tie, also at the elite threshold (more ties than places left): the elite rows written by the selection kernel
    must be the first E of a stable sort by (return descending, population row ascending)."""
    from blackbox_mpc_b200.utils import rewards
    monkeypatch.setenv("BBMPC_NO_CEM_FUSE", "1")       # the split kernels leave the elite records in the message buffer
    monkeypatch.setenv("BBMPC_NO_GRAPH", "1")
    src = "__device__ float reward(const float* s, const float* a, const float* s2) { return floorf(3.0f * a[0]); }"
    P, E = 3000, 50
    w = workloads.make("C2", population_size=P, planning_horizon=4, bias_scale=0.1)
    w.max_iterations = 1
    w.optimizer_args = dict(num_elite=E, alpha=0.25)
    from blackbox_mpc_b200.policies.mpc_policy import MPCPolicy
    base = workloads.build_policy(w, precision="fp32")
    policy = MPCPolicy(reward_function=rewards.cuda_reward(src), env_action_space=base._optimizer._env_action_space,
                       env_observation_space=base._optimizer._env_observation_space,
                       dynamics_handler=base._trajectory_evaluator._system_dynamics_handler, optimizer_name="CEM",
                       num_agents=1, planning_horizon=4, population_size=P, max_iterations=1, num_elite=E, alpha=0.25)
    opt = policy._optimizer
    opt(torch.from_numpy(w.state), 0, False)
    returns = opt.get_tensor("returns").cpu().numpy().ravel()
    rec = opt.get_tensor("partial").cpu().numpy().reshape(E, -1)
    got_rows = rec[:, 1].view(np.int32)
    order = np.lexsort((np.arange(P), -returns))[:E]
    values, counts = np.unique(returns, return_counts=True)
    assert counts.max() > E                       # heavy ties ...
    assert (returns == returns[order[-1]]).sum() > (returns[order] == returns[order[-1]]).sum()   # ... also across the elite threshold
    np.testing.assert_array_equal(got_rows, order.astype(np.int32))
    np.testing.assert_array_equal(rec[:, 0], returns[order])
