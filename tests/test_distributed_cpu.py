"""world_size-2 (and 3) gloo tests of the N>1 host path on CPU: shard ranges, the per-iteration
all_gather of fixed-size partial messages (blackbox_mpc_b200.sharding, the function the optimizers
call between bbmpc_opt_iter_local and bbmpc_opt_iter_merge), and the merge — with the CPU
restatement standing in for the two GPU halves.  Result must equal the unsharded optimizer run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

F64 = torch.float64


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cem_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import helpers
    import oracle
    from oracle import sharded
    from blackbox_mpc_b200 import sharding
    from blackbox_mpc_b200.utils import workloads
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    try:
        P, A, H, E, iters = 101, 2, 8, 12, 3
        w = workloads.make("C2", population_size=P, planning_horizon=H, num_agents=A, bias_scale=0.1)
        w.optimizer_args = dict(num_elite=E, alpha=0.25)
        ev = helpers.oracle_evaluator(w, F64)
        state = torch.from_numpy(w.state).double()
        # unsharded run (every rank computes it; rank 0 compares)
        ref = helpers.oracle_optimizer(w, "CEM", dtype=F64, max_iterations=iters)
        draws = oracle.TorchDraws(seed=5, dtype=F64)
        ref_action = ref._optimize(state, 0, draws)
        samples_per_iter = [t["samples"] for t in ref.trace]      # draws keyed on the GLOBAL row: same on all ranks
        # sharded run
        p0, p1 = sharding.shard_range(P, rank, world)
        n = sharding.partial_floats("CEM", A, H, w.dU, num_elite=E)
        mean, var = ref._midpoint().reshape(A, -1).clone(), ref._init_variance().reshape(A, -1).clone()
        gather = torch.empty(world, n, dtype=F64)
        for it in range(iters):
            local = samples_per_iter[it][p0:p1]
            returns = ev(state, local, 0)                                                   # "iter_local"
            partial = sharded.cem_partial(local, returns, p0, E).reshape(-1).contiguous()
            assert partial.numel() == n
            sharding.all_gather_partials(partial, gather)                                   # the collective
            mean, var = sharded.cem_merge(gather.reshape(world, A, E, -1), E, mean, var, 0.25)  # "iter_merge"
        got_action = mean.reshape(A, H, w.dU)[:, 0]
        np.testing.assert_allclose(got_action.numpy(), ref_action.numpy(), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(mean.numpy(), ref.trace[-1]["mean"].reshape(A, -1).numpy(), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(var.numpy(), ref.trace[-1]["variance"].reshape(A, -1).numpy(), rtol=1e-12, atol=1e-13)
        # bit-identical on every rank (identical merge arithmetic on identical gathered bytes)
        probe = [torch.empty_like(mean) for _ in range(world)]
        dist.all_gather(probe, mean)
        assert all(torch.equal(probe[0], q) for q in probe)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_cem_over_gloo_equals_unsharded(world, tmp_path):
    mp.spawn(_cem_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok{r}" for r in range(world)]


def _bench_reference_worker(rank, world, port, out_dir):
    """bench.py --impl reference under a 2-rank launch: rank 0 alone prints, the others exit 0."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(port))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", str(world),
                        "--steps", "1", "--warmup", "0", "--workload", "C2", "--cpu-budget", "2"],
                       capture_output=True, text=True, env=env, timeout=300)
    with open(os.path.join(out_dir, f"out{rank}"), "w") as f:
        f.write(f"{r.returncode}\n{r.stdout}")


def test_bench_reference_arm_rank0_only(tmp_path):
    import json
    mp.spawn(_bench_reference_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    rc0, out0 = open(tmp_path / "out0").read().split("\n", 1)
    rc1, out1 = open(tmp_path / "out1").read().split("\n", 1)
    assert rc0 == "0" and rc1 == "0" and out1.strip() == ""
    line = json.loads(out0.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["n_gpus"] == 2


def test_sharded_cmaes_selection_equals_unsharded():
    """CMA-ES exchange (csrc/cmaes.cu: local top-E records -> exact global top-E): for any shard count the merged
    elite rows, in rank order, are the first E rows of the unsharded descending argsort, ties to the lower row."""
    import torch
    from oracle import sharded
    from blackbox_mpc_b200.sharding import shard_range, partial_floats
    g = torch.Generator().manual_seed(3)
    P, N, E = 203, 12, 16
    x = torch.randn(P, N, generator=g, dtype=torch.float64)
    rewards = torch.randn(P, generator=g, dtype=torch.float64)
    rewards[17] = rewards[101] = rewards[5] = rewards.max() + 1.0        # a three-way tie at the top
    ref_order = torch.sort(-rewards, stable=True).indices[:E]
    assert ref_order[:3].tolist() == [5, 17, 101]
    for world in (1, 2, 3, 8):
        parts = []
        for r in range(world):
            p0, p1 = shard_range(P, r, world)
            parts.append(sharded.cmaes_partial(x[p0:p1], rewards[p0:p1], p0, E))
        assert parts[0].numel() == partial_floats("CMA-ES", 1, 2, 6, num_elite=E)   # N = A*H*dU = 12
        rows, x_sorted = sharded.cmaes_merge(torch.stack(parts), E)
        assert rows.tolist() == ref_order.tolist(), world
        assert torch.equal(x_sorted, x[ref_order])
