"""Population sharding for one-process-per-GPU runs (the reference is single-device).

Rank g of G evaluates the contiguous global population rows [g*P//G, (g+1)*P//G) — the same split
bbmpc_opt_set_shard computes on the C side (csrc/optimizers.cu set_shard).  Per optimizer
iteration every rank reduces its slice to one fixed-size fp32 "partial" message; the messages of
all ranks are all-gathered (the only data-path collective) and merged with identical arithmetic on
every rank.  Pure host logic: importable and testable without a GPU (gloo)."""
from __future__ import annotations

from typing import Tuple


def shard_range(population_size: int, rank: int, world: int) -> Tuple[int, int]:
    """[p0, p1) of the global population owned by `rank`; sizes differ by at most one row."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad shard {rank}/{world}")
    return population_size * rank // world, population_size * (rank + 1) // world


def partial_floats(kind: str, num_agents: int, planning_horizon: int, dim_u: int, num_elite: int = 0) -> int:
    """Floats per rank per iteration (mirror of partial_floats() in csrc/optimizers.cu)."""
    hu = planning_horizon * dim_u
    if kind == "CEM":
        return num_agents * num_elite * (2 + hu)      # E x (reward, global row, sequence)
    if kind in ("PI2", "RandomSearch", "PSO"):
        return num_agents * (2 + hu)                  # (max | best, sum | row, weighted | best sequence)
    if kind == "SPSA":
        return num_agents * hu                        # gradient partial sums
    if kind == "CMA-ES":
        return num_elite * (2 + num_agents * hu)      # E x (summed reward, global row, x[N])
    raise ValueError(kind)


def all_gather_partials(partial, gather_buf, group=None):
    """The per-iteration exchange: gather_buf[g] <- rank g's partial (torch.distributed; NCCL over
    NVLink on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(gather_buf.view(-1), partial.view(-1), group=group)  # flat: concatenation along dim 0
    return gather_buf
