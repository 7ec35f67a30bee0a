"""CPU restatement of the sharded-population protocol — TEST INFRASTRUCTURE.

The reference is single-device; sharding is this repo's own design, so what is pinned here is the
invariant "sharded == unsharded": the per-rank partial message + merge of csrc/optimizers.cu
(topk_partial_kernel / cem_refit_kernel, pi2_partial_kernel / pi2_merge_kernel) restated in
torch-CPU, checked against the plain single-device restatement of optimizers/cem.py:98-125 and
optimizers/pi2.py:79-87 in reference_port.py."""
import torch


def cem_partial(samples_local, returns_local, p0, num_elite):
    """samples_local [Pl,A,H,dU], returns_local [Pl,A] -> [A, E, 2+HU] records
    (reward, global row as float64-exact integer, sequence), sorted (reward desc, row asc).
    Slots beyond Pl hold reward -inf."""
    Pl, A = returns_local.shape
    HU = samples_local.shape[2] * samples_local.shape[3]
    out = torch.zeros(A, num_elite, 2 + HU, dtype=samples_local.dtype)
    out[:, :, 0] = -float("inf")
    out[:, :, 1] = float(2 ** 31 - 1)
    for a in range(A):
        order = torch.sort(-returns_local[:, a], stable=True).indices[:num_elite]
        k = order.numel()
        out[a, :k, 0] = returns_local[order, a]
        out[a, :k, 1] = (order + p0).to(out.dtype)
        out[a, :k, 2:] = samples_local[order, a].reshape(k, HU)
    return out


def cem_merge(partials, num_elite, mean, var, alpha):
    """partials [G, A, E, 2+HU] -> global top-E (reward desc, global row asc), mean / ddof-0
    variance over the elites, alpha blend (cem.py:112-125).  mean, var [A, HU]."""
    G, A, E, rec = partials.shape
    cand = partials.permute(1, 0, 2, 3).reshape(A, G * E, rec)
    new_mean, new_var = torch.empty_like(mean), torch.empty_like(var)
    for a in range(A):
        r, row = cand[a, :, 0], cand[a, :, 1]
        # lexicographic (reward desc, row asc): stable sort by row, then stable sort by -reward
        o1 = torch.sort(row, stable=True).indices
        o2 = torch.sort(-r[o1], stable=True).indices
        sel = o1[o2][:num_elite]
        el = cand[a, sel, 2:]
        nm = el.mean(dim=0)
        nv = ((el - nm) ** 2).mean(dim=0)
        new_mean[a] = alpha * mean[a] + (1 - alpha) * nm
        new_var[a] = alpha * var[a] + (1 - alpha) * nv
    return new_mean, new_var


def pi2_partial(samples_local, rewards_local, lamda):
    """-> [A, 2+HU]: (max reward, sum_p e_p, sum_p e_p * x_p), e_p = exp((r_p - max)/lambda)."""
    Pl, A = rewards_local.shape
    HU = samples_local.shape[2] * samples_local.shape[3]
    out = torch.zeros(A, 2 + HU, dtype=samples_local.dtype)
    for a in range(A):
        mx = rewards_local[:, a].max()
        e = torch.exp((rewards_local[:, a] - mx) / lamda)
        out[a, 0], out[a, 1] = mx, e.sum()
        out[a, 2:] = (e[:, None] * samples_local[:, a].reshape(Pl, HU)).sum(dim=0)
    return out


def pi2_merge(partials, lamda):
    """partials [G, A, 2+HU] -> new mean [A, HU] (log-sum-exp merge of the ranks)."""
    gmax = partials[:, :, 0].max(dim=0).values                     # [A]
    scale = torch.exp((partials[:, :, 0] - gmax[None, :]) / lamda)  # [G, A]
    eta = (partials[:, :, 1] * scale).sum(dim=0)
    acc = (partials[:, :, 2:] * scale[:, :, None]).sum(dim=0)
    return acc / eta[:, None]


def cmaes_partial(x_local, rewards_local, p0, num_elite):
    """x_local [Pl, N] (clipped samples), rewards_local [Pl] (summed over agents, cma_es.py:158) ->
    [E, 2+N] records (reward, global row, x), sorted (reward desc, row asc); unused slots hold -inf."""
    Pl, N = x_local.shape
    out = torch.zeros(num_elite, 2 + N, dtype=x_local.dtype)
    out[:, 0] = -float("inf")
    out[:, 1] = float(2 ** 31 - 1)
    order = torch.sort(-rewards_local, stable=True).indices[:num_elite]
    k = order.numel()
    out[:k, 0] = rewards_local[order]
    out[:k, 1] = (order + p0).to(out.dtype)
    out[:k, 2:] = x_local[order]
    return out


def cmaes_merge(partials, num_elite):
    """partials [G, E, 2+N] -> the E globally best rows in rank order (reward desc, global row asc): exactly the
    first E rows of the reference's full argsort (cma_es.py:159), which are the only ones with non-zero
    recombination weight (:62-68).  Returns (rows [E] int64, x_sorted [E, N])."""
    G, E, rec = partials.shape
    cand = partials.reshape(G * E, rec)
    o1 = torch.sort(cand[:, 1], stable=True).indices
    o2 = torch.sort(-cand[o1, 0], stable=True).indices
    sel = o1[o2][:num_elite]
    return cand[sel, 1].to(torch.int64), cand[sel, 2:]
