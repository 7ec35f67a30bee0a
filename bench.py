#!/usr/bin/env python
"""bench.py — MPC steps/sec on HalfCheetah CEM (pop 10 000, H 30, 5x(3x200) MLP ensemble, 5 iters).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C4]

One "step" = one MPCPolicy.act(): 5 CEM iterations of (sample -> fused rollout -> top-k refit) plus
the executed-action predict/reward tail (reference: policies/mpc_policy.py:124-172).  Prints ONE
JSON line (rank 0).  `value` is timed with CUDA events with the observation already on the device
(bbmpc_opt_call); `e2e` goes through MPCPolicy.act with numpy buffers (H2D + D2H inside the timed
region).  `roofline` is the rollout kernel's algorithmic FLOP rate (CUDA events recorded around
every rollout launch, on its stream, inside the timed region) against the measured dense-bf16 peak
of MEASURED_PEAKS.json.  `cpu_baseline` / `--impl reference` time the CPU restatement of the
reference's TF2 graph (oracle/, torch-CPU fp32, all host threads): TensorFlow 2.0 itself cannot be
installed in this image, see DESIGN.md.

For N>1 launch under torchrun (one rank per GPU): the population is sharded over the ranks; per CEM
iteration the ranks exchange their elite records through peer memory (CUDA IPC buffers pulled over NVLink
inside the merge-side kernel; BBMPC_P2P=0 or a failed IPC set-up falls back to one NCCL all_gather).
`run.exchange` names the path used and `run.ranks_agree` whether all ranks computed the same action; `config` is
the workload only and is identical in both arms.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
print_json = print
METRIC = "mpc_steps_per_sec"
UNIT = "steps/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    except Exception:
        d = dict(FALLBACK_PEAKS)
        d["_source"] = "fallback"
        return d


def workload_config(w, extra=None):
    cfg = {
        "workload": f"{w.name}: HalfCheetah {w.optimizer_name}" if w.reward == "halfcheetah" else f"{w.name}: Pendulum {w.optimizer_name}",
        "population_size": w.population_size, "planning_horizon": w.planning_horizon,
        "max_iterations": w.max_iterations, "num_agents": w.num_agents,
        "dynamics": (f"{w.n_members}x MLP {w.layers}" if w.dynamics == "mlp" else w.dynamics),
        "dS": w.dS, "dU": w.dU,
    }
    cfg.update(w.optimizer_args)
    # identical in both arms (the driver compares the two dicts); what differs per arm is reported under "run"
    cfg["l2"] = "GPU arm: flushed (256 MB write) before every timed step"
    cfg["parallelism"] = "GPU arm: population rows sharded over n_gpus; reference arm: all host threads of rank 0"
    if extra:
        cfg.update(extra)
    return cfg


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    _REASONS = {
        0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
        0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------- CPU arm
def cpu_reference_steps_per_sec(w, steps: int, warmup: int, budget_s: float):
    """Times the CPU restatement of the reference's TF2 graph (oracle/, unfused torch-CPU fp32 ops,
    all host threads).  Each step is one act() on a population sample sized to the time budget; a
    sample of P_s rows out of P is scaled as steps/s = (P_s / P) / t (work is linear in P)."""
    import copy
    import numpy as np
    import torch
    import oracle
    from oracle import build as ob
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = w.population_size

    def time_act(ws, n):
        opt = ob.optimizer(ws, dtype=torch.float32)
        draws = oracle.TorchDraws(seed=0, dtype=torch.float32)
        ts = []
        for i in range(n):
            t0 = time.perf_counter()
            oracle.policy_act(opt, ws.state, i, draws)
            ts.append(time.perf_counter() - t0)
        return ts

    # calibrate on a small sample
    cal = copy.copy(w)
    cal.population_size = max(min(P, 500), w.optimizer_args.get("num_elite", 1))
    t_cal = min(time_act(cal, 2))
    per_row = t_cal / cal.population_size
    n_total = max(1, steps + warmup)
    rows = int(budget_s / n_total / per_row)
    rows = max(min(P, rows), cal.population_size)
    if rows < P:
        rows = max(cal.population_size, rows // 100 * 100)
    ws = copy.copy(w)
    ws.population_size = rows
    ts = time_act(ws, n_total)[warmup:]
    t = statistics.median(ts)
    value = (rows / P) / t
    sample = (f"{len(ts)} act() call(s) of the CPU restatement of the reference TF2 graph (oracle/, torch-CPU fp32, "
              f"{cores} threads) on {rows} of {P} population rows, all {w.max_iterations} iterations, H={w.planning_horizon}; "
              f"median {t:.3f} s per call" + ("" if rows == P else "; scaled linearly in rows"))
    return value, cores, sample, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from blackbox_mpc_b200.utils import workloads
    w = workloads.make(args.workload)
    value, cores, sample, t = cpu_reference_steps_per_sec(w, args.steps, args.warmup, args.cpu_budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(json.dumps(line))


# ----------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from blackbox_mpc_b200 import _lib
    from blackbox_mpc_b200.utils import workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a): the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        sys.stderr.write(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE\n")

    w = workloads.make(args.workload, population_size=args.population)
    policy = workloads.build_policy(w, precision=args.precision)
    opt = policy._optimizer
    if world > 1:
        opt.shard(rank, world, group=None)
    engine = policy._trajectory_evaluator.engine()
    eff_prec = engine.effective_precision
    obs = w.state.astype(np.float32)                      # [A, dS] host
    d_state = torch.from_numpy(obs).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step(t):
        return opt(d_state, t, False)

    def host_step(t):
        return policy.act(obs, t)

    def timed(fn, steps, warmup, profile=False):
        for i in range(warmup):
            fn(i)
        barrier()
        launches0 = engine.launch_count
        if profile:
            engine.profile_enable(True)
        evs = []
        for i in range(steps):
            flush.zero_()                                  # evict L2 between timed steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(warmup + i)
            e1.record()
            evs.append((e0, e1))
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        kern = engine.profile_read() if profile else (0.0, 0)
        if profile:
            engine.profile_enable(False)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), engine.launch_count - launches0, kern

    with ClockSampler(local_rank) as clocks:
        # one act() = one CUDA-graph launch (bbmpc_opt_call replays its captured kernels; sharded runs stay eager)
        total_ms, launches, _ = timed(device_step, args.steps, args.warmup)
        # end to end through the public API: host numpy in, host numpy out, every step
        e2e_ms, _, _ = timed(host_step, args.steps, max(3, args.warmup // 2))
        # rollout-kernel time for the roofline: a separate profiled pass (event pairs around every rollout launch on its
        # stream; profiling runs the act() eagerly, the kernel itself is the same)
        prof_steps = min(args.steps, 20)
        _, _, (kern_ms, kern_n) = timed(device_step, prof_steps, 3, profile=True)
    exchange = "none (1 GPU)"
    ranks_agree = None
    if world > 1:
        exchange = "peer memory (CUDA IPC, NVLink P2P loads in the merge-side kernel)" if getattr(opt, "_p2p", False) else "NCCL all_gather per iteration"
        a_me, _, _ = opt(d_state, 10 ** 6, False)
        a_all = torch.empty(world, *a_me.shape, device=dev)
        dist.all_gather_into_tensor(a_all, a_me.contiguous())
        ranks_agree = bool((a_all == a_all[0:1]).all().item())
    A = w.num_agents
    h2d = A * w.dS * 4
    d2h = (A * w.dU + A * w.dS + A) * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step
    rows_local = opt.get_tensor("returns").numel()         # rows this rank rolls out per iteration
    flops_per_launch = w.flops_per_row_step() * rows_local * w.planning_horizon
    roofline = None
    if kern_n > 0 and flops_per_launch > 0:
        avg_ms = kern_ms / kern_n
        achieved = flops_per_launch / (avg_ms * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        traffic = None
        tp = os.path.join(ROOT, "profiles", "rollout_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": engine.last_rollout_kernel,
            "kernel_ms_avg": avg_ms, "kernel_launches": kern_n, "kernel_launches_per_step": kern_n / prof_steps,
            "kernel_share_of_step": (kern_ms / prof_steps) / (total_ms / args.steps),
            "algorithmic_flops_per_launch": flops_per_launch, "peak_source": f"{peaks['_source']} bf16_tflops_sustained (dense cuBLAS bf16)",
            "note": ("operands are split into bf16 hi+lo and contracted in 3 tensor-core passes with fp32 accumulation "
                     "(fp32-grade parity with the reference); algorithmic FLOPs are counted once, so the ceiling of frac is 1/3"
                     if eff_prec == "bf16x3" else f"precision {eff_prec}"),
            "hbm_algorithmic_bytes_per_launch": rows_local * (w.planning_horizon * w.dU * 4 + 4),
        }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(w),
        "run": {"precision": eff_prec, "exchange": exchange, "ranks_agree": ranks_agree},
        "clocks": clocks.summary(),
        "e2e": {"value": 1e3 / (e2e_ms / args.steps), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "api": "MPCPolicy.act(numpy obs) -> numpy (action, next_obs, reward)"},
        "gpu_launches": int(launches),
        "act_launch": ("one CUDA-graph launch per act() (kernels counted inside the graph)" if world == 1 and not os.environ.get("BBMPC_NO_GRAPH")
                       else "eager kernel launches"),
        "roofline": roofline,
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _ = cpu_reference_steps_per_sec(w, 1, 0, args.cpu_budget)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    elif world > 1:
        line["cpu_baseline"] = None
    print_json(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _json_only_stdout():
    """The driver reads ONE JSON line from stdout; libraries (NCCL's version banner, torch warnings) may
    print there too.  Point fd 1 at stderr for the whole run and return a writer for the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return lambda text: os.write(real, (text + "\n").encode())


def main():
    emit = _json_only_stdout()
    global print_json
    print_json = emit
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--population", type=int, default=None, help="override population (parity/debug only)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (ncu runs)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the CPU arm")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.cpu_budget == 20.0:
            args.cpu_budget = 120.0
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
