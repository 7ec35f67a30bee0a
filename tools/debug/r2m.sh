#!/bin/bash
# round-2 evidence run: full GPU tests, bench lines per workload, ncu --set full of the pipelined rollout (6 launches),
# compute-sanitizer memcheck + racecheck of one small C4 act()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2m_tests.log
timeout 300 python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; cut -c1-600 gpurun_out/r2_bench_c4.json
for wl in C1 C2 C3 C5; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --cpu-budget 5 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; cut -c1-400 gpurun_out/r2_bench_$wl.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-500 gpurun_out/r2_bench_ref.json
BBMPC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_pipe -s 20 -c 6 -f -o gpurun_out/r2_pipe_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
cat > /tmp/san.py <<'PY'
import numpy as np
from blackbox_mpc_b200.utils import workloads
w = workloads.make("C4", population_size=600, planning_horizon=6, bias_scale=0.1)
w.max_iterations = 2
p = workloads.build_policy(w)
for t in range(2):
    a, n, r = p.act(w.state[0], t)
print("kernel", p._optimizer._engine.last_rollout_kernel, a)
PY
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 900 compute-sanitizer --tool memcheck python /tmp/san.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 900 compute-sanitizer --tool racecheck python /tmp/san.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck.log
BBMPC_TC_PIPE=0 BBMPC_NO_GRAPH=1 timeout 900 compute-sanitizer --tool memcheck python /tmp/san.py > gpurun_out/r2_sanitizer_memcheck_tc.log 2>&1; echo "memcheck(tc) rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_tc.log
