#!/bin/bash
# round-2 evidence run: GPU tests, smoke, bench lines per workload + reference arm, ncu launch list and --set full of the
# pipelined rollout (6 launches), compute-sanitizer memcheck + racecheck of one small C4 act()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2f_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; cut -c1-200 gpurun_out/r2_bench_c4.json
for wl in C1 C2 C3 C5; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 3 --cpu-budget 5 > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; cut -c1-120 gpurun_out/r2_bench_$wl.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cut -c1-120 gpurun_out/r2_bench_ref.json
BBMPC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu1.log 2>&1; echo "ncu launches rc=$?"
BBMPC_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_pipe -s 20 -c 6 -f -o gpurun_out/r2_pipe_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu2.log 2>&1; echo "ncu full rc=$?"
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool memcheck python tools/debug/san_act.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck.log
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool racecheck python tools/debug/san_act.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2_sanitizer_racecheck.log
BBMPC_TC_PIPE=0 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool memcheck python tools/debug/san_act.py > gpurun_out/r2_sanitizer_memcheck_tc.log 2>&1; echo "memcheck(tc) rc=$?"; tail -2 gpurun_out/r2_sanitizer_memcheck_tc.log
