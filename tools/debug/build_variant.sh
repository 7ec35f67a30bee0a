#!/bin/bash
# build_variant.sh <suffix> <extra nvcc flags...>: alternative libbbmpc_<suffix>.so for A/B runs through BBMPC_LIB
set -e
SUF=$1; shift
D=blackbox_mpc_b200/csrc; O=/tmp/bbmpc_var_$SUF; mkdir -p $O
for f in context rollout_simt rollout_tc rollout_pipe optimizers cmaes user_reward; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -I include -c $D/$f.cu -o $O/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o blackbox_mpc_b200/libbbmpc_$SUF.so $O/*.o -Xlinker --exclude-libs,ALL -cudart static -ldl
echo built blackbox_mpc_b200/libbbmpc_$SUF.so
