#!/bin/bash
# builds libtt_b200 variants that differ only in the v4 actor's sweep interleave (TT_TC4_LEAD, tt_tc4_layout.cuh) into
# profiles/bin/libtt_lead<L>.so; on the GPU box copy one over ddpg-trucktrailer_b200/libtt_b200.so before timing.
cd "$(dirname "$0")/.." || exit 1
C=ddpg-trucktrailer_b200/csrc
mkdir -p profiles/bin
for L in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=default --expt-relaxed-constexpr \
       -DTT_TC4_LEAD=$L -c $C/tt_actor_tc4.cu -o profiles/bin/tc4_lead$L.o &
done
wait
for L in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o profiles/bin/libtt_lead$L.so \
       $C/tt_lib.o $C/tt_env.o $C/tt_agent.o $C/tt_actor_tc.o profiles/bin/tc4_lead$L.o $C/tt_replay.o $C/tt_rollout.o -lcudart && echo built lead $L
done
