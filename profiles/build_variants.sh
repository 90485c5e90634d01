#!/bin/bash
# usage: build_variants.sh NAME=FLAGS ...   e.g.  lead13=-DTT_TC4_LEAD=13 noW2=-DTT_ABLATE=1
# builds libtt_b200 variants that differ only in compile-time switches of the v4 actor kernel (TT_TC4_LEAD: sweep interleave,
# tt_tc4_layout.cuh; TT_ABLATE: timing-only ablations, tt_actor_tc4.cu) into profiles/bin/libtt_<NAME>.so; on the GPU box copy
# one over ddpg-trucktrailer_b200/libtt_b200.so before timing.
cd "$(dirname "$0")/.." || exit 1
C=ddpg-trucktrailer_b200/csrc
mkdir -p profiles/bin
for V in "$@"; do
  NAME=${V%%=*}; FLAGS=${V#*=}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=default --expt-relaxed-constexpr \
       $FLAGS -c $C/tt_actor_tc4.cu -o profiles/bin/tc4_$NAME.o &
done
wait
for V in "$@"; do
  NAME=${V%%=*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o profiles/bin/libtt_$NAME.so \
       $C/tt_lib.o $C/tt_env.o $C/tt_agent.o $C/tt_actor_tc.o profiles/bin/tc4_$NAME.o $C/tt_replay.o $C/tt_rollout.o -lcudart && echo built $NAME
done
