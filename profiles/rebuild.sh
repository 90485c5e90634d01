#!/bin/bash
# rebuild libtt_b200.so from anywhere and print the tensor-core kernels' register / spill lines
cd "$(dirname "$0")/.." && python -c "
import sys; sys.path.insert(0,'.')
from ddpg_trucktrailer_b200 import build; print(build.build(force=True))" 2>&1 | grep -i "error\|libtt"
grep -A3 "actor_tc4" ddpg-trucktrailer_b200/csrc/ptxas.log | grep -i "used\|spill" | head -4
