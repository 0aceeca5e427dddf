#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02i.log 2>&1; tail -3 $O/pytest_gpu_r02i.log
cd "$(dirname "$0")/.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for T in 1 4 16; do
  echo "--- T=$T"; DBDE_B200_PROFILE=1 scratch/dropin_mt_b200 2048 2048 60 0 $T 2>&1 | sort | uniq -c | sort -rn | head -8
done
nproc; 
