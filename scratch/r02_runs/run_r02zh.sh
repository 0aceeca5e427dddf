#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/../.."
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02zh.log 2>&1; tail -3 $O/pytest_gpu_r02zh.log
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_soak.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_soak || exit 1
for T in 1 1 4 8 16; do timeout 60 scratch/dropin_mt_b200 2048 2048 100 0 $T 2>&1 | cut -c1-200; done
timeout 60 scratch/dropin_mt_b200 2048 2048 100 1 1
timeout 60 scratch/dropin_mt_b200 1001 1003 200 0 1
DBDE_B200_PROFILE=1 timeout 60 scratch/dropin_mt_b200 2048 2048 100 0 1 2>&1 | grep "gpu:" | cut -c1-220
timeout 60 scratch/dropin_soak 10 20
