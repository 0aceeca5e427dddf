#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for i in 1 2; do DBDE_B200_PROFILE=1 timeout 40 scratch/dropin_mt_b200 2048 2048 100 0 1 2>&1 | cut -c1-260; done
DBDE_B200_H2D_DMA_KB=4096 DBDE_B200_PROFILE=1 timeout 40 scratch/dropin_mt_b200 2048 2048 100 0 1 2>&1 | cut -c1-260
DBDE_B200_PROFILE=1 timeout 40 scratch/dropin_mt_b200 1001 1003 200 0 1 2>&1 | cut -c1-260
nvidia-smi --query-gpu=clocks.sm,clocks.mem --format=csv
