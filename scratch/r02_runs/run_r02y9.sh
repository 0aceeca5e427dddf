#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for T in 8 8 8 6 6 12 12 16 16 4 4; do timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T 2>&1 | cut -c1-200; done
echo "--- turns off"
for T in 8 8 8 6 12 16 16 4; do DBDE_B200_SUBMIT_TURNS=0 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T 2>&1 | cut -c1-200; done
