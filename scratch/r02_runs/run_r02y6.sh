#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for T in 6 8 10 12; do
for c in 100 8 4; do for i in 1 2; do DBDE_B200_COPY_CROWD=$c timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T 2>&1 | sed "s/^/crowd=$c /" | cut -c1-200; done; done
done
for T in 8; do for i in 1 2; do DBDE_B200_COPY_THREADS=0 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T 2>&1 | sed "s/^/nopool /" | cut -c1-200; done; done
for T in 8; do for i in 1 2; do DBDE_B200_PROFILE=1 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T 2>&1 | grep -v "gpu:" | sort | head -5 | cut -c1-200; done; done
