#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for i in 1 2 3 4 5 6; do timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 16 2>&1 | cut -c1-200; done
for i in 1 2 3; do DBDE_B200_COPY_CROWD=100 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 16 2>&1 | sed 's/^/crowd100 /' | cut -c1-200; done
for i in 1 2 3; do timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 8 2>&1 | cut -c1-200; done
uptime
