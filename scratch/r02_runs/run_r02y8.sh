#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for T in 8 8 8 6 6 12 12 8 8; do DBDE_B200_PROFILE=1 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T > /tmp/out.txt 2>&1; grep fps /tmp/out.txt | cut -c1-160; grep "h2d_wait_jobs" /tmp/out.txt | sed 's/.*sites/sites/' | sort -t= -k2 -n | tail -3 | cut -c1-260; done
