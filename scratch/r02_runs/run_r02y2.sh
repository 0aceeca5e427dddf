#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
for v in 1 0 1 0; do echo "--- streaming=$v"; DBDE_B200_H2D_STREAMING=$v DBDE_B200_PROFILE=1 timeout 40 scratch/dropin_mt_b200 2048 2048 150 0 1 2>&1 | cut -c1-260; done
for v in 1 0; do echo "--- streaming=$v T=16"; DBDE_B200_H2D_STREAMING=$v timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 16 2>&1 | cut -c1-260; done
for v in 1 0; do echo "--- streaming=$v T=4"; DBDE_B200_H2D_STREAMING=$v timeout 60 scratch/dropin_mt_b200 2048 2048 100 0 4 2>&1 | cut -c1-260; done
