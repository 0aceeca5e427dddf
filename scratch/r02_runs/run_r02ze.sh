#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_soak.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_soak || exit 1
timeout 80 scratch/dropin_soak 12 40
timeout 60 scratch/dropin_soak 3 15
DBDE_B200_COPY_THREADS=0 timeout 60 scratch/dropin_soak 20 15
