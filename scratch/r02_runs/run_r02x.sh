#!/bin/bash
cd "$(dirname "$0")/../.."
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
echo "--- copy threads 0, T=1";  DBDE_B200_COPY_THREADS=0 timeout 40 scratch/dropin_mt_b200 2048 2048 60 0 1;  echo "rc=$?"
echo "--- copy threads 0, T=4";  DBDE_B200_COPY_THREADS=0 timeout 40 scratch/dropin_mt_b200 2048 2048 60 0 4;  echo "rc=$?"
echo "--- copy threads 0, T=16"; DBDE_B200_COPY_THREADS=0 timeout 40 scratch/dropin_mt_b200 2048 2048 30 0 16; echo "rc=$?"
echo "--- copy threads 0, T=32"; DBDE_B200_COPY_THREADS=0 timeout 40 scratch/dropin_mt_b200 2048 2048 20 0 32; echo "rc=$?"
echo "--- default, T=32";        timeout 40 scratch/dropin_mt_b200 2048 2048 20 0 32; echo "rc=$?"
which gdb pstack eu-stack 2>/dev/null
