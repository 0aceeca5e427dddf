#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/.."
timeout 60 scratch/dma_probe > $O/dma_probe_r02k.txt 2>&1; cat $O/dma_probe_r02k.txt
LIBDIR=$PWD/dbce-video-cpp_b200
g++ -O2 -std=c++14 -pthread -Iinclude scratch/dropin_mt.cpp -L$LIBDIR -ldbde_b200 -Wl,-rpath,$LIBDIR -o scratch/dropin_mt_b200 || exit 1
run() { timeout 60 "$@" 2>&1 | sort | uniq -c | sort -rn | head -4 | cut -c1-200; }
echo "--- T=1 default"; DBDE_B200_PROFILE=1 run scratch/dropin_mt_b200 2048 2048 60 0 1
echo "--- T=1 h2d dma 4096"; DBDE_B200_H2D_DMA_KB=4096 DBDE_B200_PROFILE=1 run scratch/dropin_mt_b200 2048 2048 60 0 1
echo "--- T=1 h2d dma 512, d2h 512"; DBDE_B200_H2D_DMA_KB=512 DBDE_B200_D2H_DMA_KB=512 DBDE_B200_PROFILE=1 run scratch/dropin_mt_b200 2048 2048 60 0 1
echo "--- T=1 d2h dma 2048"; DBDE_B200_D2H_DMA_KB=2048 DBDE_B200_PROFILE=1 run scratch/dropin_mt_b200 2048 2048 60 0 1
for T in 4 8 16; do echo "--- T=$T default"; timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 $T; done
echo "--- T=16 crowd=100 (pool never stands back)"; DBDE_B200_COPY_CROWD=100 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 16
echo "--- T=16 copy threads 2"; DBDE_B200_COPY_THREADS=2 timeout 60 scratch/dropin_mt_b200 2048 2048 60 0 16
echo "--- T=16 noise"; timeout 60 scratch/dropin_mt_b200 2048 2048 60 1 16
echo "--- T=16 1001"; timeout 60 scratch/dropin_mt_b200 1001 1003 200 0 16
