#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/../.."
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02zi.log 2>&1; tail -2 $O/pytest_gpu_r02zi.log
for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "400 10 micro 2304 2304" "1000 10 micro 2048 2048"; do echo "--- $cfg"; timeout 120 python scratch/enc_only.py $cfg 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dbde_decode_scan -c 3 python scratch/enc_only.py 1000 1 mix 1001 1003 2>&1 | grep -E "gpu__time|dbde_decode_scan" | head -6
