#!/bin/bash
O=gpurun_out
cd "$(dirname "$0")/../.."
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02zc.log 2>&1; tail -3 $O/pytest_gpu_r02zc.log
{
for cfg in "300 10 micro 1388 1040" "300 10 micro 1628 1236" "1000 10 mix 1001 1003" "1000 10 micro 2048 2048" "1000 10 micro 1920 1080" "300 10 micro 1100 1000" "300 10 micro 1500 1000" "300 10 mix 1388 1040"; do echo "--- $cfg"; timeout 120 python scratch/enc_only.py $cfg 2>&1 | tail -2; done
} > $O/ab_r02zc.log 2>&1
cat $O/ab_r02zc.log
