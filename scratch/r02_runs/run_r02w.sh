#!/bin/bash
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02w.log 2>&1; tail -3 $O/pytest_gpu_r02w.log
{
for cfg in "300 10 mix 2049 1003" "300 10 micro 4100 1003" "300 10 micro 2456 2054" "300 10 micro 3384 2704" "500 10 noise 2048 2048" "1000 10 mix 1001 1003" "1000 10 micro 2048 2048" "300 10 micro 1628 1236" "300 10 micro 1388 1040"; do echo "--- $cfg"; timeout 120 python scratch/enc_only.py $cfg 2>&1 | tail -2; done
} > $O/ab_r02w.log 2>&1
cat $O/ab_r02w.log
