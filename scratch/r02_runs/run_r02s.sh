#!/bin/bash
O=gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02s.log 2>&1; tail -3 $O/pytest_gpu_r02s.log
{
for cfg in "400 10 micro 2560 2160" "1500 10 micro 1280 1024" "400 10 micro 2304 2304" "1500 10 micro 1392 1040" "1000 10 micro 1920 1080" "400 10 mix 2304 2304" "400 10 low 2560 2160" "400 10 noise 2304 2304"; do echo "--- $cfg"; timeout 120 python scratch/enc_only.py $cfg 2>&1 | tail -3; done
echo "=== DBDE_B200_NO_LINEAR=1"
for cfg in "400 10 micro 2560 2160" "400 10 mix 2304 2304"; do echo "--- $cfg"; DBDE_B200_NO_LINEAR=1 timeout 120 python scratch/enc_only.py $cfg 2>&1 | tail -2; done
} > $O/camera_sizes_r02s.log 2>&1
cat $O/camera_sizes_r02s.log
