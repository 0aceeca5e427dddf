#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02p.log 2>&1; tail -3 $O/pytest_gpu_r02p.log
{
for lib in dbce-video-cpp_b200/libdbde_b200.so scratch/libdbde_var3.so scratch/libdbde_copy8.so dbce-video-cpp_b200/libdbde_b200.so; do
  echo "=== $lib"
  for cfg in "1000 20 micro 2048 2048" "1000 10 mix 1001 1003" "300 10 low 4096 4096" "500 10 noise 2048 2048" "500 10 mix 2048 2048"; do echo "--- $cfg"; CHECK=0 timeout 120 env DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -1; done
done
} > $O/ab_r02p.log 2>&1
cat $O/ab_r02p.log
