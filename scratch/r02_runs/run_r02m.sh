#!/bin/bash
O=gpurun_out
{
for lib in dbce-video-cpp_b200/libdbde_b200.so scratch/libdbde_spread1.so scratch/libdbde_spread2.so; do
  echo "=== $lib"
  for cfg in "1000 10 mix 1001 1003" "1000 10 micro 1001 1003" "1000 10 micro 2048 2048" "300 10 low 4096 4096" "500 10 mix 2048 2048" "500 10 noise 2048 2048"; do echo "--- $cfg"; timeout 120 env DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -1; done
done
} > $O/ab_r02m.log 2>&1
cat $O/ab_r02m.log
