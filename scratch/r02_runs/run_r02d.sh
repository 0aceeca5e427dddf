#!/bin/bash
O=gpurun_out
{
for lib in scratch/libdbde_head.so dbce-video-cpp_b200/libdbde_b200.so scratch/libdbde_head.so dbce-video-cpp_b200/libdbde_b200.so; do
  echo "=== $lib"
  for cfg in "300 10 low 4096 4096" "1000 10 micro 2048 2048"; do echo "--- $cfg"; DBDE_B200_LIB=$lib python scratch/enc_only.py $cfg 2>&1 | tail -2; done
done
} > $O/ab_r02d.log 2>&1
for path in direct staged; do
DBDE_B200_ODD_DECODE=$path ncu --set full --import-source on --clock-control none -k regex:dbde_ -c 3 -f -o $O/prof_r02d_mix1001_$path python scratch/enc_only.py 1000 1 mix 1001 1003 > $O/ncu_full_r02d_$path.log 2>&1
done
cat $O/ab_r02d.log
